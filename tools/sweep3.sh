#!/bin/bash
for nb in 32 48 64; do for c in 2 1; do
  echo "nbranch=$nb ctas_per_sm=$c"
  PSOAP_CTAS_PER_SM=$c timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --nbranch $nb 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(' value=%.3f e2e=%.3f ms=%.1f step_tflops=%.2f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['step_tflops_per_gpu']))"
done; done
