#!/bin/bash
for c in 2 1; do
  echo "CTAS_PER_SM=$c"
  PSOAP_CTAS_PER_SM=$c timeout 100 python tools/bench_syrk.py 4096 10
  PSOAP_CTAS_PER_SM=$c timeout 100 python tools/bench_syrk.py 8192 10
  PSOAP_CTAS_PER_SM=$c timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(' farm value=%.3f e2e=%.3f ms=%.1f step_tflops=%.2f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['step_tflops_per_gpu']))"
  PSOAP_CTAS_PER_SM=$c timeout 200 python tools/time_lnlike.py 2>&1 | tail -3 | cut -c1-110
done
