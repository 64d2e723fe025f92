#!/bin/bash
set -x
mkdir -p gpurun_out
for t in 0 1; do for m in 2048 4096 6144 8192; do PSOAP_TAIL=$t timeout 120 python tools/bench_syrk.py $m 20 512; done; done > gpurun_out/r2_tail_syrk.txt 2>&1
cat gpurun_out/r2_tail_syrk.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_tail_tests.log 2>&1
tail -5 gpurun_out/r2_tail_tests.log
for t in 0 1; do PSOAP_TAIL=$t timeout 300 python tools/time_lnlike.py; done > gpurun_out/r2_tail_lnlike.txt 2>&1
cat gpurun_out/r2_tail_lnlike.txt
for t in 0 1; do PSOAP_FARM_TAIL=$t timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | head -c 260; echo; done > gpurun_out/r2_tail_farm.txt 2>&1
cat gpurun_out/r2_tail_farm.txt
