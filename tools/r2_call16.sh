#!/bin/bash
# round 2: farm priorities A/B on one rank's share of an 8-GPU run, cost-model fit, C6 workload, full GPU test suite
set -x
mkdir -p gpurun_out
for prio in 0 1; do PSOAP_FARM_PRIO=$prio python tools/farm_subset_time.py 8 32; PSOAP_FARM_PRIO=$prio python tools/farm_subset_time.py 1 32; done > gpurun_out/r2_prio_ab.txt 2>&1
cat gpurun_out/r2_prio_ab.txt
python tools/farm_cost_fit.py > gpurun_out/r2_cost_fit.txt 2>&1
cat gpurun_out/r2_cost_fit.txt
python bench.py --workload C6 --steps 5 --warmup 3 > gpurun_out/r2_bench_c6.json 2> gpurun_out/r2_bench_c6.err
head -c 700 gpurun_out/r2_bench_c6.json; tail -c 300 gpurun_out/r2_bench_c6.err
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_tests_full.log 2>&1
tail -5 gpurun_out/r2_tests_full.log
