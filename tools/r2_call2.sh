#!/bin/bash
# round 2, GPU call 2: potrf_diag7 lab, then parity tests and single-chunk timings with PSOAP_POTRF=7
set -x
mkdir -p gpurun_out
timeout 120 tools/potrf7_lab.bin > gpurun_out/r2_potrf7_lab.txt 2>&1
cat gpurun_out/r2_potrf7_lab.txt
PSOAP_POTRF=7 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "lnlike_golden or tile_boundaries or predict_golden or farm_vs_oracle or vs_reference_cpu or package_default or lnlike_vs_oracle" > gpurun_out/r2_tests2.log 2>&1
tail -5 gpurun_out/r2_tests2.log
PSOAP_POTRF=7 python tools/time_lnlike.py > gpurun_out/r2_time_lnlike_p7.txt 2>&1
cat gpurun_out/r2_time_lnlike_p7.txt
PSOAP_POTRF=7 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_p7.json 2> gpurun_out/r2_bench_p7.err
head -c 400 gpurun_out/r2_bench_p7.json
ncu --set full --clock-control none --import-source on -k regex:fill_ -c 6 -o gpurun_out/r2_fill2 python tools/fill_once.py 300 1 > gpurun_out/r2_ncu_fill2.log 2>&1
