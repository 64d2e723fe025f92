#!/bin/bash
# round 2, GPU call 3: potrf_diag7 + trsm7 lab, parity tests and timings with the new chain as default
set -x
mkdir -p gpurun_out
timeout 120 tools/potrf7_lab.bin > gpurun_out/r2_potrf7_lab3.txt 2>&1
cat gpurun_out/r2_potrf7_lab3.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "lnlike_golden or tile_boundaries or predict_golden or farm_vs_oracle or vs_reference_cpu or package_default or lnlike_vs_oracle or calibration or repeatable" > gpurun_out/r2_tests5.log 2>&1
tail -5 gpurun_out/r2_tests5.log
python tools/time_lnlike.py > gpurun_out/r2_time_lnlike_c7b.txt 2>&1
cat gpurun_out/r2_time_lnlike_c7b.txt
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c7b.json 2> gpurun_out/r2_bench_c7.err
head -c 300 gpurun_out/r2_bench_c7b.json
