#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2_tests_full2.log 2>&1
tail -5 gpurun_out/r2_tests_full2.log
python tools/fill_once.py 300 20 > gpurun_out/r2_fill_time2.txt 2>&1; python tools/fill_once.py 100 20 >> gpurun_out/r2_fill_time2.txt 2>&1
cat gpurun_out/r2_fill_time2.txt
python tools/time_predict.py > gpurun_out/r2_time_predict.txt 2>&1
cat gpurun_out/r2_time_predict.txt
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
head -c 300 gpurun_out/r2_bench_b.json; tail -c 300 gpurun_out/r2_bench_b.err
python tools/time_lnlike.py > gpurun_out/r2_time_lnlike_b.txt 2>&1
cat gpurun_out/r2_time_lnlike_b.txt
