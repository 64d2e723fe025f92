#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2_tests_y2.log 2>&1
tail -4 gpurun_out/r2_tests_y2.log
for w in C1 C2 C3 C5; do python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err; head -c 200 gpurun_out/r2_bench_$w.json; echo; done
python tools/time_lnlike.py --big > gpurun_out/r2_time_lnlike_final.txt 2>&1; cat gpurun_out/r2_time_lnlike_final.txt
python tools/time_predict.py > gpurun_out/r2_time_predict_final.txt 2>&1; cat gpurun_out/r2_time_predict_final.txt
python tools/timeline.py SB2 20 200 > gpurun_out/timeline_n4000.txt 2>&1
python tools/timeline.py SB2 20 100 > gpurun_out/timeline_n2000.txt 2>&1
