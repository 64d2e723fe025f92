#!/bin/bash
mkdir -p gpurun_out
for w in C5; do python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err; head -c 200 gpurun_out/r2_bench_$w.json; echo; done
python tools/time_lnlike.py --big > gpurun_out/r2_time_lnlike_final.txt 2>&1; cat gpurun_out/r2_time_lnlike_final.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; head -c 200 gpurun_out/r2_bench_c4.json; echo
