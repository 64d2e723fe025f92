#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:syrk3 -s 2 -c 1 -f -o gpurun_out/syrk3_m4096_k1024_r02 python tools/bench_syrk.py 4096 2 1024 0 > gpurun_out/ev_s3.log 2>&1
for m in 2048 4096 6144 8192; do python tools/bench_syrk.py $m 20 1024 0; done
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; head -c 200 gpurun_out/r2_bench_c4.json; echo
python bench.py --workload C6 --steps 10 --warmup 3 > gpurun_out/r2_bench_C6.json 2> gpurun_out/r2_bench_C6.err; head -c 200 gpurun_out/r2_bench_C6.json; echo
