#!/bin/bash
# round 2, last call: full GPU suite, smoke, bench lines and single-call timings with the committed library
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2_tests_final.log 2>&1
tail -4 gpurun_out/r2_tests_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
for w in C1 C2 C3 C6; do python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err; head -c 200 gpurun_out/r2_bench_$w.json; echo; done
python tools/time_lnlike.py --big > gpurun_out/r2_time_lnlike_final.txt 2>&1; cat gpurun_out/r2_time_lnlike_final.txt
python tools/time_predict.py > gpurun_out/r2_time_predict_final.txt 2>&1; cat gpurun_out/r2_time_predict_final.txt
python bench.py > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; head -c 300 gpurun_out/r2_bench_c4.json; echo
