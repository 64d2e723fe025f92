#!/bin/bash
# round 2, last call: full GPU suite, smoke, the default bench line and the reference arm with the committed library
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2_tests_final.log 2>&1
tail -4 gpurun_out/r2_tests_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; head -c 200 gpurun_out/r2_bench_ref.json; echo
python bench.py > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; head -c 300 gpurun_out/r2_bench_c4.json; echo
