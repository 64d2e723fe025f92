#!/bin/bash
# round 2, GPU call 1: new parity-by-size tests, bench with parity/fill/ranks, full CPU pass, fill ncu captures
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "vs_reference_cpu or c4_chunks or package_default or fill_lower_entrywise" > gpurun_out/r2_tests1.log 2>&1
tail -5 gpurun_out/r2_tests1.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
tail -c 600 gpurun_out/r2_bench_a.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_a.json 2> gpurun_out/r2_bench_ref_a.err
python -m oracle.cpu_farm --config C4 --full --steps 1 --warmup 0 > gpurun_out/r2_cpu_full_c4.json 2> gpurun_out/r2_cpu_full_c4.err
python tools/fill_once.py 300 20 > gpurun_out/r2_fill_time.txt 2>&1
cat gpurun_out/r2_fill_time.txt
ncu --set full --clock-control none --import-source on -k regex:fill_ -s 4 -c 4 -o gpurun_out/r2_fill python tools/fill_once.py 300 1 > gpurun_out/r2_ncu_fill.log 2>&1
python tools/time_lnlike.py > gpurun_out/r2_time_lnlike_a.txt 2>&1
cat gpurun_out/r2_time_lnlike_a.txt
nproc
