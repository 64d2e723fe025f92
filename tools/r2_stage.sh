#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python tools/time_lnlike.py
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "lnlike_golden or tile_boundaries or predict or farm_vs_oracle or vs_reference_cpu or package_default or lnlike_vs_oracle or calibration or repeatable" 2>&1 | tail -5
timeout 300 python bench.py --workload C1 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | head -c 200; echo
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_stage.csv python tools/time_lnlike.py --one SB2 20 200 > /dev/null 2>&1
python tools/summarize_ncu.py launches gpurun_out/launches_stage.csv
} > gpurun_out/r2_stage.txt 2>&1
cat gpurun_out/r2_stage.txt
