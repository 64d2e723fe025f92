#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python tools/time_lnlike.py
for g in 1 4; do echo "GROUP=$g"; PSOAP_GROUP=$g timeout 300 python tools/time_lnlike.py | head -5; done
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "lnlike_golden or tile_boundaries or predict or farm_vs_oracle or vs_reference_cpu or package_default or lnlike_vs_oracle or calibration or repeatable" 2>&1 | tail -5
for w in C1 C2; do timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | head -c 200; echo; done
timeout 200 python tools/timeline.py SB2 20 200 > gpurun_out/timeline_n4000_b.txt 2>&1
} > gpurun_out/r2_lanes.txt 2>&1
cat gpurun_out/r2_lanes.txt
