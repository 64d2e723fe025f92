#!/bin/bash
mkdir -p gpurun_out
{
for sp in 0 1 2 3; do echo "SPLIT=$sp"; PSOAP_SPLIT=$sp timeout 300 python tools/time_lnlike.py | head -4; done
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "lnlike_golden or tile_boundaries or predict or farm_vs_oracle or vs_reference_cpu or package_default or lnlike_vs_oracle or calibration or repeatable" 2>&1 | tail -5
} > gpurun_out/r2_split.txt 2>&1
cat gpurun_out/r2_split.txt
