#!/bin/bash
mkdir -p gpurun_out
{
for pdl in 1 130; do for h in 0 1; do echo "PDL=$pdl HANDOVER=$h"; PSOAP_PDL=$pdl PSOAP_HANDOVER=$h timeout 120 python tools/time_lnlike.py; done; done
PSOAP_PDL=1 timeout 120 python tools/timeline.py SB2 20 200 > gpurun_out/timeline_n4000_d.txt 2>&1
} > gpurun_out/r2_lanes2.txt 2>&1
cat gpurun_out/r2_lanes2.txt
