#!/bin/bash
mkdir -p gpurun_out
{
for k in 0 8 16 24; do echo "LAST7=$k"; PSOAP_FARM_TAIL=0 PSOAP_FARM_LAST7=$k timeout 300 python tools/farm_subset_time.py 8; done
for k in 0 16; do echo "LAST7=$k TAIL=1"; PSOAP_FARM_TAIL=1 PSOAP_FARM_LAST7=$k timeout 300 python tools/farm_subset_time.py 8; done
for k in 0 16; do echo "LAST7=$k full farm"; PSOAP_FARM_TAIL=0 PSOAP_FARM_LAST7=$k timeout 300 python tools/farm_subset_time.py 1; done
for k in 0 8 16; do echo "direct LAST7=$k (T>=192 only uses chain 3)"; PSOAP_POTRF=3 PSOAP_TAIL=0 PSOAP_LAST7=$k timeout 300 python tools/time_lnlike.py; done
} > gpurun_out/r2_last7.txt 2>&1
cat gpurun_out/r2_last7.txt
