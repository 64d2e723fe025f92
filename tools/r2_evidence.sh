#!/bin/bash
# round 2 evidence: launch lists (ncu time-only pass) and ncu --set full captures of the kernels as shipped
set -x
mkdir -p gpurun_out
LL="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 900 $LL --log-file gpurun_out/launches_farm_c4_32chunks_r02.csv python tools/farm_once.py 1 32 > gpurun_out/ev_farm32.log 2>&1
timeout 300 $LL --log-file gpurun_out/launches_lnlike_n2000_r02.csv python tools/time_lnlike.py --one SB2 20 100 > gpurun_out/ev_n2000.log 2>&1
timeout 300 $LL --log-file gpurun_out/launches_lnlike_n4000_r02.csv python tools/time_lnlike.py --one SB2 20 200 > gpurun_out/ev_n4000.log 2>&1
timeout 300 $LL --log-file gpurun_out/launches_lnlike_n9000_r02.csv python tools/time_lnlike.py --one SB2 30 300 > gpurun_out/ev_n9000.log 2>&1
FULL="ncu --set full --clock-control none --import-source on"
timeout 300 $FULL -k regex:syrk3 -s 2 -c 1 -f -o gpurun_out/syrk3_m4096_k512_r02 python tools/bench_syrk.py 4096 2 512 0 > gpurun_out/ev_s1.log 2>&1
timeout 300 $FULL -k regex:syrk3 -s 2 -c 1 -f -o gpurun_out/syrk3_m4096_k512_tail_r02 python tools/bench_syrk.py 4096 2 512 1 > gpurun_out/ev_s2.log 2>&1
timeout 300 $FULL -k regex:potrf_diag7 -s 40 -c 1 -f -o gpurun_out/potrf7_n4000_r02 python tools/time_lnlike.py --one SB2 20 200 > gpurun_out/ev_p7.log 2>&1
timeout 300 $FULL -k regex:trsm7 -s 40 -c 1 -f -o gpurun_out/trsm7_n4000_r02 python tools/time_lnlike.py --one SB2 20 200 > gpurun_out/ev_t7.log 2>&1
timeout 300 $FULL -k regex:potrf_diag3 -s 40 -c 1 -f -o gpurun_out/potrf3_farm_r02 python tools/farm_once.py 1 32 > gpurun_out/ev_p3.log 2>&1
timeout 300 $FULL -k regex:trsm3 -s 40 -c 1 -f -o gpurun_out/trsm3_farm_r02 python tools/farm_once.py 1 32 > gpurun_out/ev_t3.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -8
tail -2 gpurun_out/ev_*.log
