#!/bin/bash
# usage: r2_scale.sh N  — the bench line at N GPUs of one box
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_c4_${N}gpu.json 2> gpurun_out/r2_bench_c4_${N}gpu.err
tail -c 600 gpurun_out/r2_bench_c4_${N}gpu.json | head -c 300; echo
