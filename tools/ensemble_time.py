"""Throughput of batched multi-proposal evaluation (ChunkFarm(n_proposals=K)) on one workload:
python tools/ensemble_time.py [C2] [K ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from psoap_b200 import synthetic
from psoap_b200.farm import ChunkFarm

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
Ks = [int(a) for a in sys.argv[2:]] or [1, 4, 8]
model, chunks = synthetic.config_chunks(cfg)
p = synthetic.default_params(model)
flops = sum(c["N"] ** 3 / 3.0 + 2.0 * c["N"] ** 2 for c in chunks)
for K in Ks:
    farm = ChunkFarm(model, chunks, n_proposals=K)
    P = np.tile(p, (K, 1)) * (1.0 + 1e-3 * np.arange(K)[:, None])
    for _ in range(2):
        farm.lnprob_many(P)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        farm.lnprob_many(P)
    dt = (time.perf_counter() - t0) / reps
    print("%s K=%d: %.2f ms per launch, %.1f evals/s, %.1f TFLOP/s" % (cfg, K, dt * 1e3, K / dt, K * flops / dt * 1e-12), flush=True)
    farm.close()
    del farm
    torch.cuda.empty_cache()
