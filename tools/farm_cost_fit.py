"""Cost model of the static partition (psoap_b200/farm.py chunk_cost): time of a farm of 64 equal chunks for several
sizes, fitted as t(N) = a (N^3 + ALPHA2 N^2 + ALPHA1 N).  python tools/farm_cost_fit.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psoap_b200 import synthetic  # noqa: E402
from psoap_b200.farm import ChunkFarm  # noqa: E402

p = synthetic.default_params("SB2")
rows = []
for n_pix in (80, 100, 150, 200, 250, 300):
    chunks = [synthetic.make_chunk("SB2", 20, n_pix, seed=9000 + i, wl0=5000.0 + 3 * i) for i in range(64)]
    farm = ChunkFarm("SB2", chunks, nbranch=32)
    for _ in range(2):
        farm.lnprob(p)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        farm.lnprob(p)
    dt = (time.perf_counter() - t0) / 4 / 64
    N = 20 * n_pix
    rows.append((N, dt))
    print("N=%5d: %.1f us per chunk inside a 64-chunk farm, %.2f TFLOP/s" % (N, dt * 1e6, (N ** 3 / 3 + 2 * N * N) / dt * 1e-12), flush=True)
    farm.close()
N = np.array([r[0] for r in rows], dtype=float)
t = np.array([r[1] for r in rows])
A = np.stack([N ** 3, N ** 2, N], axis=1)
coef, *_ = np.linalg.lstsq(A / t[:, None], np.ones(len(t)), rcond=None)   # relative least squares
print("t(N) = %.3e N^3 + %.3e N^2 + %.3e N   ->  ALPHA2 = %.1f, ALPHA1 = %.3e" % (coef[0], coef[1], coef[2], coef[1] / coef[0], coef[2] / coef[0]))
