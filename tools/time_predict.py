"""Time psoap_predict (device-resident inputs) at the sizes the reference's scripts use: python tools/time_predict.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psoap_b200 import _lib, covariance, synthetic  # noqa: E402

out = []
for name, ne, npx, m in (("retrieve SB2 (psoap_retrieve_SB2.py:76-105): n=9000, 600 points per component", 30, 300, 600),
                         ("predict SB2 (psoap_predict_SB2.py:75): n=m=2000", 8, 250, 2000),
                         ("predict SB2: n=m=4000", 20, 200, 4000)):
    ch = synthetic.make_chunk("SB2", ne, npx, seed=3)
    p = synthetic.default_params("SB2")
    vel = synthetic.host_velocities("SB2", p[:7], ch["date1D"])
    lw = [torch.from_numpy(np.ascontiguousarray(ch["lwl"] - vel[c][ch["epoch"]] / synthetic.c_kms)).cuda() for c in range(2)]
    fl, sg = torch.from_numpy(ch["fl"]).cuda(), torch.from_numpy(ch["sigma"]).cuda()
    lo, hi = float(ch["lwl"].min()), float(ch["lwl"].max())
    grids = [torch.linspace(lo, hi, m, dtype=torch.float64, device="cuda") for _ in range(2)]
    n, M = ch["N"], 2 * m

    def run():
        return covariance._predict_device(0, lw, fl, sg, grids, p[7::2], p[8::2], 1.0)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    flops = n ** 3 / 3.0 + float(n) * n * M + float(n) * M * M   # eliminate n columns of an (n + M) lower triangle
    r = dict(case=name, n=n, M=M, ms=ms, tflops=flops / ms * 1e-9, workspace_gb=_lib.load().psoap_predict_workspace_bytes(2, 0, n, m) / 1e9)
    print(json.dumps(r), flush=True)
    out.append(r)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/time_predict.json", "w"), indent=1)
