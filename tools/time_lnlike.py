"""Time the single-chunk likelihood (device-resident inputs, CUDA events) for a few sizes."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psoap_b200 import _lib, synthetic  # noqa: E402


def time_chunk(model, n_epochs, n_pix, reps=5):
    lib = _lib.load()
    ch = synthetic.make_chunk(model, n_epochs, n_pix, seed=1)
    p = synthetic.default_params(model)
    n_orb = _lib.N_ORB[model]
    vel = synthetic.host_velocities(model, p[:n_orb], ch["date1D"])
    ncomp = vel.shape[0]
    lw = [torch.from_numpy(ch["lwl"] - vel[c][ch["epoch"]] / synthetic.c_kms).cuda() for c in range(ncomp)]
    fl, sg = torch.from_numpy(ch["fl"]).cuda(), torch.from_numpy(ch["sigma"]).cuda()
    N = ch["N"]
    nbytes = lib.psoap_lnlike_workspace_bytes(N)
    ws = torch.empty(nbytes + 256, dtype=torch.uint8, device="cuda")
    res = torch.empty(4, dtype=torch.float64, device="cuda")
    ptrs = [_lib.ptr(v) for v in lw] + [_lib.vp(None)] * (3 - ncomp)
    amps, ls = _lib.dbl_array(p[n_orb::2]), _lib.dbl_array(p[n_orb + 1::2])

    def run():
        _lib.check(lib.psoap_lnlike(ncomp, N, ptrs[0], ptrs[1], ptrs[2], _lib.ptr(fl), _lib.ptr(sg), amps, ls, 1.0,
                                    _lib.ptr(ws), nbytes, _lib.ptr(res), _lib.stream_ptr()))
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    times = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = float(np.median(times))
    flops = N ** 3 / 3.0 + 2.0 * N ** 2
    return dict(model=model, N=N, ms=ms, ms_min=min(times), tflops=flops / ms * 1e-9, lnlike=float(res[0].item()))


if __name__ == "__main__":
    if "--one" in sys.argv:  # python tools/time_lnlike.py --one SB2 20 200  (for ncu launch lists)
        i = sys.argv.index("--one")
        print(json.dumps(time_chunk(sys.argv[i + 1], int(sys.argv[i + 2]), int(sys.argv[i + 3]), reps=1)))
        sys.exit(0)
    out = []
    for model, ne, npx in [("SB2", 20, 100), ("SB2", 20, 200), ("SB1", 20, 200), ("SB2", 20, 300), ("SB2", 30, 300),
                           ("ST3", 40, 250)] + ([("SB2", 64, 256)] if "--big" in sys.argv else []):
        r = time_chunk(model, ne, npx)
        print(json.dumps(r), flush=True)
        out.append(r)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/time_lnlike.json", "w"), indent=1)
