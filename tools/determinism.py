"""Run the same likelihood repeatedly (sync between calls) and print the distinct values seen."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psoap_b200 import _lib, synthetic  # noqa: E402

lib = _lib.load()
for n_epochs, n_pix in [(20, 200), (20, 300), (30, 300)]:
    ch = synthetic.make_chunk("SB2", n_epochs, n_pix, seed=1)
    p = synthetic.default_params("SB2")
    vel = synthetic.host_velocities("SB2", p[:7], ch["date1D"])
    lw = [torch.from_numpy(ch["lwl"] - vel[c][ch["epoch"]] / synthetic.c_kms).cuda() for c in range(2)]
    fl, sg = torch.from_numpy(ch["fl"]).cuda(), torch.from_numpy(ch["sigma"]).cuda()
    N = ch["N"]
    nbytes = lib.psoap_lnlike_workspace_bytes(N)
    ws = torch.empty(nbytes + 256, dtype=torch.uint8, device="cuda")
    res = torch.empty(4, dtype=torch.float64, device="cuda")
    vals = []
    for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
        _lib.check(lib.psoap_lnlike(2, N, _lib.ptr(lw[0]), _lib.ptr(lw[1]), _lib.vp(None), _lib.ptr(fl), _lib.ptr(sg),
                                    _lib.dbl_array(p[7::2]), _lib.dbl_array(p[8::2]), 1.0, _lib.ptr(ws), nbytes,
                                    _lib.ptr(res), _lib.stream_ptr()))
        torch.cuda.synchronize()
        vals.append(repr(float(res[0].item())))
    print("N=%d distinct=%d %s" % (N, len(set(vals)), sorted(set(vals))), flush=True)
