"""One launch sequence of both fill kernels at the size of C4's largest chunk (for ncu captures):
python tools/fill_once.py [N_pix] [reps]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psoap_b200 import _lib, synthetic  # noqa: E402

n_pix = int(sys.argv[1]) if len(sys.argv) > 1 else 300
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
lib = _lib.load()
ch = synthetic.make_chunk("SB2", 20, n_pix, seed=4255, wl0=5000.0 + 4.0 * 255)
p = synthetic.default_params("SB2")
vel = synthetic.host_velocities("SB2", p[:7], ch["date1D"])
lw = [torch.from_numpy(np.ascontiguousarray(ch["lwl"] - vel[c][ch["epoch"]] / synthetic.c_kms)).cuda() for c in range(2)]
amp, l = _lib.dbl_array(p[7::2]), _lib.dbl_array(p[8::2])
N = ch["N"]
for kind, name, nbytes in ((0, "fill_lower", 4.0 * N * N), (1, "fill_full", 8.0 * N * N)):
    t = ctypes.c_double()
    _lib.check(lib.psoap_bench_fill(kind, 2, N, _lib.ptr(lw[0]), _lib.ptr(lw[1]), None, amp, l, reps, ctypes.byref(t)))
    print("%s N=%d: %.1f us  %.0f GB/s algorithmic  %.0f Gexp/s" % (name, N, t.value * 1e3, nbytes / t.value * 1e-6,
                                                                      N * (N - 1.0) / t.value * 1e-6))
