"""Where the SM slots of a farm evaluation go, from GPU global-timer stamps of every CTA (lab build of the library with
-DPSOAP_TIMELINE, see tools/timeline.py).

  python tools/timeline_farm.py [nchunks=32] [nbranch=32] [config=C4]

Prints, per kernel, the CTA count, the summed CTA residence time and its share of the evaluation's slot-time
(wall x 148 SMs x 2 CTA slots: the trailing update runs 2 CTAs per SM), and per SM the fraction of the wall time with
0 / 1 / >= 2 CTAs resident.
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import timeline  # noqa: E402
from psoap_b200 import _lib  # noqa: E402


def main():
    timeline.build()
    _lib.LIB_PATH = timeline.TL_LIB
    import torch
    from psoap_b200 import synthetic
    from psoap_b200.farm import ChunkFarm
    nchunks = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    nbranch = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    lib = _lib.load()
    lib.psoap_debug_timeline.restype = ctypes.c_int
    lib.psoap_debug_timeline.argtypes = [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
    model, chunks = synthetic.config_chunks(sys.argv[3] if len(sys.argv) > 3 else "C4")
    chunks = chunks[::max(1, len(chunks) // nchunks)]
    p = synthetic.default_params(model)
    farm = ChunkFarm(model, chunks, nbranch=nbranch)
    for _ in range(2):
        farm.lnprob(p)
    torch.cuda.synchronize()
    cap = 1 << 20
    buf = (ctypes.c_longlong * (7 * cap))()
    lib.psoap_debug_timeline(buf, cap)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); farm.lnprob(p); e1.record(); torch.cuda.synchronize()
    n = lib.psoap_debug_timeline(buf, cap)
    a = np.ctypeslib.as_array(buf)[:7 * min(n, cap)].reshape(-1, 7).astype(np.float64)
    t0, t1 = a[:, 1].min(), a[:, 2].max()
    wall = (t1 - t0) * 1e-3
    flops = sum(c["N"] ** 3 / 3.0 for c in chunks)
    print("# %d chunks on %d branches: %.2f ms (events), %.2f ms first CTA start to last CTA end, %d CTA records%s, %.2f TFLOP/s"
          % (len(chunks), nbranch, e0.elapsed_time(e1), wall * 1e-3, n, " (TRUNCATED)" if n >= cap else "", flops / (wall * 1e-6) * 1e-12))
    print("# %-12s %8s %12s %10s %10s" % ("kernel", "CTAs", "sum_ms", "slot_share", "avg_us"))
    slots = wall * 148 * 2
    for k in sorted(set(a[:, 3].astype(int))):
        r = a[a[:, 3] == k]
        d = (r[:, 2] - r[:, 1]) * 1e-3
        w = (r[:, 1] - r[:, 0]) * 1e-3    # resident but waiting in griddepcontrol.wait
        print("%-14s %8d %12.2f %9.1f%% %10.2f   (pre-staged wait: %.2f ms)" % (timeline.NAMES.get(k, str(k)), len(r), d.sum() * 1e-3, 100 * d.sum() / slots,
                                                  d.mean(), w.sum() * 1e-3))
    # per-SM residency histogram from a sweep over the CTA intervals (entry .. end)
    hist = np.zeros(4)
    for sm in range(148):
        r = a[a[:, 6] == sm]
        ev = np.concatenate([np.stack([r[:, 0], np.ones(len(r))], 1), np.stack([r[:, 2], -np.ones(len(r))], 1)])
        ev = ev[np.argsort(ev[:, 0], kind="stable")]
        cur, last = 0, t0
        for t, dlt in ev:
            hist[min(cur, 3)] += max(0.0, t - last)
            last = max(last, t)
            cur += int(dlt)
        hist[min(cur, 3)] += max(0.0, t1 - last)
    hist /= hist.sum()
    print("# SM-time with 0 / 1 / 2 / >=3 CTAs resident: %.1f%% / %.1f%% / %.1f%% / %.1f%%" % tuple(100 * hist))


if __name__ == "__main__":
    main()
