"""Per-launch timeline of ONE likelihood evaluation from GPU global-timer stamps (lab build of the library with
-DPSOAP_TIMELINE: every CTA records entry / start-after-PDL-wait / end).

  python tools/timeline.py SB2 20 200 [env=value ...]     # N = 20 x 200 = 4000

Prints one line per kernel launch: first CTA in, first CTA running, last CTA out (microseconds from the first stamp),
so the critical path and the gaps between the chain links can be read off.  The stamps perturb the run by < 1 %.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
for kv in sys.argv[4:]:
    k, v = kv.split("=")
    os.environ[k] = v
from psoap_b200 import _build, _lib  # noqa: E402

TL_LIB = os.path.join(ROOT, "tools", "libpsoap_tl.so")
NAMES = {1: "potrf7", 2: "trsm7", 3: "syrk<1>", 4: "syrk<2>", 5: "potrf3", 6: "trsm3", 8: "syrk<1>col"}


def build():
    src = os.path.join(_build.CSRC, "api.cu")
    if os.path.exists(TL_LIB) and os.path.getmtime(TL_LIB) > max(
            os.path.getmtime(os.path.join(_build.CSRC, f)) for f in _build.SOURCES + _build.HEADERS):
        return
    cmd = ["/usr/local/cuda/bin/nvcc"] + [f for f in _build.NVCC_FLAGS if f not in ("-Xptxas", "-v")] + ["-DPSOAP_TIMELINE", src, "-o", TL_LIB]
    subprocess.run(cmd, check=True)


def main():
    build()
    if "--build-only" in sys.argv:
        return
    _lib.LIB_PATH = TL_LIB
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import time_lnlike
    lib = _lib.load()
    lib.psoap_debug_timeline.restype = ctypes.c_int
    lib.psoap_debug_timeline.argtypes = [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
    model, ne, npx = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    cap = 1 << 20
    buf = (ctypes.c_longlong * (7 * cap))()
    r = time_lnlike.time_chunk(model, ne, npx, reps=3)        # warm
    lib.psoap_debug_timeline(buf, cap)                         # reset
    r = time_lnlike.time_chunk(model, ne, npx, reps=1)        # 2 warm-up + 1 timed evaluation
    n = lib.psoap_debug_timeline(buf, cap)
    a = np.ctypeslib.as_array(buf)[:7 * min(n, cap)].reshape(-1, 7)
    # keep the LAST evaluation: split at the largest gaps between potrf launches is fragile; use the launch count instead
    order = np.argsort(a[:, 1], kind="stable")
    a = a[order]
    launches, open_ = [], {}
    for rec in a:
        key = (int(rec[3]), int(rec[5]))
        L = open_.get(key)
        if L is None or L["n"] >= key[1]:
            L = dict(kernel=key[0], nblocks=key[1], n=0, t_in=rec[0], t_go=rec[1], t_out=rec[2])
            open_[key] = L
            launches.append(L)
        L["n"] += 1
        L["t_in"] = min(L["t_in"], rec[0]); L["t_go"] = min(L["t_go"], rec[1]); L["t_out"] = max(L["t_out"], rec[2])
    per_eval = len(launches) // 3
    launches = launches[-per_eval:]
    t0 = min(L["t_in"] for L in launches)
    print("# %s N=%d: %.3f ms per evaluation (CUDA events), %d instrumented launches per evaluation" % (model, r["N"], r["ms"], per_eval))
    print("# %-11s %6s %10s %10s %10s %9s" % ("kernel", "CTAs", "in_us", "go_us", "out_us", "run_us"))
    for L in launches:
        print("%-13s %6d %10.2f %10.2f %10.2f %9.2f" % (NAMES.get(L["kernel"], str(L["kernel"])), L["nblocks"],
              (L["t_in"] - t0) * 1e-3, (L["t_go"] - t0) * 1e-3, (L["t_out"] - t0) * 1e-3, (L["t_out"] - L["t_go"]) * 1e-3))


if __name__ == "__main__":
    main()
