"""Run the dominant kernel alone (for ncu captures): python tools/bench_syrk.py [m] [reps] [K] [tail_split]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psoap_b200 import _lib  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
K = int(sys.argv[3]) if len(sys.argv) > 3 else 256
lib = _lib.load()
ms, fl = ctypes.c_double(), ctypes.c_double()
tail = int(sys.argv[4]) if len(sys.argv) > 4 else 0
_lib.check(lib.psoap_bench_syrk_split(m, K, reps, tail, ctypes.byref(ms), ctypes.byref(fl)))
print("m=%d K=%d tail_split=%d avg_ms=%.4f tflops=%.2f" % (m, K, tail, ms.value, fl.value / ms.value * 1e-9))
