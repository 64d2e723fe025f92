"""Time the operator-surface covariance fill (fill_V11_f_g, both triangles) on the device: python tools/time_fill.py [N]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from psoap_b200 import matrix_functions as mf

N = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
rng = np.random.default_rng(0)
n_pix = 300
z = np.log(5000.0) + (np.arange(N) % n_pix) * 2.8 / 2.99792458e5 + rng.normal(size=N) * 1e-7
zf = torch.tensor(z, device="cuda"); zg = torch.tensor(z + 3e-5, device="cuda")
mat = torch.empty((N, N), dtype=torch.float64, device="cuda")
for _ in range(3):
    mf.fill_V11_f_g(mat, zf, zg, 0.1, 5.0, 0.05, 7.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 20
for _ in range(reps):
    mf.fill_V11_f_g(mat, zf, zg, 0.1, 5.0, 0.05, 7.0)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
pairs = N * (N - 1) / 2
print("fill_V11_f_g N=%d: %.1f us, %.0f G exp/s, %.2f TB/s written (8N^2 bytes)" % (N, ms * 1e3, 2 * pairs / ms * 1e-6, 8.0 * N * N / ms * 1e-9))
