"""Time a one-chunk farm (graph replay vs direct issue, PSOAP_FARM_DIRECT): python tools/farm_one_chunk.py n_epochs n_pix"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from psoap_b200 import synthetic
from psoap_b200.farm import ChunkFarm
ne, npx = int(sys.argv[1]), int(sys.argv[2])
ch = synthetic.make_chunk("SB2", ne, npx, seed=1)
p = synthetic.default_params("SB2")
farm = ChunkFarm("SB2", [ch])
for _ in range(3):
    farm.lnprob(p)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    farm.lnprob(p)
dt = (time.perf_counter() - t0) / 10
print("N=%d PSOAP_FARM_DIRECT=%s: %.3f ms per evaluation" % (ch["N"], os.environ.get("PSOAP_FARM_DIRECT", "auto"), dt * 1e3))
