"""Time the farm on a subset of C4 (what one rank of a W-GPU run holds): python tools/farm_subset_time.py [W] [nbranch]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psoap_b200 import synthetic  # noqa: E402
from psoap_b200.farm import ChunkFarm, lpt_partition, chunk_cost  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
model, chunks = synthetic.config_chunks("C4")
parts = lpt_partition([chunk_cost(c["N"]) for c in chunks], world)
mine = [chunks[i] for i in sorted(parts[0])]
p = synthetic.default_params(model)
nbranch = int(sys.argv[2]) if len(sys.argv) > 2 else 32
farm = ChunkFarm(model, mine, nbranch=nbranch)
for _ in range(3):
    farm.lnprob(p)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    farm.lnprob(p)
dt = (time.perf_counter() - t0) / 5
fl = sum(c["N"] ** 3 / 3 for c in mine)
print("rank 0 of %d: %d chunks, nbranch %d, %.2f ms per evaluation, %.2f TFLOP/s" % (world, len(mine), nbranch, dt * 1e3, fl / dt * 1e-12))
