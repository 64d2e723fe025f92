"""One (or a few) full C4 farm evaluation(s) — for ncu launch lists.  python tools/farm_once.py [nevals] [nchunks]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psoap_b200 import synthetic  # noqa: E402
from psoap_b200.farm import ChunkFarm  # noqa: E402

nev = int(sys.argv[1]) if len(sys.argv) > 1 else 1
model, chunks = synthetic.config_chunks("C4")
if len(sys.argv) > 2:
    step = max(1, len(chunks) // int(sys.argv[2]))
    chunks = chunks[::step]
p = synthetic.default_params(model)
farm = ChunkFarm(model, chunks)
for _ in range(nev):
    v = farm.lnprob(p)
torch.cuda.synchronize()
print("chunks", len(chunks), "lnprob", v, "launches/eval", farm.launches_per_eval)
