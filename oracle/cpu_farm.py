"""CPU baseline of the chunk farm: the reference's process-per-chunk layout (psoap/sample_parallel.py:258-278,
:371-390) on the host cores — Cython fill (oracle/_ref, the reference's own compiled code, when present; the C
restatement otherwise) + scipy/LAPACK Cholesky, one worker process per chunk slot.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (bench.py's cpu_baseline and --impl reference legs).  Runs as its own
process (python -m oracle.cpu_farm ...) so that the fork-based pool never shares a CUDA context.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_G = {}


def _init(model, p, chunks, use_ref, blas_threads):
    from oracle import oracle as orc
    try:
        from threadpoolctl import threadpool_limits
        _G["limit"] = threadpool_limits(limits=blas_threads, user_api="blas")
    except Exception:  # pragma: no cover
        pass
    _G.update(model=model, p=p, chunks=chunks, use_ref=use_ref, orc=orc, V11={})


def _eval(i):
    orc, ch = _G["orc"], _G["chunks"][i]
    N = len(ch["fl"])
    V11 = _G["V11"].get(N)
    if V11 is None:
        V11 = _G["V11"][N] = np.empty((N, N), dtype=np.float64)  # sample_parallel.py:163: allocated once per worker
    t0 = time.perf_counter()
    v = orc.chunk_lnprob(_G["model"], _G["p"], ch, V11=V11, use_ref_fill=_G["use_ref"])
    return i, v, time.perf_counter() - t0


def run(config="C4", sample=8, steps=1, warmup=0, workers=None):
    from oracle import oracle as orc
    from psoap_b200 import synthetic
    model, chunks = synthetic.config_chunks(config)
    n_total = len(chunks)
    sample = max(1, min(sample, n_total))
    idx = [int(round(k * (n_total - 1) / max(1, sample - 1))) for k in range(sample)] if sample > 1 else [n_total // 2]
    idx = sorted(set(idx))
    sub = [chunks[i] for i in idx]
    p = synthetic.default_params(model)
    cores = os.cpu_count() or 1
    workers = max(1, min(workers or cores, len(sub)))
    blas_threads = max(1, cores // workers)
    use_ref = orc.ref_matrix_functions() is not None
    # cost model used to scale the sample to the full configuration: N^3/3 + 2 N^2 flops + fill N^2 ncomp
    cost = lambda ch: len(ch["fl"]) ** 3 / 3.0
    scale = sum(cost(c) for c in chunks) / sum(cost(c) for c in sub)
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(workers, initializer=_init, initargs=(model, p, sub, use_ref, blas_threads)) as pool:
        for s in range(warmup + steps):
            t0 = time.perf_counter()
            res = pool.map(_eval, range(len(sub)), chunksize=1)
            dt = time.perf_counter() - t0
            if s >= warmup:
                times.append(dt)
    lnl = [r[1] for r in sorted(res)]
    t = float(np.mean(times))
    return dict(config=config, model=model, n_chunks=n_total, sample_chunks=idx, sample_N=[len(c["fl"]) for c in sub],
                seconds_per_sample_eval=t, scale_to_full=scale, evals_per_s=1.0 / (t * scale), cores=cores,
                workers=workers, blas_threads_per_worker=blas_threads, kind="reference" if use_ref else "port",
                per_chunk_seconds=[r[2] for r in sorted(res)], lnlike_sample_sum=float(np.sum(lnl)), steps=steps,
                warmup=warmup)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C4")
    ap.add_argument("--sample", type=int, default=8)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--workers", type=int, default=None)
    a = ap.parse_args()
    print(json.dumps(run(a.config, a.sample, a.steps, a.warmup, a.workers)))
