"""CPU baseline of the chunk farm: the reference's chunk-parallel layout (psoap/sample_parallel.py:258-278,
:371-390) on the host cores — Cython fill (oracle/_ref, the reference's own compiled code, when present; the C
restatement otherwise) + scipy/LAPACK Cholesky.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (bench.py's cpu_baseline and --impl reference legs).  Runs as its own
process (python -m oracle.cpu_farm ...) so that the fork-based pool never shares a CUDA context.

How the cores are kept busy.  The reference forks one process per chunk and lets the OS time-slice them; with more
chunks than cores that is a pool of `cores` busy workers.  Here: `workers` = all host cores, ONE BLAS thread each
(the fill is single-threaded, so any other split idles cores during the fill), and the sampled chunks are handed out
DYNAMICALLY, largest first (imap_unordered, chunksize 1), at least 4 chunks per worker, so no worker waits for the
slowest chunk.  A single-chunk configuration (C1-C3, C5) runs one process with all cores as BLAS threads, which is
what `lnlike_*` does when called directly.

What is reported.  `seconds_per_sample_eval` is the measured wall time of one pass over the sample;
`per_chunk_seconds` (+ fill / LAPACK split) are measured inside the workers.  A sample is scaled to the whole
configuration with the per-chunk seconds themselves (piecewise-linear in N over the sampled sizes), never with an
assumed N^3 law; `--full` evaluates every chunk (no scaling at all) and is the run that validates the sampled figure
(profiles/).  `evals_per_s_from_cpu_seconds` = workers / (sum of per-chunk seconds, scaled) is the same number
recomputed without the pool's wall clock.
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


def _init(model, p, chunks, use_ref, blas_threads, nmax):
    from oracle import oracle as orc
    try:
        from threadpoolctl import threadpool_limits
        _G["limit"] = threadpool_limits(limits=blas_threads, user_api="blas")
    except Exception:  # pragma: no cover
        pass
    # sample_parallel.py:161-163: ONE scratch matrix per worker, allocated once; smaller chunks use a leading view
    _G.update(model=model, p=p, chunks=chunks, use_ref=use_ref, orc=orc, buf=np.empty(nmax * nmax, dtype=np.float64))


def _eval(i):
    orc, ch = _G["orc"], _G["chunks"][i]
    N = len(ch["fl"])
    V11 = _G["buf"][:N * N].reshape(N, N)
    timers = {}
    t0 = time.perf_counter()
    v = orc.chunk_lnprob(_G["model"], _G["p"], ch, V11=V11, use_ref_fill=_G["use_ref"], timers=timers)
    return i, v, time.perf_counter() - t0, timers.get("fill", 0.0), timers.get("lapack", 0.0)


def sample_indices(n_total, sample):
    """`sample` indices spread evenly over the configuration's chunks (mid-points of equal strides)."""
    sample = max(1, min(sample, n_total))
    return sorted(set(int((k + 0.5) * n_total / sample) for k in range(sample)))


def run(config="C4", sample=64, steps=1, warmup=0, workers=None, full=False, blas_threads=None):
    from oracle import oracle as orc
    from psoap_b200 import synthetic
    model, chunks = synthetic.config_chunks(config)
    n_total = len(chunks)
    idx = list(range(n_total)) if full else sample_indices(n_total, sample)
    sub = [chunks[i] for i in idx]
    Ns_all = np.array([len(c["fl"]) for c in chunks], dtype=np.float64)
    Ns = np.array([len(c["fl"]) for c in sub], dtype=np.float64)
    p = synthetic.default_params(model)
    cores = os.cpu_count() or 1
    workers = max(1, min(workers or cores, len(sub)))
    blas_threads = blas_threads or max(1, cores // workers)
    use_ref = orc.ref_matrix_functions() is not None
    order = sorted(range(len(sub)), key=lambda k: -Ns[k])       # largest first
    ctx = mp.get_context("fork")
    times, res = [], {}
    with ctx.Pool(workers, initializer=_init, initargs=(model, p, sub, use_ref, blas_threads, int(Ns.max()))) as pool:
        for s in range(warmup + steps):
            t0 = time.perf_counter()
            out = list(pool.imap_unordered(_eval, order, chunksize=1))
            dt = time.perf_counter() - t0
            if s >= warmup:
                times.append(dt)
                res = {r[0]: r for r in out}
    rows = [res[k] for k in range(len(sub))]
    per_chunk = np.array([r[2] for r in rows])
    t = float(np.mean(times))
    # scale the sample to the whole configuration with the measured per-chunk seconds (interpolated in N)
    if len(sub) == n_total:
        cpu_seconds_full = float(per_chunk.sum())
    else:
        o = np.argsort(Ns)
        cpu_seconds_full = float(np.interp(Ns_all, Ns[o], per_chunk[o]).sum())
    scale = cpu_seconds_full / float(per_chunk.sum())
    return dict(config=config, model=model, n_chunks=n_total, sample_chunks=idx, sample_N=[int(n) for n in Ns],
                seconds_per_sample_eval=t, scale_to_full=scale, evals_per_s=1.0 / (t * scale), cores=cores,
                workers=workers, blas_threads_per_worker=blas_threads, kind="reference" if use_ref else "port",
                extrapolated=len(sub) != n_total, scheduling="dynamic pool, largest chunk first, chunksize 1",
                per_chunk_seconds=[float(x) for x in per_chunk],
                per_chunk_fill_seconds=[float(r[3]) for r in rows], per_chunk_lapack_seconds=[float(r[4]) for r in rows],
                fill_fraction=float(sum(r[3] for r in rows) / per_chunk.sum()),
                lapack_fraction=float(sum(r[4] for r in rows) / per_chunk.sum()),
                cpu_seconds_full=cpu_seconds_full, evals_per_s_from_cpu_seconds=workers / cpu_seconds_full,
                pool_efficiency=float(per_chunk.sum() / (workers * t)),
                lnlike_per_chunk=[float(r[1]) for r in rows], lnlike_sample_sum=float(np.sum([r[1] for r in rows])),
                params=[float(x) for x in p], steps=steps, warmup=warmup)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C4")
    ap.add_argument("--sample", type=int, default=64)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--workers", type=int, default=None)
    ap.add_argument("--full", action="store_true", help="evaluate every chunk of the configuration (no scaling)")
    ap.add_argument("--blas-threads", type=int, default=None, help="BLAS threads per worker (default: cores // workers)")
    a = ap.parse_args()
    print(json.dumps(run(a.config, a.sample, a.steps, a.warmup, a.workers, a.full, a.blas_threads)))
