"""CPU: host-side logic — chunk partitioning, the cross-rank reduction (gloo, world_size 2), synthetic inputs."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lpt_partition_covers_and_balances():
    from psoap_b200.farm import chunk_cost, lpt_partition
    Ns = [2000 + (4000 * i) // 255 for i in range(256)]
    costs = [chunk_cost(n) for n in Ns]
    for g in (1, 2, 4, 8):
        parts = lpt_partition(costs, g)
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(256))
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) / (sum(loads) / g) < 1.02  # LPT is within 2 % of perfect here
        assert parts == lpt_partition(costs, g)      # deterministic


def test_epoch_index_matches_mask_broadcast():
    from psoap_b200.data import epoch_index
    rng = np.random.default_rng(0)
    mask = rng.uniform(size=(5, 17)) > 0.2
    v = rng.normal(size=5)
    expect = (v[:, np.newaxis] * np.ones_like(mask))[mask]  # psoap/data.py:61
    assert np.array_equal(v[epoch_index(mask)], expect)


def test_synthetic_chunks_are_seeded_and_shaped():
    from psoap_b200 import synthetic
    a = synthetic.make_chunk("SB2", 5, 40, seed=3, mask_frac=0.1)
    b = synthetic.make_chunk("SB2", 5, 40, seed=3, mask_frac=0.1)
    assert all(np.array_equal(a[k], b[k]) for k in ("lwl", "fl", "sigma", "mask", "date1D", "epoch"))
    assert a["N"] == a["mask"].sum() == len(a["fl"]) < 200
    model, chunks = synthetic.config_chunks("C4")
    Ns = [c["N"] for c in chunks]
    assert model == "SB2" and len(chunks) == 256 and min(Ns) == 2000 and max(Ns) == 6000
    assert synthetic.config_chunks("C2")[1][0]["N"] == 9000


def test_synthetic_velocities_match_oracle(oracle):
    from psoap_b200 import synthetic
    dates = np.linspace(-3.0, 70.0, 23)
    for model in ("SB1", "SB2", "ST3"):
        p = synthetic.ORBIT_PARAMS[model]
        assert np.allclose(synthetic.host_velocities(model, p, dates), oracle.get_velocities(model, p, dates),
                           rtol=0, atol=1e-10)


def _worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    from psoap_b200 import synthetic
    from psoap_b200.farm import chunk_cost, lpt_partition
    chunks = [synthetic.make_chunk("SB2", 3, 20 + 5 * i, seed=50 + i, mask_frac=0.05 * (i % 2)) for i in range(7)]
    p = synthetic.default_params("SB2")
    if os.environ.get("PSOAP_TEST_INF") == "1":
        p[1] = 4e5  # |v| >= c on every chunk
    parts = lpt_partition([chunk_cost(c["N"]) for c in chunks], world)
    vec = torch.zeros(len(chunks), dtype=torch.float64)
    for i in parts[rank]:
        vec[i] = orc.chunk_lnprob("SB2", p, chunks[i])  # the oracle stands in for the GPU evaluation
    dist.all_reduce(vec, op=dist.ReduceOp.SUM)         # the one collective of the path
    total = float(np.sum(vec.numpy()))
    ref_total, ref_vec = orc.farm_lnprob("SB2", p, chunks)
    ok = np.array_equal(vec.numpy(), ref_vec) and total == ref_total
    with open(os.path.join(tmp, f"rank{rank}.txt"), "w") as fh:
        fh.write("ok" if ok else f"mismatch {total} {ref_total}")
    dist.destroy_process_group()


@pytest.mark.parametrize("inf_case", ["0", "1"])
def test_two_rank_gloo_reduction_equals_serial_sum(tmp_path, inf_case):
    import torch.multiprocessing as mp
    os.environ["PSOAP_TEST_INF"] = inf_case
    port = 29500 + (os.getpid() % 2000) + int(inf_case)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(tmp_path / f"rank{r}.txt").read() == "ok"


def test_param_registry_round_trip():
    from psoap_b200 import utils
    pars = dict(q=0.2, K=5.0, e=0.2, omega=10.0, P=10.0, T0=0.0, gamma=5.0, amp_f=0.1, l_f=5.0, amp_g=0.05, l_g=7.0)
    fix = ["gamma", "T0"]
    p = utils.convert_dict("SB2", fix, **pars)
    assert len(p) == 9 and p[0] == 0.2 and p[-1] == 7.0
    p_orb, p_GP = utils.convert_vector(p, "SB2", fix, **pars)
    assert list(p_orb) == [0.2, 5.0, 0.2, 10.0, 10.0, 0.0, 5.0] and list(p_GP) == [0.1, 5.0, 0.05, 7.0]
    assert utils.n_params_orb == {"SB1": 6, "SB2": 7, "ST1": 11, "ST2": 12, "ST3": 13}


def test_priors_match_reference_bounds():
    from psoap_b200 import sample
    ok = ([0.2, 5.0, 0.2, 10.0, 10.0, 0.0, 5.0], [0.1, 5.0, 0.05, 7.0])
    assert sample.prior_SB2(*ok) == 0.0
    for idx, val in [(0, -0.1), (1, -1.0), (2, 1.5), (3, 451.0), (3, -91.0), (4, -1.0)]:
        po = list(ok[0]); po[idx] = val
        assert sample.prior_SB2(po, ok[1]) == -np.inf
    assert sample.prior_SB2(ok[0], [0.1, -5.0, 0.05, 7.0]) == -np.inf
    assert sample.prior_SB1([5.0, 0.2, 10.0, 10.0, 0.0, 5.0], [0.1, 5.0]) == 0.0


def test_mh_sampler_recovers_a_gaussian():
    from psoap_b200.sample import MHSampler
    mean, sig = np.array([1.0, -2.0]), np.array([0.5, 2.0])
    lnp = lambda p: float(-0.5 * np.sum(((p - mean) / sig) ** 2))
    s = MHSampler(np.diag((1.2 * sig) ** 2), 2, lnp, seed=3)
    s.run_mcmc(mean, 20000)
    assert 0.2 < s.acceptance_fraction < 0.6
    assert np.allclose(s.flatchain[2000:].mean(axis=0), mean, atol=0.1)
    assert np.allclose(s.flatchain[2000:].std(axis=0), sig, rtol=0.1)
    assert s.lnprobability.shape == (20000,)
    s2 = MHSampler(np.diag((1.2 * sig) ** 2), 2, lnp, seed=3)
    s2.run_mcmc(mean, 100)
    assert np.array_equal(s2.flatchain, s.flatchain[:100])  # same seed, same chain: the replicated-rank contract


def test_load_chunk_npz_matches_reference_chunk_semantics(tmp_path):
    """sample.load_chunk: the reference's chunk datasets (data.py:149-197) -> the masked, flattened attributes a
    Chunk exposes after apply_mask() (data.py:120-147), honouring the epoch limit."""
    from psoap_b200 import sample
    rng = np.random.default_rng(4)
    wl = np.tile(np.linspace(5000.0, 5010.0, 30), (6, 1))
    fl, sigma = rng.normal(1.0, 0.01, wl.shape), np.full(wl.shape, 0.01)
    date = np.tile(np.arange(6.0)[:, None], (1, 30))
    mask = rng.uniform(size=wl.shape) > 0.1
    f = str(tmp_path / "chunk_22_5000_5010.npz")
    np.savez(f, wl=wl, fl=fl, sigma=sigma, date=date, mask=mask)
    ch = sample.load_chunk(f, limit=4)
    m = mask[:4]
    assert np.array_equal(ch["lwl"], np.log(wl[:4])[m]) and np.array_equal(ch["fl"], fl[:4][m])
    assert np.array_equal(ch["date1D"], np.arange(4.0)) and ch["mask"].shape == (4, 30)
    assert len(ch["sigma"]) == m.sum()


def test_chunk_container_roundtrip(tmp_path):
    """Chunk mirrors psoap/data.py:120-197: attributes, apply_mask flattening, file naming, limit on open."""
    from psoap_b200 import data
    rng = np.random.default_rng(5)
    n_epochs, n_pix = 6, 9
    wl = 5000.0 * np.exp(np.arange(n_pix) * 2.8 / 2.99792458e5)[None, :] * np.ones((n_epochs, 1))
    date = np.sort(rng.uniform(0, 60, n_epochs))[:, None] * np.ones((1, n_pix))
    mask = rng.uniform(size=wl.shape) > 0.2
    ch = data.Chunk(wl, 1 + 0.01 * rng.normal(size=wl.shape), np.full(wl.shape, 0.04), date, mask)
    assert (ch.n_epochs, ch.n_pix) == (n_epochs, n_pix) and np.array_equal(ch.date1D, date[:, 0])
    fname = ch.save(22, 5160.2, 5190.7, prefix=str(tmp_path) + "/", fmt="npz")
    assert fname.endswith("chunk_22_5160_5191.npz")           # constants.py:39 format
    back = data.Chunk.open(22, 5160.2, 5190.7, limit=4, prefix=str(tmp_path) + "/")
    assert back.wl.shape == (4, n_pix) and back.wl.dtype == np.float64 and back.mask.dtype == bool
    back.apply_mask()
    assert back.N == int(mask[:4].sum())
    assert np.array_equal(back.lwl, np.log(wl[:4])[mask[:4]])
    fc = back.as_farm_chunk()
    assert set(fc) == {"lwl", "fl", "sigma", "mask", "date1D"} and len(fc["date1D"]) == 4
    assert np.allclose(data.redshift(np.array([5000.0]), 30.0), 5000.0 * np.sqrt((2.99792458e5 + 30) / (2.99792458e5 - 30)))


def test_chunks_dat_table(tmp_path):
    from psoap_b200 import data
    rows = [(22, 5160.25, 5190.5), (23, 5200.0, 5230.0)]
    f = str(tmp_path / "chunks.dat")
    data.write_chunks_dat(rows, f)
    assert open(f).readline().split() == ["order", "wl0", "wl1"]
    assert data.read_chunks_dat(f) == rows
    (tmp_path / "bad.dat").write_text("a b c\n1 2 3\n")
    with pytest.raises(ValueError):
        data.read_chunks_dat(str(tmp_path / "bad.dat"))


def test_chain_diagnostics():
    """utils.get_labels / gelman_rubin / estimate_covariance (psoap/utils.py:87-201)."""
    from psoap_b200 import utils
    assert utils.get_labels("SB1", ["gamma", "e"]) == [r"$K$", r"$\omega$", r"$P$", r"$T_0$", r"$a_f$", r"$l_f$"]
    assert len(utils.get_labels("ST3", [])) == len(utils.registered_params["ST3"])
    rng = np.random.default_rng(1)
    good = [rng.normal(size=(400, 2)) for _ in range(3)]
    mean, std, rhat = utils.gelman_rubin(good)
    assert np.all(rhat < 1.05) and np.allclose(std, 1.0, atol=0.1) and np.allclose(mean, 0.0, atol=0.15)
    bad = [good[0], good[1] + np.array([3.0, 0.0])]
    assert utils.gelman_rubin(bad)[2][0] > 1.1 and utils.gelman_rubin(bad)[2][1] < 1.05
    with pytest.raises(AssertionError):
        utils.gelman_rubin([good[0][:399]])
    fc = rng.normal(size=(2000, 3)) * np.array([1.0, 2.0, 0.5])
    oj = utils.estimate_covariance(fc)
    assert np.allclose(oj, 2.38 ** 2 / 3 * np.cov(fc, rowvar=0))


# ------------------------------------------------------------------------------------- property-based checks
from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(min_value=1, max_value=9000), min_size=1, max_size=80), st.integers(min_value=1, max_value=8))
def test_lpt_partition_properties(Ns, nparts):
    """Every chunk is assigned exactly once, empty parts only when there are fewer chunks than parts, and the largest
    load obeys Graham's bound for greedy list scheduling: max <= sum/m + (1 - 1/m) * largest item."""
    from psoap_b200.farm import chunk_cost, lpt_partition
    costs = [chunk_cost(n) for n in Ns]
    parts = lpt_partition(costs, nparts)
    assert len(parts) == nparts
    assert sorted(i for p in parts for i in p) == list(range(len(Ns)))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) <= (sum(costs) / nparts + (1.0 - 1.0 / nparts) * max(costs)) * (1 + 1e-12)
    if len(Ns) >= nparts:
        assert all(len(p) > 0 for p in parts)


@settings(max_examples=60, deadline=None)
@given(st.integers(min_value=1, max_value=12), st.integers(min_value=1, max_value=40), st.integers(min_value=0, max_value=2 ** 31 - 1))
def test_epoch_index_properties(n_epochs, n_pix, seed):
    """epoch_index(mask)[k] is the row of the k-th kept pixel of the row-major flattening (data.py:61), for any mask
    including all-false rows and the empty mask."""
    from psoap_b200.data import epoch_index
    rng = np.random.default_rng(seed)
    mask = rng.uniform(size=(n_epochs, n_pix)) > rng.uniform()
    ep = epoch_index(mask)
    assert ep.dtype == np.int32 and ep.shape == (int(mask.sum()),)
    assert np.array_equal(ep, np.nonzero(mask)[0])
    assert np.all(np.diff(ep) >= 0)


@settings(max_examples=40, deadline=None)
@given(st.sampled_from(["SB1", "SB2", "ST1", "ST2", "ST3"]), st.integers(min_value=0, max_value=2 ** 31 - 1))
def test_convert_vector_dict_round_trip(model, seed):
    """convert_dict picks the fitted parameters in registry order; convert_vector puts them back next to the fixed
    ones and splits at gamma (utils.py:27-85) — for any choice of fixed parameters."""
    from psoap_b200 import utils
    rng = np.random.default_rng(seed)
    names = utils.registered_params[model]
    values = {n: float(rng.normal()) for n in names}
    fixed = [n for n in names if rng.uniform() < 0.3]
    p = utils.convert_dict(model, fixed, **values)
    assert len(p) == len(names) - len(fixed)
    p_orb, p_GP = utils.convert_vector(p, model, fixed, **values)
    assert len(p_orb) == utils.n_params_orb[model] and len(p_orb) + len(p_GP) == len(names)
    assert np.array_equal(np.concatenate([p_orb, p_GP]), np.array([values[n] for n in names]))
    assert names[len(p_orb) - 1] == "gamma"


def test_yaml_config_and_user_prior_override(tmp_path):
    """sample_parallel.py:11-17 (config.yaml) and :362-369 (a prior.py in the run directory replaces the default
    prior): host logic only, with a stand-in for the farm."""
    from psoap_b200 import sample, utils
    cfg = tmp_path / "config.yaml"
    cfg.write_text("model: SB2\nsoften: 1.0\nsamples: 3\nfix_params: [gamma]\n"
                   "parameters: {q: 0.2, K: 5.0, e: 0.2, omega: 10.0, P: 10.0, T0: 0.0, gamma: 5.0, amp_f: 0.1, l_f: 5.0,"
                   " amp_g: 0.05, l_g: 7.0}\n")
    config = sample.load_config(str(cfg))
    assert config["model"] == "SB2" and config["parameters"]["l_g"] == 7.0 and config["fix_params"] == ["gamma"]
    with pytest.raises(FileNotFoundError):
        sample.load_config(str(tmp_path / "missing.yaml"))
    assert sample.load_user_prior(str(tmp_path)) is None
    (tmp_path / "prior.py").write_text("import numpy as np\n\ndef prior(p):\n    return -np.inf if p[1] > 6.0 else -0.5 * p[0] ** 2\n")
    user = sample.load_user_prior(str(tmp_path))

    class FakeFarm:
        def lnprob(self, p):
            return float(np.sum(p))
    pars, fix = config["parameters"], config["fix_params"]
    p0 = utils.convert_dict("SB2", fix, **pars)
    full = np.concatenate(utils.convert_vector(p0, "SB2", fix, **pars))
    default = sample.make_lnprob(FakeFarm(), "SB2", fix, pars)
    custom = sample.make_lnprob(FakeFarm(), "SB2", fix, pars, user_prior=user)
    assert default(p0) == np.sum(full)
    assert custom(p0) == np.sum(full) - 0.5 * p0[0] ** 2
    p1 = p0.copy(); p1[1] = 7.0
    assert custom(p1) == -np.inf and np.isfinite(default(p1))
    p2 = p0.copy(); p2[0] = -0.1            # q < 0: the default bounds reject it, the user prior does not
    assert default(p2) == -np.inf and np.isfinite(custom(p2))
