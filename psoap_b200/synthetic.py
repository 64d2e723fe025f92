"""Seeded synthetic spectroscopic chunks.

Restates the recipe of the reference's scripts/fake/make_fake_{SB1,SB2,ST3}.py without the private LkCa14
template: a common barycentric ln-wavelength grid per epoch, Keplerian velocities per epoch, per-component
absorption-line templates evaluated in each component's rest frame, flux-ratio mixing, Gaussian noise.
Host-side numpy only; used by bench.py, the tests and the golden-vector generator to make INPUTS.
Nothing here is on the likelihood path.
"""
import numpy as np

c_kms = 2.99792458e5

# scripts/fake/make_fake_SB1.py:13-18, make_fake_SB2.py:18-24, make_fake_ST3.py:14-26
ORBIT_PARAMS = {
    "SB1": [5.0, 0.2, 10.0, 10.0, 0.0, 5.0],                                            # K e omega P T0 gamma
    "SB2": [0.2, 5.0, 0.2, 10.0, 10.0, 0.0, 5.0],                                       # q K e omega P T0 gamma
    "ST3": [0.4, 5.0, 0.2, 10.0, 10.0, 0.0, 0.2, 4.0, 0.2, 80.0, 100.0, 3.0, 5.0],
}
GP_PARAMS = {"SB1": [0.1, 5.0], "SB2": [0.1, 5.0, 0.05, 7.0], "ST3": [0.1, 5.0, 0.05, 7.0, 0.03, 6.0]}
NCOMP = {"SB1": 1, "SB2": 2, "ST1": 1, "ST2": 2, "ST3": 3}
# flux fractions of the components (SB2 ratio=0.2, make_fake_SB2.py:124-125; ST3 alpha=0.5 beta=0.3,
# make_fake_ST3.py:131-132) and per-pixel noise (make_fake_SB1.py:86-88, SB2 :130-131, ST3 :163-165)
FLUX_FRAC = {"SB1": [1.0], "SB2": [1 / 1.2, 0.2 / 1.2], "ST3": [0.5, 0.3, 0.2]}
NOISE = {"SB1": 1.0 / 25, "SB2": 1.0 / (60 / np.sqrt(2.5)), "ST3": 1.0 / 40}


def _true_anomaly(t, T0, P, e):
    """Newton solve of Kepler's equation, vectorised over dates (input generation only)."""
    t = np.mod(t - T0, P)
    M = 2 * np.pi * t / P
    E = M.copy()
    for _ in range(60):
        E = E - (E - e * np.sin(E) - M) / (1 - e * np.cos(E))
    th = 2 * np.arctan(np.sqrt((1 + e) / (1 - e)) * np.tan(E / 2.0))
    return np.where(E < np.pi, th, th + 2 * np.pi)


def _v(K, e, omega_deg, f):
    return K * (np.cos(omega_deg * np.pi / 180 + f) + e * np.cos(omega_deg * np.pi / 180))


def host_velocities(model, p_orb, dates):
    """[ncomp, n_epochs] Keplerian radial velocities (same conventions as the reference's orbit.py)."""
    dates = np.asarray(dates, dtype=np.float64)
    if model == "SB1":
        K, e, om, P, T0, gam = p_orb
        return np.atleast_2d(_v(K, e, om, _true_anomaly(dates, T0, P, e)) + gam)
    if model == "SB2":
        q, K, e, om, P, T0, gam = p_orb
        f = _true_anomaly(dates, T0, P, e)
        return np.vstack((_v(K, e, om, f) + gam, _v(K / q, e, om + 180, f) + gam))
    if model == "ST3":
        q_in, K_in, e_in, om_in, P_in, T0_in, q_out, K_out, e_out, om_out, P_out, T0_out, gam = p_orb
        fi = _true_anomaly(dates, T0_in, P_in, e_in)
        fo = _true_anomaly(dates, T0_out, P_out, e_out)
        v3 = _v(K_out, e_out, om_out, fo)
        return np.vstack((_v(K_in, e_in, om_in, fi) + v3 + gam,
                          _v(K_in / q_in, e_in, om_in + 180, fi) + v3 + gam,
                          _v(K_out / q_out, e_out, om_out + 180, fo) + gam))
    raise KeyError(model)


def _template(rng, lwl_lo, lwl_hi, dpix):
    """A rest-frame absorption-line template: ~1 Gaussian line per 20 px over a padded range."""
    pad = 40.0 / c_kms
    lo, hi = lwl_lo - pad, lwl_hi + pad
    n_lines = max(1, int((hi - lo) / dpix / 20))
    centers = rng.uniform(lo, hi, n_lines)
    depths = rng.uniform(0.05, 0.5, n_lines)
    widths = rng.uniform(2.0, 4.0, n_lines) * dpix

    def f(lwl):
        out = np.ones_like(lwl)
        for c, d, w in zip(centers, depths, widths):
            out -= d * np.exp(-0.5 * ((lwl - c) / w) ** 2)
        return out
    return f


def make_chunk(model, n_epochs, n_pix, seed, mask_frac=0.0, wl0=5000.0, dv_pix=2.8, p_orb=None):
    """One synthetic chunk.

    Returns a dict with the arrays a reference `Chunk` exposes after `apply_mask()` (data.py:120-147):
    lwl, fl, sigma (1-D, length N = number of unmasked pixels, epoch-major), mask [n_epochs, n_pix] bool,
    date1D [n_epochs], plus `epoch` (int32 [N], epoch index of every kept pixel) and N.
    """
    rng = np.random.default_rng(seed)
    p_orb = ORBIT_PARAMS[model] if p_orb is None else p_orb
    dpix = dv_pix / c_kms
    lwl_grid = np.log(wl0) + np.arange(n_pix) * dpix
    lwl2d = np.tile(lwl_grid, (n_epochs, 1))
    dates = np.sort(rng.uniform(0.0, 60.0, n_epochs))
    vel = host_velocities(model, p_orb, dates)
    ncomp = NCOMP[model]
    fl2d = np.zeros((n_epochs, n_pix))
    for c in range(ncomp):
        tmpl = _template(rng, lwl_grid[0], lwl_grid[-1], dpix)
        rest = lwl2d - vel[c][:, None] / c_kms
        fl2d += FLUX_FRAC[model][c] * tmpl(rest)
    sig = NOISE[model]
    fl2d = fl2d + rng.normal(0.0, sig, size=fl2d.shape)
    sigma2d = sig * np.ones((n_epochs, n_pix))
    mask = np.ones((n_epochs, n_pix), dtype=bool)
    if mask_frac > 0:
        mask &= rng.uniform(size=mask.shape) >= mask_frac
    epoch2d = np.tile(np.arange(n_epochs, dtype=np.int32)[:, None], (1, n_pix))
    return dict(model=model, lwl=np.ascontiguousarray(lwl2d[mask]), fl=np.ascontiguousarray(fl2d[mask]),
                sigma=np.ascontiguousarray(sigma2d[mask]), mask=mask, date1D=dates,
                epoch=np.ascontiguousarray(epoch2d[mask]), N=int(mask.sum()), n_epochs=n_epochs, n_pix=n_pix)


def default_params(model):
    """Full registered parameter vector (orbital then GP), psoap/utils.py:4-8 order."""
    return np.array(list(ORBIT_PARAMS[model]) + list(GP_PARAMS[model]), dtype=np.float64)


# BASELINE.json configs
def config_chunks(name):
    if name == "C1":
        return "SB1", [make_chunk("SB1", 20, 200, seed=1)]
    if name == "C2":
        return "SB2", [make_chunk("SB2", 30, 300, seed=2)]
    if name == "C3":
        return "ST3", [make_chunk("ST3", 40, 250, seed=3)]
    if name == "C4":
        return "SB2", [make_chunk("SB2", 20, 100 + (200 * i) // 255, seed=4000 + i, wl0=5000.0 + 4.0 * i)
                       for i in range(256)]
    if name == "C5":
        return "SB2", [make_chunk("SB2", 64, 512, seed=5)]
    if name == "C6":   # the reference's own chunk size: 20 epochs x 80 px (scripts/psoap_generate_chunks.py:8-9), many chunks
        return "SB2", [make_chunk("SB2", 20, 80, seed=6000 + i, wl0=5000.0 + 1.5 * i) for i in range(512)]
    raise KeyError(name)
