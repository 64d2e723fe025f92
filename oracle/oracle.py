"""CPU oracle for PSOAP's GP log-likelihood / prediction hot path (numpy + scipy + oracle C library).

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; the product package psoap_b200/ never does and fails loudly
when its CUDA library is missing.

Parity status: PINNED against outputs of the unmodified reference run in the build container
(tests/golden/make_golden.py -> tests/golden/*.npz, checked by tests/test_oracle.py).  The reference's own
test-suite holds no vector for this path.

Third-party arithmetic: the reference reaches LAPACK dpotrf/dpotrs through scipy.linalg.cho_factor /
cho_solve (psoap/covariance.py:4; scipy unpinned in requirements.txt:2, 1.18.1 + OpenBLAS 0.3.31.dev in this
image).  `lnlike_*`/`predict_*` below call the same scipy entry points; psoap_oracle.c additionally restates
dpotf2/dpotrs in scalar C as a dependency-free cross-check.

Every function cites the reference lines it restates (paths relative to /root/reference).
"""
import ctypes
import os
import subprocess
import sys
import time
import types

import numpy as np
from scipy.linalg import cho_factor, cho_solve
from scipy.optimize import fsolve

HERE = os.path.dirname(os.path.abspath(__file__))
c_kms = 2.99792458e5  # psoap/constants.py:13

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


# --------------------------------------------------------------------------------------------------
# C restatement (oracle/psoap_oracle.c)
# --------------------------------------------------------------------------------------------------
def build_c(force=False):
    """gcc -O2 -ffp-contract=off: the flags of the reference's Cython build (no fast-math, no FMA)."""
    src = os.path.join(HERE, "psoap_oracle.c")
    out_dir = os.path.join(HERE, "_build")
    so = os.path.join(out_dir, "libpsoap_oracle.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", src, "-o", so, "-lm"])
    return so


_clib = None


def clib():
    global _clib
    if _clib is None:
        L = ctypes.CDLL(build_c())
        L.oracle_fill_V11.argtypes = [_dp, ctypes.c_long, ctypes.c_int, ctypes.c_int, _dp, _dp, _dp, _dp, _dp]
        L.oracle_fill_V11.restype = None
        L.oracle_fill_V12_f.argtypes = [_dp, ctypes.c_long, ctypes.c_int, ctypes.c_int, _dp, _dp,
                                        ctypes.c_double, ctypes.c_double]
        L.oracle_fill_V12_f.restype = None
        L.oracle_replicate_wls.argtypes = [_dp, _dp, _ip, ctypes.c_long, _dp, ctypes.c_int, ctypes.c_int]
        L.oracle_replicate_wls.restype = None
        L.oracle_potrf.argtypes = [_dp, ctypes.c_long, ctypes.c_int]
        L.oracle_potrf.restype = ctypes.c_int
        L.oracle_potrs.argtypes = [_dp, ctypes.c_long, ctypes.c_int, _dp]
        L.oracle_potrs.restype = None
        L.oracle_lnlike.argtypes = [_dp, ctypes.c_int, ctypes.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp,
                                    ctypes.c_double]
        L.oracle_lnlike.restype = ctypes.c_double
        _clib = L
    return _clib


def _p(a):
    return a.ctypes.data_as(_dp)


def _vec(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# --------------------------------------------------------------------------------------------------
# The reference's compiled Cython fill (oracle/_ref, built by oracle/build_ref.py) — optional
# --------------------------------------------------------------------------------------------------
_ref_mf = None


def ref_matrix_functions():
    """Import the reference's compiled matrix_functions from oracle/_ref (None if it was never built).

    The .pyx does `import psoap.constants as C` at import time but never uses it (it hard-codes c_kms,
    matrix_functions.pyx:16-17); on the GPU box /root/reference does not exist, so an empty stub package is
    registered for that import only."""
    global _ref_mf
    if _ref_mf is not None:
        return _ref_mf
    sys.path.insert(0, HERE)
    try:
        import build_ref
    finally:
        sys.path.pop(0)
    so = build_ref.build()
    if so is None or not os.path.exists(so):
        return None
    import importlib.machinery
    import importlib.util
    stubbed = []
    if "psoap" not in sys.modules:
        pkg = types.ModuleType("psoap")
        pkg.__path__ = []
        sys.modules["psoap"] = pkg
        stubbed.append("psoap")
    if "psoap.constants" not in sys.modules:
        cst = types.ModuleType("psoap.constants")
        cst.c_kms = c_kms
        sys.modules["psoap.constants"] = cst
        sys.modules["psoap"].constants = cst
        stubbed.append("psoap.constants")
    try:
        loader = importlib.machinery.ExtensionFileLoader("psoap.matrix_functions", so)
        spec = importlib.util.spec_from_loader("psoap.matrix_functions", loader, origin=so)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
    finally:
        for name in stubbed:
            sys.modules.pop(name, None)
    _ref_mf = mod
    return mod


# --------------------------------------------------------------------------------------------------
# Fill operators — psoap/matrix_functions.pyx
# --------------------------------------------------------------------------------------------------
def _check_mat(mat):
    if not (isinstance(mat, np.ndarray) and mat.dtype == np.float64 and mat.ndim == 2 and mat.flags.c_contiguous):
        raise ValueError("mat must be a C-contiguous float64 2-D array")


def fill_V11_f(mat, lwl_f, amp_f, l_f):
    """matrix_functions.pyx:21-57."""
    _check_mat(mat)
    f = _vec(lwl_f)
    clib().oracle_fill_V11(_p(mat), mat.shape[1], len(mat), 1, _p(f), None, None,
                           _p(np.array([amp_f], dtype=np.float64)), _p(np.array([l_f], dtype=np.float64)))


def fill_V11_f_g(mat, lwl_f, lwl_g, amp_f, l_f, amp_g, l_g):
    """matrix_functions.pyx:101-144."""
    _check_mat(mat)
    f, g = _vec(lwl_f), _vec(lwl_g)
    clib().oracle_fill_V11(_p(mat), mat.shape[1], len(mat), 2, _p(f), _p(g), None,
                           _p(np.array([amp_f, amp_g], dtype=np.float64)),
                           _p(np.array([l_f, l_g], dtype=np.float64)))


def fill_V11_f_g_h(mat, lwl_f, lwl_g, lwl_h, amp_f, l_f, amp_g, l_g, amp_h, l_h):
    """matrix_functions.pyx:151-201."""
    _check_mat(mat)
    f, g, h = _vec(lwl_f), _vec(lwl_g), _vec(lwl_h)
    clib().oracle_fill_V11(_p(mat), mat.shape[1], len(mat), 3, _p(f), _p(g), _p(h),
                           _p(np.array([amp_f, amp_g, amp_h], dtype=np.float64)),
                           _p(np.array([l_f, l_g, l_h], dtype=np.float64)))


def fill_V12_f(mat, lwl_f, lwl_predict, amp_f, l_f):
    """matrix_functions.pyx:63-94: M=len(lwl_f) rows, N=len(lwl_predict) columns."""
    _check_mat(mat)
    f, p = _vec(lwl_f), _vec(lwl_predict)
    clib().oracle_fill_V12_f(_p(mat), mat.shape[1], len(f), len(p), _p(f), _p(p), float(amp_f), float(l_f))


# --------------------------------------------------------------------------------------------------
# Doppler shift — psoap/data.py
# --------------------------------------------------------------------------------------------------
def lredshift(lwl, v):
    """data.py:25-38."""
    return lwl + v / c_kms


def replicate_wls(lwls, velocities, mask):
    """data.py:40-63: per component, blue-shift the masked flattened ln-wavelengths by that epoch's velocity."""
    n_components, n_epochs = velocities.shape
    n_good_pix = np.sum(mask)
    lwls_out = np.empty((n_components, n_good_pix), dtype=np.float64)
    for i in range(n_components):
        lwls_out[i] = lredshift(lwls, (-velocities[i][:, np.newaxis] * np.ones_like(mask))[mask])
    return lwls_out


# --------------------------------------------------------------------------------------------------
# Orbits — psoap/orbit.py (get_velocities of SB1/SB2/ST1/ST2/ST3)
# --------------------------------------------------------------------------------------------------
def _true_anomaly(t, T0, P, e):
    """orbit.py:47-72 (and :210-250 for the triple's inner/outer orbits)."""
    t = (t - T0) % P  # Python modulus: result in [0, P)
    f = lambda E: E - e * np.sin(E) - 2 * np.pi * t / P
    E0 = 2 * np.pi * t / P
    E = fsolve(f, E0)[0]
    th = 2 * np.arctan(np.sqrt((1 + e) / (1 - e)) * np.tan(E / 2.))
    if E < np.pi:
        return th
    return th + 2 * np.pi


def _v(K, e, omega_deg, f):
    """orbit.py:74-81 / :134-140 / :252-263 / :444-449: K (cos(omega + f) + e cos omega), omega in degrees."""
    return K * (np.cos(omega_deg * np.pi / 180 + f) + e * np.cos(omega_deg * np.pi / 180))


def get_velocities(model, p_orb, dates):
    """orbit.py get_velocities(): SB1 :94-115, SB2 :148-170, ST1 :295-320, ST2 :390-417, ST3 :463-487.
    p_orb is in utils.registered_params order (utils.py:4-8).  Returns [ncomp, n_epochs]."""
    dates = np.atleast_1d(dates)
    if model in ("SB1", "SB2"):
        if model == "SB1":
            K, e, omega, P, T0, gamma = p_orb
            q = None
        else:
            q, K, e, omega, P, T0, gamma = p_orb
        assert (e >= 0.0) and (e < 1.0), "Eccentricity must be between [0, 1)"  # orbit.py:33
        vA = np.array([_v(K, e, omega, _true_anomaly(t, T0, P, e)) + gamma for t in dates])
        if model == "SB1":
            return np.atleast_2d(vA)
        vB = np.array([_v(K / q, e, omega + 180, _true_anomaly(t, T0, P, e)) + gamma for t in dates])
        return np.vstack((vA, vB))
    if model == "ST1":
        K_in, e_in, omega_in, P_in, T0_in, K_out, e_out, omega_out, P_out, T0_out, gamma = p_orb
        q_in = q_out = None
    elif model == "ST2":
        q_in, K_in, e_in, omega_in, P_in, T0_in, K_out, e_out, omega_out, P_out, T0_out, gamma = p_orb
        q_out = None
    elif model == "ST3":
        q_in, K_in, e_in, omega_in, P_in, T0_in, q_out, K_out, e_out, omega_out, P_out, T0_out, gamma = p_orb
    else:
        raise KeyError(model)
    assert (e_in >= 0.0) and (e_in < 1.0), "Inner eccentricity must be between [0, 1)"
    assert (e_out >= 0.0) and (e_out < 1.0), "Outer eccentricity must be between [0, 1)"
    f_in = [_true_anomaly(t, T0_in, P_in, e_in) for t in dates]
    f_out = [_true_anomaly(t, T0_out, P_out, e_out) for t in dates]
    # orbit.py:265-275: vA = v1 + v3 + gamma
    vA = np.array([_v(K_in, e_in, omega_in, fi) + _v(K_out, e_out, omega_out, fo) + gamma
                   for fi, fo in zip(f_in, f_out)])
    if model == "ST1":
        return np.atleast_2d(vA)
    # orbit.py:345-363: vB = v2 + v3 + gamma, v2 with K_in/q_in and omega_in + 180
    vB = np.array([_v(K_in / q_in, e_in, omega_in + 180, fi) + _v(K_out, e_out, omega_out, fo) + gamma
                   for fi, fo in zip(f_in, f_out)])
    if model == "ST2":
        return np.vstack((vA, vB))
    # orbit.py:444-460: vC = K_out/q_out (cos(omega_out + 180 + f_out) + e_out cos(...)) + gamma
    vC = np.array([_v(K_out / q_out, e_out, omega_out + 180, fo) + gamma for fo in f_out])
    return np.vstack((vA, vB, vC))


# --------------------------------------------------------------------------------------------------
# Likelihood — psoap/covariance.py:299-379
# --------------------------------------------------------------------------------------------------
_FILLS = {1: "fill_V11_f", 2: "fill_V11_f_g", 3: "fill_V11_f_g_h"}


def _lnlike(V11, lwls, fl, sigma, amps, ls, mu_GP, use_ref_fill, cho_kwargs, timers=None):
    """`timers` (measurement only): dict that receives the seconds spent in the fill and in LAPACK."""
    if any(a < 0.0 for a in amps) or any(l < 0.0 for l in ls):
        return -np.inf
    ncomp = len(lwls)
    args = list(lwls) + [x for pair in zip(amps, ls) for x in pair]
    mod = ref_matrix_functions() if use_ref_fill else None
    t0 = time.perf_counter()
    if mod is not None:
        getattr(mod, _FILLS[ncomp])(V11, *args)
    else:
        globals()[_FILLS[ncomp]](V11, *args)
    V11[np.diag_indices_from(V11)] += sigma ** 2
    t1 = time.perf_counter()
    try:
        factor, flag = cho_factor(V11, **cho_kwargs)
    except np.linalg.LinAlgError:
        return -np.inf
    logdet = np.sum(2 * np.log((np.diag(factor))))
    out = -0.5 * (np.dot((fl - mu_GP).T, cho_solve((factor, flag), (fl - mu_GP))) + logdet)
    if timers is not None:
        timers["fill"] = timers.get("fill", 0.0) + (t1 - t0)
        timers["lapack"] = timers.get("lapack", 0.0) + (time.perf_counter() - t1)
    return out


def lnlike_f(V11, wl_f, fl, sigma, amp_f, l_f, mu_GP=1., use_ref_fill=False, timers=None):
    """covariance.py:299-331 (cho_factor with defaults: copy, finite check)."""
    return _lnlike(V11, [wl_f], fl, sigma, [amp_f], [l_f], mu_GP, use_ref_fill, {}, timers)


def lnlike_f_g(V11, wl_f, wl_g, fl, sigma, amp_f, l_f, amp_g, l_g, mu_GP=1., use_ref_fill=False, timers=None):
    """covariance.py:333-354 (cho_factor(overwrite_a=True, lower=False, check_finite=False), :348)."""
    return _lnlike(V11, [wl_f, wl_g], fl, sigma, [amp_f, amp_g], [l_f, l_g], mu_GP, use_ref_fill,
                   dict(overwrite_a=True, lower=False, check_finite=False), timers)


def lnlike_f_g_h(V11, wl_f, wl_g, wl_h, fl, sigma, amp_f, l_f, amp_g, l_g, amp_h, l_h, mu_GP=1.,
                 use_ref_fill=False, timers=None):
    """covariance.py:356-376."""
    return _lnlike(V11, [wl_f, wl_g, wl_h], fl, sigma, [amp_f, amp_g, amp_h], [l_f, l_g, l_h], mu_GP,
                   use_ref_fill, {}, timers)


lnlike = {"SB1": lnlike_f, "SB2": lnlike_f_g, "ST1": lnlike_f, "ST2": lnlike_f_g, "ST3": lnlike_f_g_h}  # :379


def lnlike_c(V11, lwls, fl, sigma, amps, ls, mu_GP=1.):
    """All-C path (psoap_oracle.c: fill + scalar dpotf2/dpotrs) — independent of scipy/LAPACK."""
    lw = [_vec(x) for x in lwls] + [None] * (3 - len(lwls))
    ptrs = [(_p(x) if x is not None else None) for x in lw]
    fl, sigma = _vec(fl), _vec(sigma)
    a, l = np.array(amps, dtype=np.float64), np.array(ls, dtype=np.float64)
    return clib().oracle_lnlike(_p(V11), len(fl), len(lwls), ptrs[0], ptrs[1], ptrs[2], _p(fl), _p(sigma),
                                _p(a), _p(l), float(mu_GP))


# --------------------------------------------------------------------------------------------------
# Prediction — psoap/covariance.py:25-297
# --------------------------------------------------------------------------------------------------
def _K11(lwl, amp, l):
    m = np.empty((len(lwl), len(lwl)), dtype=np.float64)
    fill_V11_f(m, lwl, amp, l)
    return m


def _K12(lwl_rows, lwl_cols, amp, l):
    m = np.empty((len(lwl_rows), len(lwl_cols)), dtype=np.float64)
    fill_V12_f(m, lwl_rows, lwl_cols, amp, l)
    return m


def predict_f(lwl_known, fl_known, sigma_known, lwl_predict, amp_f, l_f, mu_GP=1.0):
    """covariance.py:25-54.  The reference body raises NameError (`wl_predict` undefined at :38); this is the
    evident intent with `lwl_predict` substituted."""
    M = len(lwl_known)
    V11 = _K11(lwl_known, amp_f, l_f) + sigma_known ** 2 * np.eye(M)
    V12 = _K12(lwl_known, lwl_predict, amp_f, l_f)
    V22 = _K11(lwl_predict, amp_f, l_f)
    factor, flag = cho_factor(V11)
    mu = mu_GP + np.dot(V12.T, cho_solve((factor, flag), (fl_known - mu_GP)))
    Sigma = V22 - np.dot(V12.T, cho_solve((factor, flag), V12))
    return (mu, Sigma)


def _predict_components(lwls, fl, sigma, lwls_predict, mus, amps, ls, get_Sigma=True):
    """Common body of predict_f_g (covariance.py:81-148) and predict_f_g_h (:190-251)."""
    n_pix_predict = len(lwls_predict[0])
    mu_cat = np.hstack([mu * np.ones(n_pix_predict) for mu in mus])
    B = sum(_K11(lw, a, l) for lw, a, l in zip(lwls, amps, ls))
    B[np.diag_indices_from(B)] += sigma ** 2
    factor, flag = cho_factor(B)
    ncomp = len(lwls)
    A = np.zeros((ncomp * n_pix_predict, ncomp * n_pix_predict))
    for c in range(ncomp):
        s = slice(c * n_pix_predict, (c + 1) * n_pix_predict)
        A[s, s] = _K11(lwls_predict[c], amps[c], ls[c])
    C = np.vstack([_K12(lwls_predict[c], lwls[c], amps[c], ls[c]) for c in range(ncomp)])
    mu = mu_cat + np.dot(C, cho_solve((factor, flag), fl - 1.0))  # hard-coded 1.0, covariance.py:140,:248
    if get_Sigma:
        Sigma = A - np.dot(C, cho_solve((factor, flag), C.T))
        return mu, Sigma
    return mu


def predict_f_g(lwl_f, lwl_g, fl_fg, sigma_fg, lwl_f_predict, lwl_g_predict, mu_f, amp_f, l_f, mu_g, amp_g, l_g,
                get_Sigma=True):
    """covariance.py:81-148."""
    assert len(lwl_f) == len(lwl_g), "Input wavelengths must be the same length."
    assert len(lwl_f_predict) == len(lwl_g_predict), "Prediction wavelengths must be the same length."
    return _predict_components([lwl_f, lwl_g], fl_fg, sigma_fg, [lwl_f_predict, lwl_g_predict], [mu_f, mu_g],
                               [amp_f, amp_g], [l_f, l_g], get_Sigma)


def predict_f_g_h(lwl_f, lwl_g, lwl_h, fl_fgh, sigma_fgh, lwl_f_predict, lwl_g_predict, lwl_h_predict, mu_f, mu_g,
                  mu_h, amp_f, l_f, amp_g, l_g, amp_h, l_h):
    """covariance.py:190-251."""
    assert len(lwl_f) == len(lwl_g), "Input wavelengths must be the same length."
    assert len(lwl_f) == len(lwl_h), "Input wavelengths must be the same length."
    assert len(lwl_f_predict) == len(lwl_g_predict), "Prediction wavelengths must be the same length."
    assert len(lwl_f_predict) == len(lwl_h_predict), "Prediction wavelengths must be the same length."
    return _predict_components([lwl_f, lwl_g, lwl_h], fl_fgh, sigma_fgh,
                               [lwl_f_predict, lwl_g_predict, lwl_h_predict], [mu_f, mu_g, mu_h],
                               [amp_f, amp_g, amp_h], [l_f, l_g, l_h], True)


def _predict_sum(lwls, fl, sigma, lwls_predict, amps, ls, nugget, mean_fn):
    V11 = sum(_K11(lp, a, l) for lp, a, l in zip(lwls_predict, amps, ls))
    if nugget:
        V11[np.diag_indices_from(V11)] += nugget
    V12 = sum(_K12(lp, lw, a, l) for lp, lw, a, l in zip(lwls_predict, lwls, amps, ls))
    V22 = sum(_K11(lw, a, l) for lw, a, l in zip(lwls, amps, ls))
    V22[np.diag_indices_from(V22)] += sigma ** 2
    factor, flag = cho_factor(V22)
    mu = mean_fn(V12, factor, flag)
    Sigma = V11 - np.dot(V12, cho_solve((factor, flag), V12.T))
    return mu, Sigma


def predict_f_g_sum(lwl_f, lwl_g, fl_fg, sigma_fg, lwl_f_predict, lwl_g_predict, mu_fg, amp_f, l_f, amp_g, l_g):
    """covariance.py:151-187 (nugget 1e-8 on V11 at :165; mean uses fl - 1.0 at :184)."""
    assert len(lwl_f) == len(lwl_g), "Input wavelengths must be the same length."
    return _predict_sum([lwl_f, lwl_g], fl_fg, sigma_fg, [lwl_f_predict, lwl_g_predict], [amp_f, amp_g],
                        [l_f, l_g], 1e-8,
                        lambda V12, factor, flag: mu_fg + np.dot(V12, cho_solve((factor, flag), (fl_fg - 1.0))))


def predict_f_g_h_sum(lwl_f, lwl_g, lwl_h, fl_fgh, sigma_fgh, lwl_f_predict, lwl_g_predict, lwl_h_predict, mu_fgh,
                      amp_f, l_f, amp_g, l_g, amp_h, l_h):
    """covariance.py:253-297.  Reference quirk kept: the mean multiplies by V12.T (:294), so it only runs when
    M == N; no nugget (:270 commented out); mean uses fl - mu_fgh."""
    assert len(lwl_f) == len(lwl_g), "Input wavelengths must be the same length."
    return _predict_sum([lwl_f, lwl_g, lwl_h], fl_fgh, sigma_fgh, [lwl_f_predict, lwl_g_predict, lwl_h_predict],
                        [amp_f, amp_g, amp_h], [l_f, l_g, l_h], 0.0,
                        lambda V12, factor, flag: mu_fgh + np.dot(V12.T, cho_solve((factor, flag),
                                                                                 (fl_fgh - mu_fgh))))


# --------------------------------------------------------------------------------------------------
# Chunk farm — psoap/sample_parallel.py:168-198 (Worker.lnprob) and :371-390 (master lnprob, no prior)
# --------------------------------------------------------------------------------------------------
n_params_orb = {"SB1": 6, "SB2": 7, "ST1": 11, "ST2": 12, "ST3": 13}  # utils.py:14 (index of gamma + 1)


def chunk_lnprob(model, p_full, chunk, V11=None, use_ref_fill=False, timers=None):
    """sample_parallel.py:168-198 for one chunk.  `chunk` has lwl, fl, sigma (masked 1-D), mask, date1D.
    p_full is the full registered parameter vector (orbital then GP)."""
    p_orb, p_GP = p_full[:n_params_orb[model]], p_full[n_params_orb[model]:]
    velocities = get_velocities(model, p_orb, chunk["date1D"])
    if np.any(np.abs(np.array(velocities)) >= c_kms):  # :186-187
        return -np.inf
    lwls = replicate_wls(chunk["lwl"], velocities, chunk["mask"])
    N = len(chunk["fl"])
    if V11 is None:
        V11 = np.empty((N, N), dtype=np.float64)  # :163
    return lnlike[model](V11, *lwls, chunk["fl"], chunk["sigma"], *p_GP, use_ref_fill=use_ref_fill, timers=timers)


def farm_lnprob(model, p_full, chunks, use_ref_fill=False):
    """sample_parallel.py:378-387: gather one float per chunk, np.sum in chunk order."""
    lnps = np.empty(len(chunks))
    for i, ch in enumerate(chunks):
        lnps[i] = chunk_lnprob(model, p_full, ch, use_ref_fill=use_ref_fill)
    return np.sum(lnps), lnps


# --------------------------------------------------------------------------------------------------
# Calibration — psoap/covariance.py:560-711
# --------------------------------------------------------------------------------------------------
def _cheb_design(x0, x1, x, fl_cal, order):
    """covariance.py:584-595 / :669-680."""
    from numpy.polynomial import Chebyshev as Ch
    T = []
    for i in range(0, order + 1):
        coeff = [0 for j in range(i)] + [1]
        T += [Ch(coeff, domain=[x0, x1])(x)]
    T = np.array(T)
    return fl_cal[:, np.newaxis] * T.T


def _calibration_solve(A, B, C, D, fl_fixed, mu_GP):
    """covariance.py:598-624 (identical body at :686-711)."""
    B_cho = cho_factor(B)
    fl_prime = mu_GP + np.dot(C, cho_solve(B_cho, (fl_fixed.flatten() - mu_GP)))
    C_prime = A - np.dot(C, cho_solve(B_cho, C.T))
    CP_cho = cho_factor(C_prime)
    left = np.dot(D.T, cho_solve(CP_cho, D))
    right = np.dot(D.T, cho_solve(CP_cho, fl_prime))
    X = cho_solve(cho_factor(left), right)
    return np.dot(D, X), X


def optimize_calibration(lwl0, lwl1, lwl_cal, fl_cal, fl_fixed, A, B, C, order=1, mu_GP=1.0):
    """covariance.py:560-624."""
    return _calibration_solve(A, B, C, _cheb_design(lwl0, lwl1, lwl_cal, fl_cal, order), fl_fixed, mu_GP)


def optimize_calibration_static(wl0, wl1, wl_cal, fl_cal, sigma_cal, wl_fixed, fl_fixed, sigma_fixed, amp, l_f, order=1,
                                mu_GP=1.0):
    """covariance.py:627-711."""
    A, B = _K11(wl_cal, amp, l_f), _K11(wl_fixed, amp, l_f)
    C = _K12(wl_cal, wl_fixed, amp, l_f)
    A[np.diag_indices_from(A)] += sigma_cal ** 2
    B[np.diag_indices_from(B)] += sigma_fixed ** 2
    return _calibration_solve(A, B, C, _cheb_design(wl0, wl1, wl_cal, fl_cal, order), fl_fixed, mu_GP)


def optimize_GP_f(wl_known, fl_known, sigma_known, amp_f, l_f, mu_GP=1.0):
    """covariance.py:408-424."""
    from scipy.optimize import minimize
    N = len(wl_known)
    V11 = np.empty((N, N), dtype=np.float64)
    func = lambda x: -lnlike_f(V11, wl_known, fl_known, sigma_known, x[0], x[1], mu_GP)
    return minimize(func, np.array([amp_f, l_f]), method="Nelder-Mead")["x"]
