#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) in the build container.

The reference's own tests pin nothing on this path, so these fixtures are the pin: inputs are
seeded synthetic arrays, outputs come from psoap.matrix_functions (compiled in place by oracle/build_ref.py),
psoap.covariance, psoap.orbit and psoap.data imported from /root/reference with two import shims
(an empty `h5py` module because data.py:4 imports it and it is not installed; `np.float = float` because
covariance.py:102 uses the alias removed in numpy 1.24).  The farm cases compose the reference functions
exactly as sample_parallel.Worker.lnprob does (:181-193) because that module cannot be imported
(emcee/astropy absent, yaml.load without Loader at import).

Run here only:  python tests/golden/make_golden.py
"""
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def import_reference():
    import build_ref
    so = build_ref.build()
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    np.float = float
    sys.path.insert(0, "/root/reference")
    import psoap
    psoap.__path__.append(os.path.dirname(so))
    warnings.simplefilter("ignore")
    from psoap import covariance, matrix_functions, orbit, data
    return covariance, matrix_functions, orbit, data


def main():
    covariance, mf, orbit, data = import_reference()
    from psoap_b200 import synthetic as syn

    # ---------------------------------------------------------------- fills (matrix_functions.pyx)
    rng = np.random.default_rng(11)
    N, M = 70, 45
    lw = np.log(5000.0) + rng.uniform(0, 300.0, size=(3, N)) / syn.c_kms  # unsorted
    lwp = np.log(5000.0) + rng.uniform(0, 300.0, size=M) / syn.c_kms
    amp, l = np.array([0.1, 0.05, 0.03]), np.array([5.0, 7.0, 6.0])
    out = dict(lwl=lw, lwl_predict=lwp, amp=amp, l=l)
    m = np.empty((N, N)); mf.fill_V11_f(m, lw[0], amp[0], l[0]); out["V11_f"] = m.copy()
    m = np.empty((N, N)); mf.fill_V11_f_g(m, lw[0], lw[1], amp[0], l[0], amp[1], l[1]); out["V11_f_g"] = m.copy()
    m = np.empty((N, N)); mf.fill_V11_f_g_h(m, lw[0], lw[1], lw[2], amp[0], l[0], amp[1], l[1], amp[2], l[2])
    out["V11_f_g_h"] = m.copy()
    m = np.empty((N, M)); mf.fill_V12_f(m, lw[0], lwp, amp[0], l[0]); out["V12_f"] = m.copy()
    # far-apart wavelengths: entries in exp's denormal/underflow range
    lwf = np.log(5000.0) + np.linspace(0, 400.0, 33) / syn.c_kms
    m = np.empty((33, 33)); mf.fill_V11_f(m, lwf, 0.5, 5.0); out["far_lwl"] = lwf; out["far_V11_f"] = m.copy()
    np.savez_compressed(os.path.join(HERE, "fills.npz"), **out)

    # ---------------------------------------------------------------- orbits (orbit.py get_velocities)
    dates = np.sort(np.random.default_rng(12).uniform(-30.0, 200.0, 40))
    out = dict(dates=dates)
    psets = {
        "SB1": [[5.0, 0.2, 10.0, 10.0, 0.0, 5.0], [31.0, 0.83, 250.0, 3.7, 12.5, -20.0], [12.0, 0.0, 0.0, 55.0, 7.0, 0.0]],
        "SB2": [[0.2, 5.0, 0.2, 10.0, 10.0, 0.0, 5.0], [0.9, 40.0, 0.6, 130.0, 25.0, 190.0, 14.0]],
        "ST1": [[5.0, 0.2, 10.0, 10.0, 0.0, 4.0, 0.2, 80.0, 100.0, 3.0, 5.0]],
        "ST2": [[0.4, 5.0, 0.2, 10.0, 10.0, 0.0, 4.0, 0.2, 80.0, 100.0, 3.0, 5.0]],
        "ST3": [[0.4, 5.0, 0.2, 10.0, 10.0, 0.0, 0.2, 4.0, 0.2, 80.0, 100.0, 3.0, 5.0],
                [0.7, 22.0, 0.45, 300.0, 6.3, 2.0, 0.5, 9.0, 0.7, 45.0, 410.0, -50.0, -3.0]],
    }
    for model, plist in psets.items():
        for k, p in enumerate(plist):
            out[f"{model}_{k}_p"] = np.array(p)
            out[f"{model}_{k}_v"] = orbit.models[model](*p, dates).get_velocities()
    np.savez_compressed(os.path.join(HERE, "orbits.npz"), **out)

    # ---------------------------------------------------------------- replicate_wls / lnlike / farm
    out = {}
    cases = [("SB1", 6, 40, 0.0, 101), ("SB2", 8, 50, 0.02, 102), ("ST3", 6, 40, 0.05, 103), ("SB2", 5, 37, 0.10, 104)]
    for k, (model, ne, npix, mfrac, seed) in enumerate(cases):
        ch = syn.make_chunk(model, ne, npix, seed, mask_frac=mfrac)
        p = syn.default_params(model)
        n_orb = len(syn.ORBIT_PARAMS[model])
        vel = orbit.models[model](*p[:n_orb], ch["date1D"]).get_velocities()
        lwls = data.replicate_wls(ch["lwl"], vel, ch["mask"])
        V11 = np.empty((ch["N"], ch["N"]))
        lnp = covariance.lnlike[model](V11, *lwls, ch["fl"], ch["sigma"], *p[n_orb:])
        pre = f"case{k}_"
        out[pre + "model"] = model
        for key in ("lwl", "fl", "sigma", "mask", "date1D", "epoch"):
            out[pre + key] = ch[key]
        out[pre + "p"] = p
        out[pre + "vel"] = vel
        out[pre + "lwls"] = lwls
        out[pre + "lnlike"] = lnp
        # mu_GP != default, and sentinels
        out[pre + "lnlike_mu09"] = covariance.lnlike[model](V11, *lwls, ch["fl"], ch["sigma"], *p[n_orb:], mu_GP=0.9)
        pg = p[n_orb:].copy(); pg[0] = -0.1
        out[pre + "lnlike_negamp"] = covariance.lnlike[model](V11, *lwls, ch["fl"], ch["sigma"], *pg)
        # non positive definite: duplicated pixel with zero noise
        lw2 = lwls.copy(); lw2[:, 1] = lw2[:, 0]
        out[pre + "lnlike_nonpd"] = covariance.lnlike[model](V11, *lw2, ch["fl"], np.zeros_like(ch["sigma"]), *p[n_orb:])
    np.savez_compressed(os.path.join(HERE, "lnlike.npz"), **out)

    # ---------------------------------------------------------------- predict_* (covariance.py:81-297)
    out = {}
    ch = syn.make_chunk("ST3", 5, 30, 201)
    p = syn.default_params("ST3")
    vel = orbit.models["ST3"](*p[:13], ch["date1D"]).get_velocities()
    lwls = data.replicate_wls(ch["lwl"], vel, ch["mask"])
    n = ch["N"]
    mgrid = 40
    grid = np.linspace(ch["lwl"].min(), ch["lwl"].max(), mgrid)
    lwp = np.array([grid, grid + 1e-5, grid - 2e-5])
    amp, l = p[13::2], p[14::2]
    out.update(lwls=lwls, fl=ch["fl"], sigma=ch["sigma"], lwl_predict=lwp, amp=amp, l=l)
    mu, Sig = covariance.predict_f_g(lwls[0], lwls[1], ch["fl"], ch["sigma"], lwp[0], lwp[1], 0.7, amp[0], l[0], 0.3,
                                     amp[1], l[1])
    out["fg_mu"], out["fg_Sigma"] = mu, Sig
    out["fg_mu_only"] = covariance.predict_f_g(lwls[0], lwls[1], ch["fl"], ch["sigma"], lwp[0], lwp[1], 0.7, amp[0],
                                               l[0], 0.3, amp[1], l[1], get_Sigma=False)
    mu, Sig = covariance.predict_f_g_sum(lwls[0], lwls[1], ch["fl"], ch["sigma"], lwp[0], lwp[1], 1.0, amp[0], l[0],
                                         amp[1], l[1])
    out["fgsum_mu"], out["fgsum_Sigma"] = mu, Sig
    mu, Sig = covariance.predict_f_g_h(lwls[0], lwls[1], lwls[2], ch["fl"], ch["sigma"], lwp[0], lwp[1], lwp[2], 0.5,
                                       0.3, 0.2, amp[0], l[0], amp[1], l[1], amp[2], l[2])
    out["fgh_mu"], out["fgh_Sigma"] = mu, Sig
    # predict_f_g_h_sum only runs when M == N (mean uses V12.T, covariance.py:294): predict at the data pixels
    mu, Sig = covariance.predict_f_g_h_sum(lwls[0], lwls[1], lwls[2], ch["fl"], ch["sigma"], lwls[0], lwls[1], lwls[2],
                                           1.0, amp[0], l[0], amp[1], l[1], amp[2], l[2])
    out["fghsum_mu"], out["fghsum_Sigma"] = mu, Sig
    try:
        covariance.predict_f(lwls[0], ch["fl"], ch["sigma"], lwp[0], amp[0], l[0])
        out["predict_f_raises"] = "no"
    except NameError:
        out["predict_f_raises"] = "NameError"
    np.savez_compressed(os.path.join(HERE, "predict.npz"), **out)

    # ---------------------------------------------------------------- calibration (covariance.py:560-711)
    out = {}
    rng = np.random.default_rng(31)
    ch = syn.make_chunk("SB1", 4, 60, 301)
    lwl2d = ch["lwl"].reshape(4, 60); fl2d = ch["fl"].reshape(4, 60); sg2d = ch["sigma"].reshape(4, 60)
    fl_cal = fl2d[0] * (1.0 + 0.03 * np.linspace(-1, 1, 60))           # a tilted epoch to calibrate
    lwl_fixed, fl_fixed, sg_fixed = lwl2d[1:].flatten(), fl2d[1:].flatten(), sg2d[1:].flatten()
    lwl0, lwl1 = lwl2d.min(), lwl2d.max()
    amp, l = 0.1, 5.0
    out.update(lwl_cal=lwl2d[0], fl_cal=fl_cal, sigma_cal=sg2d[0], lwl_fixed=lwl_fixed, fl_fixed=fl_fixed,
               sigma_fixed=sg_fixed, lwl0=lwl0, lwl1=lwl1, amp=amp, l=l)
    for order in (1, 2):
        fl_cor, X = covariance.optimize_calibration_static(lwl0, lwl1, lwl2d[0], fl_cal, sg2d[0], lwl_fixed, fl_fixed,
                                                           sg_fixed, amp, l, order=order, mu_GP=1.0)
        out[f"static_o{order}_fl"], out[f"static_o{order}_X"] = fl_cor, X
    A = np.empty((60, 60)); mf.fill_V11_f(A, lwl2d[0], amp, l); A[np.diag_indices_from(A)] += sg2d[0] ** 2
    B = np.empty((180, 180)); mf.fill_V11_f(B, lwl_fixed, amp, l); B[np.diag_indices_from(B)] += sg_fixed ** 2
    Cm = np.empty((60, 180)); mf.fill_V12_f(Cm, lwl2d[0], lwl_fixed, amp, l)
    fl_cor, X = covariance.optimize_calibration(lwl0, lwl1, lwl2d[0], fl_cal, fl_fixed, A, B, Cm, order=1, mu_GP=1.0)
    out["general_fl"], out["general_X"] = fl_cor, X
    out["gp_fit"] = covariance.optimize_GP_f(lwl2d[1], fl2d[1], sg2d[1], 0.2, 8.0)
    np.savez_compressed(os.path.join(HERE, "calibration.npz"), **out)
    for f in ("fills", "orbits", "lnlike", "predict", "calibration"):
        print(f, os.path.getsize(os.path.join(HERE, f + ".npz")), "bytes")


if __name__ == "__main__":
    main()
