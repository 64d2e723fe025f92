"""Parameter registry and vector <-> (orbital, GP) conversion (mirror of psoap/utils.py:4-85).

Host-side glue only: it produces the <= 19 doubles per proposal that the chunk farm consumes.
"""
import numpy as np

# psoap/utils.py:4-8
registered_params = {
    "SB1": ["K", "e", "omega", "P", "T0", "gamma", "amp_f", "l_f"],
    "SB2": ["q", "K", "e", "omega", "P", "T0", "gamma", "amp_f", "l_f", "amp_g", "l_g"],
    "ST1": ["K_in", "e_in", "omega_in", "P_in", "T0_in", "K_out", "e_out", "omega_out", "P_out", "T0_out", "gamma",
            "amp_f", "l_f"],
    "ST2": ["q_in", "K_in", "e_in", "omega_in", "P_in", "T0_in", "K_out", "e_out", "omega_out", "P_out", "T0_out",
            "gamma", "amp_f", "l_f"],
    "ST3": ["q_in", "K_in", "e_in", "omega_in", "P_in", "T0_in", "q_out", "K_out", "e_out", "omega_out", "P_out",
            "T0_out", "gamma", "amp_f", "l_f", "amp_g", "l_g", "amp_h", "l_h"],
}
registered_models = registered_params.keys()
# psoap/utils.py:14: number of orbital parameters = position of gamma + 1
n_params_orb = {model: (registered_params[model].index("gamma") + 1) for model in registered_params}


def convert_vector(p, model, fix_params, **kwargs):
    """psoap/utils.py:27-69: unroll the vector of fitted values into the full (orbital, GP) parameter vectors,
    back-filling the fixed parameters from `kwargs`."""
    reg_params = registered_params[model]
    fit_ind = [i for (i, param) in enumerate(reg_params) if param not in fix_params]
    fix_ind = [reg_params.index(param) for param in fix_params]
    par_vec = np.empty(len(reg_params), dtype=np.float64)
    par_vec[fit_ind] = p
    par_vec[fix_ind] = np.array([kwargs[name] for name in fix_params])
    ind_split = n_params_orb[model]
    return (par_vec[:ind_split], par_vec[ind_split:])


def convert_dict(model, fix_params, **kwargs):
    """psoap/utils.py:72-85: dictionary of parameter values -> vector of the fitted ones, registry order."""
    fit_params = [param for param in registered_params[model] if param not in fix_params]
    return np.array([kwargs[name] for name in fit_params], dtype=np.float64)


# psoap/utils.py:17-19 (plot labels; the reference registers none for ST1/ST2)
registered_labels = {
    "SB1": [r"$K$", r"$e$", r"$\omega$", r"$P$", r"$T_0$", r"$\gamma$", r"$a_f$", r"$l_f$"],
    "SB2": [r"$q$", r"$K$", r"$e$", r"$\omega$", r"$P$", r"$T_0$", r"$\gamma$", r"$a_f$", r"$l_f$", r"$a_g$", r"$l_g$"],
    "ST3": [r"$q_\mathrm{in}$", r"$K_\mathrm{in}$", r"$e_\mathrm{in}$", r"$\omega_\mathrm{in}$", r"$P_\mathrm{in}$",
            r"$T_{0,\mathrm{in}}$", r"$q_\mathrm{out}$", r"$K_\mathrm{out}$", r"$e_\mathrm{out}$",
            r"$\omega_\mathrm{out}$", r"$P_\mathrm{out}$", r"$T_{0,\mathrm{out}}$", r"$\gamma$", r"$a_f$", r"$l_f$",
            r"$a_g$", r"$l_g$", r"$a_h$", r"$l_h$"],
}


def get_labels(model, fix_params):
    """psoap/utils.py:87-97: labels of the fitted parameters, registry order."""
    reg_params = registered_params[model]
    reg_labels = registered_labels[model]
    return [reg_labels[i] for (i, param) in enumerate(reg_params) if param not in fix_params]


def gelman_rubin(samplelist, verbose=False):
    """psoap/utils.py:99-163 (BDA3 p.284, split chains).  The reference prints an astropy table and returns
    nothing; this returns (mean, std_hat, R_hat) so the numbers can be used, and prints only when asked."""
    full_iterations = len(samplelist[0])
    assert full_iterations % 2 == 0, "Number of iterations must be even. Try cutting off a different number of burn in samples."
    shape = samplelist[0].shape
    for flatchain in samplelist:
        assert len(flatchain) == full_iterations, "Not all chains have the same number of iterations!"
        assert flatchain.shape == shape, "Not all flatchains have the same shape!"
    n = full_iterations // 2
    m = 2 * len(samplelist)
    nparams = samplelist[0].shape[-1]
    chains = np.empty((n, m, nparams))
    for k, flatchain in enumerate(samplelist):
        chains[:, 2 * k, :] = flatchain[:n]
        chains[:, 2 * k + 1, :] = flatchain[n:]
    avg_phi_j = np.mean(chains, axis=0, dtype="f8")
    avg_phi = np.mean(chains, axis=(0, 1), dtype="f8")
    B = n / (m - 1.0) * np.sum((avg_phi_j - avg_phi) ** 2, axis=0, dtype="f8")
    s2j = 1.0 / (n - 1.0) * np.sum((chains - avg_phi_j) ** 2, axis=0, dtype="f8")
    W = 1.0 / m * np.sum(s2j, axis=0, dtype="f8")
    var_hat = (n - 1.0) / n * W + B / n
    std_hat = np.sqrt(var_hat)
    R_hat = np.sqrt(var_hat / W)
    if verbose:
        print("Value:", avg_phi)
        print("Uncertainty:", std_hat)
        print("R_hat: {}".format(R_hat))
        if np.any(R_hat >= 1.1):
            print("You might consider running the chain for longer. Not all R_hats are less than 1.1.")
    return avg_phi, std_hat, R_hat


def estimate_covariance(flatchain, ndim=0):
    """psoap/utils.py:168-201 without the matplotlib figure: the 'optimal' MH jump covariance 2.38^2/d * cov."""
    d = flatchain.shape[1] if ndim == 0 else ndim
    cov = np.cov(flatchain, rowvar=0)
    return 2.38 ** 2 / d * cov
