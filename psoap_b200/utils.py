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
    """Split-chain potential scale reduction (BDA3 p.284), the statistic psoap/utils.py:99-163 prints.  Takes the
    same list of equally shaped flatchains [iterations, n_params] (even number of iterations); returns
    (mean, std_hat, R_hat) per parameter instead of printing an astropy table, and prints only when asked."""
    chains = np.asarray(samplelist, dtype=np.float64)
    if chains.ndim != 3:
        raise AssertionError("flatchains must all have the same shape [iterations, n_params]")
    n_iter = chains.shape[1]
    if n_iter % 2:
        raise AssertionError("the number of iterations must be even; cut a different number of burn-in samples")
    half = n_iter // 2
    # every chain is split in two: [2 * n_chains, half, n_params]
    split = np.concatenate([chains[:, :half], chains[:, half:]], axis=0)
    n_split = split.shape[0]
    chain_mean = split.mean(axis=1)
    grand_mean = chain_mean.mean(axis=0)
    between = half / (n_split - 1.0) * ((chain_mean - grand_mean) ** 2).sum(axis=0)
    within = split.var(axis=1, ddof=1).mean(axis=0)
    var_hat = (half - 1.0) / half * within + between / half
    std_hat = np.sqrt(var_hat)
    R_hat = np.sqrt(var_hat / within)
    if verbose:
        print("value", grand_mean, "uncertainty", std_hat, "R_hat", R_hat)
        if np.any(R_hat >= 1.1):
            print("not every R_hat is below 1.1: consider running the chains for longer")
    return grand_mean, std_hat, R_hat


def estimate_covariance(flatchain, ndim=0):
    """psoap/utils.py:168-201 without the matplotlib figure: the 'optimal' MH jump covariance 2.38^2/d * cov."""
    d = flatchain.shape[1] if ndim == 0 else ndim
    cov = np.cov(flatchain, rowvar=0)
    return 2.38 ** 2 / d * cov
