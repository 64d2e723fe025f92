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
