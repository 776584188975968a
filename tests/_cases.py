"""Shared helpers: load golden fixtures and build the oracle energy each one used."""

import os

import numpy as np
import torch

from oracle import energies as E

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {}
    for k in z.files:
        a = z[k]
        out[k] = torch.from_numpy(a) if a.ndim > 0 else a.item()
    return out


def langevin_noise(g):
    """Noise of a Langevin golden: stored, or re-drawn from the recorded CPU seed."""
    if "noise" in g:
        return g["noise"]
    gen = torch.Generator().manual_seed(int(g["noise_seed"]))
    noise = torch.stack([torch.randn_like(g["x0"], generator=gen) for _ in range(int(g["k"]))])
    assert abs(noise.double().sum().item() - g["noise_checksum"]) < 1e-6, "CPU generator stream changed"
    return noise


def mlp_from(g, activation):
    ws, bs = [], []
    i = 0
    while f"w{i}" in g:
        ws.append(g[f"w{i}"])
        bs.append(g[f"b{i}"])
        i += 1
    return E.MLP(ws, bs, activation)


def energy_for(name, g):
    """The oracle energy matching a golden file name."""
    if "doublewell" in name or name in ("langevin_single_chain", "langevin_scheduled", "hmc_mass_float", "hmc_far_start"):
        return E.DoubleWell(g.get("barrier_height", 2.0), g.get("b", 1.0))
    if "harmonic" in name or name == "hmc_mass_vec":
        return E.Harmonic(g["kspring"])
    if "rastrigin" in name:
        return E.Rastrigin(g["a"])
    if "gaussian" in name:
        return E.Gaussian(g["mean"], g["cov"])
    if "mog" in name:
        return E.MixtureOfGaussians(g["means"], g["sigmas"], g["weights"])
    if "mlp_tanh" in name or "mlp_deep" in name:
        return mlp_from(g, "tanh")
    if "mlp" in name:
        return mlp_from(g, "silu")
    raise KeyError(name)


def mass_of(g):
    m = g["mass"]
    if torch.is_tensor(m):
        return m
    return None if m < 0 else float(m)


LANGEVIN_CASES = [
    "langevin_doublewell", "langevin_doublewell_odd", "langevin_doublewell_k100", "langevin_harmonic",
    "langevin_rastrigin", "langevin_gaussian_c1", "langevin_gaussian_d16", "langevin_single_chain",
    "langevin_scheduled", "langevin_mlp_silu", "langevin_mlp_tanh", "langevin_mlp_d128", "langevin_mog",
    "langevin_mlp_d784", "langevin_mlp_deep",
]
# cases whose closed-form gradient is bit-identical to the reference's autograd on CPU
LANGEVIN_BITEXACT_CLOSED = {
    "langevin_doublewell", "langevin_doublewell_odd", "langevin_doublewell_k100", "langevin_harmonic",
    "langevin_rastrigin", "langevin_single_chain", "langevin_scheduled",
}
HMC_CASES = ["hmc_doublewell", "hmc_rastrigin", "hmc_rastrigin_diag", "hmc_gaussian", "hmc_mass_float",
             "hmc_mass_vec", "hmc_far_start"]
LEAPFROG_CASES = ["leapfrog_nomass", "leapfrog_safe", "leapfrog_mass_float", "leapfrog_mass_vec", "leapfrog_extreme"]


def langevin_kwargs(name, g):
    kw = {}
    if name == "langevin_doublewell_odd":
        kw.update(clamp=(float(g["clamp"][0]), float(g["clamp"][1])), thin=3, return_trajectory=True,
                  return_diagnostics=True)
    if name == "langevin_single_chain":
        kw.update(return_diagnostics=True)
    return kw


def langevin_schedule(name, g):
    if name == "langevin_scheduled":
        return [float(v) for v in g["h_values"]], [float(v) for v in g["ns_values"]]
    return float(g["h"]), float(g["ns"])

# noise-free descent samplers: (golden, momentum or None)
DESCENT_CASES = ["gd_doublewell", "gd_rastrigin_traj", "nesterov_doublewell_sched", "nesterov_harmonic_traj", "gd_mlp"]


def descent_setup(name, g):
    """(oracle energy, step sizes, momentum, kwargs) of a descent golden."""
    if name == "gd_mlp":
        en = mlp_from(load("langevin_mlp_d784"), "silu")
    elif "doublewell" in name:
        en = E.DoubleWell(2.0, 1.0)
    elif "rastrigin" in name:
        en = E.Rastrigin(g["a"])
    else:
        en = E.Harmonic(g["kspring"])
    hs = [float(v) for v in g["h_values"]] if "h_values" in g else {"gd_doublewell": 0.01, "gd_rastrigin_traj": 0.001,
                                                                     "nesterov_harmonic_traj": 0.05, "gd_mlp": 0.05}[name]
    mu = float(g["momentum"]) if "momentum" in g else None
    kw = {}
    if name == "gd_rastrigin_traj":
        kw = dict(thin=3, return_trajectory=True, return_diagnostics=True)
    if name == "nesterov_harmonic_traj":
        kw = dict(thin=2, return_trajectory=True, return_diagnostics=True)
    return en, hs, mu, kw

# Heun SDE integrator behind LangevinDynamics: (golden, oracle energy, sample kwargs)
HEUN_CASES = ["heun_doublewell", "heun_rastrigin_traj"]


def heun_setup(name):
    if name == "heun_doublewell":
        return E.DoubleWell(2.0, 1.0), {}
    return E.Rastrigin(10.0), dict(thin=3, return_trajectory=True)
