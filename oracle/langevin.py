"""Oracle for `LangevinDynamics.sample`.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows torchebm/samplers/langevin_dynamics.py:125-188 (the K-step loop) and the
Euler-Maruyama step it drives, torchebm/core/base_integrator.py:673-731 with the
1-stage tableau a=(()), b=(1.0,), c=(0.0,) (torchebm/integrators/euler_maruyama.py:55-65):

    d      = -gradient(x)                          langevin_dynamics.py:154
    x1     = x + h * (1.0 * d)                     base_integrator.py:387-397
    eps    = randn_like(x)   (drawn even if ns==0) base_integrator.py:721-725
    dw     = eps * h**0.5                          base_integrator.py:728
    x'     = x1 + (2.0 * ns**2)**0.5 * dw          base_integrator.py:729
    clamp_ (optional)                              langevin_dynamics.py:166-167

`h`, `h**0.5` and `(2 ns^2)**0.5` are Python doubles that torch rounds to fp32 when
they meet the fp32 tensor.
"""

from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple, Union

import torch

from .energies import Energy


def _as_schedule(v: Union[float, Sequence[float]], n: int):
    if isinstance(v, (int, float)):
        return [float(v)] * n
    v = list(v)
    assert len(v) == n, (len(v), n)
    return [float(t) for t in v]


def em_step(x, grad, h: float, ns: float, eps):
    """One reference Euler-Maruyama step given the gradient and the noise."""
    d = -grad
    x1 = x + h * (1.0 * d)
    dw = eps * (h**0.5)
    return x1 + (2.0 * (ns**2)) ** 0.5 * dw


def heun_step(x, grad_fn, h: float, ns: float, eps):
    """One reference Heun (improved Euler) SDE step: `LangevinDynamics(integrator="heun")`, i.e. the generic RK path of
    core/base_integrator.py:300-347,387-397,673-731 with the tableau of integrators/heun.py (a = ((), (1,)),
    b = (1/2, 1/2)): k1 = f(x); k2 = f(x + h * (1 * k1)); x1 = x + h * (k1/2 + k2/2); then the additive noise of the EM
    step.  The einsum scalings by 1 and 1/2 are exact, so only the sums and products below round."""
    k1 = -grad_fn(x)
    k2 = -grad_fn(x + h * k1)
    x1 = x + h * (0.5 * k1 + 0.5 * k2)
    dw = eps * (h**0.5)
    return x1 + (2.0 * (ns**2)) ** 0.5 * dw


@torch.no_grad()
def sample(
    energy: Energy,
    x: torch.Tensor,
    n_steps: int,
    step_size: Union[float, Sequence[float]],
    noise_scale: Union[float, Sequence[float]] = 1.0,
    *,
    clamp: Optional[Tuple[float, float]] = None,
    thin: int = 1,
    return_trajectory: bool = False,
    return_diagnostics: bool = False,
    noise: Optional[torch.Tensor] = None,
    generator: Optional[torch.Generator] = None,
    closed_form: bool = False,
    scheme: str = "euler_maruyama",
):
    """`noise` is `[n_steps, *x.shape]` (injected) or None (draw `randn_like` per step with
    `generator`, the reference's own draw order: SURVEY.md section 8c)."""
    if thin < 1:
        raise ValueError("thin must be >= 1")
    hs = _as_schedule(step_size, n_steps)
    sigmas = _as_schedule(noise_scale, n_steps)
    n = x.shape[0]
    n_kept = n_steps // thin
    traj = torch.empty((n, n_kept, *x.shape[1:]), dtype=x.dtype, device=x.device) if return_trajectory else None
    diag: Optional[Dict[str, torch.Tensor]] = None
    if return_diagnostics:
        diag = {
            "mean": torch.empty(n_kept, *x.shape[1:], dtype=x.dtype, device=x.device),
            "var": torch.empty(n_kept, *x.shape[1:], dtype=x.dtype, device=x.device),
            "energy": torch.empty(n_kept, dtype=x.dtype, device=x.device),
        }
    keep = 0
    grad_fn = energy.gradient_closed if closed_form else energy.gradient
    for i in range(n_steps):
        g = grad_fn(x) if scheme == "euler_maruyama" else None
        eps = noise[i] if noise is not None else torch.randn_like(x, generator=generator)
        x = em_step(x, g, hs[i], sigmas[i], eps) if scheme == "euler_maruyama" else heun_step(x, grad_fn, hs[i], sigmas[i], eps)
        if clamp is not None:
            x = x.clamp_(*clamp)
        if (i + 1) % thin == 0:
            if traj is not None:
                traj[:, keep] = x
            if diag is not None:
                if n > 1:
                    diag["mean"][keep] = x.mean(dim=0)
                    diag["var"][keep] = x.var(dim=0, unbiased=False).clamp_(min=1e-10, max=1e10)
                else:
                    diag["mean"][keep] = x.squeeze(0)
                    diag["var"][keep].zero_()
                diag["energy"][keep] = energy.energy(x).mean()
            keep += 1
    out = traj if return_trajectory else x
    return (out, diag) if return_diagnostics else out
