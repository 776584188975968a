"""Oracle for `HamiltonianMonteCarlo.sample`.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows torchebm/samplers/hmc.py:244-312 (proposal loop), :92-134 (momentum draw),
:136-159 (kinetic energy) and drives oracle.leapfrog.integrate with safe=True
(hmc.py:258-265).  Draw order per proposal: `normal_` for the
momentum, then `torch.rand(N)` for the Metropolis test.
"""

from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Union

import torch

from .energies import Energy
from . import leapfrog


def kinetic(p, mass):
    # hmc.py:148-159
    if mass is None:
        return 0.5 * torch.sum(p.square(), dim=-1)
    if isinstance(mass, float):
        return 0.5 * torch.sum(p.square(), dim=-1) / mass
    return 0.5 * torch.sum(p.square() / mass.view((1,) * (p.ndim - 1) + (-1,)), dim=-1)


@torch.no_grad()
def sample(
    energy: Energy,
    x: torch.Tensor,
    n_steps: int,
    step_size: Union[float, Sequence[float]],
    n_leapfrog_steps: int,
    mass: Optional[Union[float, torch.Tensor]] = None,
    *,
    thin: int = 1,
    return_trajectory: bool = False,
    return_diagnostics: bool = False,
    noise_p: Optional[torch.Tensor] = None,  # [n_steps, N, D] standard normals
    noise_u: Optional[torch.Tensor] = None,  # [n_steps, N] uniforms in [0, 1)
    generator: Optional[torch.Generator] = None,
    closed_form: bool = False,
):
    if thin < 1:
        raise ValueError("thin must be >= 1")
    hs = [float(step_size)] * n_steps if isinstance(step_size, (int, float)) else [float(h) for h in step_size]
    n, d = x.shape
    n_kept = n_steps // thin
    traj = torch.empty((n, n_kept, d), dtype=x.dtype, device=x.device) if return_trajectory else None
    diag: Optional[Dict[str, torch.Tensor]] = None
    if return_diagnostics:
        diag = {
            "mean": torch.empty(n_kept, d, dtype=x.dtype, device=x.device),
            "var": torch.empty(n_kept, d, dtype=x.dtype, device=x.device),
            "energy": torch.empty(n_kept, dtype=x.dtype, device=x.device),
            "acceptance_rate": torch.empty(n_kept, dtype=x.dtype, device=x.device),
        }
    grad_fn = energy.gradient_closed if closed_form else energy.gradient
    drift = lambda x_: -grad_fn(x_)
    keep = 0
    for i in range(n_steps):
        # hmc.py:245 / :92-134
        if noise_p is not None:
            p = noise_p[i].clone()
        else:
            p = torch.empty_like(x).normal_(generator=generator)
        if mass is not None:
            if isinstance(mass, float):
                p.mul_(math.sqrt(mass))
            else:
                p.mul_(torch.sqrt(mass).view(1, -1))
        e0 = energy.energy(x).clamp_(min=-1e10, max=1e10)
        k0 = kinetic(p, mass).clamp_(min=0.0, max=1e10)
        h0 = e0 + k0
        xp, pp = leapfrog.integrate(drift, x, p, hs[i], n_leapfrog_steps, mass=mass, safe=True)
        e1 = energy.energy(xp).clamp_(min=-1e10, max=1e10)
        k1 = kinetic(pp, mass).clamp_(min=0.0, max=1e10)
        h1 = e1 + k1
        dh = (h0 - h1).clamp_(min=-50.0, max=50.0)
        acc_prob = torch.exp(dh).clamp_(max=1.0)
        if noise_u is not None:
            u = noise_u[i]
        else:
            u = torch.rand(n, device=x.device, generator=generator)
        accepted = u < acc_prob
        x = torch.where(accepted.view(-1, 1), xp, x)
        if (i + 1) % thin == 0:
            if traj is not None:
                traj[:, keep, :] = x
            if diag is not None:
                diag["mean"][keep] = x.mean(dim=0)
                diag["var"][keep] = (
                    x.var(dim=0, unbiased=False).clamp_(min=1e-10, max=1e10)
                    if n > 1
                    else torch.zeros(d, dtype=x.dtype, device=x.device)
                )
                diag["energy"][keep] = energy.energy(x).clamp_(min=-1e10, max=1e10).mean()
                diag["acceptance_rate"][keep] = accepted.float().mean()
            keep += 1
    out = traj if return_trajectory else x
    return (out, diag) if return_diagnostics else out
