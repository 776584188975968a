"""Oracle for `GradientDescentSampler.sample` and `NesterovSampler.sample`.  TEST INFRASTRUCTURE ONLY (see
oracle/__init__.py).

Follows torchebm/samplers/gradient_descent.py:123-138 (x <- torch.sub(x, grad, alpha=eta)) and :238-276
(lookahead = torch.add(x, v, alpha=mu); v.mul_(mu).sub_(grad, alpha=eta); x = x + v), with the same torch ops so that
the fused-multiply-add rounding of ATen's `alpha` kernels is reproduced.
"""

from __future__ import annotations

from typing import Dict, Optional, Sequence, Union

import torch

from .energies import Energy


@torch.no_grad()
def sample(energy: Energy, x: torch.Tensor, n_steps: int, step_size: Union[float, Sequence[float]], momentum: Optional[float] = None,
           *, thin: int = 1, return_trajectory: bool = False, return_diagnostics: bool = False, closed_form: bool = False):
    if thin < 1:
        raise ValueError("thin must be >= 1")
    hs = [float(step_size)] * n_steps if isinstance(step_size, (int, float)) else [float(h) for h in step_size]
    grad_fn = energy.gradient_closed if closed_form else energy.gradient
    n_kept = n_steps // thin
    traj = torch.empty(x.shape[0], n_kept, *x.shape[1:], dtype=x.dtype, device=x.device) if return_trajectory else None
    diag: Optional[Dict[str, torch.Tensor]] = {"energy": torch.empty(n_kept, dtype=x.dtype, device=x.device)} if return_diagnostics else None
    v = torch.zeros_like(x) if momentum is not None else None
    keep = 0
    for i in range(n_steps):
        eta = hs[i]
        if momentum is None:
            x = torch.sub(x, grad_fn(x), alpha=eta)
        else:
            lookahead = torch.add(x, v, alpha=momentum)
            v.mul_(momentum).sub_(grad_fn(lookahead), alpha=eta)
            x = x + v
        if (i + 1) % thin == 0:
            if traj is not None:
                traj[:, keep] = x
            if diag is not None:
                diag["energy"][keep] = energy.energy(x).mean()
            keep += 1
    out = traj if return_trajectory else x
    return (out, diag) if return_diagnostics else out
