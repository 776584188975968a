"""Oracle for `LeapfrogIntegrator.integrate`.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows torchebm/integrators/leapfrog.py:157-187 and the safe-mode helpers in
torchebm/core/base_integrator.py:846-889 (`_unpack_state`, `_safe_clamp_` at +-1e6,
`_sanitize_state_` = nan_to_num_(nan=0.0), which also maps +-inf to +-FLT_MAX).
The force is evaluated twice per step exactly like the reference; the CUDA path
reuses it (bit-equal unless a NaN was sanitised).
"""

from __future__ import annotations

from typing import Callable, Optional, Union

import torch

SAFE_CLAMP = 1e6  # base_integrator.py:847


def integrate(
    drift: Callable[[torch.Tensor], torch.Tensor],
    x: torch.Tensor,
    p: torch.Tensor,
    step_size: float,
    n_steps: int,
    mass: Optional[Union[float, torch.Tensor]] = None,
    safe: bool = False,
):
    if n_steps <= 0:
        raise ValueError("n_steps must be positive")
    # base_integrator.py:870-871: a non-tensor step size becomes a 0-d tensor of x's dtype
    h = torch.tensor(step_size, device=x.device, dtype=x.dtype)
    for _ in range(n_steps):
        force = drift(x)
        if safe:
            force = force.clamp_(min=-SAFE_CLAMP, max=SAFE_CLAMP)
        p_half = p + 0.5 * h * force
        if mass is None:
            x = x + h * p_half
        elif isinstance(mass, float):
            x = x + h * p_half / max(mass, 1e-10)
        else:
            safe_mass = torch.clamp(mass, min=1e-10)
            x = x + h * p_half / safe_mass.view((1,) * (x.ndim - 1) + (-1,))
        force_new = drift(x)
        if safe:
            force_new = force_new.clamp_(min=-SAFE_CLAMP, max=SAFE_CLAMP)
        p = p_half + 0.5 * h * force_new
        if safe:
            x.nan_to_num_(nan=0.0)
            p.nan_to_num_(nan=0.0)
    return x, p
