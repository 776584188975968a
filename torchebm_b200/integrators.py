"""Integrator-level boundary: `EulerMaruyamaIntegrator` and `LeapfrogIntegrator` with the reference's
call signatures (torchebm/core/base_integrator.py:673-731, torchebm/integrators/leapfrog.py:63-187,
registry torchebm/integrators/integrator_utils.py:8-111).

At this level `drift` is an opaque Python callable, so only the update arithmetic can run in the
library: `EulerMaruyamaIntegrator.step` evaluates the caller's drift and then launches one fused
update kernel.  A drift built with `energy_drift(model)` is a tagged callable that carries an energy
descriptor; `LeapfrogIntegrator.integrate` recognises it and runs all steps in one fused kernel.
For untagged drifts the leapfrog loop follows the reference step by step with tensor ops on the
caller's device (compatibility path for the integrator KATs, which use fp64 lambdas).
"""

from __future__ import annotations

from typing import Callable, Dict, Optional, Type, Union

import torch

from . import ops
from .core import TorchEBMModule, energy_descriptor


class BaseIntegrator(TorchEBMModule):
    def __init__(self, device=None, dtype=None):
        super().__init__(device=device, dtype=dtype)

    @staticmethod
    def _resolve_drift(drift):
        if drift is None:
            raise ValueError("drift must be provided")
        return drift


class EnergyDrift:
    """drift(x, t) = -grad E(x) for a model the library recognises; carries the model so integrators can fuse."""

    def __init__(self, model):
        self.model = model

    def __call__(self, x: torch.Tensor, t: Optional[torch.Tensor] = None) -> torch.Tensor:
        return -self.model.gradient(x)


def energy_drift(model) -> EnergyDrift:
    return EnergyDrift(model)


class BaseSDERungeKuttaIntegrator(BaseIntegrator):
    pass


class EulerMaruyamaIntegrator(BaseSDERungeKuttaIntegrator):
    """x' = (x + h*drift(x,t)) + (2 D)^0.5 * (noise * h^0.5), D = noise_scale^2 (base_integrator.py:719-729)."""

    def step(self, state: Dict[str, torch.Tensor], step_size, *, drift=None, diffusion=None, noise=None,
             noise_scale=None, t=None, generator=None) -> Dict[str, torch.Tensor]:
        x = state["x"]
        if t is None:
            t = torch.zeros(x.size(0), device=x.device, dtype=x.dtype)
        d = self._resolve_drift(drift)(x, t)
        stochastic = diffusion is not None or noise_scale is not None
        if stochastic and noise is None:
            noise = torch.randn_like(x, generator=generator)
        fused = (x.is_cuda and x.dtype == torch.float32 and diffusion is None and not torch.is_tensor(step_size)
                 and not torch.is_tensor(noise_scale) and d.shape == x.shape and d.dtype == x.dtype
                 and (noise is None or (noise.shape == x.shape and noise.dtype == x.dtype)))   # broadcastable-only drifts: eager
        if fused:
            return {"x": ops.euler_maruyama_step(x, d, noise, float(step_size),
                                                 None if noise_scale is None else float(noise_scale))}
        x_new = x + step_size * d
        if stochastic:
            diffusion_val = diffusion if diffusion is not None else noise_scale**2
            x_new = x_new + (2.0 * diffusion_val) ** 0.5 * (noise * (step_size**0.5))
        return {"x": x_new}


class HeunIntegrator(BaseSDERungeKuttaIntegrator):
    """Heun / improved Euler (integrators/heun.py: a = ((), (1,)), b = (1/2, 1/2)) through the reference's generic
    Runge-Kutta step (core/base_integrator.py:300-347,387-397,673-731):
    k1 = f(x, t); k2 = f(x + h*k1, t + h); x' = (x + h*(k1/2 + k2/2)) + (2 D)^0.5 * (noise * h^0.5).
    `LangevinDynamics(integrator="heun")` runs it as one fused burst for the elementwise energies; this step-level form
    serves every other drift."""

    def step(self, state: Dict[str, torch.Tensor], step_size, *, drift=None, diffusion=None, noise=None,
             noise_scale=None, t=None, generator=None) -> Dict[str, torch.Tensor]:
        x = state["x"]
        if t is None:
            t = torch.zeros(x.shape[0], device=x.device, dtype=x.dtype)
        f = self._resolve_drift(drift)
        k1 = f(x, t)
        k2 = f(x + step_size * k1, t + 1.0 * step_size)
        x_new = x + step_size * (0.5 * k1 + 0.5 * k2)
        if diffusion is not None or noise_scale is not None:
            if noise is None:
                noise = torch.randn_like(x, generator=generator)
            diffusion_val = diffusion if diffusion is not None else noise_scale**2
            x_new = x_new + (2.0 * diffusion_val) ** 0.5 * (noise * (step_size**0.5))
        return {"x": x_new}


class BaseSymplecticIntegrator(BaseIntegrator):
    separable: bool = True
    _SAFE_CLAMP: float = 1e6


class LeapfrogIntegrator(BaseSymplecticIntegrator):
    separable = True

    def step(self, state, step_size=None, mass=None, *, drift=None, safe: bool = False):
        return self.integrate(state, step_size=step_size, n_steps=1, mass=mass, drift=drift, safe=safe)

    def integrate(self, state: Dict[str, torch.Tensor], step_size=None, n_steps: int = None, mass=None, *,
                  drift=None, safe: bool = False, inference_mode: bool = False) -> Dict[str, torch.Tensor]:
        if n_steps is None or n_steps <= 0:
            raise ValueError("n_steps must be positive")
        if inference_mode:
            with torch.inference_mode():
                return self.integrate(state, step_size=step_size, n_steps=n_steps, mass=mass, drift=drift, safe=safe)
        drift_fn = self._resolve_drift(drift)
        x, p = state["x"], state["p"]
        if (isinstance(drift_fn, EnergyDrift) and x.is_cuda and x.dtype == torch.float32 and x.ndim == 2
                and not torch.is_tensor(step_size)):
            desc = energy_descriptor(drift_fn.model, x.shape[1], x.device)
            if desc is not None and desc.kind != "mlp":
                xo, po = ops.leapfrog(desc, x, p, float(step_size), n_steps, mass=mass, safe=safe)
                return {"x": xo, "p": po}
        # opaque drift: reference order of operations (leapfrog.py:160-185)
        h = step_size if torch.is_tensor(step_size) else torch.tensor(step_size, device=x.device, dtype=x.dtype)
        t = torch.zeros(x.size(0), device=x.device, dtype=x.dtype)
        for _ in range(n_steps):
            force = drift_fn(x, t)
            if safe:
                force = force.clamp(min=-self._SAFE_CLAMP, max=self._SAFE_CLAMP)
            p_half = p + 0.5 * h * force
            if mass is None:
                x = x + h * p_half
            elif isinstance(mass, float):
                x = x + h * p_half / max(mass, 1e-10)
            else:
                x = x + h * p_half / torch.clamp(mass, min=1e-10).view((1,) * (x.ndim - 1) + (-1,))
            force_new = drift_fn(x, t)
            if safe:
                force_new = force_new.clamp(min=-self._SAFE_CLAMP, max=self._SAFE_CLAMP)
            p = p_half + 0.5 * h * force_new
            if safe:
                x = x.nan_to_num(nan=0.0)
                p = p.nan_to_num(nan=0.0)
        return {"x": x, "p": p}


_REGISTRY = {"euler_maruyama": EulerMaruyamaIntegrator, "heun": HeunIntegrator, "leapfrog": LeapfrogIntegrator}

# Integrator classes whose arithmetic the fused bursts reproduce: this package's and, when importable, the reference's
# own (exact types only -- a subclass may override `step`).  dropin.py registers its reference-derived classes here.
_EM_TYPES = {EulerMaruyamaIntegrator}
_HEUN_TYPES = {HeunIntegrator}
_LEAPFROG_TYPES = {LeapfrogIntegrator}


def register_known_integrators(em=(), heun=(), leapfrog=()) -> None:
    _EM_TYPES.update(em)
    _HEUN_TYPES.update(heun)
    _LEAPFROG_TYPES.update(leapfrog)


def sde_scheme_of(integrator) -> Optional[str]:
    """"euler_maruyama" / "heun" when `integrator` is one whose step the fused Langevin burst implements, else None."""
    t = type(integrator)
    if t in _EM_TYPES:
        return "euler_maruyama"
    if t in _HEUN_TYPES:
        return "heun"
    return None


def is_plain_leapfrog(integrator) -> bool:
    return type(integrator) in _LEAPFROG_TYPES


def get_integrator(name: str, device=None, dtype=None) -> BaseIntegrator:
    if name not in _REGISTRY:
        raise ValueError(f"Unknown integrator '{name}'. Available: {sorted(_REGISTRY)}")
    return _REGISTRY[name](device=device, dtype=dtype)


def resolve_integrator(integrator, *, default: str, family: Type[BaseIntegrator], owner: str, device=None, dtype=None):
    """integrator_utils.py:55-111: None / name -> construct; instance -> validate family and device/dtype."""
    if isinstance(integrator, BaseIntegrator):
        if not isinstance(integrator, family):
            raise TypeError(f"{owner} requires a {family.__name__}; got {type(integrator).__name__}")
        if integrator.device != device or integrator.dtype != dtype:
            raise ValueError(
                f"{owner} device/dtype ({device}, {dtype}) does not match the integrator's "
                f"({integrator.device}, {integrator.dtype}). Construct the integrator with matching device/dtype; "
                f"no implicit .to() is performed.")
        return integrator
    if integrator is not None and not isinstance(integrator, str):
        raise TypeError(f"{owner}: integrator must be None, a registry name or a {family.__name__} instance")
    resolved = get_integrator(integrator if integrator is not None else default, device=device, dtype=dtype)
    if not isinstance(resolved, family):
        raise TypeError(f"{owner} requires a {family.__name__}; got {type(resolved).__name__}")
    return resolved
