"""The drop-in layer: this package's fused samplers, integrators and loss built ON TOP of the reference's own classes.

Importable only when the reference package (`torchebm`) is (see `_ref.py`).  Every class here subclasses the
reference class it accelerates and keeps its constructor untouched, so it satisfies the reference's own contract tests
(tests/samplers/test_api_contract.py: `sample()` prefix, ctor order with `integrator` last, `BaseSampler` subclass) and
plugs into the reference's unmodified `ContrastiveDivergence`, trainers and schedulers:

    torchebm.samplers.LangevinDynamics          -> dropin.LangevinDynamics          (FusedLangevinMixin.sample)
    torchebm.samplers.HamiltonianMonteCarlo     -> dropin.HamiltonianMonteCarlo     (FusedHMCMixin.sample)
    torchebm.samplers.GradientDescentSampler    -> dropin.GradientDescentSampler    (FusedDescentMixin.sample)
    torchebm.samplers.NesterovSampler           -> dropin.NesterovSampler
    torchebm.integrators.EulerMaruyamaIntegrator -> dropin.EulerMaruyamaIntegrator  (fused update kernel in `step`)
    torchebm.integrators.LeapfrogIntegrator     -> dropin.LeapfrogIntegrator        (fused `integrate` for tagged drifts)
    torchebm.losses.ContrastiveDivergence       -> dropin.ContrastiveDivergence     (device-side replay buffer, one-call negatives)

What the fused path does not cover -- CPU devices, fp16 / fp64 states, `model_kwargs` conditioning, energies the
library has no kernel for, other integrators -- goes to the reference's own method through `super()`: inside a reference
install nothing that worked before stops working.  `install()` rebinds the names inside the `torchebm` package so that
existing `from torchebm.samplers import LangevinDynamics` code picks the fused classes up; `uninstall()` restores them.
"""

from __future__ import annotations

from typing import Dict, Optional

import torch

from . import ops
from ._ref import reference
from .integrators import EnergyDrift, register_known_integrators
from .losses import FusedPCDMixin
from .samplers import FusedDescentMixin, FusedHMCMixin, FusedLangevinMixin

_ref = reference()
if _ref is None:   # pragma: no cover
    raise ImportError("torchebm_b200.dropin needs the reference package `torchebm` on sys.path")

import torchebm.core as _RC  # noqa: E402
import torchebm.integrators as _RI  # noqa: E402
import torchebm.losses as _RL  # noqa: E402
import torchebm.samplers as _RS  # noqa: E402

from .core import autograd_gradient, energy_descriptor  # noqa: E402

_ORIGINAL = {
    ("samplers", "LangevinDynamics"): _RS.LangevinDynamics,
    ("samplers", "HamiltonianMonteCarlo"): _RS.HamiltonianMonteCarlo,
    ("samplers", "GradientDescentSampler"): _RS.GradientDescentSampler,
    ("samplers", "NesterovSampler"): _RS.NesterovSampler,
    ("integrators", "EulerMaruyamaIntegrator"): _RI.EulerMaruyamaIntegrator,
    ("integrators", "LeapfrogIntegrator"): _RI.LeapfrogIntegrator,
    ("losses", "ContrastiveDivergence"): _RL.ContrastiveDivergence,
}
_RefLangevin, _RefHMC, _RefGD, _RefNesterov = (_ORIGINAL[("samplers", n)] for n in
                                               ("LangevinDynamics", "HamiltonianMonteCarlo", "GradientDescentSampler",
                                                "NesterovSampler"))
_RefEM, _RefLeapfrog = _ORIGINAL[("integrators", "EulerMaruyamaIntegrator")], _ORIGINAL[("integrators", "LeapfrogIntegrator")]
_RefCD = _ORIGINAL[("losses", "ContrastiveDivergence")]


# ---- integrators -------------------------------------------------------------------------------------------------

class EulerMaruyamaIntegrator(_RefEM):
    """core/base_integrator.py:673-731 with the update arithmetic in one library kernel (`ebm_euler_maruyama_step_f32`)
    when the state is fp32 on CUDA and the coefficients are Python scalars; the reference's own step otherwise."""

    def step(self, state: Dict[str, torch.Tensor], step_size, *, drift=None, diffusion=None, noise=None,
             noise_scale=None, t=None, generator=None) -> Dict[str, torch.Tensor]:
        x = state["x"]
        if not (x.is_cuda and x.dtype == torch.float32 and diffusion is None and not torch.is_tensor(step_size)
                and not torch.is_tensor(noise_scale)):
            return super().step(state, step_size, drift=drift, diffusion=diffusion, noise=noise, noise_scale=noise_scale,
                                t=t, generator=generator)
        if t is None:
            t = torch.zeros(x.size(0), device=x.device, dtype=x.dtype)
        d = self._resolve_drift(drift)(x, t)
        if noise_scale is not None and noise is None:
            noise = torch.randn_like(x, generator=generator)
        if d.shape != x.shape or d.dtype != x.dtype or (noise is not None and (noise.shape != x.shape or noise.dtype != x.dtype)):
            x_new = x + step_size * d     # a drift / noise that only broadcasts against x: the reference's arithmetic
            if noise_scale is not None:
                x_new = x_new + (2.0 * noise_scale ** 2) ** 0.5 * (noise * (step_size ** 0.5))
            return {"x": x_new}
        return {"x": ops.euler_maruyama_step(x, d, noise, float(step_size), None if noise_scale is None else float(noise_scale))}


class LeapfrogIntegrator(_RefLeapfrog):
    """integrators/leapfrog.py:116-187; a drift built with `torchebm_b200.energy_drift(model)` carries the model, and
    all steps of `integrate` then run in one fused kernel."""

    def integrate(self, state, step_size=None, n_steps=None, mass=None, *, drift=None, safe: bool = False,
                  inference_mode: bool = False):
        x = state["x"]
        if (isinstance(drift, EnergyDrift) and n_steps is not None and n_steps > 0 and not inference_mode and x.is_cuda
                and x.dtype == torch.float32 and x.ndim == 2 and not torch.is_tensor(step_size)):
            desc = energy_descriptor(drift.model, x.shape[1], x.device)
            if desc is not None and desc.kind != "mlp":
                xo, po = ops.leapfrog(desc, x, state["p"], float(step_size), n_steps, mass=mass, safe=safe)
                return {"x": xo, "p": po}
        return super().integrate(state, step_size=step_size, n_steps=n_steps, mass=mass, drift=drift, safe=safe,
                                 inference_mode=inference_mode)


# the fused bursts reproduce the arithmetic of these integrator classes (exact types only)
register_known_integrators(em=(_RefEM, EulerMaruyamaIntegrator), heun=(_RI.HeunIntegrator,),
                           leapfrog=(_RefLeapfrog, LeapfrogIntegrator))


# ---- samplers ----------------------------------------------------------------------------------------------------

class _ReferenceFallback:
    """`_sample_unfused` = the reference's own `sample()` (the next `sample` in the MRO after the fused mixin)."""

    _fused_mixin: type = object

    def _model_gradient(self, x, model_kwargs):
        # core/base_sampler.py:79-92 needs `model.gradient`; a plain nn.Module energy gets the same autograd gradient
        # (core/base_model.py:84-127) instead of an AttributeError
        if not hasattr(self.model, "gradient"):
            return autograd_gradient(self.model, x, model_kwargs)
        return super()._model_gradient(x, model_kwargs)

    def _sample_unfused(self, x, dim, n_steps, n_samples, thin, return_trajectory, return_diagnostics, reset_schedulers,
                        model_kwargs, generator):
        return super(self._fused_mixin, self).sample(
            x=x, dim=dim, n_steps=n_steps, n_samples=n_samples, thin=thin, return_trajectory=return_trajectory,
            return_diagnostics=return_diagnostics, reset_schedulers=reset_schedulers, model_kwargs=model_kwargs,
            generator=generator)


class LangevinDynamics(_ReferenceFallback, FusedLangevinMixin, _RefLangevin):
    """`torchebm.samplers.LangevinDynamics` (langevin_dynamics.py:16-188) with the K-step loop fused into one kernel."""

    _fused_mixin = FusedLangevinMixin


class HamiltonianMonteCarlo(_ReferenceFallback, FusedHMCMixin, _RefHMC):
    """`torchebm.samplers.HamiltonianMonteCarlo` (hmc.py:19-315) with all proposals of a call in one kernel."""

    _fused_mixin = FusedHMCMixin


class GradientDescentSampler(_ReferenceFallback, FusedDescentMixin, _RefGD):
    """`torchebm.samplers.GradientDescentSampler` (gradient_descent.py:16-140)."""

    _fused_mixin = FusedDescentMixin


class NesterovSampler(_ReferenceFallback, FusedDescentMixin, _RefNesterov):
    """`torchebm.samplers.NesterovSampler` (gradient_descent.py:143-276)."""

    _fused_mixin = FusedDescentMixin


# ---- loss --------------------------------------------------------------------------------------------------------

class ContrastiveDivergence(FusedPCDMixin, _RefCD):
    """`torchebm.losses.ContrastiveDivergence` (contrastive_divergence.py:13-223) with the replay-buffer gather /
    exploration noise / FIFO write-back as library kernels on the registered buffer and the negatives drawn in one
    library call when the sampler offers it.  The loss value (`compute_loss`) is the reference's own autograd code."""

    def forward(self, x: torch.Tensor, *args, model_kwargs: Optional[dict] = None,
                generator: Optional[torch.Generator] = None, **kwargs):
        buf = self.replay_buffer
        on_cuda = torch.device(self.device).type == "cuda" and self.dtype == torch.float32 and (buf is None or buf.is_cuda)
        if not on_cuda:
            return super().forward(x, *args, model_kwargs=model_kwargs, generator=generator, **kwargs)
        model_kwargs = self._prepare_model_kwargs(model_kwargs)
        pred_samples = self.sample_negatives(x, model_kwargs=model_kwargs, generator=generator)
        kwargs.setdefault("energy_reg_weight", self.energy_reg_weight)
        kwargs.setdefault("add_noise_to_real", self.add_noise_to_real)
        kwargs.setdefault("noise_scale", self.noise_scale)
        loss = self.compute_loss(x, pred_samples, *args, model_kwargs=model_kwargs, generator=generator, **kwargs)
        return loss, pred_samples

    def _get_start_points_unfused(self, x, generator):
        return _RefCD.get_start_points(self, x, generator=generator)

    def _update_buffer_unfused(self, samples):
        return _RefCD.update_buffer(self, samples)


_FUSED = {
    ("samplers", "LangevinDynamics"): LangevinDynamics,
    ("samplers", "HamiltonianMonteCarlo"): HamiltonianMonteCarlo,
    ("samplers", "GradientDescentSampler"): GradientDescentSampler,
    ("samplers", "NesterovSampler"): NesterovSampler,
    ("integrators", "EulerMaruyamaIntegrator"): EulerMaruyamaIntegrator,
    ("integrators", "LeapfrogIntegrator"): LeapfrogIntegrator,
    ("losses", "ContrastiveDivergence"): ContrastiveDivergence,
}
_MODULES = {"samplers": _RS, "integrators": _RI, "losses": _RL}


def install() -> None:
    """Rebind the accelerated names inside the `torchebm` package (and at its top level when it re-exports them)."""
    import torchebm

    for (mod, name), cls in _FUSED.items():
        setattr(_MODULES[mod], name, cls)
        if getattr(torchebm, name, None) is _ORIGINAL[(mod, name)]:
            setattr(torchebm, name, cls)


def uninstall() -> None:
    import torchebm

    for (mod, name), cls in _ORIGINAL.items():
        setattr(_MODULES[mod], name, cls)
        if getattr(torchebm, name, None) is _FUSED[(mod, name)]:
            setattr(torchebm, name, cls)
