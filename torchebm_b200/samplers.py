"""`LangevinDynamics` and `HamiltonianMonteCarlo` with the reference's API, running fused CUDA bursts.

Mirrors torchebm/core/base_sampler.py:10-155, torchebm/samplers/langevin_dynamics.py:16-188 and
torchebm/samplers/hmc.py:19-315: same constructor and `sample()` signatures, same return shapes,
same errors, same scheduler semantics, same RNG consumption (with `rng="torch"`, the default, a burst
draws exactly the Philox stream the reference's per-step `randn_like` / `normal_` / `rand` calls would,
so equal seeds give equal chains).  What differs is where the loop runs: one kernel launch does all
`n_steps` steps (`thin` steps per launch when diagnostics are requested).

Energies the library recognises (core.energy_descriptor) take the fused path.  Any other `nn.Module`
energy keeps its own PyTorch forward/autograd for the gradient -- it is the caller's code -- and only the
Euler-Maruyama update arithmetic runs in the library (the integrator-level boundary,
core/base_integrator.py:673-731).  CPU tensors, non-fp32 dtypes and a missing CUDA library raise.
"""

from __future__ import annotations

import warnings
from abc import ABC, abstractmethod
from typing import Any, Dict, Optional, Tuple, Union

import torch
from torch import nn

from . import _lib, ops
from .core import BaseScheduler, EnergyDescriptor, Schedulable, TorchEBMModule, autograd_gradient, energy_descriptor
from .integrators import (BaseSDERungeKuttaIntegrator, BaseSymplecticIntegrator, EulerMaruyamaIntegrator,
                          HeunIntegrator, LeapfrogIntegrator, resolve_integrator)


class BaseSampler(Schedulable, TorchEBMModule, ABC):
    """base_sampler.py:10-155."""

    def __init__(self, model: nn.Module, dtype: torch.dtype = torch.float32,
                 device: Optional[Union[str, torch.device]] = None):
        super().__init__(device=device, dtype=dtype)
        self.model = model

    def _init_state(self, x, dim, n_samples, generator=None) -> torch.Tensor:
        if x is not None:
            return x.to(device=self.device, dtype=self.dtype)
        if dim is None:
            raise ValueError("dim must be provided when x is None")
        shape = (dim,) if isinstance(dim, int) else tuple(dim)
        return torch.randn(n_samples, *shape, dtype=self.dtype, device=self.device, generator=generator)

    def _model_gradient(self, x, model_kwargs):
        if model_kwargs:
            return self.model.gradient(x, model_kwargs=model_kwargs)
        if hasattr(self.model, "gradient"):
            return self.model.gradient(x)
        return autograd_gradient(self.model, x)

    def _model_energy(self, x, model_kwargs):
        if model_kwargs:
            return self.model(x, **model_kwargs)
        return self.model(x)

    # ---- fused-path plumbing ------------------------------------------------------------------
    def _require_cuda_fp32(self) -> None:
        if self.device.type != "cuda":
            raise RuntimeError(
                f"{type(self).__name__} runs on a CUDA device only (got device={self.device}); "
                "torchebm_b200 has no CPU path")
        if self.dtype != torch.float32:
            raise TypeError(f"{type(self).__name__} supports dtype=torch.float32 only, got {self.dtype}")
        _lib.load()

    def _rng_state(self, generator: Optional[torch.Generator]) -> Tuple[torch.Generator, int, int]:
        """(generator, seed, offset) of the Philox stream this call draws from; no host sync."""
        if generator is None:
            idx = ops.device_index(self.device)
            generator = torch.cuda.default_generators[idx]
        elif generator.device.type != "cuda":
            raise RuntimeError(
                f"Expected a 'cuda' device type for generator but found '{generator.device.type}'")
        return generator, generator.initial_seed(), generator.get_offset()

    def _descriptor(self, x: torch.Tensor, model_kwargs: dict) -> Optional[EnergyDescriptor]:
        if model_kwargs or x.ndim != 2:
            return None
        return energy_descriptor(self.model, x.shape[1], x.device)

    @abstractmethod
    def sample(self, x=None, dim=None, n_steps=100, n_samples=1, thin=1, return_trajectory=False,
               return_diagnostics=False, reset_schedulers=True, *, generator=None): ...


def _batch_diag(x: torch.Tensor):
    if x.shape[0] > 1:
        return x.mean(dim=0), x.var(dim=0, unbiased=False).clamp_(min=1e-10, max=1e10)
    return x.squeeze(0), torch.zeros_like(x.squeeze(0))


class LangevinDynamics(BaseSampler):
    """langevin_dynamics.py:16-188.  Extra keyword `rng`: "torch" (default, reference-identical stream) or
    "native" (cheaper layout-native Philox stream)."""

    def __init__(self, model, step_size: Union[float, BaseScheduler] = 1e-3,
                 noise_scale: Union[float, BaseScheduler] = 1.0, decay: float = 0.0,
                 clamp: Optional[Tuple[float, float]] = None, dtype: torch.dtype = torch.float32,
                 device: Optional[Union[str, torch.device]] = None,
                 integrator: Union[str, BaseSDERungeKuttaIntegrator, None] = None, rng: str = "torch"):
        super().__init__(model=model, dtype=dtype, device=device)
        self._register_param("step_size", step_size, positive=True)
        self._register_param("noise_scale", noise_scale, positive=True)
        if clamp is not None and clamp[0] >= clamp[1]:
            raise ValueError(f"clamp min must be < max, got {clamp}")
        if rng not in ("torch", "native"):
            raise ValueError("rng must be 'torch' or 'native'")
        self.clamp = clamp
        self.decay = decay
        self.rng = rng
        self.integrator = resolve_integrator(integrator, default="euler_maruyama", family=BaseSDERungeKuttaIntegrator,
                                             owner="LangevinDynamics", device=self.device, dtype=self.dtype)

    @torch.no_grad()
    def sample(self, x: Optional[torch.Tensor] = None, dim: Optional[Union[int, Tuple[int, ...]]] = None,
               n_steps: int = 100, n_samples: int = 1, thin: int = 1, return_trajectory: bool = False,
               return_diagnostics: bool = False, reset_schedulers: bool = True, *,
               model_kwargs: Optional[Dict[str, Any]] = None, generator: Optional[torch.Generator] = None):
        if thin < 1:
            raise ValueError("thin must be >= 1")
        self._require_cuda_fp32()
        if reset_schedulers:
            self.reset_schedulers()
        gen, seed, offset = self._rng_state(generator)
        x = self._init_state(x, dim, n_samples, generator)
        offset = gen.get_offset()  # after _init_state, which may have drawn from the same generator
        model_kwargs = self._prepare_model_kwargs(model_kwargs)
        n = x.shape[0]
        data_shape = x.shape[1:]
        n_kept = n_steps // thin
        scheme = {EulerMaruyamaIntegrator: "euler_maruyama", HeunIntegrator: "heun"}.get(type(self.integrator))
        desc = self._descriptor(x, model_kwargs) if scheme is not None else None
        if desc is not None and scheme == "heun" and desc.kind not in ("double_well", "harmonic", "rastrigin"):
            desc = None   # the fused Heun burst exists for the elementwise energies; others step through the integrator
        if desc is None:
            return self._sample_opaque(x, n_steps, thin, return_trajectory, return_diagnostics, model_kwargs, generator)

        x = x.contiguous()
        rng_mode = _lib.RNG_MODES[self.rng]
        numel = x.numel()
        traj = torch.empty((n, n_kept, *data_shape), dtype=self.dtype, device=self.device) if return_trajectory else None
        if n_steps <= 0:
            out = traj if return_trajectory else x
            return (out, self._empty_diag(data_shape)) if return_diagnostics else out
        vals, constant = self._advance_schedules(("step_size", "noise_scale"), n_steps)
        hs, nss = vals["step_size"], vals["noise_scale"]

        if not return_diagnostics:
            out = ops.langevin_burst(desc, x, n_steps, hs, nss, clamp=self.clamp, rng_mode=rng_mode, seed=seed,
                                     offset=offset, traj=traj, thin=thin, scheme=scheme)
            gen.set_offset(offset + ops.rng_consumed_langevin(self.device, numel, n_steps, rng_mode))
            return traj if return_trajectory else out

        # diagnostics: one launch per kept sample, statistics from the device-resident state
        diag = self._empty_diag(data_shape, n_kept)
        cur = x
        done = 0
        for j in range(n_kept):
            h_j = hs if constant else hs[done:done + thin]
            ns_j = nss if constant else nss[done:done + thin]
            cur = ops.langevin_burst(desc, cur, thin, h_j, ns_j, clamp=self.clamp, rng_mode=rng_mode, seed=seed,
                                     offset=offset, scheme=scheme)
            offset += ops.rng_consumed_langevin(self.device, numel, thin, rng_mode)
            done += thin
            if traj is not None:
                traj[:, j] = cur
            diag["mean"][j], diag["var"][j] = _batch_diag(cur)
            diag["energy"][j] = ops.energy(desc, cur).mean()
        rest = n_steps - done
        if rest > 0:
            h_j = hs if constant else hs[done:]
            ns_j = nss if constant else nss[done:]
            cur = ops.langevin_burst(desc, cur, rest, h_j, ns_j, clamp=self.clamp, rng_mode=rng_mode, seed=seed,
                                     offset=offset, scheme=scheme)
            offset += ops.rng_consumed_langevin(self.device, numel, rest, rng_mode)
        gen.set_offset(offset)
        out = traj if return_trajectory else cur
        return out, diag

    @torch.no_grad()
    def sample_from_buffer(self, buffer: torch.Tensor, indices: torch.Tensor, ptr: int, n_steps: int, *,
                           reset_schedulers: bool = True, generator: Optional[torch.Generator] = None):
        """Persistent-CD negatives in one library call: start points `buffer[indices]`, `n_steps` Langevin steps,
        FIFO write-back into `buffer` at `ptr` (get_start_points + sample + update_buffer of
        core/base_loss.py:266-337,390-426 and losses/contrastive_divergence.py:127-139, without exploration noise).
        Same scheduler and generator semantics as `sample(x=buffer[indices], n_steps=n_steps, generator=generator)`.
        Returns `(negatives, new_ptr)`, or None when this sampler / energy has no library kernel for it (the caller
        then takes the three-call path)."""
        if type(self.integrator) is not EulerMaruyamaIntegrator or buffer.ndim != 2 or not buffer.is_cuda:
            return None   # (the one-call PCD path is Euler-Maruyama only)
        self._require_cuda_fp32()
        desc = energy_descriptor(self.model, buffer.shape[1], buffer.device)
        if desc is None or n_steps <= 0:
            return None
        if reset_schedulers:
            self.reset_schedulers()
        gen, seed, offset = self._rng_state(generator)
        rng_mode = _lib.RNG_MODES[self.rng]
        vals, _ = self._advance_schedules(("step_size", "noise_scale"), n_steps)
        out, new_ptr = ops.pcd_langevin_burst(desc, buffer, indices, ptr, n_steps, vals["step_size"], vals["noise_scale"],
                                              clamp=self.clamp, rng_mode=rng_mode, seed=seed, offset=offset)
        gen.set_offset(offset + ops.rng_consumed_langevin(self.device, out.numel(), n_steps, rng_mode))
        return out, new_ptr

    def _empty_diag(self, data_shape, n_kept: int = 0):
        return {
            "mean": torch.empty(n_kept, *data_shape, dtype=self.dtype, device=self.device),
            "var": torch.empty(n_kept, *data_shape, dtype=self.dtype, device=self.device),
            "energy": torch.empty(n_kept, dtype=self.dtype, device=self.device),
        }

    def _sample_opaque(self, x, n_steps, thin, return_trajectory, return_diagnostics, model_kwargs, generator):
        """Energies with no library kernel: their own gradient + the integrator's fused update per step
        (langevin_dynamics.py:157-185 verbatim in structure)."""
        n = x.shape[0]
        data_shape = x.shape[1:]
        n_kept = n_steps // thin
        traj = torch.empty((n, n_kept, *data_shape), dtype=self.dtype, device=self.device) if return_trajectory else None
        diag = self._empty_diag(data_shape, n_kept) if return_diagnostics else None
        drift = lambda x_, t_: -self._model_gradient(x_, model_kwargs)
        keep = 0
        for i in range(n_steps):
            x = self.integrator.step(state={"x": x}, step_size=self.get_scheduled_value("step_size"),
                                     noise_scale=self.get_scheduled_value("noise_scale"), drift=drift,
                                     generator=generator)["x"]
            if self.clamp is not None:
                x = x.clamp_(*self.clamp)
            self.step_schedulers()
            if (i + 1) % thin == 0:
                if traj is not None:
                    traj[:, keep] = x
                if diag is not None:
                    diag["mean"][keep], diag["var"][keep] = _batch_diag(x)
                    diag["energy"][keep] = self._model_energy(x, model_kwargs).mean()
                keep += 1
        out = traj if return_trajectory else x
        return (out, diag) if return_diagnostics else out


class HamiltonianMonteCarlo(BaseSampler):
    """hmc.py:19-315.  Extra keyword `rng` as in LangevinDynamics."""

    def __init__(self, model, step_size: Union[float, BaseScheduler] = 1e-3, n_leapfrog_steps: int = 10,
                 mass: Optional[Union[float, torch.Tensor]] = None, dtype: torch.dtype = torch.float32,
                 device: Optional[Union[str, torch.device]] = None,
                 integrator: Union[str, BaseSymplecticIntegrator, None] = None, rng: str = "torch"):
        super().__init__(model=model, dtype=dtype, device=device)
        self._register_param("step_size", step_size, positive=True)
        if n_leapfrog_steps <= 0:
            raise ValueError("n_leapfrog_steps must be positive")
        if rng not in ("torch", "native"):
            raise ValueError("rng must be 'torch' or 'native'")
        self.n_leapfrog_steps = n_leapfrog_steps
        self.mass = mass.to(self.device) if (mass is not None and not isinstance(mass, float)) else mass
        self.rng = rng
        integ = resolve_integrator(integrator, default="leapfrog", family=BaseSymplecticIntegrator,
                                   owner="HamiltonianMonteCarlo", device=self.device, dtype=self.dtype)
        if not integ.separable:
            raise TypeError("HamiltonianMonteCarlo requires a separable symplectic integrator")
        self.integrator = integ

    @torch.no_grad()
    def sample(self, x: Optional[torch.Tensor] = None, dim: Optional[int] = None, n_steps: int = 100,
               n_samples: int = 1, thin: int = 1, return_trajectory: bool = False, return_diagnostics: bool = False,
               reset_schedulers: bool = True, *, model_kwargs: Optional[Dict[str, Any]] = None,
               generator: Optional[torch.Generator] = None):
        if thin < 1:
            raise ValueError("thin must be >= 1")
        self._require_cuda_fp32()
        if reset_schedulers:
            self.reset_schedulers()
        model_kwargs = self._prepare_model_kwargs(model_kwargs)
        if x is None and dim is None:
            if hasattr(self.model, "mean") and isinstance(self.model.mean, torch.Tensor):
                dim = self.model.mean.shape[0]
            else:
                raise ValueError("dim must be provided when x is None and cannot be inferred from model")
        gen, seed, _ = self._rng_state(generator)
        x = self._init_state(x, dim, n_samples, generator)
        offset = gen.get_offset()
        if x.ndim != 2:
            raise ValueError(f"HamiltonianMonteCarlo expects a 2-D state [n_samples, dim], got {tuple(x.shape)}")
        n, d = x.shape
        n_kept = n_steps // thin
        desc = self._descriptor(x, model_kwargs) if type(self.integrator) is LeapfrogIntegrator else None
        if desc is None or (desc.kind == "mlp" and max(desc.c.dim, desc.c.hidden1, desc.c.hidden2) > 128):
            # no fused HMC kernel for this energy (custom models, MLP energies wider than 128, conditioning, custom
            # symplectic integrators): the integrator-level path, hmc.py:244-312 step for step
            return self._sample_opaque(x, n_steps, thin, return_trajectory, return_diagnostics, model_kwargs, generator)
        x = x.contiguous()
        rng_mode = _lib.RNG_MODES[self.rng]
        traj = torch.empty((n, n_kept, d), dtype=self.dtype, device=self.device) if return_trajectory else None
        diag = None
        if return_diagnostics:
            diag = {k: torch.empty(n_kept, d, dtype=self.dtype, device=self.device) for k in ("mean", "var")}
            diag["energy"] = torch.empty(n_kept, dtype=self.dtype, device=self.device)
            diag["acceptance_rate"] = torch.empty(n_kept, dtype=self.dtype, device=self.device)
        if n_steps <= 0:
            out = traj if return_trajectory else x
            return (out, diag) if return_diagnostics else out
        vals, constant = self._advance_schedules(("step_size",), n_steps)
        hs = vals["step_size"]

        if not return_diagnostics:
            out = ops.hmc_burst(desc, x, n_steps, self.n_leapfrog_steps, hs, mass=self.mass, rng_mode=rng_mode,
                                seed=seed, offset=offset, traj=traj, thin=thin)
            gen.set_offset(offset + ops.rng_consumed_hmc(self.device, n, d, n_steps, rng_mode))
            return traj if return_trajectory else out

        cur = x
        done = 0
        acc = torch.zeros(n_steps, dtype=torch.int32, device=self.device)
        e_out = torch.empty(n, dtype=self.dtype, device=self.device)
        for j in range(n_kept):
            h_j = hs if constant else hs[done:done + thin]
            cur = ops.hmc_burst(desc, cur, thin, self.n_leapfrog_steps, h_j, mass=self.mass, rng_mode=rng_mode,
                                seed=seed, offset=offset, accept_count=acc[done:done + thin], energy_out=e_out)
            offset += ops.rng_consumed_hmc(self.device, n, d, thin, rng_mode)
            done += thin
            if traj is not None:
                traj[:, j, :] = cur
            diag["mean"][j] = cur.mean(dim=0)
            diag["var"][j] = (cur.var(dim=0, unbiased=False).clamp_(min=1e-10, max=1e10) if n > 1
                              else torch.zeros(d, dtype=self.dtype, device=self.device))
            diag["energy"][j] = e_out.mean()
            diag["acceptance_rate"][j] = acc[done - 1].to(self.dtype) / n
        rest = n_steps - done
        if rest > 0:
            h_j = hs if constant else hs[done:]
            cur = ops.hmc_burst(desc, cur, rest, self.n_leapfrog_steps, h_j, mass=self.mass, rng_mode=rng_mode,
                                seed=seed, offset=offset)
            offset += ops.rng_consumed_hmc(self.device, n, d, rest, rng_mode)
        gen.set_offset(offset)
        out = traj if return_trajectory else cur
        return out, diag

    def _kinetic(self, p: torch.Tensor) -> torch.Tensor:
        """hmc.py:136-159."""
        if self.mass is None:
            return 0.5 * torch.sum(p.square(), dim=-1)
        if isinstance(self.mass, float):
            return 0.5 * torch.sum(p.square(), dim=-1) / self.mass
        return 0.5 * torch.sum(p.square() / self.mass.view((1,) * (p.ndim - 1) + (-1,)), dim=-1)

    def _sample_opaque(self, x, n_steps, thin, return_trajectory, return_diagnostics, model_kwargs, generator):
        """Energies with no fused HMC kernel: their own energy / gradient (the library's MLP kernels when the model is
        an `MLPEnergy`, autograd otherwise) driven through the integrator, proposal by proposal, in the reference's
        order of operations and draws (hmc.py:244-312: `normal_` for the momentum, `rand(N)` for the accept test)."""
        n, d = x.shape
        n_kept = n_steps // thin
        traj = torch.empty((n, n_kept, d), dtype=self.dtype, device=self.device) if return_trajectory else None
        diag = None
        if return_diagnostics:
            diag = {k: torch.empty(n_kept, d, dtype=self.dtype, device=self.device) for k in ("mean", "var")}
            diag["energy"] = torch.empty(n_kept, dtype=self.dtype, device=self.device)
            diag["acceptance_rate"] = torch.empty(n_kept, dtype=self.dtype, device=self.device)
        drift = lambda x_, t_: -self._model_gradient(x_, model_kwargs)
        keep = 0
        for i in range(n_steps):
            p = torch.empty_like(x).normal_(generator=generator)
            if self.mass is not None:
                p = p * (self.mass ** 0.5 if isinstance(self.mass, float) else torch.sqrt(self.mass).view(1, -1))
            h0 = self._model_energy(x, model_kwargs).clamp(min=-1e10, max=1e10) + self._kinetic(p).clamp(min=0.0, max=1e10)
            prop = self.integrator.integrate({"x": x, "p": p}, step_size=self.get_scheduled_value("step_size"),
                                             n_steps=self.n_leapfrog_steps, mass=self.mass, drift=drift, safe=True)
            xp, pp = prop["x"], prop["p"]
            h1 = self._model_energy(xp, model_kwargs).clamp(min=-1e10, max=1e10) + self._kinetic(pp).clamp(min=0.0, max=1e10)
            acc_prob = torch.exp((h0 - h1).clamp(min=-50.0, max=50.0)).clamp(max=1.0)
            u = torch.rand(n, device=self.device, dtype=self.dtype, generator=generator)
            accepted = u < acc_prob
            x = torch.where(accepted.view(-1, 1), xp, x)
            if (i + 1) % thin == 0:
                if traj is not None:
                    traj[:, keep, :] = x
                if diag is not None:
                    diag["mean"][keep] = x.mean(dim=0)
                    diag["var"][keep] = (x.var(dim=0, unbiased=False).clamp_(min=1e-10, max=1e10) if n > 1
                                         else torch.zeros(d, dtype=self.dtype, device=self.device))
                    diag["energy"][keep] = self._model_energy(x, model_kwargs).clamp(min=-1e10, max=1e10).mean()
                    diag["acceptance_rate"][keep] = accepted.to(self.dtype).mean()
                keep += 1
            self.step_schedulers()
        out = traj if return_trajectory else x
        return (out, diag) if return_diagnostics else out


class _DescentSampler(BaseSampler):
    """Shared body of the noise-free samplers (samplers/gradient_descent.py:16-281): elementwise library energies run as
    one fused burst (`ops.descent_burst`); every other energy steps through its own gradient (the library's kernels for
    recognised energies, autograd otherwise) with the reference's ATen update ops."""

    _momentum: Optional[float] = None

    @torch.no_grad()
    def sample(self, x: Optional[torch.Tensor] = None, dim: Optional[Union[int, Tuple[int, ...]]] = None,
               n_steps: int = 100, n_samples: int = 1, thin: int = 1, return_trajectory: bool = False,
               return_diagnostics: bool = False, reset_schedulers: bool = True, *,
               model_kwargs: Optional[Dict[str, Any]] = None, generator: Optional[torch.Generator] = None):
        if thin < 1:
            raise ValueError("thin must be >= 1")
        self._require_cuda_fp32()
        if reset_schedulers:
            self.reset_schedulers()
        x = self._init_state(x, dim, n_samples, generator)
        model_kwargs = self._prepare_model_kwargs(model_kwargs)
        n, data_shape = x.shape[0], x.shape[1:]
        n_kept = n_steps // thin
        traj = torch.empty((n, n_kept, *data_shape), dtype=self.dtype, device=self.device) if return_trajectory else None
        diag = {"energy": torch.empty(n_kept, dtype=self.dtype, device=self.device)} if return_diagnostics else None
        desc = self._descriptor(x, model_kwargs)
        mu = self._momentum
        if desc is not None and desc.kind in ("double_well", "harmonic", "rastrigin") and n_steps > 0 and not return_diagnostics:
            vals, _ = self._advance_schedules(("step_size",), n_steps)
            out = ops.descent_burst(desc, x.contiguous(), n_steps, vals["step_size"], momentum=mu, traj=traj, thin=thin)
            return traj if return_trajectory else out
        # step by step (diagnostics need the energy of every kept state; other energies have no fused descent kernel)
        v = torch.zeros_like(x) if mu is not None else None
        keep = 0
        for i in range(n_steps):
            eta = self.get_scheduled_value("step_size")
            if mu is None:
                x = torch.sub(x, self._model_gradient(x, model_kwargs), alpha=eta)
            else:
                lookahead = torch.add(x, v, alpha=mu)
                v.mul_(mu).sub_(self._model_gradient(lookahead, model_kwargs), alpha=eta)
                x = x + v
            if (i + 1) % thin == 0:
                if traj is not None:
                    traj[:, keep] = x
                if diag is not None:
                    diag["energy"][keep] = self._model_energy(x, model_kwargs).mean()
                keep += 1
            self.step_schedulers()
        out = traj if return_trajectory else x
        return (out, diag) if return_diagnostics else out


class GradientDescentSampler(_DescentSampler):
    """samplers/gradient_descent.py:16-140: x <- x - eta * grad E(x)."""

    def __init__(self, model, step_size: Union[float, BaseScheduler] = 1e-3, dtype: torch.dtype = torch.float32,
                 device: Optional[Union[str, torch.device]] = None):
        super().__init__(model=model, dtype=dtype, device=device)
        self._register_param("step_size", step_size, positive=True)


class NesterovSampler(_DescentSampler):
    """samplers/gradient_descent.py:143-276: v <- mu v - eta grad E(x + mu v); x <- x + v."""

    def __init__(self, model, step_size: Union[float, BaseScheduler] = 1e-3, momentum: float = 0.9,
                 dtype: torch.dtype = torch.float32, device: Optional[Union[str, torch.device]] = None):
        super().__init__(model=model, dtype=dtype, device=device)
        if not (0 <= momentum < 1):
            raise ValueError("momentum must be in [0, 1)")
        self.momentum = momentum
        self._momentum = float(momentum)
        self._register_param("step_size", step_size, positive=True)
