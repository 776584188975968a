"""`LangevinDynamics` and `HamiltonianMonteCarlo` with the reference's API, running fused CUDA bursts.

Mirrors torchebm/core/base_sampler.py:10-155, torchebm/samplers/langevin_dynamics.py:16-188 and
torchebm/samplers/hmc.py:19-315: same constructor and `sample()` signatures, same return shapes,
same errors, same scheduler semantics, same RNG consumption (with `rng = "torch"`, the default, a burst
draws exactly the Philox stream the reference's per-step `randn_like` / `normal_` / `rand` calls would,
so equal seeds give equal chains).  What differs is where the loop runs: one kernel launch does all
`n_steps` steps, diagnostics included.

The fused bodies live in mixins (`FusedLangevinMixin`, `FusedHMCMixin`, `FusedDescentMixin`) that are combined with
one of two bases:
  * this module's standalone classes (no reference package needed; requests outside the fused kernels' scope step
    through the integrator-level path, and CPU tensors, non-fp32 dtypes or a missing CUDA library raise), or
  * the reference's own classes when `torchebm` is importable (`torchebm_b200/dropin.py`): everything the fused path
    does not cover is then handed to the reference's own `sample()` through `super()`.
The mixin decides BEFORE it consumes any scheduler or generator state; `sampler.last_path` says which way the last call
went ("fused" / "unfused").  `rng` ("torch" or "native") is an attribute, not a constructor argument: the constructor
signatures are the reference's, `integrator` last (tests/samplers/test_api_contract.py:199-201).
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Dict, Optional, Tuple, Union

import torch
from torch import nn

from . import _lib, ops
from .core import (BaseScheduler, EnergyDescriptor, Schedulable, TorchEBMModule, advance_schedules, autograd_gradient,
                   energy_descriptor)
from .integrators import (BaseSDERungeKuttaIntegrator, BaseSymplecticIntegrator, EulerMaruyamaIntegrator,
                          HeunIntegrator, LeapfrogIntegrator, resolve_integrator, sde_scheme_of, is_plain_leapfrog)

_ELEMENTWISE = ("double_well", "harmonic", "rastrigin")


def _generator_state(device: torch.device, generator: Optional[torch.Generator]) -> Tuple[torch.Generator, int, int]:
    """(generator, seed, offset) of the Philox stream a call draws from; no host sync."""
    if generator is None:
        generator = torch.cuda.default_generators[ops.device_index(device)]
    elif generator.device.type != "cuda":
        raise RuntimeError(f"Expected a 'cuda' device type for generator but found '{generator.device.type}'")
    return generator, generator.initial_seed(), generator.get_offset()


def _state_width(x, dim) -> Optional[int]:
    """Row length of the 2-D state a call will run on, or None when the state is not [n, d]."""
    if x is not None:
        return int(x.shape[1]) if x.ndim == 2 else None
    if dim is None:
        return None
    shape = (dim,) if isinstance(dim, int) else tuple(dim)
    return int(shape[0]) if len(shape) == 1 else None


class _RngAttribute:
    """`sampler.rng`: "torch" (default; the reference's own Philox stream, equal seeds give equal chains) or "native"
    (one Philox block per aligned quad of consecutive elements: cheaper, statistically equivalent)."""

    _rng = "torch"

    @property
    def rng(self) -> str:
        return self._rng

    @rng.setter
    def rng(self, value: str) -> None:
        if value not in ("torch", "native"):
            raise ValueError("rng must be 'torch' or 'native'")
        object.__setattr__(self, "_rng", value)

    def with_rng(self, value: str):
        self.rng = value
        return self


def _batch_diag(x: torch.Tensor):
    if x.shape[0] > 1:
        return x.mean(dim=0), x.var(dim=0, unbiased=False).clamp_(min=1e-10, max=1e10)
    return x.squeeze(0), torch.zeros_like(x.squeeze(0))


# ======================================================================================================
# fused bodies


class FusedLangevinMixin(_RngAttribute):
    """`sample()` of langevin_dynamics.py:82-188 as one fused burst."""

    last_path: Optional[str] = None

    def _fused_plan(self, x, dim, model_kwargs):
        if self.device.type != "cuda" or self.dtype != torch.float32 or model_kwargs:
            return None
        d = _state_width(x, dim)
        scheme = sde_scheme_of(self.integrator)
        if d is None or scheme is None:
            return None
        desc = energy_descriptor(self.model, d, self.device)
        if desc is None or (scheme == "heun" and desc.kind not in _ELEMENTWISE):
            return None   # the fused Heun burst exists for the elementwise energies; others step through the integrator
        return scheme, desc

    @torch.no_grad()
    def sample(self, x: Optional[torch.Tensor] = None, dim: Optional[Union[int, Tuple[int, ...]]] = None,
               n_steps: int = 100, n_samples: int = 1, thin: int = 1, return_trajectory: bool = False,
               return_diagnostics: bool = False, reset_schedulers: bool = True, *,
               model_kwargs: Optional[Dict[str, Any]] = None, generator: Optional[torch.Generator] = None):
        if thin < 1:
            raise ValueError("thin must be >= 1")
        plan = self._fused_plan(x, dim, model_kwargs)
        if plan is None or n_steps <= 0:
            self.last_path = "unfused"
            return self._sample_unfused(x, dim, n_steps, n_samples, thin, return_trajectory, return_diagnostics,
                                        reset_schedulers, model_kwargs, generator)
        scheme, desc = plan
        _lib.load()   # a missing CUDA library is an error, never a reason to take another path
        self.last_path = "fused"
        if reset_schedulers:
            self.reset_schedulers()
        gen, seed, _ = _generator_state(self.device, generator)
        x = self._init_state(x, dim, n_samples, generator).contiguous()
        offset = gen.get_offset()  # after _init_state, which may have drawn from the same generator
        n, data_shape = x.shape[0], x.shape[1:]
        n_kept = n_steps // thin
        rng_mode = _lib.RNG_MODES[self.rng]
        traj = torch.empty((n, n_kept, *data_shape), dtype=self.dtype, device=self.device) if return_trajectory else None
        vals, _ = advance_schedules(self, ("step_size", "noise_scale"), n_steps)
        diag = None
        if return_diagnostics:
            diag = {"mean": torch.empty(n_kept, *data_shape, dtype=self.dtype, device=self.device),
                    "var": torch.empty(n_kept, *data_shape, dtype=self.dtype, device=self.device),
                    "energy": torch.empty(n_kept, dtype=self.dtype, device=self.device)}
        out = ops.langevin_burst(desc, x, n_steps, vals["step_size"], vals["noise_scale"], clamp=self.clamp,
                                 rng_mode=rng_mode, seed=seed, offset=offset, traj=traj, thin=thin, scheme=scheme,
                                 diag=diag)
        gen.set_offset(offset + ops.rng_consumed_langevin(self.device, x.numel(), n_steps, rng_mode))
        result = traj if return_trajectory else out
        return (result, diag) if return_diagnostics else result

    @torch.no_grad()
    def sample_from_buffer(self, buffer: torch.Tensor, indices: Optional[torch.Tensor], ptr: int, n_steps: int, *,
                           noise_rows: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None,
                           energy_out: Optional[torch.Tensor] = None, reset_schedulers: bool = True,
                           generator: Optional[torch.Generator] = None, gather_into=None):
        """Persistent-CD negatives in one library call: start points `buffer[indices]` (`indices=None`: row i for chain
        i, the stride-1 case where the whole buffer is replaced) plus `0.01 * noise[j]` on row `noise_rows[j]`, `n_steps`
        Langevin steps, FIFO write-back into `buffer` at `ptr` (get_start_points + sample + update_buffer of
        core/base_loss.py:266-337,390-426 and losses/contrastive_divergence.py:127-139); `energy_out[n]` receives
        E(x-) of the returned negatives.  Same scheduler and generator semantics as
        `sample(x=start_points, n_steps=n_steps, generator=generator)`.  Returns `(negatives, new_ptr)`, or None when
        this sampler / energy has no library kernel for it (the caller then takes the three-call path).
        `gather_into` (a `distributed.PeerGatherBuffer`): the negatives of this rank also land in every rank's gathered
        tensor, stored by the burst kernel itself; the caller issues `gather_into.barrier()` before reading it."""
        if sde_scheme_of(self.integrator) != "euler_maruyama" or buffer.ndim != 2 or not buffer.is_cuda:
            return None   # (the one-call PCD path is Euler-Maruyama only)
        if self.device.type != "cuda" or self.dtype != torch.float32:
            return None
        desc = energy_descriptor(self.model, buffer.shape[1], buffer.device)
        if desc is None or n_steps <= 0:
            return None
        _lib.load()
        if reset_schedulers:
            self.reset_schedulers()
        gen, seed, offset = _generator_state(self.device, generator)
        rng_mode = _lib.RNG_MODES[self.rng]
        vals, _ = advance_schedules(self, ("step_size", "noise_scale"), n_steps)
        out, new_ptr = ops.pcd_langevin_burst(desc, buffer, indices, ptr, n_steps, vals["step_size"], vals["noise_scale"],
                                              clamp=self.clamp, rng_mode=rng_mode, seed=seed, offset=offset,
                                              noise_rows=noise_rows, noise=noise, energy_out=energy_out,
                                              **({} if gather_into is None else
                                                 {"peer_ptrs": gather_into.ptrs, "multicast_ptr": gather_into.mc_ptr,
                                                  "row_offset": gather_into.rank * gather_into.rows_per_rank}))
        gen.set_offset(offset + ops.rng_consumed_langevin(self.device, out.numel(), n_steps, rng_mode))
        return out, new_ptr


class FusedHMCMixin(_RngAttribute):
    """`sample()` of hmc.py:161-315 as one fused burst (all proposals of a call in one kernel)."""

    last_path: Optional[str] = None

    def _fused_plan(self, x, dim, model_kwargs):
        if self.device.type != "cuda" or self.dtype != torch.float32 or model_kwargs:
            return None
        if not is_plain_leapfrog(self.integrator):
            return None
        if x is None and dim is None and hasattr(self.model, "mean") and isinstance(self.model.mean, torch.Tensor):
            dim = self.model.mean.shape[0]
        d = _state_width(x, dim)
        if d is None:
            return None
        desc = energy_descriptor(self.model, d, self.device)
        if desc is None:
            return None
        if desc.kind == "mlp" and (max(desc.c.dim, desc.c.hidden1, desc.c.hidden2) > 128 or desc.c.hidden3 > 0
                                   or desc.c.precision == _lib.MLP_BF16):
            return None   # no fused HMC kernel: wider MLP energies, single-pass bf16 (accept tests want fp32-grade energies)
        return desc, dim

    @torch.no_grad()
    def sample(self, x: Optional[torch.Tensor] = None, dim: Optional[int] = None, n_steps: int = 100,
               n_samples: int = 1, thin: int = 1, return_trajectory: bool = False, return_diagnostics: bool = False,
               reset_schedulers: bool = True, *, model_kwargs: Optional[Dict[str, Any]] = None,
               generator: Optional[torch.Generator] = None):
        if thin < 1:
            raise ValueError("thin must be >= 1")
        plan = self._fused_plan(x, dim, model_kwargs)
        if plan is None or n_steps <= 0:
            self.last_path = "unfused"
            return self._sample_unfused(x, dim, n_steps, n_samples, thin, return_trajectory, return_diagnostics,
                                        reset_schedulers, model_kwargs, generator)
        desc, dim = plan
        _lib.load()
        self.last_path = "fused"
        if reset_schedulers:
            self.reset_schedulers()
        gen, seed, _ = _generator_state(self.device, generator)
        x = self._init_state(x, dim, n_samples, generator).contiguous()
        offset = gen.get_offset()
        n, d = x.shape
        n_kept = n_steps // thin
        rng_mode = _lib.RNG_MODES[self.rng]
        traj = torch.empty((n, n_kept, d), dtype=self.dtype, device=self.device) if return_trajectory else None
        vals, _ = advance_schedules(self, ("step_size",), n_steps)
        diag = None
        if return_diagnostics:
            diag = {k: torch.empty(n_kept, d, dtype=self.dtype, device=self.device) for k in ("mean", "var")}
            diag["energy"] = torch.empty(n_kept, dtype=self.dtype, device=self.device)
            diag["acceptance_rate"] = torch.empty(n_kept, dtype=self.dtype, device=self.device)
        out = ops.hmc_burst(desc, x, n_steps, self.n_leapfrog_steps, vals["step_size"], mass=self.mass, rng_mode=rng_mode,
                            seed=seed, offset=offset, traj=traj, thin=thin, diag=diag)
        gen.set_offset(offset + ops.rng_consumed_hmc(self.device, n, d, n_steps, rng_mode))
        result = traj if return_trajectory else out
        return (result, diag) if return_diagnostics else result


class FusedDescentMixin:
    """`sample()` of samplers/gradient_descent.py:62-138,196-276 as one fused burst for the elementwise energies."""

    last_path: Optional[str] = None

    def _descent_momentum(self) -> Optional[float]:
        return float(self.momentum) if hasattr(self, "momentum") else None

    @torch.no_grad()
    def sample(self, x: Optional[torch.Tensor] = None, dim: Optional[Union[int, Tuple[int, ...]]] = None,
               n_steps: int = 100, n_samples: int = 1, thin: int = 1, return_trajectory: bool = False,
               return_diagnostics: bool = False, reset_schedulers: bool = True, *,
               model_kwargs: Optional[Dict[str, Any]] = None, generator: Optional[torch.Generator] = None):
        if thin < 1:
            raise ValueError("thin must be >= 1")
        desc = None
        d = _state_width(x, dim)
        if (self.device.type == "cuda" and self.dtype == torch.float32 and not model_kwargs and d is not None
                and n_steps > 0 and not return_diagnostics):
            desc = energy_descriptor(self.model, d, self.device)
            if desc is not None and desc.kind not in _ELEMENTWISE:
                desc = None
        if desc is None:   # diagnostics need the energy of every kept state; other energies have no fused descent kernel
            self.last_path = "unfused"
            return self._sample_unfused(x, dim, n_steps, n_samples, thin, return_trajectory, return_diagnostics,
                                        reset_schedulers, model_kwargs, generator)
        _lib.load()
        self.last_path = "fused"
        if reset_schedulers:
            self.reset_schedulers()
        x = self._init_state(x, dim, n_samples, generator).contiguous()
        n, data_shape = x.shape[0], x.shape[1:]
        traj = (torch.empty((n, n_steps // thin, *data_shape), dtype=self.dtype, device=self.device)
                if return_trajectory else None)
        vals, _ = advance_schedules(self, ("step_size",), n_steps)
        out = ops.descent_burst(desc, x, n_steps, vals["step_size"], momentum=self._descent_momentum(), traj=traj, thin=thin)
        return traj if return_trajectory else out


# ======================================================================================================
# standalone bases (used when the reference package is not importable)


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

    def _require_cuda_fp32(self) -> None:
        """The standalone classes have nothing to hand other devices / dtypes to (no reference package here)."""
        if self.device.type != "cuda":
            raise RuntimeError(
                f"{type(self).__name__} runs on a CUDA device only (got device={self.device}); "
                "torchebm_b200 has no CPU path")
        if self.dtype != torch.float32:
            raise TypeError(f"{type(self).__name__} supports dtype=torch.float32 only, got {self.dtype}")
        _lib.load()

    @abstractmethod
    def sample(self, x=None, dim=None, n_steps=100, n_samples=1, thin=1, return_trajectory=False,
               return_diagnostics=False, reset_schedulers=True, *, generator=None): ...


class _LangevinBase(BaseSampler):
    """Constructor of langevin_dynamics.py:53-80 and, as `_sample_unfused`, its per-step loop (:157-185) through the
    integrator-level boundary: energies with no library kernel keep their own gradient, the integrator runs the
    library's fused update."""

    def __init__(self, model, step_size: Union[float, BaseScheduler] = 1e-3,
                 noise_scale: Union[float, BaseScheduler] = 1.0, decay: float = 0.0,
                 clamp: Optional[Tuple[float, float]] = None, dtype: torch.dtype = torch.float32,
                 device: Optional[Union[str, torch.device]] = None,
                 integrator: Union[str, BaseSDERungeKuttaIntegrator, None] = None):
        super().__init__(model=model, dtype=dtype, device=device)
        self._register_param("step_size", step_size, positive=True)
        self._register_param("noise_scale", noise_scale, positive=True)
        if clamp is not None and clamp[0] >= clamp[1]:
            raise ValueError(f"clamp min must be < max, got {clamp}")
        self.clamp = clamp
        self.decay = decay
        self.integrator = resolve_integrator(integrator, default="euler_maruyama", family=BaseSDERungeKuttaIntegrator,
                                             owner="LangevinDynamics", device=self.device, dtype=self.dtype)

    def _sample_unfused(self, x, dim, n_steps, n_samples, thin, return_trajectory, return_diagnostics, reset_schedulers,
                        model_kwargs, generator):
        self._require_cuda_fp32()
        if reset_schedulers:
            self.reset_schedulers()
        x = self._init_state(x, dim, n_samples, generator)
        model_kwargs = self._prepare_model_kwargs(model_kwargs)
        n, data_shape = x.shape[0], x.shape[1:]
        n_kept = n_steps // thin
        traj = torch.empty((n, n_kept, *data_shape), dtype=self.dtype, device=self.device) if return_trajectory else None
        diag = None
        if return_diagnostics:
            diag = {"mean": torch.empty(n_kept, *data_shape, dtype=self.dtype, device=self.device),
                    "var": torch.empty(n_kept, *data_shape, dtype=self.dtype, device=self.device),
                    "energy": torch.empty(n_kept, dtype=self.dtype, device=self.device)}
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


class LangevinDynamics(FusedLangevinMixin, _LangevinBase):
    """langevin_dynamics.py:16-188."""


class _HMCBase(BaseSampler):
    """Constructor of hmc.py:55-90 and, as `_sample_unfused`, its proposal loop (:244-312) through the integrator."""

    def __init__(self, model, step_size: Union[float, BaseScheduler] = 1e-3, n_leapfrog_steps: int = 10,
                 mass: Optional[Union[float, torch.Tensor]] = None, dtype: torch.dtype = torch.float32,
                 device: Optional[Union[str, torch.device]] = None,
                 integrator: Union[str, BaseSymplecticIntegrator, None] = None):
        super().__init__(model=model, dtype=dtype, device=device)
        self._register_param("step_size", step_size, positive=True)
        if n_leapfrog_steps <= 0:
            raise ValueError("n_leapfrog_steps must be positive")
        self.n_leapfrog_steps = n_leapfrog_steps
        self.mass = mass.to(self.device) if (mass is not None and not isinstance(mass, float)) else mass
        integ = resolve_integrator(integrator, default="leapfrog", family=BaseSymplecticIntegrator,
                                   owner="HamiltonianMonteCarlo", device=self.device, dtype=self.dtype)
        if not integ.separable:
            raise TypeError("HamiltonianMonteCarlo requires a separable symplectic integrator")
        self.integrator = integ

    def _kinetic(self, p: torch.Tensor) -> torch.Tensor:
        """hmc.py:136-159."""
        if self.mass is None:
            return 0.5 * torch.sum(p.square(), dim=-1)
        if isinstance(self.mass, float):
            return 0.5 * torch.sum(p.square(), dim=-1) / self.mass
        return 0.5 * torch.sum(p.square() / self.mass.view((1,) * (p.ndim - 1) + (-1,)), dim=-1)

    def _sample_unfused(self, x, dim, n_steps, n_samples, thin, return_trajectory, return_diagnostics, reset_schedulers,
                        model_kwargs, generator):
        """Energies with no fused HMC kernel: their own energy / gradient (the library's MLP kernels when the model is
        an `MLPEnergy`, autograd otherwise) driven through the integrator, proposal by proposal, in the reference's
        order of operations and draws (hmc.py:244-312: `normal_` for the momentum, `rand(N)` for the accept test)."""
        self._require_cuda_fp32()
        if reset_schedulers:
            self.reset_schedulers()
        model_kwargs = self._prepare_model_kwargs(model_kwargs)
        if x is None and dim is None:
            if hasattr(self.model, "mean") and isinstance(self.model.mean, torch.Tensor):
                dim = self.model.mean.shape[0]
            else:
                raise ValueError("dim must be provided when x is None and cannot be inferred from model")
        x = self._init_state(x, dim, n_samples, generator)
        if x.ndim != 2:
            raise ValueError(f"HamiltonianMonteCarlo expects a 2-D state [n_samples, dim], got {tuple(x.shape)}")
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


class HamiltonianMonteCarlo(FusedHMCMixin, _HMCBase):
    """hmc.py:19-315."""


class _DescentBase(BaseSampler):
    """Step-by-step loops of samplers/gradient_descent.py:123-138,258-276 with the reference's ATen update ops (every
    energy steps through its own gradient: the library's kernels for recognised energies, autograd otherwise)."""

    def _sample_unfused(self, x, dim, n_steps, n_samples, thin, return_trajectory, return_diagnostics, reset_schedulers,
                        model_kwargs, generator):
        self._require_cuda_fp32()
        if reset_schedulers:
            self.reset_schedulers()
        x = self._init_state(x, dim, n_samples, generator)
        model_kwargs = self._prepare_model_kwargs(model_kwargs)
        n, data_shape = x.shape[0], x.shape[1:]
        n_kept = n_steps // thin
        traj = torch.empty((n, n_kept, *data_shape), dtype=self.dtype, device=self.device) if return_trajectory else None
        diag = {"energy": torch.empty(n_kept, dtype=self.dtype, device=self.device)} if return_diagnostics else None
        mu = self._descent_momentum()
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


class GradientDescentSampler(FusedDescentMixin, _DescentBase):
    """samplers/gradient_descent.py:16-140: x <- x - eta * grad E(x)."""

    def __init__(self, model, step_size: Union[float, BaseScheduler] = 1e-3, dtype: torch.dtype = torch.float32,
                 device: Optional[Union[str, torch.device]] = None):
        super().__init__(model=model, dtype=dtype, device=device)
        self._register_param("step_size", step_size, positive=True)


class NesterovSampler(FusedDescentMixin, _DescentBase):
    """samplers/gradient_descent.py:143-276: v <- mu v - eta grad E(x + mu v); x <- x + v."""

    def __init__(self, model, step_size: Union[float, BaseScheduler] = 1e-3, momentum: float = 0.9,
                 dtype: torch.dtype = torch.float32, device: Optional[Union[str, torch.device]] = None):
        super().__init__(model=model, dtype=dtype, device=device)
        if not (0 <= momentum < 1):
            raise ValueError("momentum must be in [0, 1)")
        self.momentum = momentum
        self._register_param("step_size", step_size, positive=True)
