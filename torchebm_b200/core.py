"""Host-side mirror of the torchebm.core pieces the sampling path touches.

Same names, constructor arguments and error behaviour as the reference so that code written against
`torchebm.core` runs against this package:

* `TorchEBMModule` (torchebm/core/base_module.py:51-176): device/dtype probe, `_prepare_model_kwargs`.
* `BaseScheduler` / `ConstantScheduler` / `LinearScheduler` / `ExponentialDecayScheduler` /
  `CosineScheduler` (torchebm/core/base_scheduler.py:73-625) and the `Schedulable` mixin
  (torchebm/core/schedulable.py:16-75).  Schedules are host floats; a fused K-step burst precomputes
  the K values and then advances the scheduler objects K times, which keeps `step_count` semantics.
* `BaseModel` and the analytic energies (torchebm/core/base_model.py) plus `MLPEnergy` and
  `MixtureOfGaussiansModel`.  Their `forward` stays PyTorch (the CD loss back-propagates through it);
  `gradient` and the sampler bursts run in the CUDA library through an `EbmEnergyDesc`.

Schedulers from the reference package itself are accepted too (duck-typed on
`step/get_value/reset`), as are the reference's own model classes (matched by class name and
attributes in `energy_descriptor`).
"""

from __future__ import annotations

import ctypes as C
import math
import warnings
from abc import ABC, abstractmethod
from contextlib import nullcontext
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import torch
from torch import nn

from . import _lib

# --------------------------------------------------------------------------------------------------
# module base


def _normalize(device: torch.device) -> torch.device:
    # base_module.py:24-27
    if device.type == "cuda" and device.index == 0:
        return torch.device("cuda")
    return device


class TorchEBMModule(nn.Module):
    def __init__(self, device: Union[str, torch.device, None] = None, dtype: Optional[torch.dtype] = None, *args, **kwargs):
        super().__init__(*args, **kwargs)
        probe_dtype = dtype if dtype is not None else torch.get_default_dtype()
        self.register_buffer("_torchebm_probe", torch.empty(0, dtype=probe_dtype, device=device), persistent=False)

    @property
    def device(self) -> torch.device:
        for p in self.parameters():
            return _normalize(p.device)
        return _normalize(self._torchebm_probe.device)

    @property
    def dtype(self) -> torch.dtype:
        for p in self.parameters():
            return p.dtype
        return self._torchebm_probe.dtype

    def _prepare_model_kwargs(self, model_kwargs: Optional[dict]) -> dict:
        if not model_kwargs:
            return {}
        if not isinstance(model_kwargs, dict):
            raise TypeError(f"model_kwargs must be a dict, got {type(model_kwargs).__name__}")
        device = self.device
        return {k: (v.to(device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in model_kwargs.items()}

    def autocast_context(self):
        return nullcontext()


# --------------------------------------------------------------------------------------------------
# schedulers


class BaseScheduler(ABC):
    def __init__(self, start_value: float):
        if not isinstance(start_value, (int, float)):
            raise TypeError(f"{type(self).__name__} received an invalid start_value")
        self.start_value = float(start_value)
        self.current_value = float(start_value)
        self.step_count = 0

    @abstractmethod
    def _compute_value(self) -> float: ...

    def step(self) -> float:
        self.step_count += 1
        self.current_value = self._compute_value()
        return self.current_value

    def reset(self) -> None:
        self.current_value = self.start_value
        self.step_count = 0

    def get_value(self) -> float:
        return self.current_value

    def state_dict(self) -> Dict[str, Any]:
        return dict(self.__dict__)

    def load_state_dict(self, state: Dict[str, Any]) -> None:
        self.__dict__.update(state)


class ConstantScheduler(BaseScheduler):
    def _compute_value(self) -> float:
        return self.start_value


class ExponentialDecayScheduler(BaseScheduler):
    def __init__(self, start_value: float, decay_rate: float, min_value: float = 0.0):
        super().__init__(start_value)
        if not 0.0 < decay_rate <= 1.0:
            raise ValueError("decay_rate must be in (0, 1]")
        if min_value < 0:
            raise ValueError("min_value must be non-negative")
        self.decay_rate = decay_rate
        self.min_value = min_value

    def _compute_value(self) -> float:
        return max(self.min_value, self.start_value * (self.decay_rate**self.step_count))


class LinearScheduler(BaseScheduler):
    def __init__(self, start_value: float, end_value: float, n_steps: int):
        super().__init__(start_value)
        if n_steps <= 0:
            raise ValueError("n_steps must be positive")
        self.end_value = end_value
        self.n_steps = n_steps
        self.step_size = (end_value - start_value) / n_steps

    def _compute_value(self) -> float:
        if self.step_count >= self.n_steps:
            return self.end_value
        return self.start_value + self.step_size * self.step_count


class CosineScheduler(BaseScheduler):
    def __init__(self, start_value: float, end_value: float, n_steps: int):
        super().__init__(start_value)
        if n_steps <= 0:
            raise ValueError("n_steps must be a positive integer")
        self.end_value = end_value
        self.n_steps = n_steps

    def _compute_value(self) -> float:
        if self.step_count >= self.n_steps:
            return self.end_value
        progress = self.step_count / self.n_steps
        return self.end_value + (self.start_value - self.end_value) * 0.5 * (1.0 + math.cos(math.pi * progress))


def _is_scheduler(obj) -> bool:
    return isinstance(obj, BaseScheduler) or all(callable(getattr(obj, a, None)) for a in ("step", "get_value", "reset"))


def _is_constant_scheduler(s) -> bool:
    return type(s).__name__ == "ConstantScheduler" and hasattr(s, "step_count") and hasattr(s, "start_value")


class Schedulable:
    """schedulable.py:16-75."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.schedulers: Dict[str, BaseScheduler] = {}

    def register_scheduler(self, name: str, scheduler: BaseScheduler) -> None:
        self.schedulers[name] = scheduler

    def _register_param(self, name: str, value, *, positive: bool = False) -> None:
        if not isinstance(value, (int, float)) and _is_scheduler(value):
            self.schedulers[name] = value
            return
        if positive and value <= 0:
            raise ValueError(f"{name} must be positive")
        self.schedulers[name] = ConstantScheduler(float(value))

    def get_schedulers(self) -> Dict[str, BaseScheduler]:
        return self.schedulers

    def get_scheduled_value(self, name: str) -> float:
        if name not in self.schedulers:
            raise KeyError(f"No scheduler registered for parameter '{name}'")
        return self.schedulers[name].get_value()

    def _subtree_schedulers(self) -> List[BaseScheduler]:
        return subtree_schedulers(self)

    def step_schedulers(self) -> None:
        for s in self._subtree_schedulers():
            s.step()

    def reset_schedulers(self) -> None:
        for s in self._subtree_schedulers():
            s.reset()

    def _advance_schedules(self, names: Sequence[str], k: int) -> Tuple[Dict[str, List[float]], bool]:
        return advance_schedules(self, names, k)


def subtree_schedulers(module: nn.Module) -> List[Any]:
    """Every scheduler `module.step_schedulers()` would step (schedulable.py:60-75: the module subtree, each
    Schedulable's own dict); works on this package's and on the reference's Schedulable alike."""
    out: List[Any] = []
    for m in module.modules():
        scheds = getattr(m, "schedulers", None)
        if isinstance(scheds, dict) and hasattr(m, "step_schedulers"):
            out.extend(scheds.values())
    return out


def advance_schedules(module: nn.Module, names: Sequence[str], k: int) -> Tuple[Dict[str, List[float]], bool]:
    """Values each of `names` takes over the next k sampler steps, advancing every scheduler in the
    subtree k times (langevin_dynamics.py:161-168: read, step, read, step, ...).

    Returns (values, constant); when every scheduler in the subtree is a ConstantScheduler the lists
    have length 1 and only the step counters are bumped."""
    all_s = subtree_schedulers(module)
    if all(_is_constant_scheduler(s) for s in all_s):
        vals = {n: [float(module.get_scheduled_value(n))] for n in names}
        for s in all_s:
            s.step_count += k
            s.current_value = s.start_value
        return vals, True
    vals = {n: [] for n in names}
    for _ in range(k):
        for n in names:
            vals[n].append(float(module.get_scheduled_value(n)))
        for s in all_s:
            s.step()
    return vals, False


# --------------------------------------------------------------------------------------------------
# energies


class BaseModel(TorchEBMModule, ABC):
    """base_model.py:10-127.  `gradient` runs the CUDA library for recognised energies on a CUDA
    device and autograd otherwise (same contract as the reference's default implementation)."""

    force_fp32_gradient: bool = False

    def __init__(self, dtype: torch.dtype = torch.float32, *args, **kwargs):
        super().__init__(dtype=dtype, *args, **kwargs)

    @abstractmethod
    def forward(self, x: torch.Tensor) -> torch.Tensor: ...

    def gradient(self, x: torch.Tensor, model_kwargs: Optional[dict] = None) -> torch.Tensor:
        if not model_kwargs and x.is_cuda and x.dtype == torch.float32 and x.ndim == 2:
            desc = energy_descriptor(self, x.shape[1], x.device)
            if desc is not None:
                from .ops import gradient as _fused_gradient

                return _fused_gradient(desc, x)
        return autograd_gradient(self, x, model_kwargs)


def autograd_gradient(model: nn.Module, x: torch.Tensor, model_kwargs: Optional[dict] = None) -> torch.Tensor:
    """base_model.py:84-127, for energies the library has no kernel for."""
    with torch.enable_grad():
        xg = x.detach().requires_grad_(True)
        energy = model(xg, **(model_kwargs or {}))
        if energy.shape != (xg.shape[0],):
            raise ValueError(f"BaseModel forward() output expected shape ({xg.shape[0]},), but got {energy.shape}.")
        if not energy.grad_fn:
            raise RuntimeError("Cannot compute gradient: `forward` method did not use the input `x` in a differentiable way.")
        (grad,) = torch.autograd.grad(energy, xg, grad_outputs=torch.ones_like(energy))
    return grad.detach()


class DoubleWellModel(BaseModel):
    def __init__(self, barrier_height: float = 2.0, b: float = 1.0, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.barrier_height = barrier_height
        self.b = b

    def forward(self, x):
        if x.ndim == 1:
            x = x.unsqueeze(0)
        return self.barrier_height * (x.pow(2) - self.b**2).pow(2).sum(dim=-1)


class HarmonicModel(BaseModel):
    def __init__(self, k: float = 1.0, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.k = k

    def forward(self, x):
        if x.ndim == 1:
            x = x.unsqueeze(0)
        return 0.5 * self.k * x.pow(2).sum(dim=-1)


class RastriginModel(BaseModel):
    def __init__(self, a: float = 10.0, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.a = a

    def forward(self, x):
        if x.ndim == 1:
            x = x.unsqueeze(0)
        n = x.shape[-1]
        return self.a * n + torch.sum(x**2 - self.a * torch.cos(2 * math.pi * x), dim=-1)


class GaussianModel(BaseModel):
    def __init__(self, mean: torch.Tensor, cov: torch.Tensor, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if mean.ndim != 1:
            raise ValueError("Mean must be a 1D tensor.")
        if cov.ndim != 2 or cov.shape[0] != cov.shape[1]:
            raise ValueError("Covariance must be a 2D square matrix.")
        if mean.shape[0] != cov.shape[0]:
            raise ValueError("Mean vector dimension must match covariance matrix dimension.")
        self.register_buffer("mean", mean.to(dtype=self.dtype, device=self.device))
        try:
            cov_inv = torch.inverse(cov)
        except RuntimeError as e:
            raise ValueError(f"Failed to invert covariance matrix: {e}. Ensure it is invertible.") from e
        self.register_buffer("cov_inv", cov_inv.to(dtype=self.dtype, device=self.device))

    def forward(self, x):
        if x.ndim == 1:
            x = x.unsqueeze(0)
        if x.ndim != 2 or x.shape[1] != self.mean.shape[0]:
            raise ValueError(f"Input x expected batch_shape (batch_size, {self.mean.shape[0]}), but got {x.shape}")
        delta = x - self.mean
        return 0.5 * torch.sum(delta * torch.matmul(delta, self.cov_inv), dim=-1)


class MixtureOfGaussiansModel(BaseModel):
    """Isotropic Gaussian mixture energy (north_star "MoG"; the reference has none):
    E(x) = -logsumexp_k(log w_k - D log sigma_k - |x - mu_k|^2 / (2 sigma_k^2))."""

    def __init__(self, means: torch.Tensor, sigmas: torch.Tensor, weights: Optional[torch.Tensor] = None, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if means.ndim != 2:
            raise ValueError("means must be [K, D]")
        k = means.shape[0]
        if weights is None:
            weights = torch.full((k,), 1.0 / k)
        if sigmas.shape != (k,) or weights.shape != (k,):
            raise ValueError("sigmas and weights must be [K]")
        self.register_buffer("means", means.to(dtype=self.dtype, device=self.device).contiguous())
        self.register_buffer("sigmas", sigmas.to(dtype=self.dtype, device=self.device).contiguous())
        self.register_buffer("weights", weights.to(dtype=self.dtype, device=self.device).contiguous())

    def forward(self, x):
        if x.ndim == 1:
            x = x.unsqueeze(0)
        d = x.shape[-1]
        diff = x.unsqueeze(1) - self.means.unsqueeze(0)
        sq = diff.pow(2).sum(dim=-1)
        logits = torch.log(self.weights) - d * torch.log(self.sigmas) - sq / (2.0 * self.sigmas**2)
        return -torch.logsumexp(logits, dim=-1)


_ACT_CODES = {nn.SiLU: _lib.ACT_SILU, nn.Tanh: _lib.ACT_TANH, nn.ReLU: _lib.ACT_RELU, nn.Softplus: _lib.ACT_SOFTPLUS}
_ACT_BY_NAME = {"silu": nn.SiLU, "tanh": nn.Tanh, "relu": nn.ReLU, "softplus": nn.Softplus}


class MLPEnergy(BaseModel):
    """`Sequential(Linear(D,H1), act, Linear(H1,H2), act, Linear(H2,1))` + `squeeze(-1)`: the MLP energies of
    examples/20-training/01-mcmc-losses/01-cd-k/main.py:20-30 and benchmarks/registry.py:375-387; with three hidden
    widths (`hidden=(H1, H2, H3)`) the deeper energy of benchmarks/distributed_fsdp2.py:43-53.
    Pass an existing `nn.Sequential` as `net` to share its parameters."""

    def __init__(self, dim: Optional[int] = None, hidden: Union[int, Sequence[int]] = 128, activation: str = "silu",
                 net: Optional[nn.Sequential] = None, precision: str = "bf16x3", *args, **kwargs):
        super().__init__(*args, **kwargs)
        if precision not in _lib.MLP_PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.MLP_PRECISIONS)}")
        #: arithmetic of the fused Langevin products: "bf16x3" (tensor cores, split operands, ~2e-5 relative),
        #: "fp32" (CUDA cores) or "bf16" (tensor cores, single pass, ~4e-3 relative)
        self.precision = precision
        #: SMs the persistent burst kernels leave free; set >= 1 when a collective runs next to the bursts on another stream
        self.sm_margin = 0
        if net is None:
            if dim is None:
                raise ValueError("dim must be given when net is None")
            widths = (hidden, hidden) if isinstance(hidden, int) else tuple(hidden)
            if len(widths) not in (2, 3):
                raise ValueError("hidden must be an int or two or three widths")
            act = _ACT_BY_NAME[activation]
            layers, prev = [], dim
            for w in widths:
                layers += [nn.Linear(prev, w), act()]
                prev = w
            net = nn.Sequential(*layers, nn.Linear(prev, 1))
        if _match_mlp(net) is None:
            raise ValueError("net must be Sequential(Linear, act, Linear, act, [Linear, act,] Linear(., 1)) with "
                             "SiLU/Tanh/ReLU/Softplus")
        self.net = net

    def forward(self, x):
        return self.net(x).squeeze(-1)


def _match_mlp(net) -> Optional[Tuple[List[nn.Linear], int]]:
    """(the Linear layers in order -- two or three hidden ones and the scalar output layer --, activation code) of a
    `Sequential(Linear, act, Linear, act, [Linear, act,] Linear(., 1))`, or None."""
    if not isinstance(net, nn.Sequential) or len(net) not in (5, 7):
        return None
    lins, acts = list(net)[0::2], list(net)[1::2]
    if not all(isinstance(l, nn.Linear) for l in lins):
        return None
    if any(type(a) is not type(acts[0]) for a in acts) or type(acts[0]) not in _ACT_CODES:
        return None
    if any(isinstance(a, nn.Softplus) and (a.beta != 1.0 or a.threshold != 20.0) for a in acts):
        return None
    if lins[-1].out_features != 1 or any(a.out_features != b.in_features for a, b in zip(lins[:-1], lins[1:])):
        return None
    if any(l.bias is None for l in lins):
        return None
    return lins, _ACT_CODES[type(acts[0])]


# --------------------------------------------------------------------------------------------------
# descriptor extraction

MLP_MAX_WIDTH = 128  # kMlpMax of csrc/ebm_mlp.cu: hidden widths, and the state width of the on-chip kernels
MLP_MAX_DIM = 4096   # kWdMaxDim of csrc/ebm_mlp_wide.cu: state width of the streamed-operand kernel

_WORKSPACES: dict = {}


def _mlp_workspace(device, nbytes: int) -> torch.Tensor:
    """Per (device, stream) scratch for the wide-MLP burst: descriptors are rebuilt at every `sample()` call, the
    workspace is not.  Bursts on one stream are ordered, so sharing it among them is safe."""
    device = torch.device(device)
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _WORKSPACES[key] = ws
    return ws


class EnergyDescriptor:
    """An `EbmEnergyDesc` plus the tensors that keep its device pointers alive."""

    def __init__(self, cdesc: _lib.EbmEnergyDesc, keep: Sequence[torch.Tensor], kind: str):
        self.c = cdesc
        self.keep = list(keep)
        self.kind = kind

    @property
    def dim(self) -> int:
        return self.c.dim


def _f32(v: float) -> float:
    return float(torch.tensor(v, dtype=torch.float32))


def _dev_f32(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def energy_descriptor(model: nn.Module, dim: int, device) -> Optional[EnergyDescriptor]:
    """Map a model onto a library energy, or None when it has to stay an opaque PyTorch module.

    Matching is by exact class name (this package's classes and the reference's share names) and requires
    that `forward` is the library one, not a user override of a subclass."""
    name = type(model).__name__
    mod = type(model).__module__
    known_module = mod.startswith("torchebm_b200") or mod.startswith("torchebm.")
    d = _lib.EbmEnergyDesc()
    d.dim = int(dim)
    if known_module and name == "DoubleWellModel":
        d.kind = _lib.ENERGY_DOUBLE_WELL
        d.p[0] = float(model.barrier_height)
        d.p[1] = float(model.b) ** 2
        return EnergyDescriptor(d, [], "double_well")
    if known_module and name == "HarmonicModel":
        d.kind = _lib.ENERGY_HARMONIC
        d.p[0] = 0.5 * float(model.k)
        return EnergyDescriptor(d, [], "harmonic")
    if known_module and name == "RastriginModel":
        d.kind = _lib.ENERGY_RASTRIGIN
        d.p[0] = float(model.a)
        d.p[1] = 2 * math.pi
        d.p[2] = float(model.a) * dim
        return EnergyDescriptor(d, [], "rastrigin")
    if known_module and name == "GaussianModel":
        if model.mean.shape[0] != dim:
            return None
        mean, cinv = _dev_f32(model.mean, device), _dev_f32(model.cov_inv, device)
        d.kind = _lib.ENERGY_GAUSSIAN
        d.buf[0], d.buf[1] = mean.data_ptr(), cinv.data_ptr()
        return EnergyDescriptor(d, [mean, cinv], "gaussian")
    if known_module and name == "MixtureOfGaussiansModel":
        if model.means.shape[1] != dim:
            return None
        mu, sg, w = (_dev_f32(t, device) for t in (model.means, model.sigmas, model.weights))
        d.kind = _lib.ENERGY_MOG
        d.n_components = mu.shape[0]
        d.buf[0], d.buf[1], d.buf[2] = mu.data_ptr(), sg.data_ptr(), w.data_ptr()
        return EnergyDescriptor(d, [mu, sg, w], "mog")
    net = None
    if isinstance(model, MLPEnergy):
        net = model.net
    elif getattr(model, "_ebm_b200_mlp", None) is not None:
        net = model._ebm_b200_mlp  # set by mark_mlp_energy() after a numerical check
    if net is not None:
        m = _match_mlp(net)
        if m is None:
            return None
        lins, act = m
        l1, l2, l_out = lins[0], lins[1], lins[-1]
        deep = len(lins) == 4   # three hidden layers: on-chip tensor-core kernel (csrc/ebm_mlp_deep.cu), every width <= 128
        if l1.in_features != dim:
            return None
        if max(l.out_features for l in lins[:-1]) > MLP_MAX_WIDTH or l1.in_features > (MLP_MAX_WIDTH if deep else MLP_MAX_DIM):
            return None  # no fused kernel: integrator-level path (own autograd + fused update)
        ts = [_dev_f32(t, device) for t in (l1.weight, l1.bias, l2.weight, l2.bias, l_out.weight.reshape(-1), l_out.bias)]
        d.kind = _lib.ENERGY_MLP
        d.hidden1, d.hidden2, d.activation = l1.out_features, l2.out_features, act
        precision = getattr(model, "precision", "bf16x3")
        if (l1.in_features > MLP_MAX_WIDTH or deep) and precision == "fp32":
            precision = "bf16x3"  # wide states / three hidden layers have a tensor-core kernel only (same accuracy class as fp32)
        d.precision = _lib.MLP_PRECISIONS[precision]
        d.sm_margin = int(getattr(model, "sm_margin", 0))
        for i, t in enumerate(ts):
            d.buf[i] = t.data_ptr()
        if deep:
            w3, b3 = _dev_f32(lins[2].weight, device), _dev_f32(lins[2].bias, device)
            d.hidden3 = lins[2].out_features
            d.buf[7], d.buf[8] = w3.data_ptr(), b3.data_ptr()
            ts += [w3, b3]
        ws_bytes = int(_lib.load().ebm_workspace_bytes(C.byref(d)))
        if ws_bytes > 0 and torch.device(device).type == "cuda":  # hand-over flags + (wide states) the bf16 weight re-split
            ws = _mlp_workspace(device, ws_bytes)
            d.buf[6] = ws.data_ptr()
            ts.append(ws)
        return EnergyDescriptor(d, ts, "mlp")
    return None


def mark_mlp_energy(model: nn.Module, probe: Optional[torch.Tensor] = None) -> bool:
    """Recognise a user-defined MLP energy (any nn.Module whose only parametrised child is a matching
    `nn.Sequential` and whose forward is `net(x).squeeze(-1)`) so that samplers take the fused path.

    The structural match is confirmed numerically on CPU autograd-free forward: `model(probe)` must equal
    the Sequential's own output.  Returns True when the model was marked."""
    seqs = [m for m in model.children() if isinstance(m, nn.Sequential)]
    if len(seqs) != 1 or _match_mlp(seqs[0]) is None:
        return False
    own_params = sum(p.numel() for p in model.parameters())
    if own_params != sum(p.numel() for p in seqs[0].parameters()):
        return False
    net = seqs[0]
    if probe is None:   # deterministic probe: drawing one would advance the global generator of a seeded script
        l1 = net[0]
        probe = torch.linspace(-2.0, 2.0, 8 * l1.in_features, device=l1.weight.device,
                               dtype=l1.weight.dtype).reshape(8, l1.in_features)
    with torch.no_grad():
        a = model(probe)
        b = net(probe).squeeze(-1)
    if a.shape != b.shape or not torch.equal(a, b):
        return False
    object.__setattr__(model, "_ebm_b200_mlp", net)
    return True
