"""Oracle energies and gradients.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

`Energy.energy` restates the reference `forward` op for op; `Energy.gradient`
restates `BaseModel.gradient` (torchebm/core/base_model.py:62-127: detach ->
requires_grad -> forward under enable_grad -> autograd.grad with ones -> detach).
`Energy.gradient_closed` is the closed form the CUDA kernels implement, written
in the reference's rounding order; the tests assert the
two agree (bit-exact on CPU for the elementwise energies).
"""

from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence

import torch


class Energy:
    """A per-row scalar energy E(x): [N, D] -> [N]."""

    kind = "base"

    def energy(self, x: torch.Tensor) -> torch.Tensor:  # pragma: no cover
        raise NotImplementedError

    def gradient(self, x: torch.Tensor) -> torch.Tensor:
        # torchebm/core/base_model.py:84-127
        with torch.enable_grad():
            xg = x.detach().requires_grad_(True)
            e = self.energy(xg)
            if e.shape != (xg.shape[0],):
                raise ValueError(f"energy expected shape ({xg.shape[0]},), got {tuple(e.shape)}")
            (g,) = torch.autograd.grad(e, xg, grad_outputs=torch.ones_like(e))
        return g.detach()

    def gradient_closed(self, x: torch.Tensor) -> torch.Tensor:
        return self.gradient(x)


class DoubleWell(Energy):
    """torchebm/core/base_model.py:130-148: h * sum((x^2 - b^2)^2)."""

    kind = "double_well"

    def __init__(self, barrier_height: float = 2.0, b: float = 1.0):
        self.barrier_height = float(barrier_height)
        self.b = float(b)

    def energy(self, x):
        return self.barrier_height * (x.pow(2) - self.b**2).pow(2).sum(dim=-1)

    def gradient_closed(self, x):
        # (h * (2 * (x^2 - b^2))) * (2 * x), each product rounded (A.1)
        u = x * x - self.b**2
        return (self.barrier_height * (2.0 * u)) * (2.0 * x)


class Harmonic(Energy):
    """torchebm/core/base_model.py:213-229: 0.5 * k * sum(x^2)."""

    kind = "harmonic"

    def __init__(self, k: float = 1.0):
        self.k = float(k)

    def energy(self, x):
        return 0.5 * self.k * x.pow(2).sum(dim=-1)

    def gradient_closed(self, x):
        return (0.5 * self.k) * (2.0 * x)


class Rastrigin(Energy):
    """torchebm/core/base_model.py:297-316: a*n + sum(x^2 - a*cos(2*pi*x))."""

    kind = "rastrigin"

    def __init__(self, a: float = 10.0):
        self.a = float(a)

    def energy(self, x):
        n = x.shape[-1]
        return self.a * n + torch.sum(x**2 - self.a * torch.cos(2 * math.pi * x), dim=-1)

    def gradient_closed(self, x):
        c = 2 * math.pi
        return 2.0 * x + (self.a * torch.sin(c * x)) * c


class Gaussian(Energy):
    """torchebm/core/base_model.py:151-210: 0.5 * delta^T cov_inv delta (two bmm when N > 1)."""

    kind = "gaussian"

    def __init__(self, mean: torch.Tensor, cov: torch.Tensor):
        self.mean = mean.to(torch.float32)
        self.cov_inv = torch.inverse(cov).to(torch.float32)

    def to(self, device):
        self.mean = self.mean.to(device)
        self.cov_inv = self.cov_inv.to(device)
        return self

    def energy(self, x):
        delta = x - self.mean
        cov_inv = self.cov_inv
        if delta.shape[0] > 1:
            temp = torch.bmm(cov_inv.unsqueeze(0).expand(delta.shape[0], -1, -1), delta.unsqueeze(-1))
            return 0.5 * torch.bmm(delta.unsqueeze(1), temp).squeeze(-1).squeeze(-1)
        return 0.5 * torch.sum(delta * torch.matmul(delta, cov_inv), dim=-1)

    def gradient_closed(self, x):
        delta = x - self.mean
        sym = 0.5 * (self.cov_inv + self.cov_inv.t())
        return delta @ sym


class MixtureOfGaussians(Energy):
    """Isotropic Gaussian mixture.  NOT in the reference: `north_star`
    lists MoG, so the build defines it and the oracle is the reference's autograd `gradient`
    (base_model.py:84-127) applied to this forward.

    E(x) = -logsumexp_k( log w_k - D*log(sigma_k) - |x - mu_k|^2 / (2 sigma_k^2) )
    """

    kind = "mog"

    def __init__(self, means: torch.Tensor, sigmas: torch.Tensor, weights: Optional[torch.Tensor] = None):
        self.means = means.to(torch.float32)  # [K, D]
        self.sigmas = sigmas.to(torch.float32)  # [K]
        k = means.shape[0]
        if weights is None:
            weights = torch.full((k,), 1.0 / k)
        self.weights = weights.to(torch.float32)

    def to(self, device):
        self.means, self.sigmas, self.weights = (t.to(device) for t in (self.means, self.sigmas, self.weights))
        return self

    def logits(self, x):
        d = x.shape[-1]
        diff = x.unsqueeze(1) - self.means.unsqueeze(0)  # [N, K, D]
        sq = diff.pow(2).sum(dim=-1)  # [N, K]
        return torch.log(self.weights) - d * torch.log(self.sigmas) - sq / (2.0 * self.sigmas**2)

    def energy(self, x):
        return -torch.logsumexp(self.logits(x), dim=-1)

    def gradient_closed(self, x):
        r = torch.softmax(self.logits(x), dim=-1)  # [N, K]
        diff = x.unsqueeze(1) - self.means.unsqueeze(0)
        return (r.unsqueeze(-1) * diff / (self.sigmas**2).view(1, -1, 1)).sum(dim=1)


_ACTS = {
    "silu": torch.nn.functional.silu,
    "tanh": torch.tanh,
    "relu": torch.relu,
    "softplus": torch.nn.functional.softplus,
}


class MLP(Energy):
    """User MLP energies: Sequential(Linear(D,H1), act, Linear(H1,H2), act, Linear(H2,1)) + squeeze(-1)
    (examples/20-training/01-mcmc-losses/01-cd-k/main.py:20-30, benchmarks/registry.py:375-387).
    `weights` / `biases` are in torch `[out, in]` layout, one per Linear."""

    kind = "mlp"

    def __init__(self, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor], activation: str = "silu"):
        self.weights = [w.detach().to(torch.float32) for w in weights]
        self.biases = [b.detach().to(torch.float32) for b in biases]
        self.activation = activation

    def to(self, device):
        self.weights = [w.to(device) for w in self.weights]
        self.biases = [b.to(device) for b in self.biases]
        return self

    def energy(self, x):
        act = _ACTS[self.activation]
        h = x
        n = len(self.weights)
        for i, (w, b) in enumerate(zip(self.weights, self.biases)):
            h = torch.nn.functional.linear(h, w, b)
            if i < n - 1:
                h = act(h)
        return h.squeeze(-1)

    def gradient_closed(self, x):
        # grad = W1^T (s'(z1) * (W2^T (s'(z2) * w3)))
        zs = []
        h = x
        n = len(self.weights)
        for i in range(n - 1):
            z = torch.nn.functional.linear(h, self.weights[i], self.biases[i])
            zs.append(z)
            h = _ACTS[self.activation](z)
        delta = self.weights[-1].expand(x.shape[0], -1)  # [N, H_last]
        for i in range(n - 2, -1, -1):
            delta = delta * _act_prime(self.activation, zs[i])
            delta = delta @ self.weights[i]
        return delta


def _act_prime(name: str, z: torch.Tensor) -> torch.Tensor:
    if name == "silu":
        s = torch.sigmoid(z)
        return s * (1.0 + z * (1.0 - s))
    if name == "tanh":
        t = torch.tanh(z)
        return 1.0 - t * t
    if name == "relu":
        return (z > 0).to(z.dtype)
    if name == "softplus":
        return torch.sigmoid(z)
    raise ValueError(name)


def make_mlp(dim: int, hidden: Sequence[int] = (128, 128), activation: str = "silu", seed: int = 0) -> MLP:
    """Default `nn.Linear` init under `torch.manual_seed(seed)`."""
    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    sizes = [dim, *hidden, 1]
    layers = [torch.nn.Linear(sizes[i], sizes[i + 1]) for i in range(len(sizes) - 1)]
    torch.random.set_rng_state(gen_state)
    return MLP([l.weight for l in layers], [l.bias for l in layers], activation)
