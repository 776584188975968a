"""Golden vectors for the three-hidden-layer MLP energy (benchmarks/distributed_fsdp2.py:43-53), produced by the UNMODIFIED
reference imported from /root/reference:

    python tests/golden/make_golden_deep.py

`langevin_mlp_deep.npz`: torchebm.samplers.LangevinDynamics on a 24-48-40-32-1 Tanh energy (a BaseModel subclass whose
gradient is the reference's autograd, core/base_model.py:84-127), K = 8 steps with the noise pre-drawn from the recorded
CPU generator seed, plus energy and gradient at x0.  Same layout as the cases of make_golden.py."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import _import_reference, predraw_langevin, save  # noqa: E402


def main():
    _import_reference()
    from torchebm.core import BaseModel
    from torchebm.samplers import LangevinDynamics

    torch.manual_seed(4321)

    class DeepMLPEnergy(BaseModel):
        def __init__(self, d, widths, act):
            super().__init__()
            layers, prev = [], d
            for w in widths:
                layers += [torch.nn.Linear(prev, w), act()]
                prev = w
            self.net = torch.nn.Sequential(*layers, torch.nn.Linear(prev, 1))

        def forward(self, x):
            return self.net(x).squeeze(-1)

    n, d, k, h, ns, seed = 72, 24, 8, 0.01, 1.0, 41
    model = DeepMLPEnergy(d, (48, 40, 32), torch.nn.Tanh)
    x0 = torch.randn(n, d, generator=torch.Generator().manual_seed(seed + 1000))
    noise = predraw_langevin(x0, k, seed)
    out = LangevinDynamics(model, step_size=h, noise_scale=ns).sample(x=x0, n_steps=k, generator=torch.Generator().manual_seed(seed))
    arrs = dict(x0=x0, noise=noise, k=k, h=h, ns=ns, out=out, grad0=model.gradient(x0), energy0=model(x0).detach())
    for i, l in enumerate(m for m in model.net if isinstance(m, torch.nn.Linear)):
        arrs[f"w{i}"] = l.weight.detach()
        arrs[f"b{i}"] = l.bias.detach()
    save("langevin_mlp_deep", **arrs)


if __name__ == "__main__":
    main()
