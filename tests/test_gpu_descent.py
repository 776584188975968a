"""GradientDescentSampler / NesterovSampler: fused descent bursts vs the reference goldens.

Bars: DoubleWell / Harmonic bit-exact (closed-form gradient and ATen's single-rounding alpha updates reproduced with
fmaf); Rastrigin atol 2e-6 (device sinf); MLP (step-by-step path with the library's gradient kernel) atol 2e-5."""

import pytest
import torch

from oracle import descent as odesc
from oracle import energies as E

from . import _cases as C

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _sampler_for(name, g, hs, mu):
    import torchebm_b200 as te

    if name == "gd_mlp":
        w = C.load("langevin_mlp_d784")
        model = te.MLPEnergy(dim=784, hidden=64, activation="silu")
        lin = [l for l in model.net if isinstance(l, torch.nn.Linear)]
        with torch.no_grad():
            for i, l in enumerate(lin):
                l.weight.copy_(w[f"w{i}"])
                l.bias.copy_(w[f"b{i}"])
        model = model.to(DEV)
    elif "doublewell" in name:
        model = te.DoubleWellModel(2.0, 1.0)
    elif "rastrigin" in name:
        model = te.RastriginModel(g["a"])
    else:
        model = te.HarmonicModel(g["kspring"])
    if isinstance(hs, list):   # the schedule the golden recorded
        step = te.ExponentialDecayScheduler(start_value=0.005, decay_rate=0.97, min_value=0.0005)
    else:
        step = hs
    if mu is None:
        return te.GradientDescentSampler(model, step_size=step, device=DEV)
    return te.NesterovSampler(model, step_size=step, momentum=mu, device=DEV)


@pytest.mark.parametrize("name", C.DESCENT_CASES)
def test_descent_samplers_match_reference_golden(name):
    g = C.load(name)
    en, hs, mu, kw = C.descent_setup(name, g)
    s = _sampler_for(name, g, hs, mu)
    res = s.sample(x=g["x0"].to(DEV), n_steps=int(g["k"]), **kw)
    out = (res[0] if isinstance(res, tuple) else res).cpu()
    if "rastrigin" in name:
        torch.testing.assert_close(out, g["out"], rtol=1e-5, atol=2e-6)
    elif name == "gd_mlp":
        torch.testing.assert_close(out, g["out"], rtol=1e-4, atol=2e-5)
    else:
        assert torch.equal(out, g["out"])
    if isinstance(res, tuple):
        torch.testing.assert_close(res[1]["energy"].cpu(), g["diag_energy"], rtol=1e-5, atol=1e-5)
    if "sched" in name:  # schedulers were advanced once per step, like the reference
        assert s.schedulers["step_size"].step_count == int(g["k"])


@pytest.mark.parametrize("mu", [None, 0.8])
def test_descent_burst_large_ragged_trajectory_equals_oracle_on_cuda(mu):
    """Many blocks, numel not a multiple of 4, thinned trajectory from the fused kernel vs the oracle's torch ops."""
    import torchebm_b200 as te

    x0 = torch.randn(5001, 77, device=DEV).clamp_(-2, 2)
    model = te.DoubleWellModel(2.0, 1.0)
    s = te.GradientDescentSampler(model, step_size=0.004, device=DEV) if mu is None else \
        te.NesterovSampler(model, step_size=0.004, momentum=mu, device=DEV)
    got = s.sample(x=x0, n_steps=13, thin=4, return_trajectory=True)
    want = odesc.sample(E.DoubleWell(2.0, 1.0), x0, 13, 0.004, mu, thin=4, return_trajectory=True, closed_form=True)
    assert got.shape == (5001, 3, 77) and torch.equal(got, want)
    with pytest.raises(ValueError, match="thin must be >= 1"):
        s.sample(x=x0, thin=0)
    if mu is not None:
        with pytest.raises(ValueError, match="momentum"):
            te.NesterovSampler(model, momentum=1.0)
