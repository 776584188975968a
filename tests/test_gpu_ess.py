"""On-device effective sample size (ebm_ess_f32) against the reference's `_ess_from_chain` goldens
(benchmarks/registry.py:348-365; fixtures from tests/golden/make_golden_ess.py) and the oracle."""

import os

import numpy as np
import pytest
import torch

from oracle import ess as oess

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ess_chains.npz")


def _cases():
    z = np.load(GOLDEN)
    return {k[6:]: (torch.from_numpy(z[k]), float(z["ess_" + k[6:]])) for k in z.files if k.startswith("chain_")}


@pytest.mark.parametrize("name", sorted(_cases()))
def test_ess_matches_reference_golden(name):
    from torchebm_b200 import ops

    chain, want = _cases()[name]
    got = ops.ess(chain.to(DEV)).item()
    # direct fp64 autocovariances vs the reference's fp32 FFT: a few 1e-7 relative on these chains; 1e-4 leaves room for
    # a lag whose autocorrelation is within rounding of zero
    assert got == pytest.approx(want, rel=1e-4), (name, got, want)


def test_ess_batch_long_chain_and_diagnostics():
    import torchebm_b200 as te
    from torchebm_b200 import ops

    gen = torch.Generator().manual_seed(3)
    # a batch of chains in one launch, ragged behaviour per row (white noise, strongly correlated, constant)
    n = 1500
    rows = [torch.randn(n, generator=gen), torch.cumsum(torch.randn(n, generator=gen), 0) * 0.05, torch.full((n,), -2.0)]
    rows[1] = rows[1] - torch.linspace(0, 1, n) * rows[1][-1]
    batch = torch.stack(rows)
    got = ops.ess(batch.to(DEV)).cpu()
    want = torch.tensor([oess.ess_from_chain(r) for r in rows])
    torch.testing.assert_close(got, want.float(), rtol=1e-4, atol=1e-4)
    # longer than the shared-memory staging limit: the kernel reads the chain from global memory
    long_chain = torch.randn(70000, generator=gen) + torch.sin(torch.arange(70000) / 50.0)
    got_long = ops.ess(long_chain.to(DEV)).item()
    assert got_long == pytest.approx(oess.ess_direct(long_chain.numpy()), rel=1e-4)
    # the harness's use: energy chain of a sampler's diagnostics (registry.py:766-770)
    s = te.LangevinDynamics(te.DoubleWellModel(2.0, 1.0), step_size=0.01, noise_scale=1.0, device=DEV)
    x0 = torch.randn(64, 4, generator=gen).clamp_(-2, 2).to(DEV)
    _, diag = s.sample(x=x0, n_steps=200, return_diagnostics=True, generator=torch.Generator(DEV).manual_seed(5))
    chain = diag["energy"].detach().cpu().float()
    assert chain.shape == (200,) and torch.isfinite(chain).all()
    assert te.ess_from_diagnostics(diag).item() == pytest.approx(oess.ess_from_chain(chain), rel=1e-4)
    stacked = torch.zeros(200, 3, 1, 1, device=DEV)   # the harness's stacked layout
    stacked[:, 2, 0, 0] = diag["energy"]
    assert te.ess_from_diagnostics(stacked).item() == pytest.approx(oess.ess_from_chain(chain), rel=1e-4)
    with pytest.raises(ValueError):
        ops.ess(torch.empty(0, device=DEV))
