"""Pin the oracle against the golden vectors produced by the unmodified reference.

Bars: bit-exact (`torch.equal`) wherever the oracle runs the same torch ops as the reference
(autograd gradient, same loop); bit-exact for the closed-form gradients of the elementwise
energies (what the CUDA kernels implement); stated tolerances for the closed forms that reorder
a reduction (Gaussian, MoG, MLP).
"""

import os

import numpy as np
import pytest
import torch

from oracle import energies as E
from oracle import hmc as ohmc
from oracle import langevin as olang
from oracle import leapfrog as olf
from oracle import pcd as opcd

from . import _cases as C


@pytest.mark.parametrize("name", C.LANGEVIN_CASES)
def test_energy_and_gradient_match_reference(name):
    g = C.load(name)
    en = C.energy_for(name, g)
    assert torch.equal(en.energy(g["x0"]), g["energy0"])
    assert torch.equal(en.gradient(g["x0"]), g["grad0"])
    gc = en.gradient_closed(g["x0"])
    if name in C.LANGEVIN_BITEXACT_CLOSED:
        assert torch.equal(gc, g["grad0"])
    else:
        # reduction order differs from bmm / addmm / logsumexp backward
        torch.testing.assert_close(gc, g["grad0"], rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("name", C.LANGEVIN_CASES)
def test_langevin_autograd_path_is_bit_exact(name):
    g = C.load(name)
    en = C.energy_for(name, g)
    h, ns = C.langevin_schedule(name, g)
    kw = C.langevin_kwargs(name, g)
    res = olang.sample(en, g["x0"], int(g["k"]), h, ns, noise=C.langevin_noise(g), **kw)
    if isinstance(res, tuple):
        out, diag = res
        for key, val in diag.items():
            assert torch.equal(val, g["diag_" + key]), key
    else:
        out = res
    assert torch.equal(out, g["out"])


@pytest.mark.parametrize("name", sorted(C.LANGEVIN_BITEXACT_CLOSED))
def test_langevin_closed_form_is_bit_exact(name):
    g = C.load(name)
    en = C.energy_for(name, g)
    h, ns = C.langevin_schedule(name, g)
    kw = C.langevin_kwargs(name, g)
    res = olang.sample(en, g["x0"], int(g["k"]), h, ns, noise=C.langevin_noise(g), closed_form=True, **kw)
    out = res[0] if isinstance(res, tuple) else res
    assert torch.equal(out, g["out"])


def test_langevin_generator_path_draws_in_reference_order():
    g = C.load("langevin_doublewell")
    en = C.energy_for("langevin_doublewell", g)
    out = olang.sample(en, g["x0"], int(g["k"]), g["h"], g["ns"], generator=torch.Generator().manual_seed(1))
    assert torch.equal(out, g["out"])


@pytest.mark.parametrize("name", C.LEAPFROG_CASES)
def test_leapfrog_matches_reference(name):
    g = C.load(name)
    en = E.DoubleWell(2.0, 1.0)
    x, p = olf.integrate(lambda x_: -en.gradient(x_), g["x0"], g["p0"], float(g["h"]), int(g["L"]),
                         mass=C.mass_of(g), safe=bool(g["safe"]))
    assert torch.equal(x, g["x"]) and torch.equal(p, g["p"])
    x, p = olf.integrate(lambda x_: -en.gradient_closed(x_), g["x0"], g["p0"], float(g["h"]), int(g["L"]),
                         mass=C.mass_of(g), safe=bool(g["safe"]))
    assert torch.equal(x, g["x"], ) and torch.equal(p, g["p"])


@pytest.mark.parametrize("name", C.HMC_CASES)
def test_hmc_matches_reference(name):
    g = C.load(name)
    en = C.energy_for(name, g)
    kw = {}
    if name == "hmc_rastrigin_diag":
        kw = dict(thin=2, return_trajectory=True, return_diagnostics=True)
    res = ohmc.sample(en, g["x0"], int(g["k"]), float(g["h"]), int(g["L"]), mass=C.mass_of(g),
                      noise_p=g["noise_p"], noise_u=g["noise_u"], **kw)
    if isinstance(res, tuple):
        out, diag = res
        for key, val in diag.items():
            assert torch.equal(val, g["diag_" + key]), key
    else:
        out = res
    assert torch.equal(out, g["out"])


def test_hmc_generator_path_draws_in_reference_order():
    g = C.load("hmc_doublewell")
    en = C.energy_for("hmc_doublewell", g)
    out = ohmc.sample(en, g["x0"], int(g["k"]), float(g["h"]), int(g["L"]), generator=torch.Generator().manual_seed(21))
    assert torch.equal(out, g["out"])


def _cd_loss(en, x, neg, reg):
    # torchebm/losses/contrastive_divergence.py:184-223
    xe, ne = en.energy(x), en.energy(neg)
    loss = xe.mean() - ne.mean()
    if reg > 0:
        loss = loss + reg * (torch.mean(xe**2) + torch.mean(ne**2))
    return loss


def test_pcd_sequence_matches_reference():
    g = C.load("pcd_mlp_tanh")
    en = C.mlp_from(g, "tanh")
    gen = torch.Generator().manual_seed(int(g["seed"]))
    buf = opcd.ReplayBuffer(40, new_sample_ratio=0.25)
    for it in range(4):
        x = g["data"][it]
        if buf.buffer is None:
            buf.initialize((6,), "cpu", generator=gen)
        start = buf.get_start_points(16, generator=gen)
        neg = olang.sample(en, start, 3, 0.01, 1.0, generator=gen)
        buf.update(neg)
        assert torch.equal(neg, g["negs"][it])
        assert torch.equal(buf.buffer, g["bufs"][it])
        assert buf.ptr == int(g["ptrs"][it])
        assert torch.equal(_cd_loss(en, x, neg, 0.001), g["losses"][it])


def test_pcd_fifo_wraparound_matches_reference():
    g = C.load("pcd_doublewell_fifo")
    en = E.DoubleWell(2.0, 1.0)
    gen = torch.Generator().manual_seed(int(g["seed"]))
    buf = opcd.ReplayBuffer(50, new_sample_ratio=0.0)
    for it in range(5):
        if buf.buffer is None:
            buf.initialize((3,), "cpu", generator=gen)
        start = buf.get_start_points(16, generator=gen)
        neg = olang.sample(en, start, 2, 0.01, 1.0, generator=gen)
        buf.update(neg)
        assert torch.equal(neg, g["negs"][it])
        assert torch.equal(buf.buffer, g["bufs"][it])
        assert buf.ptr == int(g["ptrs"][it])
    assert list(g["ptrs"].tolist()) == [16, 32, 48, 14, 30]


@pytest.mark.parametrize("name", C.DESCENT_CASES)
def test_descent_oracle_matches_reference_golden(name):
    """GradientDescentSampler / NesterovSampler restatement vs the unmodified reference: bit-exact (autograd gradient;
    closed forms bit-exact too for the elementwise energies)."""
    from oracle import descent as odesc

    g = C.load(name)
    en, hs, mu, kw = C.descent_setup(name, g)
    res = odesc.sample(en, g["x0"], int(g["k"]), hs, mu, **kw)
    out = res[0] if isinstance(res, tuple) else res
    assert torch.equal(out, g["out"])
    if isinstance(res, tuple):
        assert torch.equal(res[1]["energy"], g["diag_energy"])
    if name != "gd_mlp":
        res2 = odesc.sample(en, g["x0"], int(g["k"]), hs, mu, closed_form=True, **kw)
        out2 = res2[0] if isinstance(res2, tuple) else res2
        assert torch.equal(out2, g["out"])


@pytest.mark.parametrize("name", C.HEUN_CASES)
def test_heun_oracle_matches_reference_golden(name):
    """`LangevinDynamics(integrator="heun")` restatement vs the unmodified reference, injected noise: bit-exact with the
    autograd gradient and with the closed forms."""
    g = C.load(name)
    en, kw = C.heun_setup(name)
    for closed in (False, True):
        out = olang.sample(en, g["x0"], int(g["k"]), float(g["h"]), float(g["ns"]), noise=g["noise"], scheme="heun",
                           closed_form=closed, **kw)
        assert torch.equal(out, g["out"])


def test_ess_oracle_matches_reference_golden():
    """oracle/ess.py restates benchmarks/registry.py:348-365; the goldens come from the unmodified function."""
    import numpy as np

    from oracle import ess as oess

    z = np.load(os.path.join(C.GOLDEN, "ess_chains.npz"))
    names = [k[6:] for k in z.files if k.startswith("chain_")]
    assert len(names) >= 10
    for name in names:
        chain = torch.from_numpy(z["chain_" + name])
        assert oess.ess_from_chain(chain) == float(z["ess_" + name]), name
        assert oess.ess_direct(z["chain_" + name]) == pytest.approx(float(z["ess_" + name]), rel=1e-5), name
