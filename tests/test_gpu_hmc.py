"""Parity of the fused leapfrog / HMC kernels (through the C ABI) with the oracle and the reference goldens.

Bars (fp32): leapfrog with elementwise DoubleWell is bit-exact vs the golden (same rounding order, force
reuse is bit-equal).  HMC compares chains row by row: the row energy is a reduction whose
order differs from torch's, so H0 - H1 can differ in the last ulp and flip an accept decision that sits
within ~1e-6 of the uniform draw; the test therefore requires >= 99% of the rows to agree to 1e-5 and
every row to be either the oracle's accepted or rejected state of some proposal (checked through the
acceptance counts being within 1%)."""

import pytest
import torch

from oracle import energies as E
from oracle import hmc as ohmc
from oracle import leapfrog as olf

from . import _cases as C

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _model_for(name, g):
    import torchebm_b200 as te

    if "rastrigin" in name:
        return te.RastriginModel(g["a"])
    if "gaussian" in name:
        return te.GaussianModel(g["mean"], g["cov"]).to(DEV)
    if name == "hmc_mass_vec":
        return te.HarmonicModel(g["kspring"])
    return te.DoubleWellModel(2.0, 1.0)


def _mass(g):
    m = C.mass_of(g)
    return m.to(DEV) if torch.is_tensor(m) else m


def _rows_agree(got, want, tol=1e-5):
    return ((got - want).abs().max(dim=-1).values <= tol).float().mean().item()


@pytest.mark.parametrize("name", C.LEAPFROG_CASES)
def test_leapfrog_matches_reference_golden(name):
    import torchebm_b200 as te
    from torchebm_b200 import ops

    g = C.load(name)
    desc = te.energy_descriptor(te.DoubleWellModel(2.0, 1.0), g["x0"].shape[1], torch.device(DEV))
    x, p = ops.leapfrog(desc, g["x0"].to(DEV), g["p0"].to(DEV), float(g["h"]), int(g["L"]), mass=_mass(g),
                        safe=bool(g["safe"]))
    assert torch.equal(x.cpu(), g["x"])
    assert torch.equal(p.cpu(), g["p"])
    # integrator-level API with a tagged drift takes the same kernel
    integ = te.LeapfrogIntegrator(device=DEV)
    res = integ.integrate({"x": g["x0"].to(DEV), "p": g["p0"].to(DEV)}, step_size=float(g["h"]), n_steps=int(g["L"]),
                          mass=_mass(g), drift=te.energy_drift(te.DoubleWellModel(2.0, 1.0)), safe=bool(g["safe"]))
    assert torch.equal(res["x"].cpu(), g["x"]) and torch.equal(res["p"].cpu(), g["p"])
    # and with an opaque lambda drift (compatibility path)
    dw = te.DoubleWellModel(2.0, 1.0)
    res = integ.integrate({"x": g["x0"].to(DEV), "p": g["p0"].to(DEV)}, step_size=float(g["h"]), n_steps=int(g["L"]),
                          mass=_mass(g), drift=lambda x_, t_: -dw.gradient(x_), safe=bool(g["safe"]))
    if name == "leapfrog_mass_float":
        # eager CUDA divides a tensor by a Python scalar as a * (1/b) (ATen BinaryDivTrueKernel.cu), the CPU
        # reference (and the fused kernel) divide: 1-ulp differences on this torch-op path only
        torch.testing.assert_close(res["x"].cpu(), g["x"], rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(res["p"].cpu(), g["p"], rtol=1e-6, atol=1e-6)
    else:
        assert torch.equal(res["x"].cpu(), g["x"]) and torch.equal(res["p"].cpu(), g["p"])


def test_leapfrog_properties_full_width():
    """Reversibility and energy conservation on a harmonic oscillator at D=64, N=2^18 (C4 shape)."""
    import torchebm_b200 as te
    from torchebm_b200 import ops

    n, d = 262144, 64
    desc = te.energy_descriptor(te.HarmonicModel(1.0), d, torch.device(DEV))
    x0 = torch.randn(n, d, device=DEV)
    p0 = torch.randn(n, d, device=DEV)
    x1, p1 = ops.leapfrog(desc, x0, p0, 0.01, 50)
    xb, pb = ops.leapfrog(desc, x1, -p1, 0.01, 50)
    assert (xb - x0).abs().max() < 1e-4 and (pb + p0).abs().max() < 1e-4
    h0 = 0.5 * (x0**2 + p0**2).sum(-1)
    h1 = 0.5 * (x1**2 + p1**2).sum(-1)
    assert ((h1 - h0).abs() / h0).max() < 1e-3


@pytest.mark.parametrize("name", C.HMC_CASES)
def test_hmc_injected_noise_matches_reference_golden(name):
    import torchebm_b200 as te
    from torchebm_b200 import _lib, ops

    g = C.load(name)
    model = _model_for(name, g)
    x0 = g["x0"].to(DEV)
    k, L = int(g["k"]), int(g["L"])
    desc = te.energy_descriptor(model, x0.shape[1], x0.device)
    thin = 2 if name == "hmc_rastrigin_diag" else 1
    traj = torch.empty(x0.shape[0], k // thin, x0.shape[1], device=DEV) if name == "hmc_rastrigin_diag" else None
    acc = torch.zeros(k, dtype=torch.int32, device=DEV)
    out = ops.hmc_burst(desc, x0, k, L, [float(g["h"])], mass=_mass(g), rng_mode=_lib.RNG_INJECTED,
                        noise_p=g["noise_p"].to(DEV), noise_u=g["noise_u"].to(DEV), traj=traj, thin=thin,
                        accept_count=acc)
    got = (traj if traj is not None else out).cpu()
    assert torch.isfinite(got).all()
    tol = 1e-4 if "rastrigin" in name else 2e-5
    assert _rows_agree(got, g["out"], tol) >= 0.99, _rows_agree(got, g["out"], tol)
    if name == "hmc_rastrigin_diag":
        rate = acc.cpu().float()[1::2] / x0.shape[0]
        torch.testing.assert_close(rate, g["diag_acceptance_rate"], rtol=0, atol=0.021)


def test_hmc_sampler_diagnostics_shapes_and_values():
    import torchebm_b200 as te

    g = C.load("hmc_rastrigin_diag")
    s = te.HamiltonianMonteCarlo(te.RastriginModel(10.0), step_size=float(g["h"]), n_leapfrog_steps=int(g["L"]), device=DEV)
    x0 = g["x0"].to(DEV)
    out, diag = s.sample(x=x0, n_steps=9, thin=2, return_trajectory=True, return_diagnostics=True,
                         generator=torch.Generator(DEV).manual_seed(0))
    en = E.Rastrigin(10.0)
    want, wdiag = ohmc.sample(en, x0, 9, float(g["h"]), int(g["L"]), thin=2, return_trajectory=True,
                              return_diagnostics=True, generator=torch.Generator(DEV).manual_seed(0))
    assert out.shape == (50, 4, 5) and set(diag) == {"mean", "var", "energy", "acceptance_rate"}
    assert _rows_agree(out.reshape(-1, 5), want.reshape(-1, 5), 1e-4) >= 0.97
    torch.testing.assert_close(diag["acceptance_rate"], wdiag["acceptance_rate"], rtol=0, atol=0.05)
    torch.testing.assert_close(diag["mean"], wdiag["mean"], rtol=0, atol=0.05)


@pytest.mark.parametrize("model_name,n,d,L,k,mass", [
    ("doublewell", 4096, 8, 5, 6, None),
    ("doublewell", 3000, 33, 4, 3, 2.5),
    ("rastrigin", 8192, 64, 20, 2, None),
    ("harmonic", 2000, 6, 5, 5, "vec"),
    ("gaussian", 4096, 2, 6, 6, None),
])
def test_hmc_same_seed_matches_reference_stream_on_cuda(model_name, n, d, L, k, mass):
    """rng='torch': per proposal the kernel consumes `normal_`(N*D) then `rand`(N) from torch's Philox stream,
    like the reference; the oracle on CUDA with an equal-seeded generator must agree and end at the same offset."""
    import torchebm_b200 as te

    if model_name == "doublewell":
        model, en, h = te.DoubleWellModel(2.0, 1.0), E.DoubleWell(2.0, 1.0), 0.05
    elif model_name == "rastrigin":
        model, en, h = te.RastriginModel(10.0), E.Rastrigin(10.0), 0.01
    elif model_name == "harmonic":
        model, en, h = te.HarmonicModel(2.0), E.Harmonic(2.0), 0.1
    else:
        mean, cov = torch.tensor([1.0, -1.0]), torch.tensor([[1.0, 0.8], [0.8, 1.0]])
        model, en, h = te.GaussianModel(mean, cov).to(DEV), E.Gaussian(mean, cov).to(DEV), 0.2
    if mass == "vec":
        mass = (torch.rand(d, generator=torch.Generator().manual_seed(0)) + 0.5).to(DEV)
    s = te.HamiltonianMonteCarlo(model, step_size=h, n_leapfrog_steps=L, mass=mass, device=DEV)
    x0 = torch.randn(n, d, device=DEV, generator=torch.Generator(DEV).manual_seed(3))
    g1 = torch.Generator(DEV).manual_seed(17)
    g2 = torch.Generator(DEV).manual_seed(17)
    got = s.sample(x=x0, n_steps=k, generator=g1)
    want = ohmc.sample(en, x0, k, h, L, mass=mass, generator=g2)
    assert g1.get_offset() == g2.get_offset()
    tol = 2e-4 if model_name == "rastrigin" else 2e-5
    assert _rows_agree(got, want, tol) >= 0.99, _rows_agree(got, want, tol)


def test_hmc_safe_mode_survives_far_starts():
    """tests/samplers/test_hmc.py:835-896 of the reference: starts at 1e4 / 1e6 stay finite."""
    import torchebm_b200 as te

    s = te.HamiltonianMonteCarlo(te.DoubleWellModel(2.0, 1.0), step_size=0.05, n_leapfrog_steps=5, device=DEV)
    for scale in (1e4, 1e6):
        x0 = torch.randn(256, 4, device=DEV) * scale
        out = s.sample(x=x0, n_steps=5, generator=torch.Generator(DEV).manual_seed(0))
        assert torch.isfinite(out).all()


def test_hmc_statistical_gaussian_target():
    """tests/samplers/test_hmc.py:668-703 of the reference: HMC recovers mean and covariance of a Gaussian."""
    import torchebm_b200 as te

    mean, cov = torch.tensor([1.0, -1.0]), torch.tensor([[1.0, 0.8], [0.8, 1.0]])
    s = te.HamiltonianMonteCarlo(te.GaussianModel(mean, cov).to(DEV), step_size=0.2, n_leapfrog_steps=10, device=DEV)
    out, diag = s.sample(dim=2, n_samples=20000, n_steps=100, return_diagnostics=True, thin=100,
                         generator=torch.Generator(DEV).manual_seed(0))
    torch.testing.assert_close(out.mean(0).cpu(), mean, rtol=0.15, atol=0.05)
    torch.testing.assert_close(torch.cov(out.t()).cpu(), cov, rtol=0.15, atol=0.05)
    assert 0.5 < diag["acceptance_rate"][-1].item() <= 1.0


def test_hmc_full_size_c4_properties():
    """BASELINE config 4 at full size: Rastrigin D=64, N=262144, L=20: determinism, finiteness, acceptance."""
    import torchebm_b200 as te

    s = te.HamiltonianMonteCarlo(te.RastriginModel(10.0), step_size=0.01, n_leapfrog_steps=20, device=DEV)
    x0 = torch.randn(262144, 64, device=DEV, generator=torch.Generator(DEV).manual_seed(0))
    a = s.sample(x=x0, n_steps=3, generator=torch.Generator(DEV).manual_seed(1))
    b = s.sample(x=x0, n_steps=3, generator=torch.Generator(DEV).manual_seed(1))
    assert torch.equal(a, b) and torch.isfinite(a).all()
    moved = ((a - x0).abs().max(dim=1).values > 0).float().mean().item()
    assert 0.3 < moved <= 1.0


def test_hmc_errors():
    import torchebm_b200 as te

    with pytest.raises(ValueError, match="n_leapfrog_steps must be positive"):
        te.HamiltonianMonteCarlo(te.DoubleWellModel(), n_leapfrog_steps=0)
    s = te.HamiltonianMonteCarlo(te.DoubleWellModel(), device=DEV)
    with pytest.raises(ValueError, match="thin must be >= 1"):
        s.sample(dim=2, thin=0)
    with pytest.raises(ValueError, match="dim must be provided"):
        s.sample(n_steps=2)
    mean, cov = torch.zeros(3), torch.eye(3)
    sg = te.HamiltonianMonteCarlo(te.GaussianModel(mean, cov).to(DEV), step_size=0.1, device=DEV)
    assert sg.sample(n_samples=7, n_steps=2).shape == (7, 3)  # dim inferred from model.mean (hmc.py:209-217)


def test_hmc_integrator_level_path_for_mlp_and_custom_energies():
    """Energies without a fused HMC kernel (MLP energies, arbitrary nn.Modules) go proposal by proposal through the
    integrator with the reference's draw order; same seed on CUDA => same chains as the oracle (tolerance: the MLP
    gradient comes from the library's fused kernel instead of autograd)."""
    import torchebm_b200 as te

    torch.manual_seed(4)
    d = 12
    mlp = te.MLPEnergy(dim=d, hidden=(32, 24), activation="tanh", precision="fp32").to(DEV)
    lin = [l for l in mlp.net if isinstance(l, torch.nn.Linear)]
    en = E.MLP([l.weight for l in lin], [l.bias for l in lin], "tanh")
    x0 = torch.randn(500, d, device=DEV)
    for mass in (None, 2.5, torch.rand(d, device=DEV) + 0.5):
        s = te.HamiltonianMonteCarlo(mlp, step_size=0.05, n_leapfrog_steps=4, mass=mass, device=DEV)
        got, diag = s.sample(x=x0, n_steps=6, thin=2, return_trajectory=True, return_diagnostics=True,
                             generator=torch.Generator(DEV).manual_seed(8))
        want, wdiag = ohmc.sample(en, x0, 6, 0.05, 4, mass=mass, thin=2, return_trajectory=True, return_diagnostics=True,
                                  generator=torch.Generator(DEV).manual_seed(8))
        assert got.shape == (500, 3, d)
        bad = ((got - want).abs().amax(dim=(1, 2)) > 1e-4).float().mean().item()
        assert bad < 0.01, bad   # a borderline accept decision may flip on a last-ulp energy difference
        torch.testing.assert_close(diag["acceptance_rate"], wdiag["acceptance_rate"], atol=0.01, rtol=0)

    class Quartic(torch.nn.Module):  # not a library energy: its own forward + autograd gradient
        def forward(self, x):
            return 0.25 * (x ** 4).sum(-1) + 0.5 * (x[:, :-1] * x[:, 1:]).sum(-1)

    class QuarticE(E.Energy):
        def energy(self, x):
            return 0.25 * (x ** 4).sum(-1) + 0.5 * (x[:, :-1] * x[:, 1:]).sum(-1)

    s = te.HamiltonianMonteCarlo(Quartic(), step_size=0.05, n_leapfrog_steps=5, device=DEV)
    got = s.sample(x=x0, n_steps=4, generator=torch.Generator(DEV).manual_seed(9))
    want = ohmc.sample(QuarticE(), x0, 4, 0.05, 5, generator=torch.Generator(DEV).manual_seed(9))
    bad = ((got - want).abs().amax(dim=1) > 1e-5).float().mean().item()
    assert bad < 0.01, bad


@pytest.mark.parametrize("precision", ["bf16x3", "fp32"])
@pytest.mark.parametrize("act", ["tanh", "silu"])
@pytest.mark.parametrize("mass_kind", ["none", "scalar", "vector"])
def test_hmc_fused_mlp_kernel_matches_oracle(act, mass_kind, precision):
    """MLP energies up to 128 wide have fused HMC kernels -- "bf16x3": tcgen05 tensor cores (tile of 128 chains, momentum
    in TMEM, split operands), "fp32": CUDA-core FFMA (warp owns 8 chains): injected noise vs the CPU oracle and same seed
    vs the oracle on CUDA, trajectories + diagnostics; several tiles so the persistent loop and ragged tile run."""
    import torchebm_b200 as te
    from torchebm_b200 import _lib, ops

    torch.manual_seed(6)
    n, d, L, k = 1000, 20, 5, 6
    mlp = te.MLPEnergy(dim=d, hidden=(48, 32), activation=act, precision=precision).to(DEV)
    lin = [l for l in mlp.net if isinstance(l, torch.nn.Linear)]
    en_cpu = E.MLP([l.weight.cpu() for l in lin], [l.bias.cpu() for l in lin], act)
    en_gpu = E.MLP([l.weight for l in lin], [l.bias for l in lin], act)
    mass = {"none": None, "scalar": 2.5, "vector": torch.rand(d) + 0.5}[mass_kind]
    x0 = torch.randn(n, d)
    noise_p, noise_u = torch.randn(k, n, d), torch.rand(k, n)
    desc = te.energy_descriptor(mlp, d, torch.device(DEV))
    acc = torch.zeros(k, dtype=torch.int32, device=DEV)
    e_out = torch.empty(n, device=DEV)
    got = ops.hmc_burst(desc, x0.to(DEV), k, L, [0.05], mass=mass if not torch.is_tensor(mass) else mass.to(DEV),
                        rng_mode=_lib.RNG_INJECTED, noise_p=noise_p.to(DEV), noise_u=noise_u.to(DEV), accept_count=acc,
                        energy_out=e_out)
    want, wdiag = ohmc.sample(en_cpu, x0, k, 0.05, L, mass=mass, noise_p=noise_p, noise_u=noise_u, return_diagnostics=True)
    bad = ((got.cpu() - want).abs().amax(dim=1) > 1e-4).float().mean().item()
    assert bad < 0.01, bad   # a borderline accept decision may flip on a last-ulp energy difference
    torch.testing.assert_close(acc.float().cpu() / n, wdiag["acceptance_rate"], atol=0.01, rtol=0)
    rows_ok = (got.cpu() - want).abs().amax(dim=1) <= 1e-4
    torch.testing.assert_close(e_out.cpu()[rows_ok], en_cpu.energy(want)[rows_ok], rtol=1e-4, atol=1e-4)
    # sampler API, torch RNG stream, trajectory + diagnostics
    s = te.HamiltonianMonteCarlo(mlp, step_size=0.05, n_leapfrog_steps=L, mass=mass if not torch.is_tensor(mass) else mass.to(DEV),
                                 device=DEV)
    g1, g2 = torch.Generator(DEV).manual_seed(3), torch.Generator(DEV).manual_seed(3)
    tr, diag = s.sample(x=x0.to(DEV), n_steps=k, thin=2, return_trajectory=True, return_diagnostics=True, generator=g1)
    wt, wd = ohmc.sample(en_gpu, x0.to(DEV), k, 0.05, L, mass=mass if not torch.is_tensor(mass) else mass.to(DEV), thin=2,
                         return_trajectory=True, return_diagnostics=True, generator=g2)
    assert tr.shape == (n, 3, d) and g1.get_offset() == g2.get_offset()
    bad = ((tr - wt).abs().amax(dim=(1, 2)) > 1e-4).float().mean().item()
    assert bad < 0.01, bad
    torch.testing.assert_close(diag["acceptance_rate"], wd["acceptance_rate"], atol=0.01, rtol=0)
    torch.testing.assert_close(diag["energy"], wd["energy"], atol=1e-2, rtol=1e-3)


def test_hmc_tc_balanced_proposal_split_is_bit_identical():
    """With a workspace the tensor-core HMC kernel hands every CTA an equal range of (tile, proposal) pairs, so a tile's
    proposals may be split between two SMs (only the chain state crosses the split); without one every tile runs whole.
    Counter-based draws => identical chains, acceptance counts, energies and trajectories."""
    import torchebm_b200 as te
    from torchebm_b200 import _lib, ops

    torch.manual_seed(1)
    n, d, L, k = 128 * 190 + 9, 32, 3, 7    # 191 tiles on 148 SMs
    mlp = te.MLPEnergy(dim=d, hidden=(64, 48), activation="silu").to(DEV)
    x0 = torch.randn(n, d, device=DEV)
    outs = []
    for with_ws in (True, False):
        desc = te.energy_descriptor(mlp, d, x0.device)
        assert desc.c.buf[6]
        if not with_ws:
            desc.c.buf[6] = None
        acc = torch.zeros(k, dtype=torch.int32, device=DEV)
        e_out = torch.empty(n, device=DEV)
        traj = torch.empty(n, k // 2, d, device=DEV)
        out = ops.hmc_burst(desc, x0, k, L, [0.05], rng_mode=_lib.RNG_TORCH, seed=11, offset=8, traj=traj, thin=2,
                            accept_count=acc, energy_out=e_out)
        outs.append((out, acc, e_out, traj))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    acc = outs[0][1]
    assert bool((acc <= n).all()) and int(acc.sum()) > 0.3 * n * k


@pytest.mark.parametrize("precision", ["bf16x3", "fp32"])
def test_hmc_mlp_force_is_recomputed_after_sanitising(precision):
    """leapfrog.py:160-185: the reference evaluates the drift at the top of EVERY step, so after nan_to_num_ rewrote a
    coordinate of x the next half-kick uses the force at the sanitised state, and hmc.py:266 takes E(x') there.  The
    tensor-core kernel carries the force across steps; a sanitising event makes the tile redo the proposal without
    carrying.  Rows 3 and 200 get a momentum that overflows two coordinates to +inf / -inf in the first drift: hidden
    units whose two weights share a sign see inf - inf, the whole force is NaN, the bottom half sanitises (p -> 0,
    x -> +-FLT_MAX in the two coordinates), and from there the reference continues with finite forces and ACCEPTS
    (K0 = inf clamps to 1e10).  With a stale NaN force the row would end NaN and be rejected."""
    import torchebm_b200 as te
    from torchebm_b200 import _lib, ops

    torch.manual_seed(12)
    n, d, L, k = 300, 24, 3, 2
    mlp = te.MLPEnergy(dim=d, hidden=(48, 40), activation="tanh", precision=precision).to(DEV)
    lin = [l for l in mlp.net if isinstance(l, torch.nn.Linear)]
    en_cpu = E.MLP([l.weight.cpu() for l in lin], [l.bias.cpu() for l in lin], "tanh")
    x0 = torch.randn(n, d)
    noise_p, noise_u = torch.randn(k, n, d), torch.rand(k, n)
    special = [3, 200]
    for r in special:
        noise_p[0, r, 5] = 3.0e38
        noise_p[0, r, 11] = -3.0e38
    h = 4.0   # h * p overflows for the two coordinates; every other chain just takes large steps
    desc = te.energy_descriptor(mlp, d, torch.device(DEV))
    acc = torch.zeros(k, dtype=torch.int32, device=DEV)
    got = ops.hmc_burst(desc, x0.to(DEV), k, L, [h], rng_mode=_lib.RNG_INJECTED, noise_p=noise_p.to(DEV),
                        noise_u=noise_u.to(DEV), accept_count=acc).cpu()
    want, wdiag = ohmc.sample(en_cpu, x0, k, h, L, noise_p=noise_p, noise_u=noise_u, return_diagnostics=True)
    assert torch.isfinite(want).all() and torch.isfinite(got).all()
    fmax = torch.finfo(torch.float32).max
    for r in special:   # the oracle accepted the sanitised trajectory of proposal 0
        assert want[r].abs().max().item() == fmax
        assert got[r, 5].abs().item() == fmax and got[r, 11].abs().item() == fmax, got[r]
    mag = want.abs().clamp(min=1.0)
    bad = (((got - want).abs() / mag).amax(dim=1) > 1e-3).float().mean().item()
    assert bad < 0.02, bad   # large steps: a few borderline accept decisions may flip
    assert not bool((((got[special] - want[special]).abs() / mag[special]).amax(dim=1) > 1e-3).any())
    torch.testing.assert_close(acc.float().cpu() / n, wdiag["acceptance_rate"], atol=0.02, rtol=0)
