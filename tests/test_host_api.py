"""Host-side logic that needs no GPU: scheduler semantics, API signatures (the reference's
tests/samplers/test_api_contract.py contract), energy-descriptor extraction, error behaviour."""

import inspect
import math

import pytest
import torch

import torchebm_b200 as te
from torchebm_b200 import _lib
from torchebm_b200.core import advance_schedules, energy_descriptor, mark_mlp_energy


def test_sample_signature_matches_reference_contract():
    # tests/samplers/test_api_contract.py:27-46,145-187 of the reference
    prefix = ["self", "x", "dim", "n_steps", "n_samples", "thin", "return_trajectory", "return_diagnostics",
              "reset_schedulers"]
    for cls in (te.LangevinDynamics, te.HamiltonianMonteCarlo, te.GradientDescentSampler, te.NesterovSampler):
        sig = inspect.signature(cls.sample)
        names = list(sig.parameters)
        assert names[:len(prefix)] == prefix
        assert sig.parameters["thin"].default == 1 and sig.parameters["n_samples"].default == 1
        assert sig.parameters["return_trajectory"].default is False
        assert sig.parameters["generator"].kind is inspect.Parameter.KEYWORD_ONLY
        assert sig.parameters["model_kwargs"].kind is inspect.Parameter.KEYWORD_ONLY
        assert not any(p.kind in (p.VAR_POSITIONAL, p.VAR_KEYWORD) for p in sig.parameters.values())
        ctor = list(inspect.signature(cls.__init__).parameters)
        assert ctor[1] == "model" and ctor.index("dtype") < ctor.index("device")
        if "integrator" in ctor:   # test_api_contract.py:199-201: `integrator` is the LAST constructor parameter
            assert ctor[-1] == "integrator" and ctor.index("device") < ctor.index("integrator")
        assert "rng" not in ctor   # the RNG layout is an attribute, not a constructor argument
    s = te.LangevinDynamics(te.DoubleWellModel())
    assert s.rng == "torch" and s.with_rng("native").rng == "native"
    with pytest.raises(ValueError, match="rng must be"):
        s.rng = "mt19937"


def test_constructor_validation():
    m = te.DoubleWellModel()
    with pytest.raises(ValueError, match="step_size must be positive"):
        te.LangevinDynamics(m, step_size=0.0)
    with pytest.raises(ValueError, match="noise_scale must be positive"):
        te.LangevinDynamics(m, noise_scale=-1.0)
    with pytest.raises(ValueError, match="n_leapfrog_steps must be positive"):
        te.HamiltonianMonteCarlo(m, n_leapfrog_steps=-1)
    with pytest.raises(ValueError, match="Unknown integrator"):
        te.LangevinDynamics(m, integrator="nope")
    with pytest.raises(TypeError, match="requires a BaseSDERungeKuttaIntegrator"):
        te.LangevinDynamics(m, integrator=te.LeapfrogIntegrator())
    with pytest.raises(TypeError, match="requires a BaseSymplecticIntegrator"):
        te.HamiltonianMonteCarlo(m, integrator="euler_maruyama")
    with pytest.raises(ValueError, match="does not match"):
        te.LangevinDynamics(m, dtype=torch.float32, integrator=te.EulerMaruyamaIntegrator(dtype=torch.float64))


def test_cpu_sampler_has_no_path_of_its_own():
    """A CPU sampler never runs code of this package: the standalone classes raise, the reference-derived ones hand the
    call to the reference's own `sample()` (and say so in `last_path`)."""
    s = te.LangevinDynamics(te.DoubleWellModel(), step_size=0.01)
    others = (te.GradientDescentSampler(te.DoubleWellModel(), step_size=0.01),
              te.NesterovSampler(te.DoubleWellModel(), step_size=0.01, momentum=0.5),
              te.HamiltonianMonteCarlo(te.DoubleWellModel(), step_size=0.01))
    if te.REFERENCE_DERIVED:
        import torchebm

        for smp in (s,) + others:
            assert isinstance(smp, torchebm.core.BaseSampler)
            out = smp.sample(dim=2, n_samples=3, n_steps=2)
            assert out.shape == (3, 2) and out.device.type == "cpu" and smp.last_path == "unfused"
    else:
        for smp in (s,) + others:
            with pytest.raises(RuntimeError, match="CUDA device only"):
                smp.sample(dim=2, n_steps=2)
    with pytest.raises(ValueError, match="momentum must be in"):
        te.NesterovSampler(te.DoubleWellModel(), momentum=1.5)
    # the persistent-CD one-call path declines (None) instead of touching a CPU buffer
    assert s.sample_from_buffer(torch.zeros(4, 2), torch.zeros(4, dtype=torch.long), 0, 3) is None
    with pytest.raises(ValueError, match="thin must be >= 1"):
        s.sample(dim=2, thin=0)


def test_scheduler_values_follow_reference_formulas():
    lin = te.LinearScheduler(1.0, 0.0, 10)
    exp = te.ExponentialDecayScheduler(1.0, 0.9, min_value=0.5)
    cos = te.CosineScheduler(1.0, 0.0, 10)
    for i in range(1, 13):
        assert lin.step() == (0.0 if i >= 10 else 1.0 + (-0.1) * i)
        assert exp.step() == max(0.5, 0.9**i)
        want = 0.0 if i >= 10 else 0.5 * (1.0 + math.cos(math.pi * i / 10))
        assert abs(cos.step() - want) < 1e-15
    lin.reset()
    assert lin.get_value() == 1.0 and lin.step_count == 0


def test_advance_schedules_semantics():
    # langevin_dynamics.py:161-168: step i uses the value after i .step() calls; step_count == n_steps afterwards
    s = te.LangevinDynamics(te.DoubleWellModel(), step_size=te.LinearScheduler(0.1, 0.01, 5), noise_scale=2.0)
    vals, constant = advance_schedules(s, ("step_size", "noise_scale"), 7)
    assert not constant
    assert vals["step_size"][:6] == pytest.approx([0.1, 0.082, 0.064, 0.046, 0.028, 0.01])
    assert vals["noise_scale"] == [2.0] * 7
    assert s.schedulers["step_size"].step_count == 7 and s.schedulers["noise_scale"].step_count == 7
    s2 = te.LangevinDynamics(te.DoubleWellModel(), step_size=0.01)
    vals, constant = advance_schedules(s2, ("step_size", "noise_scale"), 500)
    assert constant and vals == {"step_size": [0.01], "noise_scale": [1.0]}
    assert all(isinstance(v, te.ConstantScheduler) and v.step_count == 500 for v in s2.schedulers.values())
    s2.reset_schedulers()
    assert all(v.step_count == 0 for v in s2.schedulers.values())


def test_energy_descriptor_extraction():
    dev = torch.device("cpu")
    d = energy_descriptor(te.DoubleWellModel(1.7, 1.3), 7, dev)
    assert d.kind == "double_well" and d.c.kind == _lib.ENERGY_DOUBLE_WELL and d.c.dim == 7
    assert d.c.p[0] == torch.tensor(1.7).item() and d.c.p[1] == torch.tensor(1.3**2, dtype=torch.float32).item()
    d = energy_descriptor(te.RastriginModel(10.0), 64, dev)
    assert d.c.p[1] == torch.tensor(2 * math.pi, dtype=torch.float32).item() and d.c.p[2] == 640.0
    d = energy_descriptor(te.HarmonicModel(1.3), 3, dev)
    assert d.c.p[0] == torch.tensor(0.65, dtype=torch.float32).item()
    g = te.GaussianModel(torch.zeros(3), torch.eye(3))
    assert energy_descriptor(g, 3, dev).kind == "gaussian" and energy_descriptor(g, 4, dev) is None
    mlp = te.MLPEnergy(dim=16, hidden=(32, 24), activation="tanh")
    d = energy_descriptor(mlp, 16, dev)
    assert (d.c.hidden1, d.c.hidden2, d.c.activation) == (32, 24, _lib.ACT_TANH)
    assert d.c.buf[0] == mlp.net[0].weight.data_ptr()  # pointers into the live parameters: weights are read fresh
    assert energy_descriptor(mlp, 17, dev) is None
    assert d.c.hidden3 == 0 and not d.c.buf[7]
    # three hidden layers (benchmarks/distributed_fsdp2.py:43-53): hidden3 + W3 / b3 in buf[7..8], the output layer stays in buf[4..5]
    deep = te.MLPEnergy(dim=16, hidden=(32, 24, 40), activation="silu", precision="fp32")
    d = energy_descriptor(deep, 16, dev)
    assert (d.c.hidden1, d.c.hidden2, d.c.hidden3) == (32, 24, 40) and d.c.precision == _lib.MLP_BF16X3   # tensor-core kernel only
    assert d.c.buf[7] == deep.net[4].weight.data_ptr() and d.c.buf[8] == deep.net[4].bias.data_ptr()
    assert d.c.buf[4] == deep.net[6].weight.data_ptr() and d.c.buf[5] == deep.net[6].bias.data_ptr()
    assert energy_descriptor(te.MLPEnergy(dim=200, hidden=(32, 24, 40)), 200, dev) is None   # three layers: states up to 128 wide
    assert energy_descriptor(te.MLPEnergy(dim=16, hidden=(32, 200, 40)), 16, dev) is None
    with pytest.raises(ValueError):
        te.MLPEnergy(dim=16, hidden=(32, 24, 40, 8))
    four = torch.nn.Sequential(*[m for w in ((16, 8), (8, 8), (8, 8), (8, 8)) for m in (torch.nn.Linear(*w), torch.nn.SiLU())],
                               torch.nn.Linear(8, 1))
    with pytest.raises(ValueError):
        te.MLPEnergy(net=four)
    mixed = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.SiLU(), torch.nn.Linear(8, 8), torch.nn.Tanh(), torch.nn.Linear(8, 8),
                                torch.nn.SiLU(), torch.nn.Linear(8, 1))
    with pytest.raises(ValueError):
        te.MLPEnergy(net=mixed)

    class Custom(te.BaseModel):  # user subclass with its own forward: never matched by name
        def forward(self, x):
            return x.sum(-1)

    class DoubleWellModel(te.BaseModel):  # same class name from a foreign module: not trusted
        barrier_height, b = 2.0, 1.0

        def forward(self, x):
            return x.sum(-1)

    assert energy_descriptor(Custom(), 4, dev) is None
    assert energy_descriptor(DoubleWellModel(), 4, dev) is None


def test_mark_mlp_energy_recognises_user_modules():
    class UserEBM(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.SiLU(), torch.nn.Linear(16, 16), torch.nn.SiLU(),
                                           torch.nn.Linear(16, 1))

        def forward(self, x):
            return self.net(x).squeeze(-1)

    class NotAnMLP(UserEBM):
        def forward(self, x):
            return self.net(x).squeeze(-1) + x.sum(-1)

    m = UserEBM()
    assert energy_descriptor(m, 8, torch.device("cpu")) is None
    assert mark_mlp_energy(m)
    assert energy_descriptor(m, 8, torch.device("cpu")).kind == "mlp"
    assert not mark_mlp_energy(NotAnMLP())

    class DeepUserEBM(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 12), torch.nn.Tanh(),
                                           torch.nn.Linear(12, 16), torch.nn.Tanh(), torch.nn.Linear(16, 1))

        def forward(self, x):
            return self.net(x).squeeze(-1)

    dm = DeepUserEBM()
    assert mark_mlp_energy(dm)
    dd = energy_descriptor(dm, 8, torch.device("cpu"))
    assert dd.kind == "mlp" and dd.c.hidden3 == 16 and dd.c.activation == _lib.ACT_TANH


def test_models_forward_match_oracle_energies():
    from oracle import energies as E

    x = torch.randn(32, 6)
    assert torch.equal(te.DoubleWellModel(1.7, 1.3)(x), E.DoubleWell(1.7, 1.3).energy(x))
    assert torch.equal(te.HarmonicModel(1.3)(x), E.Harmonic(1.3).energy(x))
    assert torch.equal(te.RastriginModel(10.0)(x), E.Rastrigin(10.0).energy(x))
    means, sig, w = torch.randn(4, 6), torch.rand(4) + 0.5, torch.softmax(torch.randn(4), 0)
    assert torch.equal(te.MixtureOfGaussiansModel(means, sig, w)(x), E.MixtureOfGaussians(means, sig, w).energy(x))
    a = torch.randn(6, 6)
    cov = a @ a.t() + torch.eye(6)
    torch.testing.assert_close(te.GaussianModel(torch.zeros(6), cov)(x), E.Gaussian(torch.zeros(6), cov).energy(x),
                               rtol=1e-5, atol=1e-5)
    # CPU gradient of a model = autograd (the reference's default), bit-equal to the oracle's
    assert torch.equal(te.DoubleWellModel(1.7, 1.3).gradient(x), E.DoubleWell(1.7, 1.3).gradient(x))


def test_cd_constructor_and_nonpersistent_start_points():
    m = te.DoubleWellModel()
    cd = te.ContrastiveDivergence(m, te.LangevinDynamics(m), k_steps=3)
    x = torch.randn(4, 2)
    sp = cd.get_start_points(x)
    assert torch.equal(sp, x) and sp.data_ptr() != x.data_ptr()
    assert cd.replay_buffer is None and int(cd.buffer_ptr) == 0
    with pytest.raises(RuntimeError, match="persistent"):
        cd.mix_buffer_across_ranks()
