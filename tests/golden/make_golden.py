"""Generate tests/golden/*.npz by running the UNMODIFIED reference (torchebm @ a77aeee).

Run here (the container that mounts /root/reference), never on the GPU box:

    python tests/golden/make_golden.py

The reference is copied to a temp dir only to add the `_version.py` stub that setuptools_scm
would generate (torchebm/__init__.py:10); nothing from it is written into this repo except the
numeric input/output vectors below.  Every case stores its inputs (x0, pre-drawn noise, parameters)
and the reference's outputs.  Noise is pre-drawn with the reference's own draw order (SURVEY.md
section 8c): one `randn_like(x)` per Langevin step; per HMC proposal `normal_` then `rand(N)`.
Each case asserts that the reference, driven by the same seeded generator, consumed exactly that noise
(the recorded `out` is produced by the generator-driven reference run).
"""

import math
import os
import shutil
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def _import_reference():
    tmp = tempfile.mkdtemp(prefix="torchebm_ref_")
    shutil.copytree(os.path.join(REF, "torchebm"), os.path.join(tmp, "torchebm"))
    with open(os.path.join(tmp, "torchebm", "_version.py"), "w") as f:
        f.write('__version__ = "0.0.0+ref"\n')
    sys.path.insert(0, tmp)
    import torchebm  # noqa: F401

    return tmp


ONLY = set(sys.argv[1:])  # optional: regenerate only the named cases


def save(name, **arrs):
    if ONLY and name not in ONLY:
        return
    out = {}
    for k, v in arrs.items():
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: tuple(v.shape) for k, v in out.items()})


def predraw_langevin(x0, k, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.stack([torch.randn_like(x0, generator=g) for _ in range(k)])


def predraw_hmc(x0, k, seed):
    g = torch.Generator().manual_seed(seed)
    ps, us = [], []
    for _ in range(k):
        ps.append(torch.empty_like(x0).normal_(generator=g))
        us.append(torch.rand(x0.shape[0], generator=g))
    return torch.stack(ps), torch.stack(us)


def main():
    _import_reference()
    from torchebm.core import (BaseModel, DoubleWellModel, GaussianModel, HarmonicModel, RastriginModel,
                               ExponentialDecayScheduler, LinearScheduler)
    from torchebm.samplers import LangevinDynamics, HamiltonianMonteCarlo
    from torchebm.integrators import LeapfrogIntegrator
    from torchebm.losses import ContrastiveDivergence

    torch.manual_seed(1234)

    class MLPEnergy(BaseModel):
        def __init__(self, d, h, act):
            super().__init__()
            self.net = torch.nn.Sequential(torch.nn.Linear(d, h), act(), torch.nn.Linear(h, h), act(), torch.nn.Linear(h, 1))

        def forward(self, x):
            return self.net(x).squeeze(-1)

    class MoGEnergy(BaseModel):
        """Not a reference class: the reference's autograd `gradient` applied to the MoG forward."""

        def __init__(self, means, sigmas, weights):
            super().__init__()
            self.register_buffer("means", means)
            self.register_buffer("sigmas", sigmas)
            self.register_buffer("weights", weights)

        def forward(self, x):
            d = x.shape[-1]
            diff = x.unsqueeze(1) - self.means.unsqueeze(0)
            sq = diff.pow(2).sum(dim=-1)
            logits = torch.log(self.weights) - d * torch.log(self.sigmas) - sq / (2.0 * self.sigmas**2)
            return -torch.logsumexp(logits, dim=-1)

    # ---------------- Langevin, elementwise energies -------------------------------------------------
    def langevin_case(name, model, n, d, k, h, ns, seed, extra=None, **kw):
        x0 = torch.randn(n, d, generator=torch.Generator().manual_seed(seed + 1000))
        noise = predraw_langevin(x0, k, seed)
        s = LangevinDynamics(model, step_size=h, noise_scale=ns, clamp=kw.pop("clamp", None))
        res = s.sample(x=x0, n_steps=k, generator=torch.Generator().manual_seed(seed), **kw)
        arrs = dict(x0=x0, noise=noise, k=k, h=h if isinstance(h, float) else -1.0, ns=ns if isinstance(ns, float) else -1.0)
        if noise.numel() > 100_000:
            # too big to commit: the test re-draws it from the CPU generator seed and checks the checksum
            del arrs["noise"]
            arrs["noise_seed"] = seed
            arrs["noise_checksum"] = noise.double().sum()
        if isinstance(res, tuple):
            arrs["out"] = res[0]
            for kk, vv in res[1].items():
                arrs["diag_" + kk] = vv
        else:
            arrs["out"] = res
        arrs["grad0"] = model.gradient(x0)
        arrs["energy0"] = model(x0).detach()
        arrs.update(extra or {})
        save(name, **arrs)

    langevin_case("langevin_doublewell", DoubleWellModel(2.0, 1.0), 64, 16, 20, 0.01, 1.0, 1)
    langevin_case("langevin_doublewell_odd", DoubleWellModel(1.7, 1.3), 33, 7, 12, 0.0137, 0.73, 2,
                  clamp=(-1.5, 1.5), thin=3, return_trajectory=True, return_diagnostics=True,
                  extra=dict(barrier_height=1.7, b=1.3, clamp=np.array([-1.5, 1.5])))
    langevin_case("langevin_doublewell_k100", DoubleWellModel(2.0, 1.0), 256, 128, 100, 0.01, 1.0, 3)
    langevin_case("langevin_harmonic", HarmonicModel(k=1.3), 48, 10, 15, 0.05, 0.5, 4, extra=dict(kspring=1.3))
    langevin_case("langevin_rastrigin", RastriginModel(a=10.0), 64, 8, 20, 0.001, 1.0, 5, extra=dict(a=10.0))
    # C1: GaussianModel dim 2 (examples/10-sampling/01-mcmc/01-langevin-101/main.py:17 covariance)
    mean2 = torch.tensor([1.0, -1.0])
    cov2 = torch.tensor([[1.0, 0.8], [0.8, 1.0]])
    langevin_case("langevin_gaussian_c1", GaussianModel(mean2, cov2), 128, 2, 100, 0.01, 1.0, 6,
                  extra=dict(mean=mean2, cov=cov2))
    a = torch.randn(16, 16)
    cov16 = a @ a.t() / 16 + 0.5 * torch.eye(16)
    mean16 = torch.randn(16)
    langevin_case("langevin_gaussian_d16", GaussianModel(mean16, cov16), 40, 16, 10, 0.02, 0.7, 7,
                  extra=dict(mean=mean16, cov=cov16))
    # noise_scale tiny-but-positive and single chain (n == 1 diagnostics branch)
    langevin_case("langevin_single_chain", DoubleWellModel(2.0, 1.0), 1, 5, 6, 0.01, 1.0, 8,
                  return_diagnostics=True)

    # scheduled step size / noise scale: record the per-step values the sampler saw
    k = 12
    hs = ExponentialDecayScheduler(start_value=0.02, decay_rate=0.9, min_value=0.001)
    nss = LinearScheduler(start_value=1.0, end_value=0.1, n_steps=k)
    hv, nv = [], []
    hs.reset(); nss.reset()
    for _ in range(k):
        hv.append(hs.get_value()); nv.append(nss.get_value()); hs.step(); nss.step()
    hs.reset(); nss.reset()
    langevin_case("langevin_scheduled", DoubleWellModel(2.0, 1.0), 32, 6, k, hs, nss, 9,
                  extra=dict(h_values=np.array(hv), ns_values=np.array(nv)))

    # MLP energies (SiLU as in examples/20-training/01-mcmc-losses/01-cd-k/main.py:20-30; Tanh as in
    # tests/losses/test_contrastive_divergence.py:96-109)
    for act_name, act in (("silu", torch.nn.SiLU), ("tanh", torch.nn.Tanh)):
        m = MLPEnergy(16, 32, act)
        lin = [l for l in m.net if isinstance(l, torch.nn.Linear)]
        extra = {}
        for i, l in enumerate(lin):
            extra[f"w{i}"] = l.weight.detach()
            extra[f"b{i}"] = l.bias.detach()
        langevin_case(f"langevin_mlp_{act_name}", m, 64, 16, 10, 0.01, 1.0, 10, extra=extra)
    m = MLPEnergy(128, 128, torch.nn.SiLU)
    lin = [l for l in m.net if isinstance(l, torch.nn.Linear)]
    extra = {}
    for i, l in enumerate(lin):
        extra[f"w{i}"] = l.weight.detach()
        extra[f"b{i}"] = l.bias.detach()
    langevin_case("langevin_mlp_d128", m, 96, 128, 5, 0.01, 1.0, 11, extra=extra)

    means = torch.randn(5, 6) * 2
    sigmas = torch.rand(5) * 0.5 + 0.5
    weights = torch.softmax(torch.randn(5), 0)
    langevin_case("langevin_mog", MoGEnergy(means, sigmas, weights), 64, 6, 15, 0.01, 1.0, 12,
                  extra=dict(means=means, sigmas=sigmas, weights=weights))

    # ---------------- Leapfrog KATs -------------------------------------------------------------------
    lf = LeapfrogIntegrator()
    dw = DoubleWellModel(2.0, 1.0)
    x0 = torch.randn(32, 8)
    p0 = torch.randn(32, 8)
    mass_vec = torch.rand(8) + 0.5
    for tag, mass, safe in (("nomass", None, False), ("safe", None, True), ("mass_float", 2.5, True), ("mass_vec", mass_vec, True)):
        res = lf.integrate({"x": x0, "p": p0}, step_size=0.03, n_steps=7, mass=mass,
                           drift=lambda x_, t_: -dw.gradient(x_), safe=safe)
        save(f"leapfrog_{tag}", x0=x0, p0=p0, h=0.03, L=7, safe=int(safe),
             mass=(np.array(-1.0) if mass is None else (np.array(mass) if isinstance(mass, float) else mass)),
             x=res["x"], p=res["p"])
    # safe mode with huge states: clamp at 1e6 and inf -> FLT_MAX
    xb = torch.tensor([[1e4, -3e5, 0.5, 1e6], [2.0, 1e3, -1e2, 0.0]])
    pb = torch.tensor([[0.0, 1.0, -1.0, 2.0], [1e30, -1e30, 3e38, 0.0]])
    res = lf.integrate({"x": xb, "p": pb}, step_size=0.1, n_steps=3, drift=lambda x_, t_: -dw.gradient(x_), safe=True)
    save("leapfrog_extreme", x0=xb, p0=pb, h=0.1, L=3, safe=1, mass=np.array(-1.0), x=res["x"], p=res["p"])

    # ---------------- HMC -----------------------------------------------------------------------------
    def hmc_case(name, model, n, d, k, h, L, seed, mass=None, extra=None, x0_scale=1.0, **kw):
        x0 = torch.randn(n, d, generator=torch.Generator().manual_seed(seed + 1000)) * x0_scale
        noise_p, noise_u = predraw_hmc(x0, k, seed)
        s = HamiltonianMonteCarlo(model, step_size=h, n_leapfrog_steps=L, mass=mass)
        res = s.sample(x=x0, n_steps=k, generator=torch.Generator().manual_seed(seed), **kw)
        arrs = dict(x0=x0, noise_p=noise_p, noise_u=noise_u, k=k, h=h, L=L,
                    mass=(np.array(-1.0) if mass is None else (np.array(mass) if isinstance(mass, float) else mass)))
        if isinstance(res, tuple):
            arrs["out"] = res[0]
            for kk, vv in res[1].items():
                arrs["diag_" + kk] = vv
        else:
            arrs["out"] = res
        arrs.update(extra or {})
        save(name, **arrs)

    hmc_case("hmc_doublewell", DoubleWellModel(2.0, 1.0), 64, 8, 10, 0.05, 5, 21)
    hmc_case("hmc_rastrigin", RastriginModel(a=10.0), 64, 8, 6, 0.01, 20, 22, extra=dict(a=10.0))
    hmc_case("hmc_rastrigin_diag", RastriginModel(a=10.0), 50, 5, 9, 0.02, 4, 23, thin=2,
             return_trajectory=True, return_diagnostics=True, extra=dict(a=10.0))
    hmc_case("hmc_gaussian", GaussianModel(mean2, cov2), 128, 2, 12, 0.2, 6, 24, extra=dict(mean=mean2, cov=cov2))
    hmc_case("hmc_mass_float", DoubleWellModel(2.0, 1.0), 48, 6, 8, 0.05, 5, 25, mass=2.5)
    hmc_case("hmc_mass_vec", HarmonicModel(k=2.0), 48, 6, 8, 0.1, 5, 26, mass=torch.rand(6) + 0.5, extra=dict(kspring=2.0))
    hmc_case("hmc_far_start", DoubleWellModel(2.0, 1.0), 16, 4, 4, 0.05, 5, 27, x0_scale=1e4)

    # ---------------- persistent CD -------------------------------------------------------------------
    m = MLPEnergy(6, 8, torch.nn.Tanh)
    lin = [l for l in m.net if isinstance(l, torch.nn.Linear)]
    extra = {}
    for i, l in enumerate(lin):
        extra[f"w{i}"] = l.weight.detach()
        extra[f"b{i}"] = l.bias.detach()
    sampler = LangevinDynamics(m, step_size=0.01, noise_scale=1.0)
    cd = ContrastiveDivergence(m, sampler, k_steps=3, persistent=True, buffer_size=40, init_steps=0,
                               new_sample_ratio=0.25, energy_reg_weight=0.001)
    g = torch.Generator().manual_seed(77)
    data = torch.randn(4, 16, 6, generator=torch.Generator().manual_seed(78))
    losses, negs, bufs, ptrs = [], [], [], []
    for it in range(4):
        loss, neg = cd(data[it], generator=g)
        losses.append(loss.detach()); negs.append(neg.detach()); bufs.append(cd.replay_buffer.clone()); ptrs.append(cd._buffer_ptr_int)
    save("pcd_mlp_tanh", data=data, seed=77, losses=torch.stack(losses), negs=torch.stack(negs),
         bufs=torch.stack(bufs), ptrs=np.array(ptrs), **extra)

    # stratified-gather / FIFO arithmetic alone (buffer larger than batch, wraparound)
    cd2 = ContrastiveDivergence(DoubleWellModel(2.0, 1.0), LangevinDynamics(DoubleWellModel(2.0, 1.0), step_size=0.01),
                                k_steps=2, persistent=True, buffer_size=50, init_steps=0, new_sample_ratio=0.0)
    g = torch.Generator().manual_seed(5)
    data = torch.randn(5, 16, 3, generator=torch.Generator().manual_seed(6))
    negs, bufs, ptrs = [], [], []
    for it in range(5):
        _, neg = cd2(data[it], generator=g)
        negs.append(neg.detach()); bufs.append(cd2.replay_buffer.clone()); ptrs.append(cd2._buffer_ptr_int)
    save("pcd_doublewell_fifo", data=data, seed=5, negs=torch.stack(negs), bufs=torch.stack(bufs), ptrs=np.array(ptrs))

    # wide-state MLP energy (BASELINE configs C3/C5: dim 784): ragged single tile, narrow hidden layers.  Own
    # manual_seed so that adding it did not shift the parameter draws of the cases above.
    torch.manual_seed(4321)
    m = MLPEnergy(784, 64, torch.nn.SiLU)
    lin = [l for l in m.net if isinstance(l, torch.nn.Linear)]
    extra = {}
    for i, l in enumerate(lin):
        extra[f"w{i}"] = l.weight.detach()
        extra[f"b{i}"] = l.bias.detach()
    langevin_case("langevin_mlp_d784", m, 40, 784, 4, 0.01, 1.0, 21, extra=extra)

    # Heun SDE integrator behind LangevinDynamics: own sampler construction, same case recorder
    def heun_case(name, model, n, d, k, h, ns, seed, **kw):
        x0 = torch.randn(n, d, generator=torch.Generator().manual_seed(seed + 1000))
        noise = predraw_langevin(x0, k, seed)
        s = LangevinDynamics(model, step_size=h, noise_scale=ns, integrator="heun")
        res = s.sample(x=x0, n_steps=k, generator=torch.Generator().manual_seed(seed), **kw)
        save(name, x0=x0, noise=noise, k=k, h=h, ns=ns, out=res)

    heun_case("heun_doublewell", DoubleWellModel(2.0, 1.0), 64, 16, 15, 0.01, 1.0, 41)
    heun_case("heun_rastrigin_traj", RastriginModel(10.0), 33, 7, 12, 0.002, 0.5, 42, thin=3, return_trajectory=True)

    # noise-free descent samplers (samplers/gradient_descent.py); deterministic, so only x0 and the outputs are stored
    from torchebm.samplers import GradientDescentSampler, NesterovSampler

    def descent_case(name, sampler, n, d, k, seed, extra=None, **kw):
        x0 = torch.randn(n, d, generator=torch.Generator().manual_seed(seed))
        res = sampler.sample(x=x0, n_steps=k, **kw)
        arrs = dict(x0=x0, k=k)
        if isinstance(res, tuple):
            arrs["out"] = res[0]
            arrs["diag_energy"] = res[1]["energy"]
        else:
            arrs["out"] = res
        arrs.update(extra or {})
        save(name, **arrs)

    descent_case("gd_doublewell", GradientDescentSampler(DoubleWellModel(2.0, 1.0), step_size=0.01), 64, 16, 25, 31)
    descent_case("gd_rastrigin_traj", GradientDescentSampler(RastriginModel(10.0), step_size=0.001), 33, 7, 12, 32,
                 thin=3, return_trajectory=True, return_diagnostics=True, extra=dict(a=10.0))
    k = 70
    hs = ExponentialDecayScheduler(start_value=0.005, decay_rate=0.97, min_value=0.0005)
    hv = []
    hs.reset()
    for _ in range(k):
        hv.append(hs.get_value()); hs.step()
    hs.reset()
    descent_case("nesterov_doublewell_sched", NesterovSampler(DoubleWellModel(2.0, 1.0), step_size=hs, momentum=0.9), 40, 10, k, 33,
                 extra=dict(h_values=np.array(hv), momentum=0.9))
    descent_case("nesterov_harmonic_traj", NesterovSampler(HarmonicModel(1.5), step_size=0.05, momentum=0.5), 20, 6, 9, 34,
                 thin=2, return_trajectory=True, return_diagnostics=True, extra=dict(kspring=1.5, momentum=0.5))
    # same 784-64-64-1 energy as langevin_mlp_d784 (its weights are stored there)
    descent_case("gd_mlp", GradientDescentSampler(m, step_size=0.05), 8, 784, 3, 35)


if __name__ == "__main__":
    main()
