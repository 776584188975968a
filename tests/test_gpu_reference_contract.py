"""The drop-in claim, checked against the REFERENCE'S OWN tests and classes (needs baseline/_ref, the unmodified
reference materialised by baseline/install_reference.py; it travels to the GPU box):

 (i)   the reference's cross-sampler contract file tests/samplers/test_api_contract.py is loaded unmodified from
       baseline/_ref and every test of it is run with the fused classes installed under the reference's names
       (`torchebm_b200.install()`), on CUDA (default device), so that the Langevin / HMC / descent cases are this package's
       kernels;
 (ii)  the reference's UNMODIFIED `ContrastiveDivergence(persistent=True)` is driven with a fused sampler and must leave
       the negatives, replay buffer, FIFO pointer and generator that this package's own loss leaves;
 (iii) `sample()` runs under `torch.cuda.set_sync_debug_mode("error")` like the reference's
       tests/core/test_gpu_first.py:116-136: no host synchronisation in the hot path.
"""

import importlib.util
import inspect
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
CONTRACT = os.path.join(REF, "tests", "samplers", "test_api_contract.py")
DEV = torch.device("cuda")

needs_ref = pytest.mark.skipif(not os.path.exists(CONTRACT), reason="baseline/_ref (the unmodified reference) is not materialised")


def _te():
    import torchebm_b200 as te

    if not te.REFERENCE_DERIVED:
        pytest.skip("standalone classes in use (EBM_B200_STANDALONE): nothing to check against the reference package")
    return te


@needs_ref
def test_reference_api_contract_file_passes_with_the_fused_classes_installed():
    te = _te()
    import torchebm.samplers

    te.install()
    prev_default = torch.get_default_device()
    try:
        assert torchebm.samplers.LangevinDynamics is te.LangevinDynamics
        assert torchebm.samplers.HamiltonianMonteCarlo is te.HamiltonianMonteCarlo
        spec = importlib.util.spec_from_file_location("_ref_test_api_contract", CONTRACT)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)     # its `from torchebm.samplers import ...` now binds the fused classes
        assert mod.LangevinDynamics is te.LangevinDynamics and mod.NesterovSampler is te.NesterovSampler
        ran = fused = 0

        def run_all(skip=()):
            n = 0
            for cls_name in ("TestSignatures", "TestSampleContract"):
                cls = getattr(mod, cls_name)
                for name, fn in inspect.getmembers(cls, predicate=inspect.isfunction):
                    if not name.startswith("test_") or name in skip:
                        continue
                    for case in mod.CASES:
                        thins = [1, 2, 3] if "thin" in inspect.signature(fn).parameters else [None]
                        for thin in thins:
                            try:
                                fn(cls(), case) if thin is None else fn(cls(), case, thin)
                            except pytest.skip.Exception:
                                pass
                            n += 1
            return n

        ran += run_all()                  # as the reference runs it: default device = CPU (the reference's own code paths)
        torch.set_default_device("cuda")  # the case factories pass no device: samplers, models and states land on the GPU
        # test_returns_tensor compares `samples.device == sampler.device`, which is False on CUDA for the reference's own
        # samplers too (tensor device cuda:0 vs the module's normalised "cuda", core/base_module.py:24-27); checked below
        ran += run_all(skip=("test_returns_tensor",))
        for case in mod.CASES:
            smp = case.factory()
            out = smp.sample(n_samples=4, dim=2, n_steps=6)
            assert out.shape == (4, 2) and out.dtype == smp.dtype and out.device.type == "cuda"
        # and the cases really were this package's kernels
        for case in mod.CASES:
            if case.name in ("langevin", "hmc", "gd", "nesterov"):
                smp = case.factory()
                assert smp.device.type == "cuda" and isinstance(smp, torchebm.core.BaseSampler)
                smp.sample(n_samples=4, dim=2, n_steps=6)
                # (the descent samplers have a fused burst for the elementwise energies only; on the contract's Gaussian
                # they step through the reference's own loop)
                assert smp.last_path == ("fused" if case.name in ("langevin", "hmc") else "unfused"), case.name
                fused += smp.last_path == "fused"
        assert ran >= 2 * 8 * 10 and fused == 2
    finally:
        torch.set_default_device(prev_default)
        te.uninstall()
    import torchebm.samplers as S

    assert S.LangevinDynamics is not te.LangevinDynamics and issubclass(te.LangevinDynamics, S.LangevinDynamics)


@needs_ref
@pytest.mark.parametrize("ratio,buffer_size,batch", [(0.05, 512, 512), (0.25, 1024, 256), (0.0, 300, 300)])
def test_reference_contrastive_divergence_runs_on_a_fused_sampler(ratio, buffer_size, batch):
    """losses/contrastive_divergence.py:129-134 calls `sampler.sample(x=start_points, n_steps=k, model_kwargs=...,
    generator=...)`: the reference's own loss, unmodified, on a fused sampler, against this package's loss."""
    te = _te()
    from torchebm_b200.dropin import _RefCD

    torch.manual_seed(0)
    model = te.MLPEnergy(dim=32, hidden=(64, 48), activation="silu").to(DEV)

    def make(loss_cls):
        sampler = te.LangevinDynamics(model, step_size=0.01, noise_scale=1.0, device=DEV)
        return loss_cls(model, sampler, k_steps=5, persistent=True, buffer_size=buffer_size, init_steps=0,
                        new_sample_ratio=ratio, device=DEV)

    ref_cd, our_cd = make(_RefCD), make(te.ContrastiveDivergence)
    assert type(ref_cd).__module__.startswith("torchebm.") and isinstance(our_cd, _RefCD)
    g_ref, g_our = torch.Generator(DEV).manual_seed(11), torch.Generator(DEV).manual_seed(11)
    data = torch.randn(3, batch, 32, device=DEV)
    for it in range(3):
        loss_r, neg_r = ref_cd(data[it], generator=g_ref)
        loss_o, neg_o = our_cd(data[it], generator=g_our)
        assert ref_cd.sampler.last_path == "fused" or it == 0
        assert torch.equal(neg_r, neg_o) and torch.equal(ref_cd.replay_buffer, our_cd.replay_buffer)
        assert int(ref_cd.buffer_ptr) == int(our_cd.buffer_ptr) and g_ref.get_offset() == g_our.get_offset()
        torch.testing.assert_close(loss_r, loss_o, rtol=1e-6, atol=1e-6)
    loss_r.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())


@needs_ref
def test_reference_integrator_instances_and_base_classes_are_accepted():
    te = _te()
    import torchebm
    from torchebm.integrators import EulerMaruyamaIntegrator as RefEM, LeapfrogIntegrator as RefLF

    assert issubclass(te.LangevinDynamics, torchebm.core.BaseSampler)
    assert issubclass(te.HamiltonianMonteCarlo, torchebm.core.BaseSampler)
    assert issubclass(te.EulerMaruyamaIntegrator, torchebm.core.BaseSDERungeKuttaIntegrator)
    model = te.DoubleWellModel(2.0, 1.0)
    x0 = torch.randn(512, 16, device=DEV)
    outs = []
    for integ in (None, "euler_maruyama", RefEM(device=DEV, dtype=torch.float32), te.EulerMaruyamaIntegrator(device=DEV, dtype=torch.float32)):
        s = te.LangevinDynamics(model, step_size=0.01, device=DEV, integrator=integ)
        outs.append(s.sample(x=x0, n_steps=7, generator=torch.Generator(DEV).manual_seed(3)))
        assert s.last_path == "fused"
    assert all(torch.equal(outs[0], o) for o in outs[1:])
    h = te.HamiltonianMonteCarlo(te.RastriginModel(10.0), step_size=0.01, n_leapfrog_steps=5, device=DEV,
                                 integrator=RefLF(device=DEV, dtype=torch.float32))
    h.sample(x=x0, n_steps=3, generator=torch.Generator(DEV).manual_seed(3))
    assert h.last_path == "fused"
    # what the fused kernels do not cover is the reference's own code: fp64 state, conditioning kwargs
    s64 = te.LangevinDynamics(te.DoubleWellModel(2.0, 1.0, dtype=torch.float64), step_size=0.01, device=DEV, dtype=torch.float64)
    out = s64.sample(x=x0.double(), n_steps=3)
    assert out.dtype == torch.float64 and s64.last_path == "unfused"


@needs_ref
def test_sampling_is_sync_free_under_sync_debug_mode():
    """tests/core/test_gpu_first.py:116-136 with the fused samplers: warm up, then `set_sync_debug_mode("error")`."""
    te = _te()
    torch.manual_seed(0)
    mlp = te.MLPEnergy(dim=64, hidden=128, activation="silu").to(DEV)
    x = torch.randn(2048, 64, device=DEV)
    gen = torch.Generator(DEV).manual_seed(1)
    samplers = [
        te.LangevinDynamics(te.DoubleWellModel(2.0, 1.0), step_size=0.01, device=DEV),
        te.LangevinDynamics(te.DoubleWellModel(2.0, 1.0), step_size=te.LinearScheduler(0.02, 0.005, 20), device=DEV),
        te.LangevinDynamics(mlp, step_size=0.01, device=DEV),
        te.HamiltonianMonteCarlo(te.RastriginModel(10.0), step_size=0.01, n_leapfrog_steps=5, device=DEV),
    ]
    cd = te.ContrastiveDivergence(mlp, te.LangevinDynamics(mlp, step_size=0.01, device=DEV), k_steps=5, persistent=True,
                                  buffer_size=2048, init_steps=0, new_sample_ratio=0.05, device=DEV)
    for s in samplers:                      # warm-up outside the guard (module load, workspace allocation)
        s.sample(x=x, n_steps=3, generator=gen)
        s.sample(x=x, n_steps=4, thin=2, return_diagnostics=True, return_trajectory=True, generator=gen)
    cd.sample_negatives(x, generator=gen)
    prev = torch.cuda.get_sync_debug_mode()
    torch.cuda.set_sync_debug_mode("error")
    try:
        for s in samplers:
            s.sample(x=x, n_steps=25, generator=gen)
            s.sample(x=x, n_steps=10, thin=5, return_diagnostics=True, return_trajectory=True, generator=gen)
            assert s.last_path == "fused"
        cd.sample_negatives(x, generator=gen)
    finally:
        torch.cuda.set_sync_debug_mode(prev)
