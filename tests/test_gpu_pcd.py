"""Persistent-CD replay buffer on the device and the ContrastiveDivergence loss driving a fused sampler."""

import pytest
import torch

from oracle import energies as E
from oracle import langevin as olang
from oracle import pcd as opcd

from . import _cases as C

pytestmark = pytest.mark.gpu

DEV = "cuda"


def test_gather_scatter_kernels_match_oracle_fifo():
    from torchebm_b200 import ops

    g = torch.Generator().manual_seed(0)
    buf = torch.randn(50, 3, 4, generator=g)
    dbuf = buf.to(DEV).clone()
    ob = opcd.ReplayBuffer(50)
    ob.buffer = buf.clone()
    ptr = 0
    for b in (16, 16, 16, 16, 7, 50, 64):
        samples = torch.randn(b, 3, 4, generator=g)
        ob.update(samples)
        ptr = ops.pcd_scatter(dbuf, ptr, samples.to(DEV))
        assert ptr == ob.ptr
        assert torch.equal(dbuf.cpu(), ob.buffer)
    idx = torch.randint(0, 50, (20,), generator=g)
    rows = torch.randperm(20, generator=g)[:5]
    noise = torch.randn(5, 3, 4, generator=g)
    want = ob.buffer[idx]
    want[rows] = want[rows] + noise * 0.01
    got = ops.pcd_gather(dbuf, idx.to(DEV), rows.to(DEV), noise.to(DEV))
    assert torch.equal(got.cpu(), want)


def test_cd_persistent_sequence_matches_oracle_on_cuda():
    """Same seed: buffer init, stratified indices, exploration noise, K-step chain, FIFO write-back and the loss
    all agree with the oracle (which is pinned to the reference's own CD run in tests/test_oracle_golden.py)."""
    import torchebm_b200 as te

    g = C.load("pcd_mlp_tanh")
    model = te.MLPEnergy(dim=6, hidden=8, activation="tanh")
    lin = [l for l in model.net if isinstance(l, torch.nn.Linear)]
    with torch.no_grad():
        for i, l in enumerate(lin):
            l.weight.copy_(g[f"w{i}"])
            l.bias.copy_(g[f"b{i}"])
    model = model.to(DEV)
    sampler = te.LangevinDynamics(model, step_size=0.01, noise_scale=1.0, device=DEV)
    cd = te.ContrastiveDivergence(model, sampler, k_steps=3, persistent=True, buffer_size=40, init_steps=0,
                                  new_sample_ratio=0.25, energy_reg_weight=0.001, device=DEV)
    en = C.mlp_from(g, "tanh").to(DEV)
    g1 = torch.Generator(DEV).manual_seed(77)
    g2 = torch.Generator(DEV).manual_seed(77)
    ob = opcd.ReplayBuffer(40, new_sample_ratio=0.25)
    for it in range(4):
        x = g["data"][it].to(DEV)
        loss, neg = cd(x, generator=g1)
        if ob.buffer is None:
            ob.initialize((6,), DEV, generator=g2)
        start = ob.get_start_points(16, generator=g2)
        wneg = olang.sample(en, start, 3, 0.01, 1.0, generator=g2)
        ob.update(wneg)
        torch.testing.assert_close(neg, wneg, rtol=1e-4, atol=2e-5)
        torch.testing.assert_close(cd.replay_buffer, ob.buffer, rtol=1e-4, atol=2e-5)
        assert cd._buffer_ptr_int == ob.ptr == int(cd.buffer_ptr.item())
        assert g1.get_offset() == g2.get_offset()
        xe, ne = en.energy(x), en.energy(wneg)
        wloss = xe.mean() - ne.mean() + 0.001 * ((xe**2).mean() + (ne**2).mean())
        torch.testing.assert_close(loss, wloss, rtol=1e-4, atol=1e-5)
    assert loss.requires_grad
    loss.backward()
    assert all(p.grad is not None for p in model.parameters())


def test_cd_fifo_pointer_and_state_dict_roundtrip():
    import torchebm_b200 as te

    model = te.DoubleWellModel(2.0, 1.0)
    sampler = te.LangevinDynamics(model, step_size=0.01, device=DEV)
    cd = te.ContrastiveDivergence(model, sampler, k_steps=2, persistent=True, buffer_size=50, init_steps=0,
                                  new_sample_ratio=0.0, device=DEV)
    ptrs = []
    for it in range(5):
        cd(torch.randn(16, 3, device=DEV))
        ptrs.append(cd._buffer_ptr_int)
    assert ptrs == [16, 32, 48, 14, 30]  # tests/losses/test_contrastive_divergence.py:419-453 arithmetic
    sd = cd.state_dict()
    assert "replay_buffer" in sd and "buffer_ptr" in sd
    cd2 = te.ContrastiveDivergence(model, sampler, k_steps=2, persistent=True, buffer_size=50, init_steps=0, device=DEV)
    cd2.initialize_buffer((3,))
    cd2.load_state_dict(sd)
    assert cd2._buffer_ptr_int == 30 and torch.equal(cd2.replay_buffer, cd.replay_buffer)


def test_cd_buffer_warmup_and_full_overwrite():
    """init_steps > 0 runs the sampler over 1024-row chunks (base_loss.py:236-246); batch == buffer overwrites all."""
    import torchebm_b200 as te

    model = te.DoubleWellModel(2.0, 1.0)
    sampler = te.LangevinDynamics(model, step_size=0.01, device=DEV)
    cd = te.ContrastiveDivergence(model, sampler, k_steps=5, persistent=True, buffer_size=2048, init_steps=10,
                                  new_sample_ratio=0.05, device=DEV)
    loss, neg = cd(torch.randn(2048, 4, device=DEV), generator=torch.Generator(DEV).manual_seed(0))
    assert neg.shape == (2048, 4) and torch.isfinite(loss)
    assert cd._buffer_ptr_int == 0 and torch.equal(cd.replay_buffer, neg)


def test_c3_full_size_cd_step_properties():
    """BASELINE config 3 at full size: persistent CD, MLP 784-128-128-1, N = 65 536 = buffer_size, k = 20, through the
    size-independent properties the path offers:
    * the one-call negatives equal get_start_points -> sample -> update_buffer bit for bit (same seed), the buffer is the
      negatives and the FIFO pointer wraps to 0;
    * pure diffusion (an energy with zero weights has zero gradient): x_K - x_0 is N(0, 2 h K sigma^2) per element, which
      pins the in-kernel noise scale and the update arithmetic at full size;
    * nearly noise-free chains descend the energy (gradient direction and step sign);
    * the loss is finite and differentiable."""
    import torchebm_b200 as te

    n, d, k = 65536, 784, 20
    torch.manual_seed(0)
    model = te.MLPEnergy(dim=d, hidden=128, activation="silu").to(DEV)

    def make(ns=1.0, m=model):
        sampler = te.LangevinDynamics(m, step_size=0.01, noise_scale=ns, device=DEV).with_rng("native")
        return te.ContrastiveDivergence(m, sampler, k_steps=k, persistent=True, buffer_size=n, init_steps=0,
                                        new_sample_ratio=0.0, device=DEV)

    x = torch.randn(n, d, device=DEV)
    a, b = make(), make()
    ga, gb = torch.Generator(DEV).manual_seed(0), torch.Generator(DEV).manual_seed(0)
    loss, neg = a(x, generator=ga)
    start = b.get_start_points(x, generator=gb)
    neg_b = b.sampler.sample(x=start, n_steps=k, generator=gb)
    b.update_buffer(neg_b)
    assert torch.equal(neg, neg_b) and torch.equal(a.replay_buffer, neg) and a._buffer_ptr_int == 0
    assert torch.isfinite(neg).all() and torch.isfinite(loss)
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())

    flat = te.MLPEnergy(dim=d, hidden=128, activation="silu").to(DEV)
    with torch.no_grad():
        for p in flat.parameters():
            p.zero_()
    s = te.LangevinDynamics(flat, step_size=0.01, noise_scale=0.7, device=DEV).with_rng("native")
    x0 = torch.randn(n, d, device=DEV)
    dx = s.sample(x=x0, n_steps=k, generator=torch.Generator(DEV).manual_seed(3)) - x0
    var = 2 * 0.01 * k * 0.7 ** 2
    assert abs(dx.mean().item()) < 5e-4 and abs(dx.var().item() / var - 1.0) < 5e-3
    assert abs(dx.var(dim=0).mean().item() / var - 1.0) < 5e-3 and abs((dx[:, :392] * dx[:, 392:]).mean().item()) < 5e-4

    quiet = te.LangevinDynamics(model, step_size=0.01, noise_scale=1e-4, device=DEV).with_rng("native")
    with torch.no_grad():
        e0 = model(x0)
        e1 = model(quiet.sample(x=x0, n_steps=k, generator=torch.Generator(DEV).manual_seed(4)))
    assert (e1 < e0).float().mean().item() > 0.999


@pytest.mark.parametrize("d,buffer_size,batch", [(784, 300, 300), (784, 700, 300), (64, 33000, 33000), (64, 40000, 4096),
                                                  (64, 100, 300), (64, 640, 640)])
@pytest.mark.parametrize("rng", ["torch", "native"])
@pytest.mark.parametrize("ratio", [0.0, 0.05, 0.25])
def test_fused_pcd_negatives_equal_the_three_call_path(d, buffer_size, batch, rng, ratio):
    """`sample_negatives` (one library call: start rows read through the index inside the burst kernel, final state
    written straight back into the buffer when buffer_size == batch) must leave exactly the negatives, the buffer, the
    FIFO pointer and the generator that get_start_points -> sample -> update_buffer leave."""
    import warnings

    import torchebm_b200 as te

    torch.manual_seed(3)
    # (the 100- and 640-row cases cover the three-hidden-layer energy: its kernel reads start rows and writes back the same way)
    model = te.MLPEnergy(dim=d, hidden=(128, 96, 64) if buffer_size in (100, 640) else (128, 96), activation="silu").to(DEV)

    def make():
        sampler = te.LangevinDynamics(model, step_size=0.01, noise_scale=1.0, device=DEV).with_rng(rng)
        return te.ContrastiveDivergence(model, sampler, k_steps=4, persistent=True, buffer_size=buffer_size, init_steps=0,
                                        new_sample_ratio=ratio, device=DEV)

    a, b = make(), make()
    ga, gb = torch.Generator(DEV).manual_seed(21), torch.Generator(DEV).manual_seed(21)
    data = torch.randn(3, batch, d, device=DEV)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # buffer smaller than batch warns, like the reference
        for it in range(3):
            neg_a = a.sample_negatives(data[it], generator=ga)
            start = b.get_start_points(data[it], generator=gb)
            neg_b = b.sampler.sample(x=start, n_steps=4, generator=gb)
            b.update_buffer(neg_b)
            assert torch.equal(neg_a, neg_b)
            assert torch.equal(a.replay_buffer, b.replay_buffer)
            assert a._buffer_ptr_int == b._buffer_ptr_int and int(a.buffer_ptr) == int(b.buffer_ptr)
            assert ga.get_offset() == gb.get_offset()
    # E(x-) of the negatives from the same call (contrastive_divergence.py:184-223 needs it for the loss value)
    e_out = torch.empty(batch, device=DEV)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        neg = a.sample_negatives(data[0], generator=ga, energy_out=e_out)
    with torch.no_grad():
        torch.testing.assert_close(e_out, model(neg), rtol=1e-4, atol=1e-4)


def test_fused_pcd_falls_back_for_analytic_energies_and_matches_reference_golden():
    """Energies without an index-reading kernel run gather -> burst -> scatter inside the same C call."""
    import torchebm_b200 as te

    g = C.load("pcd_doublewell_fifo")
    model = te.DoubleWellModel(2.0, 1.0)
    sampler = te.LangevinDynamics(model, step_size=0.01, device=DEV)
    a = te.ContrastiveDivergence(model, sampler, k_steps=2, persistent=True, buffer_size=50, init_steps=0,
                                 new_sample_ratio=0.0, device=DEV)
    b = te.ContrastiveDivergence(model, te.LangevinDynamics(model, step_size=0.01, device=DEV), k_steps=2, persistent=True,
                                 buffer_size=50, init_steps=0, new_sample_ratio=0.0, device=DEV)
    ga, gb = torch.Generator(DEV).manual_seed(5), torch.Generator(DEV).manual_seed(5)
    data = g["data"].to(DEV)
    for it in range(data.shape[0]):
        neg_a = a.sample_negatives(data[it], generator=ga)
        start = b.get_start_points(data[it], generator=gb)
        neg_b = b.sampler.sample(x=start, n_steps=2, generator=gb)
        b.update_buffer(neg_b)
        assert torch.equal(neg_a, neg_b) and torch.equal(a.replay_buffer, b.replay_buffer)
        assert a._buffer_ptr_int == b._buffer_ptr_int == int(g["ptrs"][it])
