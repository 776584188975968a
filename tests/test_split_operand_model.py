"""CPU model of the tensor-core kernels' split-operand products (no GPU needed): every operand v is carried as
hi + lo in bf16 -- activations with a TRUNCATED hi (one byte permute, csrc/mlp_tc_common.cuh:split2), weights with a
rounded hi (split_bf16) -- and a product is accumulated in fp32 as hi*hi + lo*hi + hi*lo.  The tests state the accuracy
class the GPU parity tests rely on (atol 2e-5 / rtol 1e-4 per burst) from the arithmetic alone."""

import torch


def _bf16_rn(v):
    return v.to(torch.bfloat16).to(torch.float32)


def _bf16_trunc(v):
    return (v.view(torch.int32) & -65536).view(torch.float32)


def _split(v, truncate):
    hi = _bf16_trunc(v) if truncate else _bf16_rn(v)
    return hi, _bf16_rn(v - hi)


def _product(a, w):
    """a[N,K] (activations) x w[M,K]^T (weights) the way the kernels issue it: three bf16 passes, fp32 accumulation."""
    ah, al = _split(a, truncate=True)
    wh, wl = _split(w, truncate=False)
    return ah @ wh.T + al @ wh.T + ah @ wl.T


def test_split_is_exact_to_16_bits_and_residual_is_small():
    torch.manual_seed(0)
    v = torch.randn(4096) * 3
    for trunc in (True, False):
        hi, lo = _split(v, trunc)
        assert torch.equal(_bf16_rn(hi), hi) and torch.equal(_bf16_rn(lo), lo)      # both parts are bf16 values
        rel = ((hi + lo) - v).abs() / v.abs()
        assert rel.max().item() <= 2.0 ** -15                                           # hi + lo carries >= 15 mantissa bits
    big = torch.tensor([torch.finfo(torch.float32).max, -torch.finfo(torch.float32).max])
    hi, lo = _split(big, True)    # the sanitised +-FLT_MAX of safe-mode HMC: a truncated hi stays finite
    assert torch.isfinite(hi).all() and torch.isfinite(lo).all()
    assert torch.isinf(_bf16_rn(big)).all()   # ... where a rounded hi would be inf (csrc: bf16x2_rn_finite keeps the truncated part)


def test_three_pass_product_is_in_the_2e5_class():
    torch.manual_seed(1)
    for n, k, m in ((256, 128, 128), (128, 784, 128)):
        a = torch.randn(n, k).clamp_(-3, 3)
        w = (torch.rand(m, k) * 2 - 1) / k ** 0.5          # nn.Linear's default init range
        exact = (a.double() @ w.double().T)
        got = _product(a, w).double()
        scale = exact.abs().max()
        assert ((got - exact).abs().max() / scale).item() < 2e-5
        one_pass = (_bf16_rn(a) @ _bf16_rn(w).T).double()
        assert ((one_pass - exact).abs().max() / scale).item() > 5e-4   # what precision="bf16" gives up


def test_mlp_gradient_through_split_products_matches_autograd():
    """The four products of a two-hidden-layer energy's gradient, all through the model above, against autograd in fp64."""
    torch.manual_seed(2)
    d, h, n = 128, 128, 512
    lin = [torch.nn.Linear(d, h), torch.nn.Linear(h, h), torch.nn.Linear(h, 1)]
    x = torch.randn(n, d).clamp_(-3, 3)
    with torch.no_grad():
        z1 = _product(x, lin[0].weight) + lin[0].bias
        s1 = torch.sigmoid(z1); h1 = z1 * s1; d1 = s1 * (1 + z1 * (1 - s1))
        z2 = _product(h1, lin[1].weight) + lin[1].bias
        s2 = torch.sigmoid(z2); d2 = s2 * (1 + z2 * (1 - s2))
        delta2 = d2 * lin[2].weight.reshape(-1)
        t = _product(delta2, lin[1].weight.T.contiguous())
        g = _product(t * d1, lin[0].weight.T.contiguous())
    xd = x.double().requires_grad_()
    net = torch.nn.Sequential(lin[0], torch.nn.SiLU(), lin[1], torch.nn.SiLU(), lin[2]).double()
    (want,) = torch.autograd.grad(net(xd).sum(), xd)
    net.float()
    err = (g.double() - want).abs().max().item()
    assert err < 2e-5 * max(1.0, want.abs().max().item()), err
