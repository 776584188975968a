"""The fused kernels' Philox streams, through the C ABI test hook `ebm_rng_fill_f32`.

TORCH mode must reproduce torch's own CUDA `randn` / `rand` bit for bit (this is what makes
`sampler.sample(generator=g)` seed-identical to the reference PyTorch-CUDA sampler); NATIVE mode must
match the numpy restatement in oracle/philox.py."""

import numpy as np
import pytest
import torch

from oracle import philox as P

pytestmark = pytest.mark.gpu


def _gen(seed, offset=0):
    g = torch.Generator("cuda")
    g.manual_seed(seed)
    if offset:
        g.set_offset(offset)
    return g


@pytest.mark.parametrize("numel", [1, 4, 255, 1000, 65536, 303104, 303105, 1 << 20, 5_000_001])
@pytest.mark.parametrize("seed,offset", [(0, 0), (1234567890123, 0), (7, 4096)])
def test_torch_mode_normal_is_bitwise_torch_randn(numel, seed, offset):
    from torchebm_b200 import _lib, ops

    want = torch.randn(numel, device="cuda", generator=_gen(seed, offset))
    got = ops.rng_fill(numel, "cuda", _lib.RNG_TORCH, 0, seed, offset)
    assert torch.equal(got, want)


@pytest.mark.parametrize("numel", [1, 1000, 303104 * 4 + 17, 1 << 21])
def test_torch_mode_uniform_is_bitwise_torch_rand(numel):
    from torchebm_b200 import _lib, ops

    want = torch.rand(numel, device="cuda", generator=_gen(99, 8))
    got = ops.rng_fill(numel, "cuda", _lib.RNG_TORCH, 1, 99, 8)
    assert torch.equal(got, want)


@pytest.mark.parametrize("numel", [10, 1 << 16, 1 << 22])
def test_offset_increment_matches_torch_generator(numel):
    from torchebm_b200 import ops

    g = _gen(5)
    torch.randn(numel, device="cuda", generator=g)
    assert g.get_offset() == ops.torch_offset_increment("cuda", numel)
    props = torch.cuda.get_device_properties(0)
    assert ops.torch_offset_increment("cuda", numel) == P.torch_offset_increment(
        numel, props.multi_processor_count, props.max_threads_per_multi_processor)


def test_torch_mode_matches_numpy_oracle():
    """The numpy restatement of torch's layout (used by CPU-side tests) agrees with the device."""
    from torchebm_b200 import _lib, ops

    props = torch.cuda.get_device_properties(0)
    numel = 400_000
    got = ops.rng_fill(numel, "cuda", _lib.RNG_TORCH, 0, 42, 16).cpu().numpy()
    want = P.torch_cuda_randn(42, 16, numel, props.multi_processor_count, props.max_threads_per_multi_processor)
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-6)
    gu = ops.rng_fill(numel, "cuda", _lib.RNG_TORCH, 1, 42, 16).cpu().numpy()
    wu = P.torch_cuda_rand(42, 16, numel, props.multi_processor_count, props.max_threads_per_multi_processor)
    np.testing.assert_array_equal(gu, wu)


def test_native_mode_matches_numpy_oracle():
    from torchebm_b200 import _lib, ops

    numel = 100_003
    got = ops.rng_fill(numel, "cuda", _lib.RNG_NATIVE, 0, 2024, 40).cpu().numpy()
    want = P.native_randn(2024, 10, numel)
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-6)
    assert abs(got.mean()) < 0.02 and abs(got.std() - 1.0) < 0.02
