"""Pin the RNG restatement: Random123 known-answer vectors, numpy vs C, layout arithmetic."""

import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import philox as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# Random123 kat_vectors, philox4x32 with 10 rounds: (ctr, key, expected)
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


@pytest.fixture(scope="module")
def clib():
    out_dir = os.path.join(ROOT, "oracle", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libphilox_ref.so")
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", os.path.join(ROOT, "oracle", "philox_ref.c"), "-o", so])
    lib = ctypes.CDLL(so)
    lib.torch_word_for_element.restype = ctypes.c_uint32
    lib.torch_word_for_element.argtypes = [ctypes.c_uint64] * 4 + [ctypes.c_uint32] * 2
    lib.torch_offset_increment.restype = ctypes.c_uint64
    lib.torch_offset_increment.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32]
    return lib


@pytest.mark.parametrize("ctr,key,want", KAT)
def test_known_answer_vectors_numpy(ctr, key, want):
    got = P.philox4x32_10(*ctr, *key)
    assert tuple(int(x) for x in got) == want


@pytest.mark.parametrize("ctr,key,want", KAT)
def test_known_answer_vectors_c(clib, ctr, key, want):
    out = (ctypes.c_uint32 * 4)()
    clib.philox4x32_10((ctypes.c_uint32 * 4)(*ctr), (ctypes.c_uint32 * 2)(*key), out)
    assert tuple(out) == want


def test_torch_layout_numpy_matches_c(clib):
    seed, offset, numel = 0x1234567890ABCDEF, 4096, 700_000
    w, ii = P._torch_words(seed, offset, numel, 148, 2048)
    words = np.stack(w, 0)[ii, np.arange(numel)]
    for li in (0, 1, 255, 256, 303103, 303104, 303105, 606208, 699_999):
        assert int(words[li]) == clib.torch_word_for_element(seed, offset, numel, li, 148, 2048)
    for n in (1, 255, 256, 303104, 303105, 4 * 303104, 4 * 303104 + 1, 8388608):
        assert P.torch_offset_increment(n) == clib.torch_offset_increment(n, 148, 2048)
    assert P.torch_offset_increment(8388608) == 28  # C2: 65536 x 128 elements -> 7 curand_normal4 calls per thread


def test_first_normals_of_torch_seed_zero():
    """torch.manual_seed(0); torch.randn(4, device='cuda') is the well-known [-0.9247, -0.4253, -2.6438, 0.1452]."""
    x = P.torch_cuda_randn(0, 0, 4)
    np.testing.assert_allclose(x, [-0.9247, -0.4253, -2.6438, 0.1452], atol=5e-5)


def test_uniform_range_and_moments():
    u = P.torch_cuda_rand(3, 0, 200_000)
    assert u.min() >= 0.0 and u.max() < 1.0 and abs(u.mean() - 0.5) < 5e-3
    n = P.native_randn(3, 7, 200_000)
    assert abs(n.mean()) < 1e-2 and abs(n.std() - 1) < 1e-2
