"""Oracle for the RNG streams.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The random numbers of the reference's CUDA path come from a third-party dependency that is
not under /root/reference: PyTorch 2.11.0 (pip wheel, `torch>=2.10` in pyproject.toml:52-54),
which draws Philox4x32-10 through cuRAND device headers (CUDA 12.x).  This file restates the
published algorithm and torch's thread/element layout:

* Philox4x32-10: Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11);
  cuRAND's `curand_Philox4x32_10` (/usr/local/cuda/include/curand_philox4x32_x.h:160-192).
  Pinned by the Random123 known-answer vectors in tests/test_oracle_philox.py.
* `curand_init(seed, subsequence, offset)` + `curand4` (curand_kernel.h:926-1040): the j-th
  `curand4` of a fresh state returns Philox(ctr = (lo(offset/4 + j), hi(offset/4 + j),
  lo(subsequence), hi(subsequence)), key = (lo(seed), hi(seed))) when offset % 4 == 0.
* Box-Muller `_curand_box_muller` (curand_normal.h:70-87) and `_curand_uniform4`
  (curand_uniform.h:74-82).
* torch's grid-stride layout, ATen/native/cuda/DistributionTemplates.h:50-91: block 256,
  grid = min(SMs * (maxThreadsPerSM / 256), ceil(numel / 256)), T = 256 * grid; thread `idx`
  uses subsequence `idx`; its j-th `curand_normal4` feeds elements idx + T * (4 j + ii), ii = 0..3.
  Each call advances the generator offset by ((numel - 1) // (4 T) + 1) * 4.
  `torch.rand` additionally maps 1.0 -> 0.0 (DistributionTemplates.h:485-500).

The integer part (Philox words) is exact.  The float transforms use numpy float32 libm, so
normals agree with the GPU's `logf/sqrtf/__sincosf` only to a few ulp; GPU tests that need
bit-exactness compare against `torch.randn` on the device instead.
"""

from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

TWO_POW32_INV = np.float32(2.3283064e-10)
TWO_POW32_INV_2PI = np.float32(2.3283064e-10) * np.float32(6.2831855)

B200_SM_COUNT = 148
B200_MAX_THREADS_PER_SM = 2048


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy uint32 arrays (counters) with scalar keys."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def box_muller(x, y):
    """_curand_box_muller: returns (sin(v) * s, cos(v) * s) in float32."""
    x = np.asarray(x, dtype=np.uint32).astype(np.float32)
    y = np.asarray(y, dtype=np.uint32).astype(np.float32)
    u = x * TWO_POW32_INV + TWO_POW32_INV / np.float32(2)
    # device code contracts y * C + C/2 into one FMA; emulate with float64 then round once
    v = (y.astype(np.float64) * np.float64(TWO_POW32_INV_2PI) + np.float64(TWO_POW32_INV_2PI / np.float32(2))).astype(np.float32)
    s = np.sqrt(np.float32(-2.0) * np.log(u)).astype(np.float32)
    return (np.sin(v) * s).astype(np.float32), (np.cos(v) * s).astype(np.float32)


def uniform(x):
    """_curand_uniform: (0, 1]."""
    return np.asarray(x, dtype=np.uint32).astype(np.float32) * TWO_POW32_INV + TWO_POW32_INV / np.float32(2)


def torch_grid_threads(numel: int, sm_count: int = B200_SM_COUNT, max_threads_per_sm: int = B200_MAX_THREADS_PER_SM) -> int:
    """T = 256 * grid of calc_execution_policy (DistributionTemplates.h:50-62)."""
    grid = min(sm_count * (max_threads_per_sm // 256), (numel + 255) // 256)
    return 256 * max(grid, 1)


def torch_offset_increment(numel: int, sm_count: int = B200_SM_COUNT, max_threads_per_sm: int = B200_MAX_THREADS_PER_SM) -> int:
    t = torch_grid_threads(numel, sm_count, max_threads_per_sm)
    return ((numel - 1) // (t * 4) + 1) * 4


def _torch_words(seed: int, offset: int, numel: int, sm_count: int, max_threads_per_sm: int):
    assert offset % 4 == 0
    t = torch_grid_threads(numel, sm_count, max_threads_per_sm)
    li = np.arange(numel, dtype=np.uint64)
    idx = li % np.uint64(t)
    q = li // np.uint64(t)
    j = q // np.uint64(4)
    ii = (q % np.uint64(4)).astype(np.int64)
    ctr = np.uint64(offset // 4) + j
    w = philox4x32_10(ctr & MASK, ctr >> np.uint64(32), idx & MASK, idx >> np.uint64(32), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return w, ii


def torch_cuda_randn(seed: int, offset: int, numel: int, sm_count: int = B200_SM_COUNT,
                     max_threads_per_sm: int = B200_MAX_THREADS_PER_SM) -> np.ndarray:
    """What `torch.randn(numel, device='cuda', generator=g)` yields for g at (seed, offset)."""
    w, ii = _torch_words(seed, offset, numel, sm_count, max_threads_per_sm)
    n0, n1 = box_muller(w[0], w[1])
    n2, n3 = box_muller(w[2], w[3])
    out = np.stack([n0, n1, n2, n3], axis=0)
    return out[ii, np.arange(numel)]


def torch_cuda_rand(seed: int, offset: int, numel: int, sm_count: int = B200_SM_COUNT,
                    max_threads_per_sm: int = B200_MAX_THREADS_PER_SM) -> np.ndarray:
    """What `torch.rand(numel, device='cuda', generator=g)` yields: [0, 1)."""
    w, ii = _torch_words(seed, offset, numel, sm_count, max_threads_per_sm)
    u = np.stack([uniform(x) for x in w], axis=0)[ii, np.arange(numel)]
    return np.where(u == np.float32(1.0), np.float32(0.0), u).astype(np.float32)


# ---- "native" stream of the CUDA library (torchebm_b200/csrc/rng.cuh) -------------------------
# Each aligned quad of 4 consecutive elements q = li // 4 draws one Philox block with
# ctr = (lo(q), hi(q), lo(step), hi(step)), key = (lo(seed) ^ TAG0, hi(seed) ^ TAG1), step = offset/4 + k.
NATIVE_TAG0 = 0x42323030  # "B200"
NATIVE_TAG1 = 0x45424D21  # "EBM!"


def native_words(seed: int, step: int, n_quads: int):
    q = np.arange(n_quads, dtype=np.uint64)
    st = np.full(n_quads, step, dtype=np.uint64)
    return philox4x32_10(q & MASK, q >> np.uint64(32), st & MASK, st >> np.uint64(32),
                         (seed & 0xFFFFFFFF) ^ NATIVE_TAG0, ((seed >> 32) & 0xFFFFFFFF) ^ NATIVE_TAG1)


def native_randn(seed: int, step: int, numel: int) -> np.ndarray:
    nq = (numel + 3) // 4
    w = native_words(seed, step, nq)
    n0, n1 = box_muller(w[0], w[1])
    n2, n3 = box_muller(w[2], w[3])
    return np.stack([n0, n1, n2, n3], axis=1).reshape(-1)[:numel]
