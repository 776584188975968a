/* C restatement of the RNG integer path.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Philox4x32-10 as published by Salmon et al. (SC'11) and shipped in cuRAND
 * (/usr/local/cuda/include/curand_philox4x32_x.h:88-192), plus the element -> (counter, component)
 * map of torch's CUDA distribution kernels (ATen/native/cuda/DistributionTemplates.h:50-91).
 * The third-party code itself (PyTorch 2.11.0 wheel, cuRAND headers of CUDA 12.x) is not under
 * /root/reference; this file pins the algorithm with the Random123 known-answer vectors
 * (tests/test_oracle_philox.py) and is cross-checked against oracle/philox.py.
 *
 * Build: gcc -O2 -shared -fPIC oracle/philox_ref.c -o oracle/_build/libphilox_ref.so
 */
#include <stdint.h>

#define M0 0xD2511F53u
#define M1 0xCD9E8D57u
#define W0 0x9E3779B9u
#define W1 0xBB67AE85u

void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* T = 256 * min(sm * (max_threads_per_sm / 256), ceil(numel / 256)) */
uint64_t torch_grid_threads(uint64_t numel, uint32_t sm, uint32_t max_threads_per_sm) {
  uint64_t grid = (numel + 255) / 256, cap = (uint64_t)sm * (max_threads_per_sm / 256);
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  return 256 * grid;
}

uint64_t torch_offset_increment(uint64_t numel, uint32_t sm, uint32_t max_threads_per_sm) {
  uint64_t t = torch_grid_threads(numel, sm, max_threads_per_sm);
  return ((numel - 1) / (4 * t) + 1) * 4;
}

/* The 32-bit word torch's layout assigns to element li of a numel-element draw at (seed, offset). */
uint32_t torch_word_for_element(uint64_t seed, uint64_t offset, uint64_t numel, uint64_t li, uint32_t sm,
                                uint32_t max_threads_per_sm) {
  uint64_t t = torch_grid_threads(numel, sm, max_threads_per_sm);
  uint64_t q = li / t, idx = li % t, c = offset / 4 + q / 4;
  uint32_t ctr[4] = {(uint32_t)c, (uint32_t)(c >> 32), (uint32_t)idx, (uint32_t)(idx >> 32)};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t out[4];
  philox4x32_10(ctr, key, out);
  return out[q % 4];
}
