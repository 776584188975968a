// Row-structured kernels built on row_common.cuh: energy, gradient, Langevin burst (row-coupled
// energies), leapfrog, and the fused HMC proposal loop.
#pragma once
#include "langevin_elem.cuh"  // StepTable
#include "row_common.cuh"

namespace ebm {

template <int G, int EPT>
__device__ __forceinline__ void row_load(const float* base, long long row, int d, int g, bool valid, float (&x)[EPT]) {
#pragma unroll
  for (int m = 0; m < EPT; ++m) {
    const int col = g + G * m;
    x[m] = (valid && col < d) ? base[row * d + col] : 0.0f;
  }
}
template <int G, int EPT>
__device__ __forceinline__ void row_store(float* base, long long row, int d, int g, bool valid, const float (&x)[EPT]) {
#pragma unroll
  for (int m = 0; m < EPT; ++m) {
    const int col = g + G * m;
    if (valid && col < d) base[row * d + col] = x[m];
  }
}

// ---- energy / gradient -------------------------------------------------------------------------
template <class RowE, int G, int EPT>
__global__ void __launch_bounds__(kRowThreads) row_energy_grad_kernel(const RowE en, const float* __restrict__ x_in,
                                                                      long long n, int d, float* __restrict__ energy,
                                                                      float* __restrict__ grad, int scratch_stride) {
  extern __shared__ float smem[];
  const float* staged;
  RowCtx c = make_row_ctx<G>(en, smem, d, scratch_stride, staged);
  const long long groups_per_cta = kRowThreads / G;
  const long long stride = (long long)gridDim.x * groups_per_cta;
  const long long n_iter = (n + stride - 1) / stride;
  long long row = (long long)blockIdx.x * groups_per_cta + threadIdx.x / G;
  for (long long it = 0; it < n_iter; ++it, row += stride) {
    const bool valid = row < n;
    float x[EPT], g[EPT];
    row_load<G, EPT>(x_in, row, d, c.g, valid, x);
    const float e = row_eval<G, EPT>(en, x, g, c, energy != nullptr, staged);
    if (grad) row_store<G, EPT>(grad, row, d, c.g, valid, g);
    if (energy && valid && c.g == 0) energy[row] = e;
  }
}

// ---- Langevin burst for row-coupled energies (Gaussian, MoG) ----------------------------------
struct RowRng {
  uint32_t k0, k1;
  unsigned long long ctr_base;  // off/4
  unsigned long long ctr_step;  // per-step counter stride (TORCH: inc/4, NATIVE: 1)
  unsigned long long T;
  int mode;                     // 0 injected, 1 torch, 2 native
};

struct LangevinRowParams {
  const float* x_in;
  float* x_out;
  const float* noise;
  float* traj;
  long long n;
  int d, n_steps, thin, n_kept, thin_start, kept_base, has_clamp;
  float clamp_lo, clamp_hi;
  RowRng rng;
  int scratch_stride;
};

template <class RowE, int G, int EPT>
__global__ void __launch_bounds__(kRowThreads) langevin_row_kernel(const RowE en, const __grid_constant__ LangevinRowParams P,
                                                                   const __grid_constant__ StepTable tab) {
  extern __shared__ float smem[];
  const float* staged;
  RowCtx c = make_row_ctx<G>(en, smem, P.d, P.scratch_stride, staged);
  const long long groups_per_cta = kRowThreads / G;
  const long long stride = (long long)gridDim.x * groups_per_cta;
  const long long n_iter = (P.n + stride - 1) / stride;
  long long row = (long long)blockIdx.x * groups_per_cta + threadIdx.x / G;
  const long long numel = P.n * P.d;
  for (long long it = 0; it < n_iter; ++it, row += stride) {
    const bool valid = row < P.n;
    float x[EPT], g[EPT];
    row_load<G, EPT>(P.x_in, row, P.d, c.g, valid, x);
    int until_keep = P.thin_start, kept = P.kept_base;
    RngStream rs;
    rs.k0 = P.rng.k0; rs.k1 = P.rng.k1; rs.T = P.rng.T; rs.mode = P.rng.mode; rs.ctr_base = P.rng.ctr_base;
    for (int k = 0; k < P.n_steps; ++k) {
      const int ti = k & tab.mask;
      const float h = tab.h[ti], c1 = tab.c1[ti], c2 = tab.c2[ti];
      row_eval<G, EPT>(en, x, g, c, false, staged);
#pragma unroll
      for (int m = 0; m < EPT; ++m) {
        const int col = c.g + G * m;
        const bool in = valid && col < P.d;
        const long long li = row * P.d + col;
        float eps = 0.0f;
        if (in) eps = (P.rng.mode == 0) ? P.noise[(long long)k * numel + li] : normal_for_element(rs, (uint64_t)li);
        const float x1 = __fsub_rn(x[m], __fmul_rn(h, g[m]));
        float xn = __fadd_rn(x1, __fmul_rn(c2, __fmul_rn(eps, c1)));
        if (P.has_clamp) xn = clamp_torch(xn, P.clamp_lo, P.clamp_hi);
        x[m] = in ? xn : 0.0f;
      }
      rs.ctr_base += P.rng.ctr_step;
      if (P.traj && --until_keep == 0) {
        until_keep = P.thin;
        if (kept < P.n_kept) {
#pragma unroll
          for (int m = 0; m < EPT; ++m) {
            const int col = c.g + G * m;
            if (valid && col < P.d) P.traj[(row * P.n_kept + kept) * P.d + col] = x[m];
          }
        }
        ++kept;
      }
    }
    row_store<G, EPT>(P.x_out, row, P.d, c.g, valid, x);
  }
}

// ---- leapfrog ----------------------------------------------------------------------------------
struct MassSpec {
  int kind;            // EBM_MASS_*
  float safe_scalar;   // (float)max(mass, 1e-10)
  float scalar;        // (float)mass            (kinetic energy divides by the raw mass, hmc.py:151)
  float sqrt_scalar;   // (float)sqrt(mass)
  const float* vec;    // [d]
};

constexpr float kSafeClamp = 1e6f;  // base_integrator.py:847

// force = -grad E, clamped in safe mode.  clamp_ propagates NaN like torch.
template <int EPT>
__device__ __forceinline__ void to_force(float (&g)[EPT], bool safe) {
#pragma unroll
  for (int m = 0; m < EPT; ++m) {
    float f = -g[m];
    if (safe) f = clamp_torch(f, -kSafeClamp, kSafeClamp);
    g[m] = f;
  }
}

// L leapfrog steps on (x, p) with the force f at x carried in and out (leapfrog.py:160-185).
// The reference evaluates the force at the top of every step; that value equals the force computed at
// the bottom of the previous step unless x was sanitised in between, in which case it is recomputed here.
template <class RowE, int G, int EPT>
__device__ __forceinline__ void leapfrog_steps(const RowE& en, const RowCtx& c, const float* staged, float (&x)[EPT],
                                               float (&p)[EPT], float (&f)[EPT], float h, int n_steps,
                                               const MassSpec& ms, const float (&minv)[EPT], bool safe) {
  const float half_h = __fmul_rn(0.5f, h);
  for (int l = 0; l < n_steps; ++l) {
    float ph[EPT];
#pragma unroll
    for (int m = 0; m < EPT; ++m) {
      ph[m] = __fadd_rn(p[m], __fmul_rn(half_h, f[m]));
      float dx = __fmul_rn(h, ph[m]);
      if (ms.kind == 1) dx = __fdiv_rn(dx, ms.safe_scalar);
      else if (ms.kind == 2) dx = __fdiv_rn(dx, minv[m]);
      x[m] = __fadd_rn(x[m], dx);
    }
    row_eval<G, EPT>(en, x, f, c, false, staged);
    to_force<EPT>(f, safe);
    bool dirty = false;
#pragma unroll
    for (int m = 0; m < EPT; ++m) {
      p[m] = __fadd_rn(ph[m], __fmul_rn(half_h, f[m]));
      if (safe) {
        const float xs = nan_to_num0(x[m]), ps = nan_to_num0(p[m]);
        // compare bit patterns: NaN != NaN would flag, which is what we want; +-inf -> FLT_MAX also flags
        dirty |= (__float_as_uint(xs) != __float_as_uint(x[m]));
        x[m] = xs;
        p[m] = ps;
      }
    }
    // warp-uniform decision: row_eval uses full-warp shuffles, and recomputing the force of a clean
    // row reproduces the same value, so recompute for the whole warp when any of its rows was sanitised
    if (safe && __any_sync(0xffffffffu, dirty)) {
      row_eval<G, EPT>(en, x, f, c, false, staged);
      to_force<EPT>(f, safe);
    }
  }
}

template <int G, int EPT>
__device__ __forceinline__ void load_mass(const MassSpec& ms, int g, int d, float (&minv)[EPT], float (&msqrt)[EPT],
                                          float (&mraw)[EPT]) {
#pragma unroll
  for (int m = 0; m < EPT; ++m) {
    const int col = g + G * m;
    float v = 1.0f;
    if (ms.kind == 2 && col < d) v = ms.vec[col];
    mraw[m] = v;
    minv[m] = fmaxf(v, 1e-10f);  // torch.clamp(mass, min=1e-10), leapfrog.py:174
    msqrt[m] = sqrtf(v);         // torch.sqrt(mass), hmc.py:129
  }
}

struct LeapfrogParams {
  const float* x_in;
  const float* p_in;
  float* x_out;
  float* p_out;
  long long n;
  int d, n_steps, safe;
  float h;
  MassSpec mass;
  int scratch_stride;
};

template <class RowE, int G, int EPT>
__global__ void __launch_bounds__(kRowThreads) leapfrog_kernel(const RowE en, const __grid_constant__ LeapfrogParams P) {
  extern __shared__ float smem[];
  const float* staged;
  RowCtx c = make_row_ctx<G>(en, smem, P.d, P.scratch_stride, staged);
  const long long groups_per_cta = kRowThreads / G;
  const long long stride = (long long)gridDim.x * groups_per_cta;
  const long long n_iter = (P.n + stride - 1) / stride;
  long long row = (long long)blockIdx.x * groups_per_cta + threadIdx.x / G;
  float minv[EPT], msqrt[EPT], mraw[EPT];
  load_mass<G, EPT>(P.mass, c.g, P.d, minv, msqrt, mraw);
  for (long long it = 0; it < n_iter; ++it, row += stride) {
    const bool valid = row < P.n;
    float x[EPT], p[EPT], f[EPT];
    row_load<G, EPT>(P.x_in, row, P.d, c.g, valid, x);
    row_load<G, EPT>(P.p_in, row, P.d, c.g, valid, p);
    row_eval<G, EPT>(en, x, f, c, false, staged);
    to_force<EPT>(f, P.safe != 0);
    leapfrog_steps<RowE, G, EPT>(en, c, staged, x, p, f, P.h, P.n_steps, P.mass, minv, P.safe != 0);
    row_store<G, EPT>(P.x_out, row, P.d, c.g, valid, x);
    row_store<G, EPT>(P.p_out, row, P.d, c.g, valid, p);
  }
}

// ---- HMC ---------------------------------------------------------------------------------------
struct HmcParams {
  const float* x_in;
  float* x_out;
  const float* noise_p;  // INJECTED [n_prop, n, d]
  const float* noise_u;  // INJECTED [n_prop, n]
  float* traj;           // [n, n_kept, d] or null
  int* accept_count;     // [n_prop] or null
  float* energy_out;     // [n] or null
  long long n;
  int d, n_prop, n_leapfrog, thin, n_kept, thin_start, kept_base, prop_base;
  MassSpec mass;
  RowRng rng_p;          // momentum stream; per-proposal counter stride in ctr_step
  RowRng rng_u;          // uniform stream (numel = n)
  int scratch_stride;
};

struct HStepTable {
  float h[kSchedChunk];
  int mask;
};

template <int G, int EPT>
__device__ __forceinline__ float kinetic_energy(const float (&p)[EPT], const MassSpec& ms, const float (&mraw)[EPT]) {
  // hmc.py:148-159
  float s = 0.0f;
#pragma unroll
  for (int m = 0; m < EPT; ++m) {
    const float sq = __fmul_rn(p[m], p[m]);
    s += (ms.kind == 2) ? __fdiv_rn(sq, mraw[m]) : sq;
  }
  s = __fmul_rn(0.5f, group_sum<G>(s));
  if (ms.kind == 1) s = __fdiv_rn(s, ms.scalar);
  return clamp_torch(s, 0.0f, 1e10f);
}

template <class RowE, int G, int EPT>
__global__ void __launch_bounds__(kRowThreads) hmc_kernel(const RowE en, const __grid_constant__ HmcParams P,
                                                          const __grid_constant__ HStepTable tab) {
  extern __shared__ float smem[];
  const float* staged;
  RowCtx c = make_row_ctx<G>(en, smem, P.d, P.scratch_stride, staged);
  const long long groups_per_cta = kRowThreads / G;
  const long long stride = (long long)gridDim.x * groups_per_cta;
  const long long n_iter = (P.n + stride - 1) / stride;
  long long row = (long long)blockIdx.x * groups_per_cta + threadIdx.x / G;
  const long long numel = P.n * P.d;
  float minv[EPT], msqrt[EPT], mraw[EPT];
  load_mass<G, EPT>(P.mass, c.g, P.d, minv, msqrt, mraw);

  for (long long it = 0; it < n_iter; ++it, row += stride) {
    const bool valid = row < P.n;
    float x[EPT], f[EPT];
    row_load<G, EPT>(P.x_in, row, P.d, c.g, valid, x);
    float e_cur = row_eval<G, EPT>(en, x, f, c, true, staged);  // E(x) and grad at the chain state
    to_force<EPT>(f, true);
    RngStream rp, ru;
    rp.k0 = P.rng_p.k0; rp.k1 = P.rng_p.k1; rp.T = P.rng_p.T; rp.mode = P.rng_p.mode; rp.ctr_base = P.rng_p.ctr_base;
    ru.k0 = P.rng_u.k0; ru.k1 = P.rng_u.k1; ru.T = P.rng_u.T; ru.mode = P.rng_u.mode; ru.ctr_base = P.rng_u.ctr_base;
    int until_keep = P.thin_start, kept = P.kept_base;

    for (int i = 0; i < P.n_prop; ++i) {
      const float h = tab.h[i & tab.mask];
      float p[EPT], xs[EPT], fs[EPT];
#pragma unroll
      for (int m = 0; m < EPT; ++m) {
        const int col = c.g + G * m;
        const bool in = valid && col < P.d;
        const long long li = row * P.d + col;
        float eps = 0.0f;
        if (in) eps = (P.rng_p.mode == 0) ? P.noise_p[(long long)i * numel + li] : normal_for_element(rp, (uint64_t)li);
        if (P.mass.kind == 1) eps = __fmul_rn(eps, P.mass.sqrt_scalar);       // hmc.py:124
        else if (P.mass.kind == 2) eps = __fmul_rn(eps, msqrt[m]);            // hmc.py:133
        p[m] = eps;
        xs[m] = x[m];
        fs[m] = f[m];
      }
      // H0 (hmc.py:247-256)
      const float h0 = __fadd_rn(clamp_torch(e_cur, -1e10f, 1e10f), kinetic_energy<G, EPT>(p, P.mass, mraw));
      leapfrog_steps<RowE, G, EPT>(en, c, staged, x, p, f, h, P.n_leapfrog, P.mass, minv, true);
      // H1 (hmc.py:268-275); f already holds the force at the proposal
      float gtmp[EPT];
      const float e_new = row_eval<G, EPT>(en, x, gtmp, c, true, staged);
      const float h1 = __fadd_rn(clamp_torch(e_new, -1e10f, 1e10f), kinetic_energy<G, EPT>(p, P.mass, mraw));
      const float dh = clamp_torch(__fsub_rn(h0, h1), -50.0f, 50.0f);
      float a = expf(dh);
      a = (a != a) ? a : fminf(a, 1.0f);  // clamp_(max=1.0) keeps NaN
      float u = 0.0f;
      if (valid) u = (P.rng_u.mode == 0) ? P.noise_u[(long long)i * P.n + row] : uniform_for_element(ru, (uint64_t)row);
      const bool accepted = u < a;
      if (accepted) {
        e_cur = e_new;
      } else {
#pragma unroll
        for (int m = 0; m < EPT; ++m) { x[m] = xs[m]; f[m] = fs[m]; }
      }
      if (P.accept_count && valid && c.g == 0 && accepted) atomicAdd(P.accept_count + P.prop_base + i, 1);
      rp.ctr_base += P.rng_p.ctr_step;
      ru.ctr_base += P.rng_u.ctr_step;
      if (P.traj && --until_keep == 0) {
        until_keep = P.thin;
        if (kept < P.n_kept) {
#pragma unroll
          for (int m = 0; m < EPT; ++m) {
            const int col = c.g + G * m;
            if (valid && col < P.d) P.traj[(row * P.n_kept + kept) * P.d + col] = x[m];
          }
        }
        ++kept;
      }
    }
    row_store<G, EPT>(P.x_out, row, P.d, c.g, valid, x);
    if (P.energy_out && valid && c.g == 0) P.energy_out[row] = clamp_torch(e_cur, -1e10f, 1e10f);
  }
}

}  // namespace ebm
