// Row-structured kernels: one chain (row of D floats) is owned by a group of G lanes of one warp,
// lane g holding columns c = g + G*m, m < EPT.  Used where the path needs a per-row quantity:
// energies (HMC accept test, diagnostics), the Gaussian / mixture gradients, leapfrog, HMC.
#pragma once
#include "energies.cuh"
#include "rng.cuh"

namespace ebm {

constexpr int kRowThreads = 256;

template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int G>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <int G>
__device__ __forceinline__ bool group_any(bool b) {
  unsigned m = __ballot_sync(0xffffffffu, b);
  if (G == 32) return m != 0;
  const int lane = threadIdx.x & 31;
  const unsigned gm = ((G == 32) ? 0xffffffffu : ((1u << G) - 1u)) << (lane & ~(G - 1));
  return (m & gm) != 0;
}

// Per-thread view of its row group.
struct RowCtx {
  int g;           // lane index inside the group
  int d;           // row length
  float* scratch;  // shared memory private to the group (>= max(d, K) floats), may be null
};

// ---- row energies -----------------------------------------------------------------------------
// interface: template<int G,int EPT> float eval(const float (&x)[EPT], float (&grad)[EPT], RowCtx&, bool want_e)
//   fills grad (dE/dx for the owned columns; 0 for padding columns) and returns E(row) on every lane
//   of the group when want_e (else unspecified).

template <class ElemE>
struct ElemRow {
  ElemE e;
  static constexpr int kSharedPerCta = 0;
  __device__ __forceinline__ void stage(float*, int) const {}
  template <int G, int EPT>
  __device__ __forceinline__ float eval(const float (&x)[EPT], float (&grad)[EPT], const RowCtx& c, bool want_e) const {
    float s = 0.0f;
#pragma unroll
    for (int m = 0; m < EPT; ++m) {
      const bool in = (c.g + G * m) < c.d;
      grad[m] = in ? e.grad(x[m]) : 0.0f;
      if (want_e) s += in ? e.term(x[m]) : 0.0f;
    }
    if (!want_e) return 0.0f;
    return e.finish(group_sum<G>(s));
  }
};

// Gaussian: E = 0.5 * delta^T A delta (base_model.py:181-210).  Staged in shared memory:
// S = 0.5 * (A + A^T) [D, D] and mean[D]; grad = delta * S (the reference's autograd yields
// 0.5*(delta A + delta A^T); same value up to summation order).
struct GaussianRow {
  const float* mean;
  const float* cov_inv;
  int d;
  __device__ __forceinline__ int shared_floats() const { return d * d + d; }
  __device__ __forceinline__ void stage(float* sm, int tid, int nthreads) const {
    for (int i = tid; i < d * d; i += nthreads) {
      const int r = i / d, c = i - r * d;
      sm[i] = 0.5f * (cov_inv[r * d + c] + cov_inv[c * d + r]);
    }
    for (int i = tid; i < d; i += nthreads) sm[d * d + i] = mean[i];
  }
  template <int G, int EPT>
  __device__ __forceinline__ float eval(const float (&x)[EPT], float (&grad)[EPT], const RowCtx& c, bool want_e,
                                        const float* sm) const {
    const float* S = sm;
    const float* mu = sm + d * d;
    float delta[EPT];
#pragma unroll
    for (int m = 0; m < EPT; ++m) {
      const int col = c.g + G * m;
      delta[m] = (col < d) ? __fsub_rn(x[m], mu[col]) : 0.0f;
      if (col < d) c.scratch[col] = delta[m];
      grad[m] = 0.0f;
    }
    __syncwarp();
    for (int i = 0; i < d; ++i) {
      const float di = c.scratch[i];
#pragma unroll
      for (int m = 0; m < EPT; ++m) {
        const int col = c.g + G * m;
        if (col < d) grad[m] = fmaf(di, S[i * d + col], grad[m]);
      }
    }
    __syncwarp();
    if (!want_e) return 0.0f;
    float s = 0.0f;
#pragma unroll
    for (int m = 0; m < EPT; ++m) s = fmaf(delta[m], grad[m], s);
    return 0.5f * group_sum<G>(s);
  }
};

// Isotropic mixture (not in the reference; oracle/energies.py:MixtureOfGaussians):
// E = -logsumexp_k( log w_k - D log s_k - |x - mu_k|^2 / (2 s_k^2) ).
// Shared: mu[K, D], a[K] = log w_k - D log s_k, iv[K] = 1 / s_k^2.
struct MogRow {
  const float* means;
  const float* sigmas;
  const float* weights;
  int d, k;
  __device__ __forceinline__ int shared_floats() const { return k * d + 2 * k; }
  __device__ __forceinline__ void stage(float* sm, int tid, int nthreads) const {
    for (int i = tid; i < k * d; i += nthreads) sm[i] = means[i];
    for (int i = tid; i < k; i += nthreads) {
      const float s = sigmas[i];
      sm[k * d + i] = logf(weights[i]) - (float)d * logf(s);
      sm[k * d + k + i] = 1.0f / (s * s);
    }
  }
  template <int G, int EPT>
  __device__ __forceinline__ float eval(const float (&x)[EPT], float (&grad)[EPT], const RowCtx& c, bool want_e,
                                        const float* sm) const {
    const float* mu = sm;
    const float* a = sm + k * d;
    const float* iv = a + k;
    float mx = -INFINITY;
    for (int j = 0; j < k; ++j) {
      float s = 0.0f;
#pragma unroll
      for (int m = 0; m < EPT; ++m) {
        const int col = c.g + G * m;
        const float df = (col < d) ? (x[m] - mu[j * d + col]) : 0.0f;
        s = fmaf(df, df, s);
      }
      s = group_sum<G>(s);
      const float lg = a[j] - 0.5f * s * iv[j];
      if (c.g == 0) c.scratch[j] = lg;
      mx = fmaxf(mx, lg);
    }
    __syncwarp();
    float z = 0.0f;
#pragma unroll
    for (int m = 0; m < EPT; ++m) grad[m] = 0.0f;
    for (int j = 0; j < k; ++j) {
      const float w = expf(c.scratch[j] - mx);
      z += w;
      const float wi = w * iv[j];
#pragma unroll
      for (int m = 0; m < EPT; ++m) {
        const int col = c.g + G * m;
        if (col < d) grad[m] = fmaf(wi, x[m] - mu[j * d + col], grad[m]);
      }
    }
    __syncwarp();
    const float rz = 1.0f / z;
#pragma unroll
    for (int m = 0; m < EPT; ++m) grad[m] *= rz;
    return -(mx + logf(z));
  }
};

// uniform call wrapper so kernels need not care whether an energy uses staged shared memory
template <int G, int EPT, class ElemE>
__device__ __forceinline__ float row_eval(const ElemRow<ElemE>& en, const float (&x)[EPT], float (&grad)[EPT],
                                          const RowCtx& c, bool want_e, const float*) {
  return en.template eval<G, EPT>(x, grad, c, want_e);
}
template <int G, int EPT>
__device__ __forceinline__ float row_eval(const GaussianRow& en, const float (&x)[EPT], float (&grad)[EPT],
                                          const RowCtx& c, bool want_e, const float* sm) {
  return en.template eval<G, EPT>(x, grad, c, want_e, sm);
}
template <int G, int EPT>
__device__ __forceinline__ float row_eval(const MogRow& en, const float (&x)[EPT], float (&grad)[EPT],
                                          const RowCtx& c, bool want_e, const float* sm) {
  return en.template eval<G, EPT>(x, grad, c, want_e, sm);
}

template <class ElemE>
__device__ __forceinline__ int row_shared_floats(const ElemRow<ElemE>&) { return 0; }
__device__ __forceinline__ int row_shared_floats(const GaussianRow& e) { return e.shared_floats(); }
__device__ __forceinline__ int row_shared_floats(const MogRow& e) { return e.shared_floats(); }
template <class ElemE>
__device__ __forceinline__ void row_stage(const ElemRow<ElemE>&, float*, int, int) {}
__device__ __forceinline__ void row_stage(const GaussianRow& e, float* sm, int t, int n) { e.stage(sm, t, n); }
__device__ __forceinline__ void row_stage(const MogRow& e, float* sm, int t, int n) { e.stage(sm, t, n); }

// Set up the per-thread row context.  Dynamic shared memory layout:
//   [ staged energy parameters : row_shared_floats ][ per-group scratch : groups * scratch_stride ]
template <int G, class RowE>
__device__ __forceinline__ RowCtx make_row_ctx(const RowE& en, float* smem, int d, int scratch_stride,
                                               const float*& staged) {
  row_stage(en, smem, threadIdx.x, blockDim.x);
  staged = smem;
  RowCtx c;
  c.g = threadIdx.x & (G - 1);
  c.d = d;
  c.scratch = smem + row_shared_floats(en) + (threadIdx.x / G) * scratch_stride;
  __syncthreads();
  return c;
}

}  // namespace ebm
