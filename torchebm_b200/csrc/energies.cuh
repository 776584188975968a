// Energy functors.  Gradients are written in the reference's fp32 rounding order (autograd of
// torchebm/core/base_model.py forwards): every product and sum is rounded
// separately (__fmul_rn / __fadd_rn never contract into FMA), because eager PyTorch runs one kernel
// per op.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace ebm {

// ---- elementwise energies: E(x) = scale(sum_i term(x_i)) , dE/dx_i = grad(x_i) -------------------

struct DoubleWellE {  // base_model.py:143-148
  float h, b2;
  __device__ __forceinline__ float grad(float x) const {
    const float u = __fsub_rn(__fmul_rn(x, x), b2);
    return __fmul_rn(__fmul_rn(h, __fmul_rn(2.0f, u)), __fmul_rn(2.0f, x));
  }
  __device__ __forceinline__ float term(float x) const {
    const float u = __fsub_rn(__fmul_rn(x, x), b2);
    return __fmul_rn(u, u);
  }
  __device__ __forceinline__ float finish(float sum) const { return __fmul_rn(h, sum); }
};

struct HarmonicE {  // base_model.py:224-229 ; half_k = (float)(0.5*k)
  float half_k;
  __device__ __forceinline__ float grad(float x) const { return __fmul_rn(half_k, __fmul_rn(2.0f, x)); }
  __device__ __forceinline__ float term(float x) const { return __fmul_rn(x, x); }
  __device__ __forceinline__ float finish(float sum) const { return __fmul_rn(half_k, sum); }
};

struct RastriginE {  // base_model.py:308-316 ; c = (float)(2*pi), an = (float)(a*D)
  float a, c, an;
  __device__ __forceinline__ float grad(float x) const {
    const float s = sinf(__fmul_rn(c, x));
    return __fadd_rn(__fmul_rn(2.0f, x), __fmul_rn(__fmul_rn(a, s), c));
  }
  __device__ __forceinline__ float term(float x) const {
    return __fsub_rn(__fmul_rn(x, x), __fmul_rn(a, cosf(__fmul_rn(c, x))));
  }
  __device__ __forceinline__ float finish(float sum) const { return __fadd_rn(an, sum); }
};

// torch.clamp semantics: NaN propagates
__device__ __forceinline__ float clamp_torch(float x, float lo, float hi) {
  return (x != x) ? x : fminf(fmaxf(x, lo), hi);
}

// nan_to_num_(nan=0.0): NaN -> 0, +inf -> FLT_MAX, -inf -> -FLT_MAX (base_integrator.py:879-889)
__device__ __forceinline__ float nan_to_num0(float x) {
  if (x != x) return 0.0f;
  if (x == INFINITY) return 3.402823466e+38f;
  if (x == -INFINITY) return -3.402823466e+38f;
  return x;
}

}  // namespace ebm
