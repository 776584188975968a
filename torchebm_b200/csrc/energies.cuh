// Energy functors.  Gradients are written in the reference's fp32 rounding order (autograd of
// torchebm/core/base_model.py forwards): every product and sum is rounded
// separately (__fmul_rn / __fadd_rn never contract into FMA), because eager PyTorch runs one kernel
// per op.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "f32x2.cuh"

namespace ebm {

// ---- elementwise energies: E(x) = scale(sum_i term(x_i)) , dE/dx_i = grad(x_i) -------------------

// grad2 = grad on a packed element pair, same bits: multiplications by 2 are exact, so they are folded into the
// coefficients (fl(fl(h * 2u) * 2x) == fl(fl(4h * u) * x) barring overflow / underflow of the intermediate, which needs
// |u| < 1e-38 or |h u| > 1e38); products that feed a sum are kept scalar (f32x2.cuh).
struct DoubleWellE {  // base_model.py:143-148
  float h, b2;
  __device__ __forceinline__ float grad(float x) const {   // fl(fl(h * 2u) * 2x) == fl(fl(4h * u) * x): two products fewer
    const float u = __fsub_rn(__fmul_rn(x, x), b2);
    return __fmul_rn(__fmul_rn(__fmul_rn(4.0f, h), u), x);
  }
  __device__ __forceinline__ f32x2 grad2(f32x2 X) const {
    float x0, x1;
    unpack2(X, x0, x1);
    const f32x2 U = add2(pack2(__fmul_rn(x0, x0), __fmul_rn(x1, x1)), -b2);
    return mul2(mul2(U, __fmul_rn(4.0f, h)), X);
  }
  // contracted form for the native-stream bursts (no bit-parity contract with the reference's rounding order)
  __device__ __forceinline__ f32x2 grad2_fast(f32x2 X) const { return mul2(mul2(fma2(X, X, -b2), __fmul_rn(4.0f, h)), X); }
  __device__ __forceinline__ float term(float x) const {
    const float u = __fsub_rn(__fmul_rn(x, x), b2);
    return __fmul_rn(u, u);
  }
  __device__ __forceinline__ float finish(float sum) const { return __fmul_rn(h, sum); }
};

struct HarmonicE {  // base_model.py:224-229 ; half_k = (float)(0.5*k)
  float half_k;
  __device__ __forceinline__ float grad(float x) const { return __fmul_rn(__fmul_rn(2.0f, half_k), x); }   // 2*half_k is exact
  __device__ __forceinline__ f32x2 grad2(f32x2 X) const { return mul2(X, __fmul_rn(2.0f, half_k)); }
  __device__ __forceinline__ f32x2 grad2_fast(f32x2 X) const { return grad2(X); }
  __device__ __forceinline__ float term(float x) const { return __fmul_rn(x, x); }
  __device__ __forceinline__ float finish(float sum) const { return __fmul_rn(half_k, sum); }
};

struct RastriginE {  // base_model.py:308-316 ; c = (float)(2*pi), an = (float)(a*D)
  float a, c, an;
  __device__ __forceinline__ float grad(float x) const {
    const float s = sinf(__fmul_rn(c, x));
    return __fadd_rn(__fmul_rn(2.0f, x), __fmul_rn(__fmul_rn(a, s), c));
  }
  __device__ __forceinline__ f32x2 grad2(f32x2 X) const {
    float c0, c1;
    unpack2(mul2(X, c), c0, c1);
    const f32x2 ASC = mul2(mul2(pack2(sinf(c0), sinf(c1)), a), c);
    return fma2(X, 2.0f, ASC);   // fl(2x + asc): 2x is exact, one rounding like fl(fl(2x) + asc)
  }
  __device__ __forceinline__ f32x2 grad2_fast(f32x2 X) const { return grad2(X); }
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
