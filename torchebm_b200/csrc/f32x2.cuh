// Packed fp32x2 arithmetic (sm_100a FFMA2 / FMUL2 / FADD2): one issue slot does two lanes' worth of fp32 work.  The
// fused sampler kernels are instruction-issue bound, so everything elementwise runs on register pairs.
//
// Rounding: every packed op rounds each half exactly like the scalar .rn op (mul2 = __fmul_rn, add2 = __fadd_rn,
// fma2 = __fmaf_rn per half).  ONE TRAP: ptxas (CUDA 12.9) contracts a packed multiply whose result feeds a packed add
// into a single FFMA2 even though both carry .rn (scalar .rn ops are never contracted).  Wherever the reference rounds
// the product and the sum separately, do not write add2(mul2(a, b), c); use
//     fma2(mul2(a, 2*b), 0.5, c)      -- the product scaled by an exact power of two, undone inside a true fma, or
//     add2(pack2(__fmul_rn(..), __fmul_rn(..)), c)      -- a scalar multiply feeding the packed add
// (both verified in SASS: FMUL2 + FFMA2 with the 0.5 immediate, resp. FMUL + FMUL + FADD2).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ebm {

typedef unsigned long long f32x2;  // .x = low 32 bits, .y = high 32 bits

__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 bcast2(float a) { return pack2(a, a); }

__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, float b) { return mul2(a, bcast2(b)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, float b) { return add2(a, bcast2(b)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, float b, f32x2 c) { return fma2(a, bcast2(b), c); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, float c) { return fma2(a, b, bcast2(c)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, float b, float c) { return fma2(a, bcast2(b), bcast2(c)); }

// fl(c + p) for a product p that was computed as mul2(a, 2*b): one rounding of (c + a*b), the product rounded on its own
__device__ __forceinline__ f32x2 add_scaled_product2(f32x2 twice_p, f32x2 c) { return fma2(twice_p, 0.5f, c); }

// single-instruction transcendentals, flush-to-zero (the non-ftz forms wrap every MUFU in a range test and two
// conditional multiplies for denormal operands, which none of the callers can produce)
__device__ __forceinline__ float ex2_ftz(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float lg2_ftz(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sqrt_ftz(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rsqrt_ftz(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp_ftz(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sin_ftz(float x) { float r; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float cos_ftz(float x) { float r; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

}  // namespace ebm
