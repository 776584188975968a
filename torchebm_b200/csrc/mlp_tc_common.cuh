// Shared pieces of the tensor-core MLP kernels (ebm_mlp_tc.cu, ebm_mlp_tc2.cu): tile constants, launch parameters,
// packed fp32x2 epilogue arithmetic, TMEM <-> register-pair transfers, weight staging, state row loads / stores.
#pragma once
#include "api_common.cuh"
#include "umma.cuh"
#include "mlp_schedule.cuh"
#include <cuda_bf16.h>
#include <type_traits>

namespace ebm {

using namespace umma;

constexpr int kTcM = 128;          // chains per tile = TMEM lanes
constexpr int kTcW = 128;          // padded width of every layer
constexpr int kTcChunks = kTcW / 16;
constexpr int kTcEpiWarps = 16;         // 4 lane quarters x 4 column quarters
constexpr int kTcCols = kTcW / (kTcEpiWarps / 4);  // columns per epilogue thread (32)
constexpr int kTcRoleWarps = 4;          // one warpgroup: warp 0 issues the MMAs, warps 1-3 idle (setmaxnreg works per warpgroup)
constexpr int kTcThreads = 32 * (kTcRoleWarps + kTcEpiWarps);
// 640 threads launch with 96 registers each; optionally the role warpgroup shrinks and the four epilogue warpgroups grow
// (e.g. 128 * 32 + 512 * 112 = 640 * 96)
// EBM_TC_LD32: 1 = every epilogue fetches its 32 accumulator columns with one tcgen05.ld and one wait, 0 = two 16-column
// loads, each waited for where it is used
#ifndef EBM_TC_LD32
#define EBM_TC_LD32 1
#endif
#ifndef EBM_TC_ROLE_REGS
// measured on B200 (tools/mlp_probe.py): 96/96 (no rebalance) 14.95 us per tile-step, 56/104 15.3, 32/112 15.9 -- here the
// issue latency of the MMA thread matters more than the epilogue's few spills, so the default leaves the budget alone
#define EBM_TC_ROLE_REGS 96
#define EBM_TC_EPI_REGS 96
#endif
constexpr int kTcRoleRegs = EBM_TC_ROLE_REGS;
constexpr int kTcEpiRegs = EBM_TC_EPI_REGS;
constexpr int kTcMatBytes = kTcW * kTcW * 2;  // one bf16 [128 x 128] operand

struct TcParams {
  const float* W1; const float* b1; const float* W2; const float* b2; const float* w3; const float* b3;
  int d, h1, h2;
  int passes;  // 3: bf16x3 split, 1: plain bf16
  const float* x_in;
  float* x_out;
  const float* noise;
  float* traj;
  const long long* row_index;  // persistent-CD: source row of chain i in x_in (first launch of a burst only), or NULL
  float* x_out2;               // persistent-CD: second destination of the burst's final state (last launch only), or NULL
  long long n;
  int n_steps, thin, n_kept, step_base, has_clamp;   // step_base: steps of this burst done by earlier launches
  float clamp_lo, clamp_hi;
  RowRng rng;
  PhiloxKeys keys;    // round keys of (rng.k0, rng.k1)
  MlpSchedule sched;  // balanced (tile, step-range) split, mlp_schedule.cuh
  // burst-end gather fused into the final state store (last launch of a burst only), see ebm_mlp_wide.cu
  int n_peers;
  int peer_mc;   // 1: peers[0] is an NVLS multicast address (n_peers == 1)
  long long peer_off;
  float* peers[kMaxPeers];
};

struct TcSmemLayout {
  static constexpr int w1_hi = 0;
  static constexpr int w1_lo = w1_hi + kTcMatBytes;
  static constexpr int w2_hi = w1_lo + kTcMatBytes;
  static constexpr int w2_lo = w2_hi + kTcMatBytes;
  static constexpr int a_hi = w2_lo + kTcMatBytes;
  static constexpr int a_lo = a_hi + kTcMatBytes;
  static constexpr int b1 = a_lo + kTcMatBytes;
  static constexpr int b2 = b1 + kTcW * 4;
  static constexpr int w3 = b2 + kTcW * 4;
  static constexpr int bars = w3 + kTcW * 4;           // kTcChunks + 1 mbarriers
  static constexpr int tmem_slot = bars + (kTcChunks + 1) * 8;
  static constexpr int units = tmem_slot + 16;
  static constexpr int total = units + 16;
  // HMC kernel only: per-row partial sums of E(x), E(x'), K(p), K(p') per column quarter, double-buffered by proposal
  static constexpr int hmc_part = (total + 15) & ~15;
  // HMC kernel only: "this proposal must be redone without force carrying" token + the mbarrier that hands the
  // epilogue's verdict to the MMA warp (hmc_mlp_tc_kernel)
  static constexpr int hmc_redo = hmc_part + 2 * 4 * 4 * kTcM * 4;
  static constexpr int hmc_verdict = hmc_redo + 8;
  static constexpr int hmc_total = hmc_verdict + 8;
};

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// bf16x2 word (low half = a) rounded to nearest, except that a FINITE value above the largest bf16 (it would round to
// inf, and inf - inf in the residual would poison the whole row: safe-mode HMC sanitises +-inf to +-FLT_MAX,
// base_integrator.py:879-889) keeps its truncated leading part; the residual carries the rest
__device__ __forceinline__ uint32_t bf16x2_rn_finite(float a, float b) {
  const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
  uint32_t hu = *reinterpret_cast<const uint32_t*>(&h2);
  if ((hu & 0x7fffu) == 0x7f80u && fabsf(a) <= 3.402823466e+38f) hu = (hu & 0xffff0000u) | (__float_as_uint(a) >> 16);
  if ((hu & 0x7fff0000u) == 0x7f800000u && fabsf(b) <= 3.402823466e+38f) hu = (hu & 0x0000ffffu) | (__float_as_uint(b) & 0xffff0000u);
  return hu;
}

// write kTcCols consecutive columns [col0, col0 + 32) of row r of the A operand (hi and lo copies); 8 columns = one
// 16-byte core-matrix row, consecutive rows of a warp are consecutive 16-byte slots (conflict-free st.shared.v4)
__device__ __forceinline__ void store_a_cols(uint8_t* smem, int r, int col0, const float (&v)[kTcCols], bool with_lo) {
#pragma unroll
  for (int oct = 0; oct < kTcCols / 8; ++oct) {
    uint32_t ph[4], pl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = v[oct * 8 + 2 * j], b = v[oct * 8 + 2 * j + 1];
      const uint32_t hu = bf16x2_rn_finite(a, b);
      ph[j] = hu;
      const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - __uint_as_float(hu << 16), b - __uint_as_float(hu & 0xffff0000u));
      pl[j] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    const int off = core_offset(r, col0 + oct * 8, kTcM);
    *reinterpret_cast<uint4*>(smem + TcSmemLayout::a_hi + off) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    if (with_lo) *reinterpret_cast<uint4*>(smem + TcSmemLayout::a_lo + off) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
  }
}

// 16-column variant (one MMA k-step) used to publish the first half of a thread's columns early
__device__ __forceinline__ void store_a_16(uint8_t* smem, int r, int col0, const float* v, bool with_lo) {
#pragma unroll
  for (int oct = 0; oct < 2; ++oct) {
    uint32_t ph[4], pl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = v[oct * 8 + 2 * j], b = v[oct * 8 + 2 * j + 1];
      const uint32_t hu = bf16x2_rn_finite(a, b);
      ph[j] = hu;
      const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - __uint_as_float(hu << 16), b - __uint_as_float(hu & 0xffff0000u));
      pl[j] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    const int off = core_offset(r, col0 + oct * 8, kTcM);
    *reinterpret_cast<uint4*>(smem + TcSmemLayout::a_hi + off) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    if (with_lo) *reinterpret_cast<uint4*>(smem + TcSmemLayout::a_lo + off) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
  }
}
__device__ __forceinline__ void signal_one(uint8_t* smem, int chunk, int lane) {
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) mbar_arrive(smem_u32(smem + TcSmemLayout::bars + chunk * 8));
}

// all lanes of the warp have written their rows of the warp's two 16-column chunks: publish them to the MMA warp
__device__ __forceinline__ void signal_cols(uint8_t* smem, int first_chunk, int lane) {
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < kTcCols / 16; ++c) mbar_arrive(smem_u32(smem + TcSmemLayout::bars + (first_chunk + c) * 8));
  }
}

// ---- packed (fp32x2) epilogue arithmetic: the epilogue is bound by instruction issue (f32x2.cuh) --------------------
// bf16 hi/lo split of a packed pair of activations; returns the two bf16x2 words (low half = first element).
// With a lo part the hi part is the TRUNCATED value (one byte permute for the pair instead of a convert and a shift):
// the residual v - hi is exact either way and is rounded to bf16; the weights stay split by rounding, so the dropped
// lo*lo term remains unbiased (~2^-18 relative).  Without a lo part (single-pass bf16) hi is rounded to nearest.
__device__ __forceinline__ void split2(f32x2 V, uint32_t& hi, uint32_t& lo, bool with_lo) {
  float a, b;
  unpack2(V, a, b);
  if (with_lo) {
    const uint32_t ua = __float_as_uint(a), ub = __float_as_uint(b);
    hi = __byte_perm(ua, ub, 0x7632);
    const f32x2 R = fma2(pack2(__uint_as_float(ua & 0xffff0000u), __uint_as_float(ub & 0xffff0000u)), -1.0f, V);   // exact
    float ra, rb;
    unpack2(R, ra, rb);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(ra, rb);
    lo = *reinterpret_cast<const uint32_t*>(&l2);
  } else {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<const uint32_t*>(&h2);
  }
}
// 16 consecutive columns (8 pairs) of row r of the A operand, hi (and lo) copies
__device__ __forceinline__ void store_a_16p(uint8_t* smem, int r, int col0, const f32x2* v, bool with_lo) {
#pragma unroll
  for (int oct = 0; oct < 2; ++oct) {
    uint32_t ph[4], pl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split2(v[oct * 4 + j], ph[j], pl[j], with_lo);
    const int off = core_offset(r, col0 + oct * 8, kTcM);
    *reinterpret_cast<uint4*>(smem + TcSmemLayout::a_hi + off) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    if (with_lo) *reinterpret_cast<uint4*>(smem + TcSmemLayout::a_lo + off) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
  }
}
// TMEM <-> 8 packed pairs (16 consecutive fp32 columns of this thread's lane)
__device__ __forceinline__ void tmem_ld16p_nowait(uint32_t taddr, f32x2 (&v)[8]) {
  asm volatile(
      "{\n\t.reg .b32 r<16>;\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {r0, r1, r2, r3, r4, r5, r6, r7, r8, r9, r10, r11, r12, r13, r14, r15}, [%8];\n\t"
      "mov.b64 %0, {r0, r1};\n\tmov.b64 %1, {r2, r3};\n\tmov.b64 %2, {r4, r5};\n\tmov.b64 %3, {r6, r7};\n\t"
      "mov.b64 %4, {r8, r9};\n\tmov.b64 %5, {r10, r11};\n\tmov.b64 %6, {r12, r13};\n\tmov.b64 %7, {r14, r15};\n\t}\n"
      : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]), "=l"(v[4]), "=l"(v[5]), "=l"(v[6]), "=l"(v[7])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16p(uint32_t taddr, f32x2 (&v)[8]) {
  // the wait must sit between the load and the first use of its registers: keep load + wait in one asm block
  asm volatile(
      "{\n\t.reg .b32 r<16>;\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {r0, r1, r2, r3, r4, r5, r6, r7, r8, r9, r10, r11, r12, r13, r14, r15}, [%8];\n\t"
      "tcgen05.wait::ld.sync.aligned;\n\t"
      "mov.b64 %0, {r0, r1};\n\tmov.b64 %1, {r2, r3};\n\tmov.b64 %2, {r4, r5};\n\tmov.b64 %3, {r6, r7};\n\t"
      "mov.b64 %4, {r8, r9};\n\tmov.b64 %5, {r10, r11};\n\tmov.b64 %6, {r12, r13};\n\tmov.b64 %7, {r14, r15};\n\t}\n"
      : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]), "=l"(v[4]), "=l"(v[5]), "=l"(v[6]), "=l"(v[7])
      : "r"(taddr)
      : "memory");
}
// 32 columns (16 pairs) in one instruction
#define EBM_R32 "r0, r1, r2, r3, r4, r5, r6, r7, r8, r9, r10, r11, r12, r13, r14, r15, r16, r17, r18, r19, r20, r21, r22, r23, r24, r25, r26, r27, r28, r29, r30, r31"
#define EBM_MOV32                                                                                                     \
  "mov.b64 %0, {r0, r1};\n\tmov.b64 %1, {r2, r3};\n\tmov.b64 %2, {r4, r5};\n\tmov.b64 %3, {r6, r7};\n\t"               \
  "mov.b64 %4, {r8, r9};\n\tmov.b64 %5, {r10, r11};\n\tmov.b64 %6, {r12, r13};\n\tmov.b64 %7, {r14, r15};\n\t"         \
  "mov.b64 %8, {r16, r17};\n\tmov.b64 %9, {r18, r19};\n\tmov.b64 %10, {r20, r21};\n\tmov.b64 %11, {r22, r23};\n\t"     \
  "mov.b64 %12, {r24, r25};\n\tmov.b64 %13, {r26, r27};\n\tmov.b64 %14, {r28, r29};\n\tmov.b64 %15, {r30, r31};\n\t"
#define EBM_OUT32                                                                                                     \
  "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]), "=l"(v[4]), "=l"(v[5]), "=l"(v[6]), "=l"(v[7]), "=l"(v[8]), "=l"(v[9]), \
      "=l"(v[10]), "=l"(v[11]), "=l"(v[12]), "=l"(v[13]), "=l"(v[14]), "=l"(v[15])
__device__ __forceinline__ void tmem_ld32p(uint32_t taddr, f32x2 (&v)[16]) {
  asm volatile("{\n\t.reg .b32 r<32>;\n\ttcgen05.ld.sync.aligned.32x32b.x32.b32 {" EBM_R32 "}, [%16];\n\t"
               "tcgen05.wait::ld.sync.aligned;\n\t" EBM_MOV32 "}\n"
               : EBM_OUT32
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32p_nowait(uint32_t taddr, f32x2 (&v)[16]) {
  asm volatile("{\n\t.reg .b32 r<32>;\n\ttcgen05.ld.sync.aligned.32x32b.x32.b32 {" EBM_R32 "}, [%16];\n\t" EBM_MOV32 "}\n"
               : EBM_OUT32
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st16p(uint32_t taddr, const f32x2 (&v)[8]) {
  asm volatile(
      "{\n\t.reg .b32 r<16>;\n\t"
      "mov.b64 {r0, r1}, %1;\n\tmov.b64 {r2, r3}, %2;\n\tmov.b64 {r4, r5}, %3;\n\tmov.b64 {r6, r7}, %4;\n\t"
      "mov.b64 {r8, r9}, %5;\n\tmov.b64 {r10, r11}, %6;\n\tmov.b64 {r12, r13}, %7;\n\tmov.b64 {r14, r15}, %8;\n\t"
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {r0, r1, r2, r3, r4, r5, r6, r7, r8, r9, r10, r11, r12, r13, r14, r15};\n\t}\n"
      :: "r"(taddr), "l"(v[0]), "l"(v[1]), "l"(v[2]), "l"(v[3]), "l"(v[4]), "l"(v[5]), "l"(v[6]), "l"(v[7])
      : "memory");
}
__device__ __forceinline__ void tmem_st8p(uint32_t taddr, const f32x2 (&v)[4]) {
  asm volatile(
      "{\n\t.reg .b32 r<8>;\n\t"
      "mov.b64 {r0, r1}, %1;\n\tmov.b64 {r2, r3}, %2;\n\tmov.b64 {r4, r5}, %3;\n\tmov.b64 {r6, r7}, %4;\n\t"
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {r0, r1, r2, r3, r4, r5, r6, r7};\n\t}\n"
      :: "r"(taddr), "l"(v[0]), "l"(v[1]), "l"(v[2]), "l"(v[3])
      : "memory");
}

// activation and derivative of a packed pair; SiLU is fully packed (2 ex2 + 2 rcp + 6 packed ops per pair), the others
// go through the scalar form
template <int ACT>
__device__ __forceinline__ void act_fast(float z, float& h, float& dh);
template <int ACT>
__device__ __forceinline__ void act2(f32x2 Z, f32x2& H, f32x2& DH) {
  if (ACT == EBM_ACT_SILU) {
    float t0, t1;
    unpack2(mul2(Z, -1.4426950408889634f), t0, t1);
    float d0, d1;
    unpack2(add2(pack2(ex2_ftz(t0), ex2_ftz(t1)), 1.0f), d0, d1);
    const f32x2 S = pack2(rcp_ftz(d0), rcp_ftz(d1));
    H = mul2(Z, S);
    DH = fma2(H, fma2(S, -1.0f, 1.0f), S);   // s + z s (1 - s)
  } else {
    float z0, z1, h0, h1, g0, g1;
    unpack2(Z, z0, z1);
    act_fast<ACT>(z0, h0, g0);
    act_fast<ACT>(z1, h1, g1);
    H = pack2(h0, h1);
    DH = pack2(g0, g1);
  }
}

template <int ACT>
__device__ __forceinline__ void act_fast(float z, float& h, float& dh) {
  if (ACT == EBM_ACT_SILU) {
    const float s = rcp_ftz(1.0f + ex2_ftz(-1.4426950408889634f * z));
    h = z * s;
    dh = s * (1.0f + z * (1.0f - s));
  } else if (ACT == EBM_ACT_TANH) {
    const float e = ex2_ftz(-2.8853900817779268f * fabsf(z));
    const float t = copysignf((1.0f - e) * rcp_ftz(1.0f + e), z);
    h = t;
    dh = 1.0f - t * t;
  } else if (ACT == EBM_ACT_RELU) {
    h = z > 0.0f ? z : 0.0f;
    dh = z > 0.0f ? 1.0f : 0.0f;
  } else {
    h = z > 20.0f ? z : log1pf(ex2_ftz(1.4426950408889634f * z));
    dh = rcp_ftz(1.0f + ex2_ftz(-1.4426950408889634f * z));
  }
}

template <class Layout = TcSmemLayout>
__device__ __forceinline__ void tc_stage_weights(uint8_t* smem, const TcParams& P) {
  for (int i = threadIdx.x; i < kTcW * kTcW; i += blockDim.x) {
    const int r = i / kTcW, c = i - r * kTcW;
    __nv_bfloat16 hi, lo;
    split_bf16((r < P.h1 && c < P.d) ? P.W1[r * P.d + c] : 0.0f, hi, lo);
    const int off = core_offset(r, c, kTcW);
    *reinterpret_cast<__nv_bfloat16*>(smem + Layout::w1_hi + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(smem + Layout::w1_lo + off) = lo;
    split_bf16((r < P.h2 && c < P.h1) ? P.W2[r * P.h1 + c] : 0.0f, hi, lo);
    *reinterpret_cast<__nv_bfloat16*>(smem + Layout::w2_hi + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(smem + Layout::w2_lo + off) = lo;
  }
  float* b1 = reinterpret_cast<float*>(smem + Layout::b1);
  float* b2 = reinterpret_cast<float*>(smem + Layout::b2);
  float* w3 = reinterpret_cast<float*>(smem + Layout::w3);
  for (int i = threadIdx.x; i < kTcW; i += blockDim.x) {
    b1[i] = i < P.h1 ? P.b1[i] : 0.0f;
    b2[i] = i < P.h2 ? P.b2[i] : 0.0f;
    w3[i] = i < P.h2 ? P.w3[i] : 0.0f;
  }
}

// one product: D[tmem_d] = A (k-major, chunks published by the epilogue) x B, B = W^T (forward) or W (backward).
// Called by the whole (converged) MMA warp; `leader` = elect_one() issues.
__device__ __forceinline__ void tc_issue_gemm(uint8_t* smem, uint32_t tmem_d, int w_hi_off, int w_lo_off, bool backward,
                                              int ksteps, int passes, uint32_t parity, bool leader) {
  const uint32_t a_hi = smem_u32(smem + TcSmemLayout::a_hi), a_lo = smem_u32(smem + TcSmemLayout::a_lo);
  const uint32_t w_hi = smem_u32(smem + w_hi_off), w_lo = smem_u32(smem + w_lo_off);
  const uint32_t idesc = make_idesc_bf16(kTcM, kTcW, backward);
  // every epilogue thread publishes its first 16-column block half-way through its work and the second at the
  // end, so the even chunks are ready early: consume evens first, then odds (summation order is irrelevant)
  bool first = true;
  for (int idx = 0; idx < kTcChunks; ++idx) {
    const int c = (idx < kTcChunks / 2) ? 2 * idx : 2 * (idx - kTcChunks / 2) + 1;
    if (c >= ksteps) continue;
    mbar_wait(smem_u32(smem + TcSmemLayout::bars + c * 8), parity);
    tcgen05_fence_after();
    const uint32_t a_off = c * 2 * (kTcM * 16);
    // forward: B K-major (rows = outputs): next k-step = 2 core columns; backward: B MN-major: next k-step = 16 rows
    const uint32_t b_off = backward ? c * 256 : c * 2 * (kTcW * 16);
    const uint32_t b_lbo = backward ? 128 : kTcW * 16, b_sbo = backward ? kTcW * 16 : 128;
    const uint64_t ah = make_smem_desc(a_hi + a_off, kTcM * 16, 128);
    const uint64_t bh = make_smem_desc(w_hi + b_off, b_lbo, b_sbo);
    if (leader) {
      mma_bf16(tmem_d, ah, bh, idesc, !first);
      if (passes == 3) {
        const uint64_t al = make_smem_desc(a_lo + a_off, kTcM * 16, 128);
        const uint64_t bl = make_smem_desc(w_lo + b_off, b_lbo, b_sbo);
        mma_bf16(tmem_d, al, bh, idesc, true);
        mma_bf16(tmem_d, ah, bl, idesc, true);
      }
    }
    first = false;
  }
  if (leader) mma_commit(smem_u32(smem + TcSmemLayout::bars + kTcChunks * 8));
  __syncwarp();
}

// a thread's 32 consecutive columns of one row <-> global memory: eight 128-bit accesses when the row is 16-byte aligned
// and lies inside the state, scalar otherwise
__device__ __forceinline__ void tc_store_row32(float* __restrict__ dst, long long grow, int d, int col_base, bool rv,
                                               const float (&x)[kTcCols]) {
  if (!rv) return;
  float* p = dst + grow * d + col_base;
  if (col_base + kTcCols <= d && (d & 3) == 0 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
    for (int j = 0; j < kTcCols / 4; ++j)
      reinterpret_cast<float4*>(p)[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < kTcCols; ++i)
      if (col_base + i < d) p[i] = x[i];
  }
}
__device__ __forceinline__ void tc_load_row32(const float* __restrict__ src, long long grow, int d, int col_base, bool rv,
                                              float (&x)[kTcCols]) {
  const float* p = src + grow * d + col_base;
  if (rv && col_base + kTcCols <= d && (d & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
#pragma unroll
    for (int j = 0; j < kTcCols / 4; ++j) {
      const float4 t = reinterpret_cast<const float4*>(p)[j];
      x[4 * j] = t.x; x[4 * j + 1] = t.y; x[4 * j + 2] = t.z; x[4 * j + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < kTcCols; ++i) x[i] = (rv && col_base + i < d) ? p[i] : 0.0f;
  }
}

}  // namespace ebm
