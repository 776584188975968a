// MLP-energy Langevin burst for THREE hidden layers (all widths <= 128) on tcgen05 + TMEM, sm_100a only:
//   E(x) = w4 . act(W3 act(W2 act(W1 x + b1) + b2) + b3) + b4          (benchmarks/distributed_fsdp2.py:43-53)
//   dE/dx = W1^T (act'(z1) * (W2^T (act'(z2) * (W3^T (act'(z3) * w4)))))  (what autograd of base_model.py:84-127 yields)
//
// Same numerics as ebm_mlp_tc.cu (bf16 hi + lo split operands, hi*hi + lo*hi + hi*lo accumulated in fp32 tensor memory,
// precision bf16 = one pass), six [128 x 128 x 128] products per Langevin step.  Three weight matrices as hi + lo take
// 192 KB of shared memory, so there is no room for an operand buffer: every A operand (x, h1, h2, delta3, delta2,
// delta1) lives in TENSOR memory instead -- the epilogue writes packed bf16 pairs with tcgen05.st (row = lane, column j =
// elements 2j, 2j+1) and the products use tcgen05.mma's [a_tmem] form, which also costs 74 cycles per 128x128x16
// product where the shared-memory form costs 107 (tools/umma_probe.cu).
//
// One CTA = one tile of 128 chains, persistent over (tile, step-range) units (mlp_schedule.cuh), x[r, 32 cols] of a
// thread stays in registers for the whole unit.  Per step
//     G1 z1=x W1^T -> E1 -> G2 z2=h1 W2^T -> E2 -> G3 z3=h2 W3^T -> E3 -> G4 t2=d3 W3 -> E4 -> G5 t1=d2 W2 -> E5
//     -> G6 g=d1 W1 -> E6 (update) -> G1'
//   TMEM columns: R0 [0,128):   z1 -> act'(z1) in place (read by E5)
//                 R1 [128,256): z2 -> act'(z2) in place (read by E4) -> t1
//                 R2 [256,384): z3 -> t2 -> g
//                 A  [384,512): operand of the next product, hi [384,448), lo [448,512)
//   G4 and G5 overwrite a region other warps may still be reading in the epilogue that feeds them, so they wait for
//   ALL operand chunks before their first MMA; the other products start chunk by chunk underneath their epilogue.
#include "mlp_tc_common.cuh"

namespace ebm {

struct DeepSmem {
  __host__ __device__ static constexpr int w_hi(int l) { return l * 2 * kTcMatBytes; }
  __host__ __device__ static constexpr int w_lo(int l) { return l * 2 * kTcMatBytes + kTcMatBytes; }
  static constexpr int bias = 6 * kTcMatBytes;          // b1, b2, b3, w4: 4 x 128 floats
  static constexpr int bars = bias + 4 * kTcW * 4;      // kTcChunks + 1 mbarriers
  static constexpr int tmem_slot = bars + (kTcChunks + 1) * 8;
  static constexpr int units = tmem_slot + 16;
  static constexpr int total = units + 16;
};
static_assert(DeepSmem::total <= 232448, "shared memory budget of one sm_100 CTA exceeded");

constexpr int kDpRoleRegs = 32, kDpEpiRegs = 112;
constexpr uint32_t kDpR0 = 0, kDpR1 = 128, kDpR2 = 256, kDpA = 384, kDpALo = 448;

struct DeepParams {
  const float* W[3];
  const float* b[3];
  const float* w4;
  int d, h1, h2, h3;
  int passes;
  const float* x_in;
  float* x_out;
  const float* noise;
  float* traj;
  const long long* row_index;
  float* x_out2;
  long long n;
  int n_steps, thin, n_kept, step_base, has_clamp;
  float clamp_lo, clamp_hi;
  RowRng rng;
  PhiloxKeys keys;
  MlpSchedule sched;
};

__device__ __forceinline__ void dp_stage_matrix(uint8_t* smem, int l, const float* __restrict__ W, int rows, int cols) {
  for (int i = threadIdx.x; i < kTcW * kTcW; i += blockDim.x) {
    const int r = i / kTcW, c = i - r * kTcW;
    __nv_bfloat16 hi, lo;
    split_bf16((r < rows && c < cols) ? W[r * cols + c] : 0.0f, hi, lo);
    const int off = core_offset(r, c, kTcW);
    *reinterpret_cast<__nv_bfloat16*>(smem + DeepSmem::w_hi(l) + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(smem + DeepSmem::w_lo(l) + off) = lo;
  }
}

__device__ __forceinline__ void dp_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  const uint32_t acc = accumulate ? 1u : 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

// one product D[tmem_d] = A (tensor memory, chunks published by the epilogue) x B, B = W_l^T (forward) or W_l (backward).
// Called by the whole (converged) MMA warp; `leader` = elect_one() issues.
__device__ __forceinline__ void dp_issue_gemm(uint8_t* smem, uint32_t tmem, uint32_t tmem_d, int l, bool backward, int ksteps,
                                              int passes, uint32_t parity, bool leader, bool wait_all) {
  const uint32_t w_hi = smem_u32(smem + DeepSmem::w_hi(l)), w_lo = smem_u32(smem + DeepSmem::w_lo(l));
  const uint32_t idesc = make_idesc_bf16(kTcM, kTcW, backward);
  if (wait_all) {
    for (int c = 0; c < kTcChunks; ++c) mbar_wait(smem_u32(smem + DeepSmem::bars + c * 8), parity);
    tcgen05_fence_after();
  }
  bool first = true;
  for (int idx = 0; idx < kTcChunks; ++idx) {
    const int c = (idx < kTcChunks / 2) ? 2 * idx : 2 * (idx - kTcChunks / 2) + 1;   // even chunks are published first
    if (c >= ksteps) continue;
    if (!wait_all) {
      mbar_wait(smem_u32(smem + DeepSmem::bars + c * 8), parity);
      tcgen05_fence_after();
    }
    const uint32_t b_off = backward ? c * 256 : c * 2 * (kTcW * 16);
    const uint32_t b_lbo = backward ? 128 : kTcW * 16, b_sbo = backward ? kTcW * 16 : 128;
    const uint64_t bh = make_smem_desc(w_hi + b_off, b_lbo, b_sbo);
    if (leader) {
      dp_mma_ts(tmem_d, tmem + kDpA + 8 * c, bh, idesc, !first);
      if (passes == 3) {
        const uint64_t bl = make_smem_desc(w_lo + b_off, b_lbo, b_sbo);
        dp_mma_ts(tmem_d, tmem + kDpALo + 8 * c, bh, idesc, true);
        dp_mma_ts(tmem_d, tmem + kDpA + 8 * c, bl, idesc, true);
      }
    }
    first = false;
  }
  if (leader) mma_commit(smem_u32(smem + DeepSmem::bars + kTcChunks * 8));
  __syncwarp();
}

__device__ __forceinline__ void dp_st4(uint32_t taddr, const uint32_t (&w)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]),
               "r"(w[3])
               : "memory");
}
// 16 consecutive operand columns (8 packed pairs) of this thread's row -> tensor memory, hi and residual words
__device__ __forceinline__ void dp_put16p(uint32_t t_hi, uint32_t t_lo, const f32x2* v, bool with_lo) {
#pragma unroll
  for (int o = 0; o < 2; ++o) {
    uint32_t ph[4], pl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split2(v[4 * o + j], ph[j], pl[j], with_lo);
    dp_st4(t_hi + 4 * o, ph);
    if (with_lo) dp_st4(t_lo + 4 * o, pl);
  }
}
// this warp's tensor-memory stores (operand chunk, act' kept in place) have landed -> one arrival for the MMA warp
__device__ __forceinline__ void dp_signal(uint8_t* smem, int chunk, int lane) {
  tmem_st_wait();
  tcgen05_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(smem_u32(smem + DeepSmem::bars + chunk * 8));
}

template <int ACT, bool LO>
__global__ void __launch_bounds__(kTcThreads, 1) langevin_mlp_deep_kernel(const __grid_constant__ DeepParams P,
                                                                          const __grid_constant__ StepTable tab) {
  extern __shared__ __align__(128) uint8_t dp_smem_raw[];
  uint8_t* smem = dp_smem_raw;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  dp_stage_matrix(smem, 0, P.W[0], P.h1, P.d);
  dp_stage_matrix(smem, 1, P.W[1], P.h2, P.h1);
  dp_stage_matrix(smem, 2, P.W[2], P.h3, P.h2);
  {
    float* bias = reinterpret_cast<float*>(smem + DeepSmem::bias);
    for (int i = threadIdx.x; i < kTcW; i += blockDim.x) {
      bias[i] = i < P.h1 ? P.b[0][i] : 0.0f;
      bias[kTcW + i] = i < P.h2 ? P.b[1][i] : 0.0f;
      bias[2 * kTcW + i] = i < P.h3 ? P.b[2][i] : 0.0f;
      bias[3 * kTcW + i] = i < P.h3 ? P.w4[i] : 0.0f;
    }
  }
  if (threadIdx.x == 0) {
    for (int c = 0; c < kTcChunks; ++c) mbar_init(smem_u32(smem + DeepSmem::bars + c * 8), 4);  // 4 warps per chunk
    mbar_init(smem_u32(smem + DeepSmem::bars + kTcChunks * 8), 1);
    fence_mbar_init();
  }
  if (threadIdx.x == 32) mlp_units_compute(P.sched, P.n_steps, reinterpret_cast<volatile MlpUnits*>(smem + DeepSmem::units));
  if (warp == 0) tmem_alloc(smem_u32(smem + DeepSmem::tmem_slot), 512);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + DeepSmem::tmem_slot);
  const int k0 = (P.d + 15) / 16, k1 = (P.h1 + 15) / 16, k2 = (P.h2 + 15) / 16, k3 = (P.h3 + 15) / 16;
  const volatile MlpUnits* units = reinterpret_cast<const volatile MlpUnits*>(smem + DeepSmem::units);

  if (warp < kTcRoleWarps) {
    // 640 threads launch with 96 registers; the role warpgroup keeps 32, the four epilogue warpgroups take 112
    // (128 * 32 + 512 * 112 = 640 * 96): the update epilogue holds the state, an accumulator block and the noise
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kDpRoleRegs));
    if (warp == 0) {   // converged: all lanes wait, one elected lane issues (umma.cuh: elect_one)
      const bool leader = elect_one();
      uint32_t parity = 0;
      for (int tile = units->t_last; tile >= units->t_first; --tile) {
        const int n_unit_steps = mlp_unit_s1(units, tile, P.n_steps) - mlp_unit_s0(units, tile);
        for (int k = 0; k < n_unit_steps; ++k) {
          dp_issue_gemm(smem, tmem, tmem + kDpR0, 0, false, k0, P.passes, parity, leader, false); parity ^= 1;   // z1
          dp_issue_gemm(smem, tmem, tmem + kDpR1, 1, false, k1, P.passes, parity, leader, false); parity ^= 1;   // z2
          dp_issue_gemm(smem, tmem, tmem + kDpR2, 2, false, k2, P.passes, parity, leader, false); parity ^= 1;   // z3
          dp_issue_gemm(smem, tmem, tmem + kDpR2, 2, true, k3, P.passes, parity, leader, true); parity ^= 1;     // t2 over z3
          dp_issue_gemm(smem, tmem, tmem + kDpR1, 1, true, k2, P.passes, parity, leader, true); parity ^= 1;     // t1 over act'(z2)
          dp_issue_gemm(smem, tmem, tmem + kDpR2, 0, true, k1, P.passes, parity, leader, false); parity ^= 1;    // g
        }
      }
    }
  } else {
    // ---- epilogue warps: packed fp32x2 arithmetic, padded rows / columns computed like real ones and never stored ----
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kDpEpiRegs));
    const int e = warp - kTcRoleWarps;
    const int row = 32 * (warp & 3) + lane;        // TMEM lane this thread may access (hardware: warp % 4)
    const int cq = e >> 2;                          // column quarter
    const int col_base = kTcCols * cq;
    const int first_chunk = col_base / 16;
    const uint32_t lane_base = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
    const uint32_t lane_addr = lane_base + col_base;                                   // accumulator columns of this thread
    const uint32_t a_hi = lane_base + kDpA + col_base / 2, a_lo = lane_base + kDpALo + col_base / 2;   // operand columns
    const f32x2* b1 = reinterpret_cast<const f32x2*>(smem + DeepSmem::bias + 4 * col_base);
    const f32x2* b2 = b1 + kTcW / 2;
    const f32x2* b3 = b2 + kTcW / 2;
    const f32x2* w4 = b3 + kTcW / 2;
    const uint32_t acc_bar = smem_u32(smem + DeepSmem::bars + kTcChunks * 8);
    const long long numel = P.n * P.d;
    const bool quad_rng = (P.rng.mode == 2) && (P.d % 4 == 0);
    uint32_t parity = 0;

    // operand chunk blk (16 columns) of this thread <- v, then one arrival of the warp on the chunk's barrier
    auto publish = [&](int blk, const f32x2* v) {
      dp_put16p(a_hi + 8 * blk, a_lo + 8 * blk, v, LO);
      dp_signal(smem, first_chunk + blk, lane);
    };

    for (int tile = units->t_last; tile >= units->t_first; --tile) {
      const long long grow = (long long)tile * kTcM + row;
      const bool rv = grow < P.n;
      const int s0 = mlp_unit_s0(units, tile), s1 = mlp_unit_s1(units, tile, P.n_steps);
      if (s0 > 0) mlp_unit_acquire(P.sched, kTcEpiWarps);   // the unit continues the chain another CTA left in x_out
      const float* x0src = (s0 == 0) ? P.x_in : P.x_out;
      const long long row0 = (s0 == 0 && P.row_index && rv) ? P.row_index[grow] : grow;
      f32x2 X[kTcCols / 2];
      {
        float xs[kTcCols];
        tc_load_row32(x0src, row0, P.d, col_base, rv, xs);
#pragma unroll
        for (int j = 0; j < kTcCols / 2; ++j) X[j] = pack2(xs[2 * j], xs[2 * j + 1]);
      }
      publish(0, X);
      publish(1, X + 8);
      int until_keep = P.thin - ((P.step_base + s0) % P.thin), kept = (P.step_base + s0) / P.thin;
      unsigned long long ctr = P.rng.ctr_base + (unsigned long long)s0 * P.rng.ctr_step;

      for (int k = s0; k < s1; ++k) {
        const int ti = k & tab.mask;
        const float h = tab.h[ti], c12 = tab.c1[ti] * tab.c2[ti];
        f32x2 acc[16];
        // E1 / E2: z_l -> h_l (operand of the next forward product); act'(z_l) replaces z_l in place
#pragma unroll
        for (int l = 0; l < 2; ++l) {
          const uint32_t reg = lane_addr + (l == 0 ? kDpR0 : kDpR1);
          const f32x2* bl = (l == 0) ? b1 : b2;
          mbar_wait(acc_bar, parity); parity ^= 1;
          tcgen05_fence_after();
          tmem_ld32p(reg, acc);
#pragma unroll
          for (int blk = 0; blk < 2; ++blk) {
            f32x2 sd[8];
            f32x2* v = acc + 8 * blk;
#pragma unroll
            for (int j = 0; j < 8; ++j) act2<ACT>(add2(v[j], bl[8 * blk + j]), v[j], sd[j]);
            tmem_st16p(reg + 16 * blk, sd);
            publish(blk, v);
          }
        }
        // E3: z3 -> delta3 = w4 * act'(z3)
        mbar_wait(acc_bar, parity); parity ^= 1;
        tcgen05_fence_after();
        tmem_ld32p(lane_addr + kDpR2, acc);
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          f32x2* v = acc + 8 * blk;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            f32x2 hh, dh;
            act2<ACT>(add2(v[j], b3[8 * blk + j]), hh, dh);
            v[j] = mul2(dh, w4[8 * blk + j]);
          }
          publish(blk, v);
        }
        // E4: t2 -> delta2 = t2 * act'(z2) ; E5: t1 -> delta1 = t1 * act'(z1)
#pragma unroll
        for (int l = 0; l < 2; ++l) {
          const uint32_t t_reg = lane_addr + (l == 0 ? kDpR2 : kDpR1);
          const uint32_t d_reg = lane_addr + (l == 0 ? kDpR1 : kDpR0);
          mbar_wait(acc_bar, parity); parity ^= 1;
          tcgen05_fence_after();
          tmem_ld32p_nowait(t_reg, acc);
#pragma unroll
          for (int blk = 0; blk < 2; ++blk) {
            f32x2 sd[8];
            f32x2* v = acc + 8 * blk;
            tmem_ld16p(d_reg + 16 * blk, sd);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = mul2(v[j], sd[j]);
            publish(blk, v);
          }
        }
        // E6: g -> Langevin update of x; the new x is the operand of the next step's first product
        mbar_wait(acc_bar, parity); parity ^= 1;
        tcgen05_fence_after();
        const bool last = (k == s1 - 1);
        bool keep_now = false;
        if (P.traj && --until_keep == 0) { until_keep = P.thin; keep_now = kept < P.n_kept; ++kept; }
        tmem_ld32p_nowait(lane_addr + kDpR2, acc);
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          f32x2 eps[8];
          f32x2* g = acc + 8 * blk;
          const int c0 = col_base + 16 * blk;
          const long long li0 = grow * P.d + c0;
          if (quad_rng && c0 + 16 <= P.d) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const uint64_t qi = (uint64_t)(li0 + 4 * q4) >> 2;
              const uint4 w = philox4x32_10((uint32_t)qi, (uint32_t)(qi >> 32), (uint32_t)ctr, (uint32_t)(ctr >> 32), P.keys);
              normal4_fast_packed(w, eps[2 * q4], eps[2 * q4 + 1]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float ev[2];
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const int i = 2 * j + u;
                const bool in = rv && (c0 + i) < P.d;
                ev[u] = 0.0f;
                if (in) ev[u] = (P.rng.mode == 0) ? P.noise[(long long)k * numel + li0 + i]
                                                  : normal_for_element_call(P.rng.k0, P.rng.k1, ctr, P.rng.T, P.rng.mode, (uint64_t)(li0 + i));
              }
              eps[j] = pack2(ev[0], ev[1]);
            }
          }
          if (blk == 0) tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            // x' = (x - h g) + c2 c1 eps (base_integrator.py:728-729; fused roundings are within this kernel's 2e-5 class)
            f32x2 xn = fma2(eps[j], c12, fma2(g[j], -h, X[8 * blk + j]));
            if (P.has_clamp) {
              float a, b;
              unpack2(xn, a, b);
              xn = pack2(clamp_torch(a, P.clamp_lo, P.clamp_hi), clamp_torch(b, P.clamp_lo, P.clamp_hi));
            }
            X[8 * blk + j] = xn;
          }
          if (c0 + 16 > P.d) {   // columns beyond the state stay exactly zero (their noise is zero, their gradient is not used)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float a, b;
              unpack2(X[8 * blk + j], a, b);
              X[8 * blk + j] = pack2((c0 + 2 * j) < P.d ? a : 0.0f, (c0 + 2 * j + 1) < P.d ? b : 0.0f);
            }
          }
          if (keep_now) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float a, b;
              unpack2(X[8 * blk + j], a, b);
              float* dst = P.traj + (grow * P.n_kept + (kept - 1)) * P.d + c0 + 2 * j;
              if (rv && (c0 + 2 * j) < P.d) dst[0] = a;
              if (rv && (c0 + 2 * j + 1) < P.d) dst[1] = b;
            }
          }
          if (!last) publish(blk, X + 8 * blk);
        }
        ctr += P.rng.ctr_step;
      }
      {
        float xs[kTcCols];
#pragma unroll
        for (int j = 0; j < kTcCols / 2; ++j) unpack2(X[j], xs[2 * j], xs[2 * j + 1]);
        tc_store_row32(P.x_out, grow, P.d, col_base, rv, xs);
        if (P.x_out2 && s1 == P.n_steps) tc_store_row32(P.x_out2, grow, P.d, col_base, rv, xs);
      }
      if (s1 < P.n_steps) mlp_unit_release(P.sched);  // the rest of this tile's burst runs on the next CTA
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem, 512);
  }
}

// ---- E(x) and grad E(x) (ebm_energy_f32 / ebm_gradient_f32 / diagnostics): plain fp32, one warp per row ----------------
// Utility kernel, not on the burst path: activations of a row travel through a warp-private shared buffer, lane l owns
// units l, l+32, l+64, l+96 of every layer; exact expf / tanhf as in ebm_mlp.cu.
struct DeepUtilParams {
  const float* W[3];
  const float* b[3];
  const float* w4; const float* b4;
  int dims[4];   // d, h1, h2, h3
  const float* x;
  float* energy;
  float* grad;
  long long n;
};

template <int ACT>
__device__ __forceinline__ void dp_act_exact(float z, float& h, float& dh) {
  if (ACT == EBM_ACT_SILU) {
    const float s = 1.0f / (1.0f + expf(-z));
    h = z * s;
    dh = s * (1.0f + z * (1.0f - s));
  } else if (ACT == EBM_ACT_TANH) {
    const float t = tanhf(z);
    h = t;
    dh = 1.0f - t * t;
  } else if (ACT == EBM_ACT_RELU) {
    h = z > 0.0f ? z : 0.0f;
    dh = z > 0.0f ? 1.0f : 0.0f;
  } else {
    h = z > 20.0f ? z : log1pf(expf(z));
    dh = 1.0f / (1.0f + expf(-z));
  }
}

constexpr int kDpUWarps = 8;

template <int ACT>
__global__ void __launch_bounds__(32 * kDpUWarps) mlp_deep_energy_grad_kernel(const DeepUtilParams P) {
  __shared__ float act_s[kDpUWarps][4][kTcW];   // [warp][layer input 0..3][unit]
  __shared__ float dact_s[kDpUWarps][3][kTcW];  // act'(z_l)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long r = (long long)blockIdx.x * kDpUWarps + warp; r < P.n; r += (long long)gridDim.x * kDpUWarps) {
    for (int i = lane; i < kTcW; i += 32) act_s[warp][0][i] = i < P.dims[0] ? P.x[r * P.dims[0] + i] : 0.0f;
    __syncwarp();
    for (int l = 0; l < 3; ++l) {
      const int in = P.dims[l], out = P.dims[l + 1];
      for (int v = 0; v < 4; ++v) {
        const int j = lane + 32 * v;
        float z = 0.0f, hh = 0.0f, dh = 0.0f;
        if (j < out) {
          const float* w = P.W[l] + (long long)j * in;
          for (int i = 0; i < in; ++i) z = fmaf(act_s[warp][l][i], w[i], z);
          dp_act_exact<ACT>(z + P.b[l][j], hh, dh);
        }
        act_s[warp][l + 1][j] = hh;
        dact_s[warp][l][j] = dh;
      }
      __syncwarp();
    }
    if (P.energy) {
      float es = 0.0f;
      for (int j = lane; j < P.dims[3]; j += 32) es = fmaf(P.w4[j], act_s[warp][3][j], es);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) es += __shfl_xor_sync(0xffffffffu, es, o);
      if (lane == 0) P.energy[r] = es + P.b4[0];
    }
    if (P.grad) {
      // delta_3 = w4 * act'(z3); delta_l = (W_{l+1}^T delta_{l+1}) * act'(z_l); grad = W1^T delta_1   (act_s reused for deltas)
      for (int j = lane; j < kTcW; j += 32) act_s[warp][3][j] = j < P.dims[3] ? P.w4[j] * dact_s[warp][2][j] : 0.0f;
      __syncwarp();
      for (int l = 2; l >= 0; --l) {
        const int in = P.dims[l], out = P.dims[l + 1];
        for (int v = 0; v < 4; ++v) {
          const int i = lane + 32 * v;
          float s = 0.0f;
          if (i < in)
            for (int j = 0; j < out; ++j) s = fmaf(act_s[warp][l + 1][j], P.W[l][(long long)j * in + i], s);
          if (l > 0) act_s[warp][l][i] = i < in ? s * dact_s[warp][l - 1][i] : 0.0f;
          else if (i < in) P.grad[r * in + i] = s;
        }
        __syncwarp();
      }
    }
    __syncwarp();
  }
}

static int deep_check(const EbmEnergyDesc* e) {
  if (!e->buf[7] || !e->buf[8]) { set_error("three-hidden-layer MLP needs W3 in buf[7] and b3 in buf[8]"); return EBM_ERR_INVALID; }
  if (e->dim > kTcW || e->hidden1 > kTcW || e->hidden2 > kTcW || e->hidden3 > kTcW) {
    set_error("three-hidden-layer MLP energies support widths <= %d (got %d-%d-%d-%d)", kTcW, e->dim, e->hidden1, e->hidden2,
              e->hidden3);
    return EBM_ERR_UNSUPPORTED;
  }
  if (e->activation < EBM_ACT_SILU || e->activation > EBM_ACT_SOFTPLUS) { set_error("bad activation"); return EBM_ERR_INVALID; }
  return 0;
}

int mlp_deep_energy_grad_dispatch(const EbmEnergyDesc* e, const float* x, int64_t n, float* energy, float* grad, cudaStream_t st) {
  int rc = deep_check(e);
  if (rc) return rc;
  const DeviceInfo& di = device_info(current_device());
  DeepUtilParams P;
  P.W[0] = e->buf[0]; P.b[0] = e->buf[1]; P.W[1] = e->buf[2]; P.b[1] = e->buf[3]; P.W[2] = e->buf[7]; P.b[2] = e->buf[8];
  P.w4 = e->buf[4]; P.b4 = e->buf[5];
  P.dims[0] = e->dim; P.dims[1] = e->hidden1; P.dims[2] = e->hidden2; P.dims[3] = e->hidden3;
  P.x = x; P.energy = energy; P.grad = grad; P.n = n;
  long long ctas = (n + kDpUWarps - 1) / kDpUWarps;
  const long long cap = (long long)di.sm_count * 4;
  const int grid = (int)(ctas < cap ? (ctas < 1 ? 1 : ctas) : cap);
#define CALL(A) mlp_deep_energy_grad_kernel<A><<<grid, 32 * kDpUWarps, 0, st>>>(P)
  switch (e->activation) {
    case EBM_ACT_SILU: CALL(EBM_ACT_SILU); break;
    case EBM_ACT_TANH: CALL(EBM_ACT_TANH); break;
    case EBM_ACT_RELU: CALL(EBM_ACT_RELU); break;
    default: CALL(EBM_ACT_SOFTPLUS); break;
  }
#undef CALL
  return launch_status("mlp_deep_energy_grad_kernel");
}

int langevin_mlp_deep_dispatch(const LangevinCall& c, int passes) {
  const EbmEnergyDesc* e = c.e;
  int rc = deep_check(e);
  if (rc) return rc;
  const DeviceInfo& di = device_info(current_device());
  const long long numel = (long long)c.n * e->dim;
  DeepParams P;
  memset(&P, 0, sizeof(P));
  P.W[0] = e->buf[0]; P.b[0] = e->buf[1]; P.W[1] = e->buf[2]; P.b[1] = e->buf[3]; P.W[2] = e->buf[7]; P.b[2] = e->buf[8];
  P.w4 = e->buf[4];
  P.d = e->dim; P.h1 = e->hidden1; P.h2 = e->hidden2; P.h3 = e->hidden3;
  P.passes = passes;
  P.n = c.n;
  P.thin = c.thin;
  P.n_kept = c.n_steps / c.thin;
  P.has_clamp = c.clamp != nullptr;
  if (c.clamp) { P.clamp_lo = c.clamp[0]; P.clamp_hi = c.clamp[1]; }
  P.traj = c.traj;
  P.rng.mode = c.rng_mode;
  if (c.rng_mode == EBM_RNG_TORCH) {
    P.rng.T = torch_threads(di, numel);
    P.rng.k0 = (uint32_t)c.seed; P.rng.k1 = (uint32_t)(c.seed >> 32);
    P.rng.ctr_step = torch_offset_increment(di, numel) / 4;
  } else {
    P.rng.T = 1;
    P.rng.k0 = (uint32_t)c.seed ^ kNativeTag0; P.rng.k1 = (uint32_t)(c.seed >> 32) ^ kNativeTag1;
    P.rng.ctr_step = 1;
  }
  philox_expand_keys(P.keys, P.rng.k0, P.rng.k1);
  const long long tiles = (c.n + kTcM - 1) / kTcM;
  const int sms = (e->sm_margin > 0 && e->sm_margin < di.sm_count) ? di.sm_count - e->sm_margin : di.sm_count;
  const int grid = (int)(tiles < sms ? tiles : sms);
  int* flags = reinterpret_cast<int*>(const_cast<float*>(e->buf[6]));   // NULL: whole tiles per CTA
  const bool uniform = c.schedule_len == 1;
  int done = 0;
  const float* src = c.x_in;
  while (done < c.n_steps) {
    const int chunk = uniform ? c.n_steps : ((c.n_steps - done < kSchedChunk) ? (c.n_steps - done) : kSchedChunk);
    StepTable tab;
    memset(&tab, 0, sizeof(tab));
    if (uniform) { fill_step(tab, 0, c.hs[0], c.nss[0]); tab.mask = 0; }
    else { for (int i = 0; i < chunk; ++i) fill_step(tab, i, c.hs[done + i], c.nss[done + i]); tab.mask = ~0; }
    P.x_in = src;
    P.x_out = c.x_out;
    P.row_index = (done == 0) ? c.row_index : nullptr;
    P.x_out2 = (done + chunk == c.n_steps) ? c.x_out2 : nullptr;
    P.n_steps = chunk;
    P.noise = c.noise ? c.noise + (long long)done * numel : nullptr;
    P.rng.ctr_base = c.offset / 4 + (unsigned long long)done * P.rng.ctr_step;
    P.step_base = done;
    if (flags) {
      int rc0 = mlp_schedule_setup(P.sched, tiles, chunk, grid, flags, c.st);
      if (rc0) return rc0;
    } else {
      mlp_schedule_whole_tiles(P.sched, tiles, chunk, grid);
    }
#define CALL(A)                                                                                                   \
  {                                                                                                               \
    if (passes == 3) {                                                                                            \
      auto kern = langevin_mlp_deep_kernel<A, true>;                                                              \
      EBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, DeepSmem::total));         \
      EBM_CUDA(mlp_launch_persistent(kern, grid, kTcThreads, DeepSmem::total, c.st, P, tab, tiles, chunk, grid)); \
    } else {                                                                                                      \
      auto kern = langevin_mlp_deep_kernel<A, false>;                                                             \
      EBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, DeepSmem::total));         \
      EBM_CUDA(mlp_launch_persistent(kern, grid, kTcThreads, DeepSmem::total, c.st, P, tab, tiles, chunk, grid)); \
    }                                                                                                             \
  }
    switch (e->activation) {
      case EBM_ACT_SILU: CALL(EBM_ACT_SILU); break;
      case EBM_ACT_TANH: CALL(EBM_ACT_TANH); break;
      case EBM_ACT_RELU: CALL(EBM_ACT_RELU); break;
      default: CALL(EBM_ACT_SOFTPLUS); break;
    }
#undef CALL
    int rc2 = launch_status("langevin_mlp_deep_kernel");
    if (rc2) return rc2;
    done += chunk;
    src = c.x_out;
  }
  return 0;
}

}  // namespace ebm
