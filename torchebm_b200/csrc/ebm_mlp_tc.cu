// MLP-energy Langevin burst on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Same math as ebm_mlp.cu (E = w3 . act(W2 act(W1 x + b1) + b2) + b3, grad = W1^T(act'(z1) * (W2^T(act'(z2) * w3))))
// with the four [128 x 128 x 128] products per Langevin step issued as tcgen05.mma (kind::f16, bf16 operands, fp32
// accumulators in tensor memory).  fp32-grade accuracy comes from split operands: every operand v is stored as
// hi = bf16(v), lo = bf16(v - hi) and each product is accumulated as hi*hi + lo*hi + hi*lo (error ~2^-16 relative to
// the largest term; the dropped lo*lo term is ~2^-18).  precision == EBM_MLP_BF16 skips the two correction passes.
//
// One CTA = one tile of 128 chains (TMEM lane = chain), persistent over tiles, resident for all K steps:
//   warp 0      : allocates TMEM, then one elected lane issues every MMA and commits to an mbarrier (warps 1-3 idle:
//                 they complete the warpgroup that hands its registers to the epilogue via setmaxnreg);
//   warps 4..19 : epilogue.  Thread = (row r = TMEM lane, column quarter); x[r, 32 cols] lives in its registers.
// Per step and tile the dependency chain is GEMM1 -> E1 -> GEMM2 -> E2 -> GEMM3 -> E3 -> GEMM4 -> E4(update) -> GEMM1'.
// Each epilogue produces the A operand of the NEXT product in 16-column chunks (= one MMA k-step) and signals a
// per-chunk mbarrier, so the MMA of product n+1 runs underneath epilogue n; the tensor pipe is hidden behind the
// CUDA-core work (activations, Philox, operand splitting), which is the real bound of this kernel.
//   TMEM columns: [0,128) acc0, [128,256) acc1 (products alternate), [256,384) act'(z1) kept for E3.
//   SMEM: W1/W2 hi+lo in the no-swizzle core-matrix layout of umma.cuh (one copy serves the K-major forward and the
//   MN-major backward descriptor), A hi+lo operand buffer, biases, barriers: 198 KB.
#include "mlp_tc_common.cuh"

namespace ebm {


template <int ACT, bool LO>
__global__ void __launch_bounds__(kTcThreads, 1) langevin_mlp_tc_kernel(const __grid_constant__ TcParams P,
                                                                        const __grid_constant__ StepTable tab) {
  extern __shared__ __align__(128) uint8_t tc_smem_raw[];
  uint8_t* smem = tc_smem_raw;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  tc_stage_weights(smem, P);
  if (threadIdx.x == 0) {
    for (int c = 0; c < kTcChunks; ++c) mbar_init(smem_u32(smem + TcSmemLayout::bars + c * 8), 4);  // 4 warps per chunk
    mbar_init(smem_u32(smem + TcSmemLayout::bars + kTcChunks * 8), 1);
    fence_mbar_init();
  }
  if (threadIdx.x == 32) mlp_units_compute(P.sched, P.n_steps, reinterpret_cast<volatile MlpUnits*>(smem + TcSmemLayout::units));
  if (warp == 0) tmem_alloc(smem_u32(smem + TcSmemLayout::tmem_slot), 512);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + TcSmemLayout::tmem_slot);
  const int k1 = (P.d + 15) / 16, k2 = (P.h1 + 15) / 16, k3 = (P.h2 + 15) / 16;
  const volatile MlpUnits* units = reinterpret_cast<const volatile MlpUnits*>(smem + TcSmemLayout::units);

  if (warp < kTcRoleWarps) {
    if (kTcRoleRegs < 96) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kTcRoleRegs));
    if (warp == 0) {   // converged: all lanes wait, one elected lane issues (umma.cuh: elect_one)
      const bool leader = elect_one();
      uint32_t parity = 0;
      for (int tile = units->t_last; tile >= units->t_first; --tile) {
        const int n_unit_steps = mlp_unit_s1(units, tile, P.n_steps) - mlp_unit_s0(units, tile);
        for (int k = 0; k < n_unit_steps; ++k) {
          tc_issue_gemm(smem, tmem + 0, TcSmemLayout::w1_hi, TcSmemLayout::w1_lo, false, k1, P.passes, parity, leader); parity ^= 1;
          tc_issue_gemm(smem, tmem + 128, TcSmemLayout::w2_hi, TcSmemLayout::w2_lo, false, k2, P.passes, parity, leader); parity ^= 1;
          tc_issue_gemm(smem, tmem + 0, TcSmemLayout::w2_hi, TcSmemLayout::w2_lo, true, k3, P.passes, parity, leader); parity ^= 1;
          tc_issue_gemm(smem, tmem + 128, TcSmemLayout::w1_hi, TcSmemLayout::w1_lo, true, k2, P.passes, parity, leader); parity ^= 1;
        }
      }
    }
  } else {
    // ---- epilogue warps -------------------------------------------------------------------------
    // Everything elementwise runs on packed fp32x2 pairs (pair j of a block = columns 2j, 2j+1).  Padded rows (>= n) and
    // padded columns (>= d, h1, h2) are computed like real ones and never stored: the weights, biases and w3 are staged
    // with zeros there, so they cannot leak into a real output, and they stay finite (a padded state column only
    // accumulates noise).
    if (kTcEpiRegs > 96) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kTcEpiRegs));
    const int e = warp - kTcRoleWarps;
    const int row = 32 * (warp & 3) + lane;        // TMEM lane this thread may access (hardware: warp % 4)
    const int cq = e >> 2;                          // column quarter
    const int col_base = kTcCols * cq;
    const int first_chunk = col_base / 16;
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + col_base;
    const f32x2* b1 = reinterpret_cast<const f32x2*>(smem + TcSmemLayout::b1 + 4 * col_base);
    const f32x2* b2 = reinterpret_cast<const f32x2*>(smem + TcSmemLayout::b2 + 4 * col_base);
    const f32x2* w3 = reinterpret_cast<const f32x2*>(smem + TcSmemLayout::w3 + 4 * col_base);
    const uint32_t acc_bar = smem_u32(smem + TcSmemLayout::bars + kTcChunks * 8);
    const long long numel = P.n * P.d;
    const bool quad_rng = (P.d % 4 == 0);
    // NATIVE stream: the step's noise is drawn in four 8-column parts, one in front of each wait for a product, and
    // parked in TMEM columns [384, 512) until the update epilogue -- the draw costs a third of the epilogue's
    // instructions and depends on nothing, so it fills the time the tensor pipe needs to finish each product
    const bool bubble_rng = (P.rng.mode == 2) && quad_rng;
    uint32_t parity = 0;

    for (int tile = units->t_last; tile >= units->t_first; --tile) {
      const long long grow = (long long)tile * kTcM + row;
      const bool rv = grow < P.n;
      const int s0 = mlp_unit_s0(units, tile), s1 = mlp_unit_s1(units, tile, P.n_steps);
      // a unit that starts mid-burst continues the chain another CTA left in x_out
      if (s0 > 0) mlp_unit_acquire(P.sched, kTcEpiWarps);
      const float* x0src = (s0 == 0) ? P.x_in : P.x_out;
      const long long row0 = (s0 == 0 && P.row_index && rv) ? P.row_index[grow] : grow;
      f32x2 X[kTcCols / 2];
      {
        float xs[kTcCols];
        tc_load_row32(x0src, row0, P.d, col_base, rv, xs);
#pragma unroll
        for (int j = 0; j < kTcCols / 2; ++j) X[j] = pack2(xs[2 * j], xs[2 * j + 1]);
      }
      store_a_16p(smem, row, col_base, X, LO);
      store_a_16p(smem, row, col_base + 16, X + 8, LO);
      signal_cols(smem, first_chunk, lane);
      int until_keep = P.thin - ((P.step_base + s0) % P.thin), kept = (P.step_base + s0) / P.thin;
      unsigned long long ctr = P.rng.ctr_base + (unsigned long long)s0 * P.rng.ctr_step;

      auto draw_part = [&](int part) {
        if (!bubble_rng) return;
        f32x2 e8[4];
        const uint64_t q0 = (uint64_t)(grow * P.d + col_base + 8 * part) >> 2;
#pragma unroll
        for (int q4 = 0; q4 < 2; ++q4) {
          const uint64_t q = q0 + q4;
          const uint4 w = philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)ctr, (uint32_t)(ctr >> 32), P.keys);
          normal4_fast_packed(w, e8[2 * q4], e8[2 * q4 + 1]);
        }
        tmem_st8p(lane_addr + 384 + 8 * part, e8);
      };

      for (int k = s0; k < s1; ++k) {
        const int ti = k & tab.mask;
        const float h = tab.h[ti], c12 = tab.c1[ti] * tab.c2[ti];
        // E1: z1 -> h1 (A of GEMM2); act'(z1) -> TMEM columns [256, 384)
        draw_part(0);
        mbar_wait(acc_bar, parity); parity ^= 1;
        tcgen05_fence_after();
#if EBM_TC_LD32
        f32x2 acc[16];
        tmem_ld32p(lane_addr + 0, acc);
#endif
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          f32x2 sd[8];
#if EBM_TC_LD32
          f32x2* v = acc + 8 * blk;
#else
          f32x2 v[8];
          tmem_ld16p(lane_addr + 0 + 16 * blk, v);
#endif
#pragma unroll
          for (int j = 0; j < 8; ++j) act2<ACT>(add2(v[j], b1[8 * blk + j]), v[j], sd[j]);
          tmem_st16p(lane_addr + 256 + 16 * blk, sd);
          store_a_16p(smem, row, col_base + 16 * blk, &v[0], LO);
          tcgen05_fence_before();
          signal_one(smem, first_chunk + blk, lane);
        }
        // E2: z2 -> delta2 = w3 * act'(z2) (A of GEMM3)
        draw_part(1);
        mbar_wait(acc_bar, parity); parity ^= 1;
        tcgen05_fence_after();
#if EBM_TC_LD32
        tmem_ld32p(lane_addr + 128, acc);
#endif
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
#if EBM_TC_LD32
          f32x2* v = acc + 8 * blk;
#else
          f32x2 v[8];
          tmem_ld16p(lane_addr + 128 + 16 * blk, v);
#endif
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            f32x2 hh, dh;
            act2<ACT>(add2(v[j], b2[8 * blk + j]), hh, dh);
            v[j] = mul2(dh, w3[8 * blk + j]);
          }
          store_a_16p(smem, row, col_base + 16 * blk, &v[0], LO);
          tcgen05_fence_before();
          signal_one(smem, first_chunk + blk, lane);
        }
        // E3: t -> delta1 = t * act'(z1) (A of GEMM4)
        draw_part(2);
        mbar_wait(acc_bar, parity); parity ^= 1;
        tcgen05_fence_after();
        tmem_st_wait();
#if EBM_TC_LD32
        tmem_ld32p_nowait(lane_addr + 0, acc);
#endif
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          f32x2 sd[8];
#if EBM_TC_LD32
          f32x2* v = acc + 8 * blk;
#else
          f32x2 v[8];
          tmem_ld16p_nowait(lane_addr + 0 + 16 * blk, v);
#endif
          tmem_ld16p(lane_addr + 256 + 16 * blk, sd);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = mul2(v[j], sd[j]);
          store_a_16p(smem, row, col_base + 16 * blk, &v[0], LO);
          tcgen05_fence_before();
          signal_one(smem, first_chunk + blk, lane);
        }
        // E4: g -> Langevin update of x; the new x is the A operand of the next step's GEMM1
        draw_part(3);
        mbar_wait(acc_bar, parity); parity ^= 1;
        tcgen05_fence_after();
        if (bubble_rng) tmem_st_wait();
        const bool last = (k == s1 - 1);
        bool keep_now = false;
        if (P.traj && --until_keep == 0) { until_keep = P.thin; keep_now = kept < P.n_kept; ++kept; }
#if EBM_TC_LD32
        tmem_ld32p_nowait(lane_addr + 128, acc);
#endif
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          f32x2 eps[8];
#if EBM_TC_LD32
          f32x2* g = acc + 8 * blk;
#else
          f32x2 g[8];
#endif
          const int c0 = col_base + 16 * blk;
          const long long li0 = grow * P.d + c0;
          if (bubble_rng) {
            tmem_ld16p_nowait(lane_addr + 384 + 16 * blk, eps);
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
#if EBM_TC_LD32
          tmem_ld_wait();
#else
          tmem_ld16p(lane_addr + 128 + 16 * blk, g);
#endif
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
          if (!last) {
            store_a_16p(smem, row, c0, X + 8 * blk, LO);
            tcgen05_fence_before();
            signal_one(smem, first_chunk + blk, lane);
          }
        }
        ctr += P.rng.ctr_step;
      }
      {
        float xs[kTcCols];
#pragma unroll
        for (int j = 0; j < kTcCols / 2; ++j) unpack2(X[j], xs[2 * j], xs[2 * j + 1]);
        tc_store_row32(P.x_out, grow, P.d, col_base, rv, xs);
        if (P.x_out2 && s1 == P.n_steps) tc_store_row32(P.x_out2, grow, P.d, col_base, rv, xs);
        if (s1 == P.n_steps)
          for (int w = 0; w < P.n_peers; ++w) tc_store_row32(P.peers[w] + P.peer_off, grow, P.d, col_base, rv, xs);
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


// ---- HMC on MLP energies on the tensor cores ----------------------------------------------------------------------
// Same tile / role / product structure as langevin_mlp_tc_kernel; the MMA thread runs the identical four-product chain
// once per gradient evaluation, (L + 1) evaluations per proposal: the force at the bottom of leapfrog step l is the
// force at the top of step l+1, the forward half of evaluation 0 yields E(x) and that of evaluation L yields E(x').
// Epilogue differences: E2 also accumulates w3 . act(z2) (the energy) for evaluations 0 and L; E4 applies the
// leapfrog kicks / drift instead of the Langevin update, with the momentum parked in TMEM columns [384, 512); after
// evaluation L the four column-quarter warps of a row combine their partial energies through shared memory (one named
// barrier per proposal) and every thread takes the Metropolis decision of its row.  The pre-proposal state is parked
// in x_out.  Force carrying is exact as long as safe-mode sanitising leaves x alone (the reference evaluates the drift
// at the top of every step, leapfrog.py:160: at an unchanged x that is the value carried from the bottom of the previous
// step).  When nan_to_num rewrites a coordinate of x the carried force is stale, and E(x') has to be taken at the
// sanitised state: the thread that sees it posts a token, and after the proposal's last evaluation the whole tile
// REDOES the proposal the reference's way -- same momentum (counter-based draw), no carrying: per step one evaluation
// for the top half-kick and drift, one for the bottom half-kick and sanitising, and a final forward pass for E(x') --
// 2L + 1 evaluations, the MMA warp learning the verdict through one mbarrier per proposal.  Rows that never tripped
// get the same trajectory again (the recomputed forces are bit-identical to the carried ones).
struct TcHmcParams {
  MlpSchedule sched;   // balanced (tile, proposal) split (the copy the kernel reads; T.sched is unused here)
  TcParams T;          // weights, widths, passes; T.n_steps = proposals of this launch * (L + 1)
  HmcParams H;
};

__device__ __forceinline__ int hmc_redo_token(int tile, int ip, int n_prop) {   // unique per (tile, proposal) of a launch; never 0
  return (int)((((long long)tile * n_prop + ip) & 0x3fffffff) + 1);
}

template <int ACT>
__global__ void __launch_bounds__(kTcThreads, 1) hmc_mlp_tc_kernel(const __grid_constant__ TcHmcParams Q,
                                                                   const __grid_constant__ HStepTable tab) {
  extern __shared__ __align__(128) uint8_t tc_smem_raw[];
  uint8_t* smem = tc_smem_raw;
  const TcParams& P = Q.T;
  const HmcParams& H = Q.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  tc_stage_weights(smem, P);
  if (threadIdx.x == 0) {
    for (int c = 0; c < kTcChunks; ++c) mbar_init(smem_u32(smem + TcSmemLayout::bars + c * 8), 4);
    mbar_init(smem_u32(smem + TcSmemLayout::bars + kTcChunks * 8), 1);
    mbar_init(smem_u32(smem + TcSmemLayout::hmc_verdict), 1);
    *reinterpret_cast<volatile int*>(smem + TcSmemLayout::hmc_redo) = 0;
    fence_mbar_init();
  }
  // work quantum = one proposal of one tile: a tile's proposals may be split between two CTAs (mlp_schedule.cuh); only
  // the chain state crosses the split (momentum, energy and force are rebuilt at the top of every proposal)
  if (threadIdx.x == 32) mlp_units_compute(Q.sched, H.n_prop, reinterpret_cast<volatile MlpUnits*>(smem + TcSmemLayout::units));
  if (warp == 0) tmem_alloc(smem_u32(smem + TcSmemLayout::tmem_slot), 512);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + TcSmemLayout::tmem_slot);
  const int k1 = (P.d + 15) / 16, k2 = (P.h1 + 15) / 16, k3 = (P.h2 + 15) / 16;
  const volatile MlpUnits* units = reinterpret_cast<const volatile MlpUnits*>(smem + TcSmemLayout::units);
  const int L = H.n_leapfrog;

  if (warp < kTcRoleWarps) {
    if (warp == 0) {   // converged: all lanes wait, one elected lane issues (umma.cuh: elect_one)
      const bool leader = elect_one();
      uint32_t parity = 0, vpar = 0;
      auto evaluations = [&](int n_evals) {
        for (int k = 0; k < n_evals; ++k) {
          tc_issue_gemm(smem, tmem + 0, TcSmemLayout::w1_hi, TcSmemLayout::w1_lo, false, k1, P.passes, parity, leader); parity ^= 1;
          tc_issue_gemm(smem, tmem + 128, TcSmemLayout::w2_hi, TcSmemLayout::w2_lo, false, k2, P.passes, parity, leader); parity ^= 1;
          tc_issue_gemm(smem, tmem + 0, TcSmemLayout::w2_hi, TcSmemLayout::w2_lo, true, k3, P.passes, parity, leader); parity ^= 1;
          tc_issue_gemm(smem, tmem + 128, TcSmemLayout::w1_hi, TcSmemLayout::w1_lo, true, k2, P.passes, parity, leader); parity ^= 1;
        }
      };
      for (int tile = units->t_last; tile >= units->t_first; --tile) {
        const int s0 = mlp_unit_s0(units, tile), s1 = mlp_unit_s1(units, tile, H.n_prop);
        for (int ip = s0; ip < s1; ++ip) {
          evaluations(L + 1);
          // the epilogue's verdict on this proposal: redo it without force carrying?
          mbar_wait(smem_u32(smem + TcSmemLayout::hmc_verdict), vpar); vpar ^= 1;
          if (*reinterpret_cast<volatile int*>(smem + TcSmemLayout::hmc_redo) == hmc_redo_token(tile, ip, H.n_prop))
            evaluations(2 * L + 1);
        }
      }
    }
  } else {
    const int e = warp - kTcRoleWarps;
    const int row = 32 * (warp & 3) + lane;
    const int cq = e >> 2;
    const int col_base = kTcCols * cq;
    const int first_chunk = col_base / 16;
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + col_base;
    const f32x2* b1 = reinterpret_cast<const f32x2*>(smem + TcSmemLayout::b1 + 4 * col_base);
    const f32x2* b2 = reinterpret_cast<const f32x2*>(smem + TcSmemLayout::b2 + 4 * col_base);
    const f32x2* w3 = reinterpret_cast<const f32x2*>(smem + TcSmemLayout::w3 + 4 * col_base);
    float* part = reinterpret_cast<float*>(smem + TcSmemLayout::hmc_part);   // [2][4 kinds][4 cq][128 rows]
    const uint32_t acc_bar = smem_u32(smem + TcSmemLayout::bars + kTcChunks * 8);
    const bool with_lo = P.passes == 3;
    const long long numel = H.n * H.d;
    const bool quad_rng = (H.d % 4 == 0);
    const float b3 = P.b3[0];
    uint32_t parity = 0;
    auto mass_div = [&](float v, int col) {   // leapfrog.py:167-177
      if (H.mass.kind == 1) return __fdiv_rn(v, H.mass.safe_scalar);
      if (H.mass.kind == 2) return __fdiv_rn(v, fmaxf(col < H.d ? H.mass.vec[col] : 1.0f, 1e-10f));
      return v;
    };
    auto kin_term = [&](float pv, int col) {   // hmc.py:148-159 (summand)
      const float sq = __fmul_rn(pv, pv);
      return (H.mass.kind == 2) ? __fdiv_rn(sq, col < H.d ? H.mass.vec[col] : 1.0f) : sq;
    };

    for (int tile = units->t_last; tile >= units->t_first; --tile) {
      const long long grow = (long long)tile * kTcM + row;
      const bool rv = grow < H.n;
      const int s0 = mlp_unit_s0(units, tile), s1 = mlp_unit_s1(units, tile, H.n_prop);
      if (s0 > 0) mlp_unit_acquire(Q.sched, kTcEpiWarps);   // earlier proposals of this tile ran on another CTA
      float x[kTcCols];
      tc_load_row32(s0 == 0 ? H.x_in : H.x_out, grow, H.d, col_base, rv, x);
      store_a_cols(smem, row, col_base, x, with_lo);
      signal_cols(smem, first_chunk, lane);
      RngStream rp, ru;
      rp.k0 = H.rng_p.k0; rp.k1 = H.rng_p.k1; rp.T = H.rng_p.T; rp.mode = H.rng_p.mode;
      rp.ctr_base = H.rng_p.ctr_base + (unsigned long long)s0 * H.rng_p.ctr_step;
      ru.k0 = H.rng_u.k0; ru.k1 = H.rng_u.k1; ru.T = H.rng_u.T; ru.mode = H.rng_u.mode;
      ru.ctr_base = H.rng_u.ctr_base + (unsigned long long)s0 * H.rng_u.ctr_step;
      int until_keep = H.thin_start - s0, kept = H.kept_base;
      if (until_keep <= 0) {   // thin_start / kept_base describe proposal 0 of this launch
        const int passed = (-until_keep) / H.thin + 1;
        kept += passed;
        until_keep += passed * H.thin;
      }
      float e_final = 0.0f;

      for (int ip = s0; ip < s1; ++ip) {
        const float h = tab.h[ip & tab.mask];
        const float half_h = __fmul_rn(0.5f, h);
        float* pp = part + (ip & 1) * (4 * 4 * kTcM);
        float* e0p = pp + (0 * 4 + cq) * kTcM;
        float* e1p = pp + (1 * 4 + cq) * kTcM;
        float* k0p = pp + (2 * 4 + cq) * kTcM;
        float* k1p = pp + (3 * 4 + cq) * kTcM;
        const int redo_token = hmc_redo_token(tile, ip, H.n_prop);
        volatile int* redo_flag = reinterpret_cast<volatile int*>(smem + TcSmemLayout::hmc_redo);
        bool slow = false;   // second pass of this proposal: no force carrying (header comment)
        for (;;) {
        // park the pre-proposal state, draw the momentum (hmc.py:245 / :92-134) into TMEM, K(p) partial
        {
          tc_store_row32(H.x_out, grow, H.d, col_base, rv, x);
          float ksum = 0.0f;
#pragma unroll
          for (int blk = 0; blk < 2; ++blk) {
            float pv[16];
            const int c0 = col_base + 16 * blk;
            const long long li0 = grow * H.d + c0;
            if (H.rng_p.mode == 2 && quad_rng) {
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) {
                const uint64_t q = (uint64_t)(li0 + 4 * q4) >> 2;
                const uint4 w = philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)rp.ctr_base,
                                              (uint32_t)(rp.ctr_base >> 32), rp.k0, rp.k1);
                const float4 nn = normal4(w);   // accurate transform: identical to the per-element path of the other HMC kernels
                pv[4 * q4] = nn.x; pv[4 * q4 + 1] = nn.y; pv[4 * q4 + 2] = nn.z; pv[4 * q4 + 3] = nn.w;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const bool in = rv && (c0 + i) < H.d;
                float ev = 0.0f;
                if (in) ev = (H.rng_p.mode == 0) ? H.noise_p[(long long)ip * numel + li0 + i]
                                                 : normal_for_element_call(rp.k0, rp.k1, rp.ctr_base, rp.T, rp.mode, (uint64_t)(li0 + i));
                pv[i] = ev;
              }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int col = c0 + i;
              const bool in = rv && col < H.d;
              float v = in ? pv[i] : 0.0f;
              if (H.mass.kind == 1) v = __fmul_rn(v, H.mass.sqrt_scalar);                       // hmc.py:124
              else if (H.mass.kind == 2) v = __fmul_rn(v, sqrtf(col < H.d ? H.mass.vec[col] : 1.0f));   // hmc.py:133
              pv[i] = v;
              ksum += kin_term(v, col);
            }
            tmem_st16(lane_addr + 384 + 16 * blk, pv);
          }
          k0p[row] = ksum;
        }
        const int n_ev = slow ? 2 * L + 1 : L + 1;
        for (int l = 0; l < n_ev; ++l) {
          const bool want_e = (l == 0) || (l == n_ev - 1);
          // E1: z1 -> h1 ; act'(z1) -> TMEM [256, 384)   (E1-E3 on packed fp32x2 pairs, as in langevin_mlp_tc_kernel)
          mbar_wait(acc_bar, parity); parity ^= 1;
          tcgen05_fence_after();
          f32x2 acc[16];
          tmem_ld32p(lane_addr + 0, acc);
#pragma unroll
          for (int blk = 0; blk < 2; ++blk) {
            f32x2 sd[8];
            f32x2* v = acc + 8 * blk;
#pragma unroll
            for (int i = 0; i < 8; ++i) act2<ACT>(add2(v[i], b1[8 * blk + i]), v[i], sd[i]);
            tmem_st16p(lane_addr + 256 + 16 * blk, sd);
            store_a_16p(smem, row, col_base + 16 * blk, v, with_lo);
            tcgen05_fence_before();
            signal_one(smem, first_chunk + blk, lane);
          }
          // E2: z2 -> delta2 = w3 * act'(z2) ; energy partial w3 . act(z2)
          mbar_wait(acc_bar, parity); parity ^= 1;
          tcgen05_fence_after();
          f32x2 esum2 = pack2(0.0f, 0.0f);
          tmem_ld32p(lane_addr + 128, acc);
#pragma unroll
          for (int blk = 0; blk < 2; ++blk) {
            f32x2* v = acc + 8 * blk;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              f32x2 hh, dh;
              act2<ACT>(add2(v[i], b2[8 * blk + i]), hh, dh);
              if (want_e) esum2 = fma2(w3[8 * blk + i], hh, esum2);
              v[i] = mul2(dh, w3[8 * blk + i]);
            }
            store_a_16p(smem, row, col_base + 16 * blk, v, with_lo);
            tcgen05_fence_before();
            signal_one(smem, first_chunk + blk, lane);
          }
          if (want_e) {
            float ea, eb;
            unpack2(esum2, ea, eb);
            if (l == 0) e0p[row] = ea + eb; else e1p[row] = ea + eb;
          }
          // E3: t -> delta1 = t * act'(z1)
          mbar_wait(acc_bar, parity); parity ^= 1;
          tcgen05_fence_after();
          tmem_st_wait();
          tmem_ld32p_nowait(lane_addr + 0, acc);
#pragma unroll
          for (int blk = 0; blk < 2; ++blk) {
            f32x2 sd[8];
            f32x2* v = acc + 8 * blk;
            tmem_ld16p(lane_addr + 256 + 16 * blk, sd);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = mul2(v[i], sd[i]);
            store_a_16p(smem, row, col_base + 16 * blk, v, with_lo);
            tcgen05_fence_before();
            signal_one(smem, first_chunk + blk, lane);
          }
          // E4: force -> leapfrog kick(s) and drift (leapfrog.py:160-185, safe mode)
          mbar_wait(acc_bar, parity); parity ^= 1;
          tcgen05_fence_after();
          float ksum = 0.0f;
          // three straight-line variants: first evaluation (top of step 1), middle (bottom of step l + top of step l+1),
          // last (bottom of step L + kinetic energy) -- `l` is uniform, so this only removes per-element predication
          bool dirty = false;   // sanitising rewrote a coordinate of x: the carried force is stale from here on
          auto e4 = [&](auto first_c, auto last_c, unsigned todo) {
            constexpr bool kFirst = decltype(first_c)::value, kLast = decltype(last_c)::value;
#pragma unroll
            for (int blk = 0; blk < 2; ++blk) {
              if (!((todo >> blk) & 1)) continue;   // (warp-uniform only for the mass kinds; per-thread after the detector)
              float g[16], pv[16];
              tmem_ld16(lane_addr + 128 + 16 * blk, g);
              tmem_ld16(lane_addr + 384 + 16 * blk, pv);
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int col = col_base + 16 * blk + i;
                const float f = clamp_torch(-g[i], -kSafeClamp, kSafeClamp);
                float xv = x[16 * blk + i], p_ = pv[i];
                if (!kFirst) {                     // bottom of step l: second half kick, then sanitise
                  p_ = __fadd_rn(p_, __fmul_rn(half_h, f));
                  dirty = dirty || !(fabsf(xv) <= 3.402823466e+38f);
                  xv = nan_to_num0(xv);
                  p_ = nan_to_num0(p_);
                }
                if (!kLast) {                      // top of step l+1: first half kick and drift
                  p_ = __fadd_rn(p_, __fmul_rn(half_h, f));
                  xv = __fadd_rn(xv, mass_div(__fmul_rn(h, p_), col));
                } else {
                  ksum += kin_term(p_, col);
                }
                const bool in = rv && col < H.d;
                x[16 * blk + i] = in ? xv : 0.0f;
                pv[i] = in ? p_ : 0.0f;
              }
              if (!kLast) {
                tmem_st16(lane_addr + 384 + 16 * blk, pv);
                store_a_16(smem, row, col_base + 16 * blk, x + 16 * blk, with_lo);
                tcgen05_fence_before();
                signal_one(smem, first_chunk + blk, lane);
              }
            }
          };
          // Unit mass (the reference's default): the same update on packed pairs.  The safe-mode clamp and the sanitising
          // leave finite values below 1e6 untouched, so the fast path only has to DETECT anything else -- v * 0 accumulated
          // over a 16-column block is 0 unless some v is NaN / inf, one packed FMA per pair -- and the rows of a block that
          // trips the detector (a diverged chain) go through the exact scalar code above instead.  Kicks and drift are fused
          // multiply-adds here: their rounding is far inside the split-operand force error.
          auto e4_fast = [&](auto first_c, auto last_c) -> unsigned {
            constexpr bool kFirst = decltype(first_c)::value, kLast = decltype(last_c)::value;
            const f32x2 zero2 = pack2(0.0f, 0.0f);
            unsigned todo = 0;
#pragma unroll
            for (int blk = 0; blk < 2; ++blk) {
              f32x2 G[8], Pm[8], Xn[8];
              tmem_ld16p_nowait(lane_addr + 128 + 16 * blk, G);
              tmem_ld16p(lane_addr + 384 + 16 * blk, Pm);
              f32x2 det = zero2;
              float kblk = 0.0f;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float ga, gb;
                unpack2(G[i], ga, gb);
                det = fma2(G[i], zero2, det);
                const f32x2 F = pack2(fminf(fmaxf(-ga, -kSafeClamp), kSafeClamp), fminf(fmaxf(-gb, -kSafeClamp), kSafeClamp));
                f32x2 Xv = pack2(x[16 * blk + 2 * i], x[16 * blk + 2 * i + 1]);
                f32x2 Pv = Pm[i];
                if (!kFirst) {                     // bottom of step l: second half kick; sanitising = detection only
                  Pv = fma2(F, half_h, Pv);
                  det = fma2(Xv, zero2, det);
                  det = fma2(Pv, zero2, det);
                }
                if (!kLast) {                      // top of step l+1: first half kick and drift
                  Pv = fma2(F, half_h, Pv);
                  Xv = fma2(Pv, h, Xv);
                } else {
                  float pa, pb;
                  unpack2(Pv, pa, pb);
                  kblk = fmaf(pa, pa, fmaf(pb, pb, kblk));
                }
                Xn[i] = Xv;
                Pm[i] = Pv;
              }
              float da, db;
              unpack2(det, da, db);
              const bool clean = (da + db == 0.0f);
              // the tensor core needs the chunk of every row: rows that tripped the detector publish from the exact path
              if (!__all_sync(0xffffffffu, clean)) { todo |= 1u << blk; continue; }
#pragma unroll
              for (int i = 0; i < 8; ++i) unpack2(Xn[i], x[16 * blk + 2 * i], x[16 * blk + 2 * i + 1]);
              if (kLast) ksum += kblk;
              if (!kLast) {
                tmem_st16p(lane_addr + 384 + 16 * blk, Pm);
                store_a_16p(smem, row, col_base + 16 * blk, Xn, with_lo);
                tcgen05_fence_before();
                signal_one(smem, first_chunk + blk, lane);
              }
            }
            return todo;
          };
          // second pass (no carrying): even evaluations give the force of a step's top half (kick + drift), odd ones the
          // force of its bottom half (kick + sanitising); evaluation 2L is only the forward pass for E(x')
          auto e4_slow = [&](bool bottom, bool kin) {
#pragma unroll
            for (int blk = 0; blk < 2; ++blk) {   // (unrolled: x[] must keep static indices to stay in registers)
              float g[16], pv[16];
              tmem_ld16(lane_addr + 128 + 16 * blk, g);
              tmem_ld16(lane_addr + 384 + 16 * blk, pv);
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int col = col_base + 16 * blk + i;
                const float f = clamp_torch(-g[i], -kSafeClamp, kSafeClamp);
                float xv = x[16 * blk + i];
                float p_ = __fadd_rn(pv[i], __fmul_rn(half_h, f));
                if (bottom) {
                  xv = nan_to_num0(xv);
                  p_ = nan_to_num0(p_);
                  if (kin) ksum += kin_term(p_, col);
                } else {
                  xv = __fadd_rn(xv, mass_div(__fmul_rn(h, p_), col));
                }
                const bool in = rv && col < H.d;
                x[16 * blk + i] = in ? xv : 0.0f;
                pv[i] = in ? p_ : 0.0f;
              }
              tmem_st16(lane_addr + 384 + 16 * blk, pv);
              store_a_16(smem, row, col_base + 16 * blk, x + 16 * blk, with_lo);
              tcgen05_fence_before();
              signal_one(smem, first_chunk + blk, lane);
            }
          };
          if (slow) {
            if (l < 2 * L) e4_slow((l & 1) != 0, l == 2 * L - 1);
            if (l == 2 * L - 1) k1p[row] = ksum;
          } else {
            unsigned todo = 3;
            if (H.mass.kind == 0) {
              if (l == 0) todo = e4_fast(std::true_type{}, std::false_type{});
              else if (l < L) todo = e4_fast(std::false_type{}, std::false_type{});
              else todo = e4_fast(std::false_type{}, std::true_type{});
            }
            if (todo) {
              if (l == 0) e4(std::true_type{}, std::false_type{}, todo);
              else if (l < L) e4(std::false_type{}, std::false_type{}, todo);
              else e4(std::false_type{}, std::true_type{}, todo);
            }
            if (dirty) *redo_flag = redo_token;
            if (l == L) k1p[row] = ksum;
          }
        }
        // all four column quarters of every row have written their partial sums (and any redo token)
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kTcEpiWarps) : "memory");
        if (slow) break;
        if (e == 0 && lane == 0) mbar_arrive(smem_u32(smem + TcSmemLayout::hmc_verdict));
        if (*redo_flag != redo_token) break;
        slow = true;   // restart from the parked pre-proposal state; the momentum draw repeats itself (counter-based)
        tc_load_row32(H.x_out, grow, H.d, col_base, rv, x);
        store_a_cols(smem, row, col_base, x, with_lo);
        signal_cols(smem, first_chunk, lane);
        }
        float e0 = b3, e1 = b3, kk0 = 0.0f, kk1 = 0.0f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          e0 += pp[(0 * 4 + c) * kTcM + row];
          e1 += pp[(1 * 4 + c) * kTcM + row];
          kk0 += pp[(2 * 4 + c) * kTcM + row];
          kk1 += pp[(3 * 4 + c) * kTcM + row];
        }
        kk0 = __fmul_rn(0.5f, kk0); kk1 = __fmul_rn(0.5f, kk1);
        if (H.mass.kind == 1) { kk0 = __fdiv_rn(kk0, H.mass.scalar); kk1 = __fdiv_rn(kk1, H.mass.scalar); }
        const float h0 = __fadd_rn(clamp_torch(e0, -1e10f, 1e10f), clamp_torch(kk0, 0.0f, 1e10f));   // hmc.py:247-256
        const float h1 = __fadd_rn(clamp_torch(e1, -1e10f, 1e10f), clamp_torch(kk1, 0.0f, 1e10f));   // hmc.py:268-275
        const float dh = clamp_torch(__fsub_rn(h0, h1), -50.0f, 50.0f);
        float a = expf(dh);
        a = (a != a) ? a : fminf(a, 1.0f);
        float u = 0.0f;
        if (rv) u = (H.rng_u.mode == 0) ? H.noise_u[(long long)ip * H.n + grow] : uniform_for_element(ru, (uint64_t)grow);
        const bool accepted = u < a;
        e_final = accepted ? e1 : e0;
        if (!accepted) tc_load_row32(H.x_out, grow, H.d, col_base, rv, x);
        if (H.accept_count && cq == 0) {
          const unsigned m = __ballot_sync(0xffffffffu, rv && accepted);
          if (lane == 0 && m) atomicAdd(H.accept_count + H.prop_base + ip, __popc(m));
        }
        rp.ctr_base += H.rng_p.ctr_step;
        ru.ctr_base += H.rng_u.ctr_step;
        if (H.traj && --until_keep == 0) {
          until_keep = H.thin;
          if (kept < H.n_kept && rv) {
#pragma unroll
            for (int i = 0; i < kTcCols; ++i)
              if (col_base + i < H.d) H.traj[(grow * H.n_kept + kept) * H.d + col_base + i] = x[i];
          }
          ++kept;
        }
        if (ip + 1 < s1) {   // the selected state is the A operand of the next proposal's first evaluation
          store_a_cols(smem, row, col_base, x, with_lo);
          signal_cols(smem, first_chunk, lane);
        }
      }
      tc_store_row32(H.x_out, grow, H.d, col_base, rv, x);
      if (rv && H.energy_out && cq == 0 && s1 == H.n_prop) H.energy_out[grow] = clamp_torch(e_final, -1e10f, 1e10f);
      if (s1 < H.n_prop) mlp_unit_release(Q.sched);   // the rest of this tile's proposals run on the next CTA
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem, 512);
  }
}

// one chunk of proposals of ebm_hmc_burst_f32 for an MLP energy on the tensor cores (chunk loop: ebm_hmc.cu)
int hmc_mlp_tc_launch(const EbmEnergyDesc* e, const HmcParams& H, const HStepTable& tab, int passes, cudaStream_t st) {
  const DeviceInfo& di = device_info(current_device());
  TcHmcParams Q;
  memset(&Q, 0, sizeof(Q));
  Q.H = H;
  TcParams& P = Q.T;
  P.W1 = e->buf[0]; P.b1 = e->buf[1]; P.W2 = e->buf[2]; P.b2 = e->buf[3]; P.w3 = e->buf[4]; P.b3 = e->buf[5];
  P.d = e->dim; P.h1 = e->hidden1; P.h2 = e->hidden2;
  P.passes = passes;
  P.n = H.n;
  P.n_steps = H.n_prop * (H.n_leapfrog + 1);
  const long long tiles = (H.n + kTcM - 1) / kTcM;
  const int sms = (e->sm_margin > 0 && e->sm_margin < di.sm_count) ? di.sm_count - e->sm_margin : di.sm_count;
  const int grid = (int)(tiles < sms ? tiles : sms);
  int* flags = reinterpret_cast<int*>(const_cast<float*>(e->buf[6]));  // NULL: whole tiles per CTA
  if (flags) {
    int rc0 = mlp_schedule_setup(Q.sched, tiles, H.n_prop, grid, flags, st);
    if (rc0) return rc0;
  } else {
    mlp_schedule_whole_tiles(Q.sched, tiles, H.n_prop, grid);
  }
#define CALL(A)                                                                                                  \
  {                                                                                                              \
    auto kern = hmc_mlp_tc_kernel<A>;                                                                            \
    EBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmemLayout::hmc_total));  \
    EBM_CUDA(mlp_launch_persistent(kern, grid, kTcThreads, TcSmemLayout::hmc_total, st, Q, tab, tiles, H.n_prop, grid)); \
  }
  switch (e->activation) {
    case EBM_ACT_SILU: CALL(EBM_ACT_SILU); break;
    case EBM_ACT_TANH: CALL(EBM_ACT_TANH); break;
    case EBM_ACT_RELU: CALL(EBM_ACT_RELU); break;
    default: CALL(EBM_ACT_SOFTPLUS); break;
  }
#undef CALL
  return launch_status("hmc_mlp_tc_kernel");
}

int langevin_mlp_tc1_dispatch(const LangevinCall& c, int passes) {
  const EbmEnergyDesc* e = c.e;
  if (e->dim > kTcW || e->hidden1 > kTcW || e->hidden2 > kTcW) {
    set_error("tensor-core MLP kernel supports widths up to %d", kTcW);
    return EBM_ERR_UNSUPPORTED;
  }
  const DeviceInfo& di = device_info(current_device());
  const long long numel = (long long)c.n * e->dim;
  TcParams P;
  memset(&P, 0, sizeof(P));
  P.W1 = e->buf[0]; P.b1 = e->buf[1]; P.W2 = e->buf[2]; P.b2 = e->buf[3]; P.w3 = e->buf[4]; P.b3 = e->buf[5];
  P.d = e->dim; P.h1 = e->hidden1; P.h2 = e->hidden2;
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
  int* flags = reinterpret_cast<int*>(const_cast<float*>(e->buf[6]));  // NULL: whole tiles per CTA (no balancing)
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
    P.n_peers = 0;
    if (c.n_peers > 0 && done + chunk == c.n_steps) {
      P.n_peers = c.peer_mc ? 0 : c.n_peers;   // (the single-tile A/B kernel has no multicast epilogue)
      P.peer_off = c.peer_row_offset * e->dim;
      for (int w = 0; w < c.n_peers; ++w) P.peers[w] = c.peers[w];
    }
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
#define CALL(A)                                                                                               \
  {                                                                                                           \
    auto kern = passes == 3 ? langevin_mlp_tc_kernel<A, true> : langevin_mlp_tc_kernel<A, false>;             \
    EBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmemLayout::total));   \
    EBM_CUDA(mlp_launch_persistent(kern, grid, kTcThreads, TcSmemLayout::total, c.st, P, tab, tiles, chunk, grid)); \
  }
    switch (e->activation) {
      case EBM_ACT_SILU: CALL(EBM_ACT_SILU); break;
      case EBM_ACT_TANH: CALL(EBM_ACT_TANH); break;
      case EBM_ACT_RELU: CALL(EBM_ACT_RELU); break;
      default: CALL(EBM_ACT_SOFTPLUS); break;
    }
#undef CALL
    int rc = launch_status("langevin_mlp_tc_kernel");
    if (rc) return rc;
    done += chunk;
    src = c.x_out;
  }
  return 0;
}

}  // namespace ebm
