// MLP-energy Langevin burst on the tensor cores, two tile pipelines per SM (sm_100a: tcgen05 + TMEM).
//
// Same math and operand format as ebm_mlp_tc.cu (four [128 x 128 x 128] products per Langevin step, bf16 hi/lo split
// operands, fp32 accumulators in tensor memory).  What changes is the shape of the pipeline.  A single tile is a strictly
// serial chain GEMM1 -> E1 -> GEMM2 -> E2 -> GEMM3 -> E3 -> GEMM4 -> E4: the tensor pipe idles while the CUDA cores run an
// epilogue and the other way round (ncu, one tile per SM: tensor pipe 33 % active, issue slots 37 % active).  Here one
// CTA runs TWO independent tiles ("pipelines"), each with its own epilogue warps, accumulator and operand ring, and one
// MMA-issue thread serves whichever pipeline has an operand chunk ready: tile A's products run underneath tile B's
// epilogues.  The budget that makes two tiles fit next to the 128 KB of W1/W2 hi+lo:
//   TMEM (512 columns)  per pipeline: two accumulator regions of 128 columns.  In a step, region a takes z1 (E1 overwrites
//                       it in place with act'(z1), which E3 needs) and later grad; region 1-a takes z2 and then t; a
//                       flips every step.  z1 and z2 always find their region free -- those products start under the
//                       previous epilogue, chunk by chunk -- while t and grad wait until E2 / E3 have drained theirs.
//   SMEM (227 KB)       per pipeline: a ring of 6 operand slots; a slot is one MMA k-step of the A operand (16 columns x
//                       128 rows, hi 4 KB + lo 4 KB, the chunk-contiguous core-matrix layout of umma.cuh).  Epilogue
//                       warps fill slots in the order the MMA thread consumes them; tcgen05.commit on the slot's "free"
//                       mbarrier recycles it, so 6 slots carry the 8 chunks of a product.
//   registers           16 epilogue warps x 112 (8 per pipeline: TMEM lane quarter x column half, the thread's 64 state
//                       columns live in registers for the whole burst) + a role warpgroup x 32 (setmaxnreg).
// The NATIVE-stream noise term is added to the register-resident state in four instalments, one in front of each wait
// for a product (no TMEM left to park it).  Work split: mlp_schedule.cuh with one worker per PIPELINE (2 x grid), so both pipelines of every
// SM end together; hand-over of a tile between workers through x_out as before.
#include "mlp_tc_common.cuh"

namespace ebm {

constexpr int kT2Pipes = 2;
constexpr int kT2PipeWarps = 8;                                   // 4 lane quarters x 2 column halves
constexpr int kT2EpiWarps = kT2Pipes * kT2PipeWarps;
constexpr int kT2RoleWarps = 4;                                   // one warpgroup: its first warp issues the MMAs
constexpr int kT2Threads = 32 * (kT2EpiWarps + kT2RoleWarps);
// 640 threads launch with 96 registers each (5 warps x 96 per scheduler); the role warpgroup then shrinks and the four
// epilogue warpgroups grow: 32 + 4 x 112 = 5 x 96
#ifndef EBM_T2_ROLE_REGS
#define EBM_T2_ROLE_REGS 32
#define EBM_T2_EPI_REGS 112
#endif
constexpr int kT2Cols = kTcW / 2;                                 // columns per epilogue thread (64)
constexpr int kT2Blocks = kT2Cols / 16;                           // 16-column blocks per thread = chunks per column half
constexpr int kT2Slots = 6;
constexpr int kT2SlotHalf = kTcM * 16 * 2;                        // one bf16 [128 x 16] chunk: 4096 B
constexpr int kT2SlotBytes = 2 * kT2SlotHalf;                     // hi + lo
constexpr int kT2PipeBars = 2 * kT2Slots + 4;                     // slot full[6], slot free[6], region full[2], region free[2]

struct T2Smem {
  static constexpr int w1_hi = 0;
  static constexpr int w1_lo = w1_hi + kTcMatBytes;
  static constexpr int w2_hi = w1_lo + kTcMatBytes;
  static constexpr int w2_lo = w2_hi + kTcMatBytes;
  static constexpr int ring = w2_lo + kTcMatBytes;               // [pipe][slot][hi | lo]
  static constexpr int b1 = ring + kT2Pipes * kT2Slots * kT2SlotBytes;
  static constexpr int b2 = b1 + kTcW * 4;
  static constexpr int w3 = b2 + kTcW * 4;
  static constexpr int bars = w3 + kTcW * 4;
  static constexpr int tmem_slot = bars + kT2Pipes * kT2PipeBars * 8;
  static constexpr int units = tmem_slot + 16;                   // one MlpUnits per pipeline
  static constexpr int total = units + kT2Pipes * 16;
};
static_assert(T2Smem::total <= 227 * 1024, "two-pipeline tensor-core kernel exceeds the shared memory of an SM");

__device__ __forceinline__ uint32_t t2_bar_full(uint8_t* smem, int g, int slot) { return smem_u32(smem + T2Smem::bars + (g * kT2PipeBars + slot) * 8); }
__device__ __forceinline__ uint32_t t2_bar_free(uint8_t* smem, int g, int slot) { return smem_u32(smem + T2Smem::bars + (g * kT2PipeBars + kT2Slots + slot) * 8); }
__device__ __forceinline__ uint32_t t2_bar_reg_full(uint8_t* smem, int g, int x) { return smem_u32(smem + T2Smem::bars + (g * kT2PipeBars + 2 * kT2Slots + x) * 8); }
__device__ __forceinline__ uint32_t t2_bar_reg_free(uint8_t* smem, int g, int x) { return smem_u32(smem + T2Smem::bars + (g * kT2PipeBars + 2 * kT2Slots + 2 + x) * 8); }

// non-blocking probe of an mbarrier phase
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

// a thread's 64 consecutive columns of one row <-> global memory, as packed pairs
__device__ __forceinline__ void t2_load_row(const float* __restrict__ src, long long grow, int d, int col_base, bool rv,
                                            f32x2 (&X)[kT2Cols / 2]) {
  const float* p = src + grow * d + col_base;
  if ((d & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
#pragma unroll
    for (int j = 0; j < kT2Cols / 4; ++j) {
      float4 t = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (rv && col_base + 4 * j < d) t = reinterpret_cast<const float4*>(p)[j];
      X[2 * j] = pack2(t.x, t.y);
      X[2 * j + 1] = pack2(t.z, t.w);
    }
  } else {
#pragma unroll
    for (int j = 0; j < kT2Cols / 2; ++j) {
      const float a = (rv && col_base + 2 * j < d) ? p[2 * j] : 0.0f;
      const float b = (rv && col_base + 2 * j + 1 < d) ? p[2 * j + 1] : 0.0f;
      X[j] = pack2(a, b);
    }
  }
}
// `mc`: dst is an NVLS multicast address (multimem stores, langevin_elem.cuh)
__device__ __forceinline__ void t2_store_row(float* __restrict__ dst, long long grow, int d, int col_base, bool rv,
                                             const f32x2 (&X)[kT2Cols / 2], bool mc = false) {
  if (!rv) return;
  float* p = dst + grow * d + col_base;
  if ((d & 3) == 0 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
    for (int j = 0; j < kT2Cols / 4; ++j) {
      if (col_base + 4 * j < d) {
        float4 t;
        unpack2(X[2 * j], t.x, t.y);
        unpack2(X[2 * j + 1], t.z, t.w);
        if (mc) mc_store4(p + 4 * j, t.x, t.y, t.z, t.w);
        else reinterpret_cast<float4*>(p)[j] = t;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < kT2Cols / 2; ++j) {
      float a, b;
      unpack2(X[j], a, b);
      if (col_base + 2 * j < d) { if (mc) mc_store1(p + 2 * j, a); else p[2 * j] = a; }
      if (col_base + 2 * j + 1 < d) { if (mc) mc_store1(p + 2 * j + 1, b); else p[2 * j + 1] = b; }
    }
  }
}

template <int ACT, bool LO>
__global__ void __launch_bounds__(kT2Threads, 1) langevin_mlp_tc2_kernel(const __grid_constant__ TcParams P,
                                                                         const __grid_constant__ StepTable tab) {
  extern __shared__ __align__(128) uint8_t t2_smem_raw[];
  uint8_t* smem = t2_smem_raw;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  tc_stage_weights<T2Smem>(smem, P);
  if (threadIdx.x == 0) {
    for (int g = 0; g < kT2Pipes; ++g) {
      for (int s = 0; s < kT2Slots; ++s) {
        mbar_init(t2_bar_full(smem, g, s), 4);   // the four lane-quarter warps that write a chunk
        mbar_init(t2_bar_free(smem, g, s), 1);   // tcgen05.commit
      }
      for (int x = 0; x < 2; ++x) {
        mbar_init(t2_bar_reg_full(smem, g, x), 1);
        mbar_init(t2_bar_reg_free(smem, g, x), kT2PipeWarps);
      }
    }
    fence_mbar_init();
  }
  if (threadIdx.x < kT2Pipes)
    mlp_units_compute(P.sched, P.n_steps, reinterpret_cast<volatile MlpUnits*>(smem + T2Smem::units) + threadIdx.x,
                      (long long)kT2Pipes * blockIdx.x + threadIdx.x);
  if (warp == kT2EpiWarps) tmem_alloc(smem_u32(smem + T2Smem::tmem_slot), 512);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + T2Smem::tmem_slot);
  const int k1 = (P.d + 15) / 16, k2 = (P.h1 + 15) / 16, k3 = (P.h2 + 15) / 16;

  if (warp >= kT2EpiWarps) {
    // ---- MMA issue: role warp g serves pipeline g (the other two warps of the warpgroup only complete it) --------
    // The whole warp runs the loop so that every value below is warp-uniform (descriptors live in uniform registers);
    // one elected lane issues.  Per chunk the issuer does two mbarrier waits, three MMAs and a commit: it must keep up
    // with the tensor pipe (3 x 64 cycles per chunk), which a single thread serving both pipelines through a polling
    // loop did not (measured: ~1000 cycles per chunk).
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(EBM_T2_ROLE_REGS));
    const int g = warp - kT2EpiWarps;
    if (g < kT2Pipes) {
      const volatile MlpUnits* un = reinterpret_cast<const volatile MlpUnits*>(smem + T2Smem::units) + g;
      const uint32_t bars = smem_u32(smem + T2Smem::bars + g * kT2PipeBars * 8);
      const uint32_t reg_full = bars + 2 * kT2Slots * 8, reg_free = reg_full + 16;
      const uint64_t a_desc = make_smem_desc(smem_u32(smem + T2Smem::ring + g * kT2Slots * kT2SlotBytes), kTcM * 16, 128);
      // B = W^T (forward, K-major: next k-step = 2 core columns) or W (backward, MN-major: next k-step = 16 rows)
      const uint64_t w1f_h = make_smem_desc(smem_u32(smem + T2Smem::w1_hi), kTcW * 16, 128), w1f_l = make_smem_desc(smem_u32(smem + T2Smem::w1_lo), kTcW * 16, 128);
      const uint64_t w2f_h = make_smem_desc(smem_u32(smem + T2Smem::w2_hi), kTcW * 16, 128), w2f_l = make_smem_desc(smem_u32(smem + T2Smem::w2_lo), kTcW * 16, 128);
      const uint64_t w1b_h = make_smem_desc(smem_u32(smem + T2Smem::w1_hi), 128, kTcW * 16), w1b_l = make_smem_desc(smem_u32(smem + T2Smem::w1_lo), 128, kTcW * 16);
      const uint64_t w2b_h = make_smem_desc(smem_u32(smem + T2Smem::w2_hi), 128, kTcW * 16), w2b_l = make_smem_desc(smem_u32(smem + T2Smem::w2_lo), 128, kTcW * 16);
      const uint32_t idesc_f = make_idesc_bf16(kTcM, kTcW, false), idesc_b = make_idesc_bf16(kTcM, kTcW, true);
      uint32_t r = 0, ph = 0;     // ring position of the next chunk pair and the parity of its slots' use count
      uint32_t free_par = 3;      // bit x: parity to wait for on region x's "free" barrier (a fresh barrier passes parity 1)
      const bool leader = elect_one();
      // one product into accumulator region x: 8 chunks, column halves alternating (the two halves of the epilogue fill
      // their slots in step)
      auto product = [&](uint32_t x, uint64_t bh0, uint64_t bl0, uint32_t b_step, uint32_t idesc, int ksteps) {
        const uint32_t tmem_d = tmem + 256 * g + 128 * x;
        mbar_wait(reg_free + 8 * x, (free_par >> x) & 1);   // the epilogues that read the region's previous content are done
        free_par ^= 1u << x;
#pragma unroll 1   // (unrolled, the 64 chunk descriptors of a step are hoisted into registers the role warps do not have)
        for (int jj = 0; jj < kT2Blocks; ++jj) {
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            const uint32_t slot = ch * (kT2Slots / 2) + r;
            mbar_wait(bars + slot * 8, ph);
            tcgen05_fence_after();
            const int c = kT2Blocks * ch + jj;
            if (leader) {
              if (c < ksteps) {
                const uint64_t ah = a_desc + (uint64_t)(slot * (kT2SlotBytes >> 4));
                const uint64_t bh = bh0 + (uint64_t)(c * b_step), bl = bl0 + (uint64_t)(c * b_step);
                mma_bf16(tmem_d, ah, bh, idesc, (jj | ch) != 0);
                if (LO) {
                  mma_bf16(tmem_d, ah + (kT2SlotHalf >> 4), bh, idesc, true);
                  mma_bf16(tmem_d, ah, bl, idesc, true);
                }
              }
              mma_commit(bars + (kT2Slots + slot) * 8);   // slot free again once these MMAs have read it
            }
          }
          r = (r == kT2Slots / 2 - 1) ? 0 : r + 1;
          if (r == 0) ph ^= 1;
        }
        if (leader) mma_commit(reg_full + 8 * x);
        __syncwarp();
      };
      uint32_t a = 0;   // region of z1 / act'(z1) / grad in this step; z2 / t use the other one; the roles swap every step
      for (int tile = un->t_last; tile >= un->t_first; --tile) {
        const int n_unit_steps = mlp_unit_s1(un, tile, P.n_steps) - mlp_unit_s0(un, tile);
        for (int k = 0; k < n_unit_steps; ++k) {
          product(a, w1f_h, w1f_l, (2 * kTcW * 16) >> 4, idesc_f, k1);       // z1 = x W1^T     (free since E3 of the last step)
          product(a ^ 1, w2f_h, w2f_l, (2 * kTcW * 16) >> 4, idesc_f, k2);   // z2 = h1 W2^T    (free since E4 of the last step)
          product(a ^ 1, w2b_h, w2b_l, 256 >> 4, idesc_b, k3);               // t = delta2 W2   (after E2 has drained z2)
          product(a, w1b_h, w1b_l, 256 >> 4, idesc_b, k2);                   // grad = delta1 W1 (after E3 has drained act')
          a ^= 1;
        }
      }
    }
  } else {
    // ---- epilogue warps: pipeline g, TMEM lane quarter (warp % 4, fixed by the hardware), column half ch -------
    // Padded rows (>= n) and padded columns (>= d, h1, h2) are computed like real ones and never stored: weights,
    // biases and w3 are staged with zeros there, so they cannot leak into a real output, and they stay finite.
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(EBM_T2_EPI_REGS));
    const int g = warp / kT2PipeWarps;
    const int ch = (warp >> 2) & 1;
    const int row = 32 * (warp & 3) + lane;
    const int col_base = kT2Cols * ch;
    const int worker = kT2Pipes * blockIdx.x + g;
    const uint32_t t_reg0 = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + 256 * g + col_base;   // region x at + 128 x
    uint8_t* ring_row = smem + T2Smem::ring + g * kT2Slots * kT2SlotBytes + row * 16;
    const f32x2* b1 = reinterpret_cast<const f32x2*>(smem + T2Smem::b1 + 4 * col_base);
    const f32x2* b2 = reinterpret_cast<const f32x2*>(smem + T2Smem::b2 + 4 * col_base);
    const f32x2* w3 = reinterpret_cast<const f32x2*>(smem + T2Smem::w3 + 4 * col_base);
    const volatile MlpUnits* units = reinterpret_cast<const volatile MlpUnits*>(smem + T2Smem::units) + g;
    const long long numel = P.n * P.d;
    const bool fast_rng = (P.rng.mode == 2) && (P.d % 4 == 0);
    uint32_t r = 0, ph = 0;   // ring position of this column half's next chunk, parity of the slot's use count (the MMA
                              // warp of the pipeline counts the same)
    const uint32_t bars = smem_u32(smem + T2Smem::bars + g * kT2PipeBars * 8);
    const uint32_t reg_full = bars + 2 * kT2Slots * 8, reg_free = reg_full + 16;
    uint32_t full_par = 0;    // bit x: parity to wait for on region x's "full" barrier
    uint32_t za = 0;          // region of z1 / act'(z1) / grad in this step (the MMA warp counts the same)

    // next 16-column block of this thread's row -> the column half's next ring slot, then tell the MMA warp
    auto publish = [&](const f32x2* v) {
      const uint32_t slot = ch * (kT2Slots / 2) + r;
      mbar_wait(bars + (kT2Slots + slot) * 8, ph ^ 1);   // the MMAs that read the slot's previous chunk are done
      uint8_t* dst = ring_row + slot * kT2SlotBytes;
#pragma unroll
      for (int oct = 0; oct < 2; ++oct) {
        uint32_t ph4[4], pl4[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split2(v[oct * 4 + i], ph4[i], pl4[i], LO);
        *reinterpret_cast<uint4*>(dst + oct * (kTcM * 16)) = make_uint4(ph4[0], ph4[1], ph4[2], ph4[3]);
        if (LO) *reinterpret_cast<uint4*>(dst + kT2SlotHalf + oct * (kTcM * 16)) = make_uint4(pl4[0], pl4[1], pl4[2], pl4[3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + slot * 8);
      r = (r == kT2Slots / 2 - 1) ? 0 : r + 1;
      if (r == 0) ph ^= 1;
    };
    // the product in region x is complete
    auto wait_full = [&](uint32_t x) {
      mbar_wait(reg_full + 8 * x, (full_par >> x) & 1);
      full_par ^= 1u << x;
      tcgen05_fence_after();
    };
    // this warp has drained its part of region x (and, with `both`, of the other region): the next product may overwrite it
    auto release = [&](uint32_t x, bool both) {
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(reg_free + 8 * x);
        if (both) mbar_arrive(reg_free + 8 * (x ^ 1));
      }
    };

    // (the block loops of E1-E3 stay rolled: they do not index the register-resident state, and the unrolled step body
    // would be four times the instruction cache)
    for (int tile = units->t_last; tile >= units->t_first; --tile) {
      const long long grow = (long long)tile * kTcM + row;
      const bool rv = grow < P.n;
      const int s0 = mlp_unit_s0(units, tile), s1 = mlp_unit_s1(units, tile, P.n_steps);
      // a unit that starts mid-burst continues the chain another worker left in x_out
      if (s0 > 0) mlp_unit_acquire(P.sched, kT2PipeWarps, worker);
      const float* x0src = (s0 == 0) ? P.x_in : P.x_out;
      const long long row0 = (s0 == 0 && P.row_index && rv) ? P.row_index[grow] : grow;
      f32x2 X[kT2Cols / 2];
      t2_load_row(x0src, row0, P.d, col_base, rv, X);
#pragma unroll
      for (int b = 0; b < kT2Blocks; ++b) publish(X + 8 * b);
      int until_keep = P.thin - ((P.step_base + s0) % P.thin), kept = (P.step_base + s0) / P.thin;
      unsigned long long ctr = P.rng.ctr_base + (unsigned long long)s0 * P.rng.ctr_step;

      for (int k = s0; k < s1; ++k) {
        const int ti = k & tab.mask;
        const float h = tab.h[ti], c12 = tab.c1[ti] * tab.c2[ti];
        const uint32_t t_a = t_reg0 + 128 * za, t_b = t_reg0 + 128 * (za ^ 1);
        // NATIVE stream: the step's noise term c2 c1 eps is added to x in four 16-column instalments, one in front of each
        // wait for a product -- x is not read between its publication and E4, the draw is a quarter of the step's
        // instructions and depends on nothing, so it fills the time the tensor pipe needs to finish the product.
        // x' = (x + c2 c1 eps) - h g instead of (x - h g) + c2 c1 eps (base_integrator.py:728-729): same terms, the rounding
        // order is within this kernel's 2e-5 class.
        auto draw_block = [&](int b) {
          if (!fast_rng) return;
          const long long li0 = grow * P.d + col_base + 16 * b;
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) {
            const uint64_t qi = (uint64_t)(li0 + 4 * qd) >> 2;
            const uint4 w = philox4x32_10((uint32_t)qi, (uint32_t)(qi >> 32), (uint32_t)ctr, (uint32_t)(ctr >> 32), P.keys);
            f32x2 e01, e23;
            normal4_fast_packed(w, e01, e23);
            X[8 * b + 2 * qd] = fma2(e01, c12, X[8 * b + 2 * qd]);
            X[8 * b + 2 * qd + 1] = fma2(e23, c12, X[8 * b + 2 * qd + 1]);
          }
        };
        // E1: z1 -> h1 (A of GEMM2); act'(z1) replaces z1 in its accumulator region
        draw_block(0);
        wait_full(za);
#pragma unroll 1
        for (int b = 0; b < kT2Blocks; ++b) {
          f32x2 v[8], sd[8];
          tmem_ld16p(t_a + 16 * b, v);
#pragma unroll
          for (int i = 0; i < 8; ++i) act2<ACT>(add2(v[i], b1[8 * b + i]), v[i], sd[i]);
          tmem_st16p(t_a + 16 * b, sd);
          publish(v);
        }
        // E2: z2 -> delta2 = w3 * act'(z2) (A of GEMM3)
        draw_block(1);
        wait_full(za ^ 1);
#pragma unroll 1
        for (int b = 0; b < kT2Blocks; ++b) {
          f32x2 v[8];
          tmem_ld16p(t_b + 16 * b, v);
          if (b == kT2Blocks - 1) release(za ^ 1, false);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            f32x2 hh, dh;
            act2<ACT>(add2(v[i], b2[8 * b + i]), hh, dh);
            v[i] = mul2(dh, w3[8 * b + i]);
          }
          publish(v);
        }
        // E3: t -> delta1 = t * act'(z1) (A of GEMM4)
        draw_block(2);
        wait_full(za ^ 1);
        tmem_st_wait();
#pragma unroll 1
        for (int b = 0; b < kT2Blocks; ++b) {
          f32x2 v[8], sd[8];
          tmem_ld16p_nowait(t_b + 16 * b, v);
          tmem_ld16p(t_a + 16 * b, sd);
          if (b == kT2Blocks - 1) release(za, true);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = mul2(v[i], sd[i]);
          publish(v);
        }
        // E4: grad -> Langevin update of x; the new x is the A operand of the next step's GEMM1
        draw_block(3);
        wait_full(za);
        const bool last = (k == s1 - 1);
        bool keep_now = false;
        if (P.traj && --until_keep == 0) { until_keep = P.thin; keep_now = kept < P.n_kept; ++kept; }
#pragma unroll
        for (int b = 0; b < kT2Blocks; ++b) {
          f32x2 gr[8];
          tmem_ld16p(t_a + 16 * b, gr);
          if (b == kT2Blocks - 1) release(za, false);
          const int c0 = col_base + 16 * b;
          const long long li0 = grow * P.d + c0;
          if (fast_rng) {
#pragma unroll
            for (int i = 0; i < 8; ++i) X[8 * b + i] = fma2(gr[i], -h, X[8 * b + i]);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float ev[2];
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const int col = 2 * i + u;
                const bool in = rv && (c0 + col) < P.d;
                ev[u] = 0.0f;
                if (in) ev[u] = (P.rng.mode == 0) ? P.noise[(long long)k * numel + li0 + col]
                                                  : normal_for_element_call(P.rng.k0, P.rng.k1, ctr, P.rng.T, P.rng.mode, (uint64_t)(li0 + col));
              }
              // x' = (x - h g) + c2 c1 eps (base_integrator.py:728-729; fused roundings are within this kernel's 2e-5 class)
              X[8 * b + i] = fma2(pack2(ev[0], ev[1]), c12, fma2(gr[i], -h, X[8 * b + i]));
            }
          }
          if (P.has_clamp) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float xa, xb;
              unpack2(X[8 * b + i], xa, xb);
              X[8 * b + i] = pack2(clamp_torch(xa, P.clamp_lo, P.clamp_hi), clamp_torch(xb, P.clamp_lo, P.clamp_hi));
            }
          }
          if (keep_now) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float xa, xb;
              unpack2(X[8 * b + i], xa, xb);
              float* dst = P.traj + (grow * P.n_kept + (kept - 1)) * P.d + c0 + 2 * i;
              if (rv && (c0 + 2 * i) < P.d) dst[0] = xa;
              if (rv && (c0 + 2 * i + 1) < P.d) dst[1] = xb;
            }
          }
          if (!last) publish(X + 8 * b);
        }
        za ^= 1;
        ctr += P.rng.ctr_step;
      }
      t2_store_row(P.x_out, grow, P.d, col_base, rv, X);
      if (s1 == P.n_steps) {
        if (P.x_out2) t2_store_row(P.x_out2, grow, P.d, col_base, rv, X);
        for (int w = 0; w < P.n_peers; ++w) t2_store_row(P.peers[w] + P.peer_off, grow, P.d, col_base, rv, X, P.peer_mc != 0);
      }
      if (s1 < P.n_steps) mlp_unit_release(P.sched, worker);  // the rest of this tile's burst runs on the next worker
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == kT2EpiWarps) {
    __syncwarp();
    tmem_dealloc(tmem, 512);
  }
}

int langevin_mlp_tc1_dispatch(const LangevinCall& c, int passes);   // ebm_mlp_tc.cu: one tile per SM (A/B reference)

int langevin_mlp_tc_dispatch(const LangevinCall& c, int passes) {
  static const bool single = getenv("EBM_B200_TC_SINGLE") != nullptr;   // kernel tuning only
  if (single) return langevin_mlp_tc1_dispatch(c, passes);
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
  // two workers (pipelines) per CTA and never more workers than tiles: a worker's range then spans at least one whole
  // burst, so a tile is shared by at most two workers
  const long long pairs = (tiles + kT2Pipes - 1) / kT2Pipes;
  const int grid = (int)(pairs < sms ? pairs : sms);
  const int workers = kT2Pipes * grid;
  int* flags = reinterpret_cast<int*>(const_cast<float*>(e->buf[6]));  // NULL: whole tiles per worker (no balancing)
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
      P.n_peers = c.n_peers;
      P.peer_mc = c.peer_mc;
      P.peer_off = c.peer_row_offset * e->dim;
      for (int w = 0; w < c.n_peers; ++w) P.peers[w] = c.peers[w];
    }
    P.n_steps = chunk;
    P.noise = c.noise ? c.noise + (long long)done * numel : nullptr;
    P.rng.ctr_base = c.offset / 4 + (unsigned long long)done * P.rng.ctr_step;
    P.step_base = done;
    if (flags) {   // balanced split (co-residency is taken care of by the cooperative launch, mlp_schedule.cuh)
      int rc0 = mlp_schedule_setup(P.sched, tiles, chunk, workers, flags, c.st);
      if (rc0) return rc0;
    } else {
      mlp_schedule_whole_tiles(P.sched, tiles, chunk, workers);
    }
#define CALL(A)                                                                                               \
  {                                                                                                           \
    auto kern = passes == 3 ? langevin_mlp_tc2_kernel<A, true> : langevin_mlp_tc2_kernel<A, false>;           \
    EBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T2Smem::total));         \
    EBM_CUDA(mlp_launch_persistent(kern, grid, kT2Threads, T2Smem::total, c.st, P, tab, tiles, chunk, workers)); \
  }
    switch (e->activation) {
      case EBM_ACT_SILU: CALL(EBM_ACT_SILU); break;
      case EBM_ACT_TANH: CALL(EBM_ACT_TANH); break;
      case EBM_ACT_RELU: CALL(EBM_ACT_RELU); break;
      default: CALL(EBM_ACT_SOFTPLUS); break;
    }
#undef CALL
    int rc = launch_status("langevin_mlp_tc2_kernel");
    if (rc) return rc;
    done += chunk;
    src = c.x_out;
  }
  return 0;
}

}  // namespace ebm
