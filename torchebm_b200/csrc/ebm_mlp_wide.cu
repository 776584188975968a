// MLP-energy Langevin burst for WIDE inputs (dim > 128, hidden widths <= 128) on tcgen05 + TMEM, sm_100a only:
// the 784-128-128-1 energies of BASELINE configs C3 / C5 (persistent CD on image-sized states).
//
// Same math as ebm_mlp_tc.cu (bf16 hi+lo split operands, three passes per product, fp32 accumulators in tensor
// memory), but neither the chain state (128 x 784 fp32 = 392 KB per tile) nor W1 (784 x 128, 392 KB as bf16 hi+lo)
// fits on chip, so both stream:
//   * W1 (pre-split by mlp_wide_prep_kernel into 64-column chunks, each chunk one contiguous 32 KB blob already in the
//     tensor-core core-matrix layout) and W2 (hi blob, lo blob) travel through a 3-stage shared-memory ring filled by
//     cp.async.bulk (the TMA bulk engine) with mbarrier complete_tx; the ring item order per step is fixed:
//         W2hi, W2lo, W1[0], W1[1], ..., W1[NC-1]
//   * x streams through global memory (L2 resident between steps: 148 tiles x 392 KB < L2): every step each element
//     is read once and written once -- the 8*D bytes per chain-step of the streaming roofline model.
// The trick that keeps W1 traffic at ONE pass per step: the input-gradient product G = delta1 . W1 is produced in
// 64-column chunks, and chunk c of W1 (MN-major view) serves G_c of step k and then, untouched in the same ring stage
// (K-major view), the forward product z1 += x'_c . W1_c^T of step k+1 as soon as the epilogue has turned G_c into the
// updated state chunk x'_c.  Per step and tile:
//     [E1 -> GEMM2 -> E2 -> GEMM3 -> E3]   then for c in chunks: GEMM4_c -> E4_c (update, RNG, store x) -> GEMM1'_c
// Warp roles: warp 0 = TMEM allocation + single-thread MMA issue, warp 1 = bulk-copy producer, warps 2-3 idle (they
// complete the role warpgroup that gives its registers away), warps 4..19 = epilogue
// (thread = one chain row x 32 hidden columns in the 128-wide phases, x 16 state columns per chunk in the E4 phase).
// Every A operand (h1, delta2, delta1, the updated state chunk) lives in TENSOR memory, written by the epilogue with
// tcgen05.st as packed bf16 pairs (row = lane, column j = elements 2j, 2j+1) and read by tcgen05.mma's [a_tmem] form:
// with both operands in shared memory a no-swizzle 128x128x16 product needs 107 cycles (shared-memory operand
// bandwidth), with A in tensor memory 74 (tools/umma_probe.cu), and the epilogue's operand stores, proxy fences and
// the 128 KB of operand buffers disappear from shared memory, which goes to a deeper weight ring instead.
// TMEM columns: [0,128)   z1 -> act'(z1) in place -> z1 of the next step (accumulated chunk by chunk in E4)
//               [128,256) z2 -> G even chunks [128,192), G odd chunks [192,256)
//               [256,384) A operand of the 128-wide products: hi [256,320), lo [320,384)
//               [384,512) t = delta2 . W2 (read by E3) -> state-chunk operand buffers b = 0, 1: hi [384+64b, +32), lo [+32, +64)
#include "mlp_tc_common.cuh"   // packed fp32x2 epilogue arithmetic, bf16 hi/lo split, TMEM <-> register pairs

namespace ebm {

using namespace umma;

constexpr int kWdM = 128;               // chains per tile = TMEM lanes
constexpr int kWdH = 128;               // padded hidden width
constexpr int kWdChunk = 64;            // state columns per W1 ring item
constexpr int kWdStageBytes = 32768;    // ring item: W1 chunk hi (16 KB) + lo (16 KB), or W2 hi, or W2 lo
constexpr int kWdHalf = 16384;
constexpr int kWdStages = 6;
constexpr int kWdEpiWarps = 16;
constexpr int kWdRoleWarps = 4;         // one warpgroup: MMA issue, bulk-copy producer, two idle warps (setmaxnreg works per warpgroup)
constexpr int kWdThreads = 32 * (kWdRoleWarps + kWdEpiWarps);
// Register budget: 640 threads launch with 96 registers each; the role warpgroup shrinks to kWdRoleRegs and the four
// epilogue warpgroups grow to kWdEpiRegs (128 * 32 + 512 * 112 = 640 * 96), which is what keeps the update epilogue
// (48 live values + Philox state) out of local memory.
constexpr int kWdRoleRegs = 32;
constexpr int kWdEpiRegs = 112;
constexpr int kWdMaxDim = 4096;
constexpr int kWdPushBytes = 8192;      // one bulk copy of the gather pusher (two slots per pusher warp)

struct WdSmem {
  static constexpr int ring = 0;
  static constexpr int b1 = ring + kWdStages * kWdStageBytes;
  static constexpr int b2 = b1 + kWdH * 4;
  static constexpr int w3 = b2 + kWdH * 4;
  static constexpr int bars = w3 + kWdH * 4;
  // barrier indices (8 bytes each)
  static constexpr int ring_full = 0;                 // [kWdStages], tx-count
  static constexpr int ring_empty = ring_full + kWdStages;   // [kWdStages], tcgen05.commit
  static constexpr int xa_full = ring_empty + kWdStages;     // [2], 16 epilogue warps
  static constexpr int xa_empty = xa_full + 2;        // [2], tcgen05.commit
  static constexpr int g_full = xa_empty + 2;         // [2], tcgen05.commit
  static constexpr int a_chunk = g_full + 2;          // [8], 4 warps each
  static constexpr int acc_full = a_chunk + 8;        // tcgen05.commit
  static constexpr int n_bars = acc_full + 1;
  static constexpr int tmem_slot = bars + n_bars * 8;
  static constexpr int units = tmem_slot + 16;
  static constexpr int push_cnt = units + 16;         // int: epilogue warps that have finished a tile (gather pusher)
  // gather pusher, bulk form: per pusher warp two staging slots and their "loaded" mbarriers
  static constexpr int push_bar = push_cnt + 16;      // [2 warps][2 slots] x 8 bytes
  static constexpr int push_stage = (push_bar + 32 + 127) & ~127;
  static constexpr int total = push_stage + 2 * 2 * kWdPushBytes;
};
static_assert(WdSmem::total <= 232448, "shared memory budget of one sm_100 CTA exceeded");

// tensor-memory column map (see the header comment)
constexpr uint32_t kWdTmZ1 = 0, kWdTmZ2 = 128, kWdTmG = 128, kWdTmA = 256, kWdTmALo = 320, kWdTmT = 384, kWdTmXa = 384;

struct WdParams {
  const uint8_t* ws;  // workspace: NC W1 chunk blobs, W2 hi blob, W2 lo blob (mlp_wide_prep_kernel)
  const float* b1; const float* b2; const float* w3;
  int d, h1, h2, nc;
  int passes;
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
  MlpSchedule sched;
  // burst-end gather inside the burst kernel (last launch of a burst only): two otherwise idle warps copy every finished
  // tile from x_out to element offset peer_off of the gathered buffers (NVLS multicast address, or every rank's peer
  // mapping), underneath the CTA's next tile.
  int n_peers;
  int peer_mc;   // 1: peers[0] is an NVLS multicast address (n_peers == 1)
  int peer_bulk; // 1: the pusher moves the tile with bulk copies (x_out -> shared memory -> every peer) instead of 16-byte stores
  long long peer_off;
  float* peers[kMaxPeers];
};

// Timeline tracing of CTA 0 (tuning builds only: python -m torchebm_b200.build --variant trace EBM_WD_TRACE; read with
// tools/wd_trace.py).  One record = (clock64 << 16) | tag, written by the single thread that owns a role.
#ifdef EBM_WD_TRACE
constexpr int kWdTraceLen = 8192;
__device__ unsigned long long g_wd_trace[4 * kWdTraceLen];
#define WD_TR_DECL(cond) const bool tr_on = (blockIdx.x == 0) && (cond); int tr_i = 0
#define WD_TR(role, tag)                                                                                            \
  do {                                                                                                              \
    if (tr_on && tr_i < kWdTraceLen) g_wd_trace[(role) * kWdTraceLen + tr_i++] = ((unsigned long long)clock64() << 16) | (unsigned)(tag); \
  } while (0)
#else
#define WD_TR_DECL(cond)
#define WD_TR(role, tag) do {} while (0)
#endif

__device__ __forceinline__ uint32_t wd_bar(uint8_t* smem, int idx) { return smem_u32(smem + WdSmem::bars + idx * 8); }

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(bar), "r"(bytes)
               : "memory");
}
// global -> shared bulk copy on the TMA engine; completion is signalled as `bytes` transaction bytes on `bar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// ---- weight preparation ---------------------------------------------------------------------------------
// One thread = one 16-byte core-matrix row (8 consecutive k of one output row) of the hi and of the lo copy.
__global__ void mlp_wide_prep_kernel(const float* __restrict__ W1, const float* __restrict__ W2, int d, int h1, int h2,
                                     int nc, uint8_t* __restrict__ ws) {
  const int per_chunk = kWdH * (kWdChunk / 8);
  const int n_w1 = nc * per_chunk;
  const int n_w2 = kWdH * (kWdH / 8);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_w1 + n_w2; i += gridDim.x * blockDim.x) {
    float v[8];
    uint8_t* hi_dst;
    uint8_t* lo_dst;
    if (i < n_w1) {
      const int c = i / per_chunk, rem = i - c * per_chunk;
      const int oct = rem / kWdH, r = rem - oct * kWdH;  // consecutive threads -> consecutive rows -> consecutive 16 B
      const int col0 = c * kWdChunk + oct * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (r < h1 && col0 + j < d) ? W1[(long long)r * d + col0 + j] : 0.0f;
      hi_dst = ws + (size_t)c * kWdStageBytes + core_offset(r, oct * 8, kWdH);
      lo_dst = hi_dst + kWdHalf;
    } else {
      const int rem = i - n_w1;
      const int oct = rem / kWdH, r = rem - oct * kWdH;
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (r < h2 && oct * 8 + j < h1) ? W2[r * h1 + oct * 8 + j] : 0.0f;
      hi_dst = ws + (size_t)nc * kWdStageBytes + core_offset(r, oct * 8, kWdH);
      lo_dst = hi_dst + kWdStageBytes;
    }
    uint32_t ph[4], pl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162 h2v = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
      const uint32_t hu = *reinterpret_cast<const uint32_t*>(&h2v);
      ph[j] = hu;
      const __nv_bfloat162 l2v =
          __floats2bfloat162_rn(v[2 * j] - __uint_as_float(hu << 16), v[2 * j + 1] - __uint_as_float(hu & 0xffff0000u));
      pl[j] = *reinterpret_cast<const uint32_t*>(&l2v);
    }
    *reinterpret_cast<uint4*>(hi_dst) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    *reinterpret_cast<uint4*>(lo_dst) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
  }
}

// ---- epilogue helpers -----------------------------------------------------------------------------------
// 4 packed bf16x2 words = 8 consecutive K elements of this thread's row -> 4 tensor-memory columns
__device__ __forceinline__ void tmem_st4_raw(uint32_t taddr, const uint32_t (&w)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]),
               "r"(w[3])
               : "memory");
}
// 16 consecutive operand columns (8 packed pairs) of this thread's row: hi words to t_hi, residual words to t_lo
// (four 4-column stores: splitting all 8 pairs first costs 16 live registers the update phase does not have)
__device__ __forceinline__ void wd_put16p(uint32_t t_hi, uint32_t t_lo, const f32x2* v, bool with_lo) {
#pragma unroll
  for (int o = 0; o < 2; ++o) {
    uint32_t ph[4], pl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split2(v[4 * o + j], ph[j], pl[j], with_lo);
    tmem_st4_raw(t_hi + 4 * o, ph);
    if (with_lo) tmem_st4_raw(t_lo + 4 * o, pl);
  }
}
__device__ __forceinline__ void wd_put16(uint32_t t_hi, uint32_t t_lo, const float* v, bool with_lo) {
  f32x2 p[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) p[j] = pack2(v[2 * j], v[2 * j + 1]);
  wd_put16p(t_hi, t_lo, p, with_lo);
}

// this warp's tensor-memory operand stores have landed -> one arrival per warp for the MMA warp
__device__ __forceinline__ void wd_publish(uint32_t bar) {
  tmem_st_wait();
  tcgen05_fence_before();
  __syncwarp();
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 st;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "@p mbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(bar)
      : "memory");
}

template <int ACT>
__device__ __forceinline__ void wd_act(float z, float& h, float& dh) {
  if (ACT == EBM_ACT_SILU) {
    const float s = rcp_fast(1.0f + __expf(-z));
    h = z * s;
    dh = s * (1.0f + z * (1.0f - s));
  } else if (ACT == EBM_ACT_TANH) {
    const float e = __expf(-2.0f * fabsf(z));
    const float t = copysignf((1.0f - e) * rcp_fast(1.0f + e), z);
    h = t;
    dh = 1.0f - t * t;
  } else if (ACT == EBM_ACT_RELU) {
    h = z > 0.0f ? z : 0.0f;
    dh = z > 0.0f ? 1.0f : 0.0f;
  } else {
    h = z > 20.0f ? z : log1pf(__expf(z));
    dh = rcp_fast(1.0f + __expf(-z));
  }
}

// 256-bit global accesses (sm_100): one request per 32-byte sector, so a thread streaming its 64 contiguous bytes of a
// row never leaves half-used sectors behind for the (tiny, shared-memory-carved) L1 to hold on to
__device__ __forceinline__ void ldg256(const float* p, float* v) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void stg256(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

// 16 state columns of one row from global memory (zeros outside the row / the tile).
// vec: 2 = rows are 32-byte aligned (d % 8 == 0), 1 = 16-byte aligned (d % 4 == 0), 0 = scalar
__device__ __forceinline__ void wd_load_x16(const float* __restrict__ src, long long grow, int col0, int d, bool rv, int vec,
                                            float (&v)[16]) {
  if (rv && vec == 2 && col0 + 16 <= d) {
    const float* p = src + grow * d + col0;
    ldg256(p, v);
    ldg256(p + 8, v + 8);
  } else if (rv && vec == 1 && col0 + 16 <= d) {
    const float4* p = reinterpret_cast<const float4*>(src + grow * d + col0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 t = p[j];
      v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = (rv && col0 + j < d) ? src[grow * d + col0 + j] : 0.0f;
  }
}
__device__ __forceinline__ void wd_store_x16(float* __restrict__ dst, long long row_off, int col0, int d, bool rv, int vec,
                                             const float (&v)[16]) {
  if (!rv) return;
  if (vec == 2 && col0 + 16 <= d) {
    float* p = dst + row_off + col0;
    stg256(p, v);
    stg256(p + 8, v + 8);
  } else if (vec == 1 && col0 + 16 <= d) {
    float4* p = reinterpret_cast<float4*>(dst + row_off + col0);
#pragma unroll
    for (int j = 0; j < 4; ++j) p[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (col0 + j < d) dst[row_off + col0 + j] = v[j];
  }
}

// ---- MMA issue helpers (converged warp, one elected lane issues) -----------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  const uint32_t acc = accumulate ? 1u : 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// `ksteps` k-steps of 16; the A operand advances by 8 tensor-memory columns per k-step, the B descriptor by b_step bytes;
// three passes for split operands
__device__ __forceinline__ void wd_mma_block(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                             uint32_t b_step, uint32_t b_lbo, uint32_t b_sbo, uint32_t idesc, int k_begin,
                                             int k_end, int passes, bool accumulate_first, bool leader) {
  if (!leader) return;   // (the warp runs converged; one elected lane issues, see umma.cuh: elect_one)
  for (int kk = k_begin; kk < k_end; ++kk) {
    const uint64_t bh = make_smem_desc(b_hi + kk * b_step, b_lbo, b_sbo);
    mma_bf16_ts(tmem_d, a_hi + 8 * kk, bh, idesc, accumulate_first || kk > k_begin);
    if (passes == 3) {
      const uint64_t bl = make_smem_desc(b_lo + kk * b_step, b_lbo, b_sbo);
      mma_bf16_ts(tmem_d, a_lo + 8 * kk, bh, idesc, true);
      mma_bf16_ts(tmem_d, a_hi + 8 * kk, bl, idesc, true);
    }
  }
}

// Gather pusher, bulk form (one lane of a pusher warp; out of line so that its address arithmetic does not take part in
// the register allocation of the kernel's hot paths): moves the finished tile [src, src + cnt) in 8 KB pieces,
// x_out -> staging slot (cp.async.bulk, an L2 hit) -> one bulk store per peer.  A 16-byte store makes one NVLink request per
// 16 bytes and lane; a bulk store hands the SM's copy engine 8 KB at once.  Two slots per warp: the load of piece i+1
// runs under the stores of piece i.  Returns the warp's running piece count (slot and mbarrier phase of the next piece).
__device__ __noinline__ long long wd_push_tile_bulk(const WdParams& P, uint8_t* smem, int pw, const float* src, long long e0,
                                                    long long cnt, long long bulk_cnt) {
  const uint32_t stage0 = smem_u32(smem + WdSmem::push_stage + pw * 2 * kWdPushBytes);
  const uint32_t bar0 = smem_u32(smem + WdSmem::push_bar + 16 * pw);
  const char* srcb = reinterpret_cast<const char*>(src);
  const long long bytes = cnt * 4;
  auto piece = [&](long long i) { return (long long)(pw + 2 * i) * kWdPushBytes; };   // the two warps interleave pieces
  auto load = [&](long long i) {
    const long long off = piece(i);
    const uint32_t sz = (uint32_t)((bytes - off < kWdPushBytes) ? (bytes - off) : kWdPushBytes);
    const uint32_t slot = (uint32_t)((bulk_cnt + i) & 1);
    // the slot's previous stores (piece i - 2, committed before piece i - 1 was loaded) must have read it
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    mbar_expect_tx(bar0 + 8 * slot, sz);
    bulk_g2s(stage0 + slot * kWdPushBytes, srcb + off, sz, bar0 + 8 * slot);
  };
  asm volatile("fence.proxy.async;" ::: "memory");   // the epilogue's generic-proxy stores of the tile -> bulk reads
  long long n_pieces = 0;
  while (piece(n_pieces) < bytes) ++n_pieces;
  if (n_pieces > 0) load(0);
  for (long long i = 0; i < n_pieces; ++i) {
    if (i + 1 < n_pieces) load(i + 1);
    const long long off = piece(i);
    const uint32_t sz = (uint32_t)((bytes - off < kWdPushBytes) ? (bytes - off) : kWdPushBytes);
    const uint32_t slot = (uint32_t)((bulk_cnt + i) & 1);
    mbar_wait(bar0 + 8 * slot, (uint32_t)(((bulk_cnt + i) >> 1) & 1));
    for (int w = 0; w < P.n_peers; ++w) {
      char* dst = reinterpret_cast<char*>(P.peers[w] + P.peer_off + e0) + off;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(stage0 + slot * kWdPushBytes),
                   "r"(sz)
                   : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  return bulk_cnt + n_pieces;
}

template <int ACT>
__global__ void __launch_bounds__(kWdThreads, 1) langevin_mlp_wide_kernel(const __grid_constant__ WdParams P,
                                                                          const __grid_constant__ StepTable tab) {
  extern __shared__ __align__(128) uint8_t wd_smem_raw[];
  uint8_t* smem = wd_smem_raw;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  {
    float* b1 = reinterpret_cast<float*>(smem + WdSmem::b1);
    float* b2 = reinterpret_cast<float*>(smem + WdSmem::b2);
    float* w3 = reinterpret_cast<float*>(smem + WdSmem::w3);
    for (int i = threadIdx.x; i < kWdH; i += blockDim.x) {
      b1[i] = i < P.h1 ? P.b1[i] : 0.0f;
      b2[i] = i < P.h2 ? P.b2[i] : 0.0f;
      w3[i] = i < P.h2 ? P.w3[i] : 0.0f;
    }
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kWdStages; ++s) {
      mbar_init(wd_bar(smem, WdSmem::ring_full + s), 1);
      mbar_init(wd_bar(smem, WdSmem::ring_empty + s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(wd_bar(smem, WdSmem::xa_full + b), kWdEpiWarps);
      mbar_init(wd_bar(smem, WdSmem::xa_empty + b), 1);
    }
    mbar_init(wd_bar(smem, WdSmem::g_full + 0), 1);
    mbar_init(wd_bar(smem, WdSmem::g_full + 1), 1);
    for (int c = 0; c < 8; ++c) mbar_init(wd_bar(smem, WdSmem::a_chunk + c), 4);
    mbar_init(wd_bar(smem, WdSmem::acc_full), 1);
    *reinterpret_cast<volatile int*>(smem + WdSmem::push_cnt) = 0;
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(smem + WdSmem::push_bar + 8 * i), 1);
    fence_mbar_init();
  }
  if (threadIdx.x == 32) mlp_units_compute(P.sched, P.n_steps, reinterpret_cast<volatile MlpUnits*>(smem + WdSmem::units));
  if (warp == 0) tmem_alloc(smem_u32(smem + WdSmem::tmem_slot), 512);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + WdSmem::tmem_slot);
  const int NC = P.nc;
  const int K = P.n_steps;
  const int k_h1 = (P.h1 + 15) / 16, k_h2 = (P.h2 + 15) / 16;
  // this CTA's contiguous range of (tile, step) work, walked from its last tile to its first (mlp_schedule.cuh)
  const volatile MlpUnits* units = reinterpret_cast<const volatile MlpUnits*>(smem + WdSmem::units);

  if (warp < kWdRoleWarps) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kWdRoleRegs));
  if (warp == 1) {
    // ---- producer: stream the weight blobs through the ring in the fixed item order --------------------------
    if (lane == 0) {
      WD_TR_DECL(true);
      uint32_t it = 0;
      const uint8_t* w2hi = P.ws + (size_t)NC * kWdStageBytes;
      const uint8_t* w2lo = w2hi + kWdStageBytes;
      auto push = [&](const uint8_t* src) {
        const int s = it % kWdStages;
        mbar_wait(wd_bar(smem, WdSmem::ring_empty + s), ((it / kWdStages) & 1) ^ 1);
        const uint32_t full = wd_bar(smem, WdSmem::ring_full + s);
        WD_TR(3, 256 + (it & 255));
        mbar_expect_tx(full, kWdStageBytes);
        bulk_g2s(smem_u32(smem + WdSmem::ring + s * kWdStageBytes), src, kWdStageBytes, full);
        ++it;
      };
      for (int tile = units->t_last; tile >= units->t_first; --tile) {
        for (int c = 0; c < NC; ++c) push(P.ws + (size_t)c * kWdStageBytes);
        const int n_unit_steps = mlp_unit_s1(units, tile, K) - mlp_unit_s0(units, tile);
        for (int k = 0; k < n_unit_steps; ++k) {
          push(w2hi);
          push(w2lo);
          for (int c = 0; c < NC; ++c) push(P.ws + (size_t)c * kWdStageBytes);
        }
      }
    }
  } else if (warp == 0) {
    // ---- MMA issue ----------------------------------------------------------------------------------------------
    {
      const bool leader = elect_one();
      WD_TR_DECL(leader);
      const uint32_t a_hi = tmem + kWdTmA, a_lo = tmem + kWdTmALo;
      const uint32_t ring = smem_u32(smem + WdSmem::ring);
      const uint32_t idesc_fwd = make_idesc_bf16(kWdM, kWdH, false);
      const uint32_t idesc_bwd = make_idesc_bf16(kWdM, kWdH, true);
      const uint32_t core_col = kWdM * 16;  // bytes between 8-column core blocks (R = 128 rows)
      uint32_t it = 0;                        // ring item counter (consumer side)
      uint32_t xcnt = 0, a_par = 0;         // xcnt: running count of published state chunks (buffer = xcnt & 1)
      auto ring_wait = [&](uint32_t item) { mbar_wait(wd_bar(smem, WdSmem::ring_full + item % kWdStages), (item / kWdStages) & 1); };
      auto ring_release = [&](uint32_t item) { if (leader) mma_commit(wd_bar(smem, WdSmem::ring_empty + item % kWdStages)); };
      auto stage_addr = [&](uint32_t item) { return ring + (item % kWdStages) * kWdStageBytes; };
      // forward product of state chunk c: z1 (+)= xa . W1_c^T
      auto gemm1_chunk = [&](uint32_t item, int c) {
        const int valid = (P.d - c * kWdChunk) < kWdChunk ? (P.d - c * kWdChunk) : kWdChunk;
        const uint32_t w = stage_addr(item);
        const uint32_t xa_hi = tmem + kWdTmXa + 64 * (xcnt & 1);
        wd_mma_block(tmem + kWdTmZ1, xa_hi, xa_hi + 32, w, w + kWdHalf, 2 * core_col, core_col, 128, idesc_fwd, 0,
                     (valid + 15) / 16, P.passes, c > 0, leader);
      };
      auto xa_wait = [&]() {
        mbar_wait(wd_bar(smem, WdSmem::xa_full + (xcnt & 1)), (xcnt >> 1) & 1);
        tcgen05_fence_after();
      };
      auto xa_release = [&]() { if (leader) mma_commit(wd_bar(smem, WdSmem::xa_empty + (xcnt & 1))); ++xcnt; };
      // input-gradient chunk c: G[c & 1] = delta1 . W1[:, chunk c]   (B = MN-major view of the same bytes)
      auto gemm4_chunk = [&](uint32_t item, int c) {
        const int valid = (P.d - c * kWdChunk) < kWdChunk ? (P.d - c * kWdChunk) : kWdChunk;
        const int ncols = ((valid + 15) / 16) * 16;
        const uint32_t w = stage_addr(item);
        wd_mma_block(tmem + kWdTmG + 64 * (c & 1), a_hi, a_lo, w, w + kWdHalf, 256, 128, core_col,
                     make_idesc_bf16(kWdM, ncols, true), 0, k_h1, P.passes, false, leader);
        if (leader) mma_commit(wd_bar(smem, WdSmem::g_full + (c & 1)));
      };
      // a 128-wide product whose A chunks are published one 16-column chunk at a time by the epilogue
      auto gemm_chunked = [&](uint32_t tmem_d, uint32_t w_hi, uint32_t w_lo, bool backward, int ksteps) {
        for (int kk = 0; kk < ksteps; ++kk) {
          mbar_wait(wd_bar(smem, WdSmem::a_chunk + kk), a_par);
          tcgen05_fence_after();
          if (backward)
            wd_mma_block(tmem_d, a_hi, a_lo, w_hi, w_lo, 256, 128, core_col, idesc_bwd, kk, kk + 1, P.passes, kk > 0, leader);
          else
            wd_mma_block(tmem_d, a_hi, a_lo, w_hi, w_lo, 2 * core_col, core_col, 128, idesc_fwd, kk, kk + 1, P.passes, kk > 0,
                         leader);
        }
        // chunks beyond ksteps are still published by the epilogue: consume their phases
        for (int kk = ksteps; kk < 8; ++kk) mbar_wait(wd_bar(smem, WdSmem::a_chunk + kk), a_par);
        a_par ^= 1;
      };

      for (int tile = units->t_last; tile >= units->t_first; --tile) {
        const int n_unit_steps = mlp_unit_s1(units, tile, K) - mlp_unit_s0(units, tile);
        for (int c = 0; c < NC; ++c) {  // prologue: z1 of the unit's initial state
          ring_wait(it);
          xa_wait();
          gemm1_chunk(it, c);
          xa_release();
          ring_release(it);
          ++it;
        }
        if (leader) mma_commit(wd_bar(smem, WdSmem::acc_full));
        for (int k = 0; k < n_unit_steps; ++k) {
          const bool last = (k == n_unit_steps - 1);
          WD_TR(0, 1 * 256);
          ring_wait(it);
          ring_wait(it + 1);
          WD_TR(0, 2 * 256);
          const uint32_t w2hi = stage_addr(it), w2lo = stage_addr(it + 1);
          gemm_chunked(tmem + kWdTmZ2, w2hi, w2lo, false, k_h1);   // z2 = h1 . W2^T
          if (leader) mma_commit(wd_bar(smem, WdSmem::acc_full));
          WD_TR(0, 3 * 256);
          gemm_chunked(tmem + kWdTmT, w2hi, w2lo, true, k_h2);    // t = delta2 . W2
          if (leader) mma_commit(wd_bar(smem, WdSmem::acc_full));
          WD_TR(0, 4 * 256);
          ring_release(it);
          ring_release(it + 1);
          it += 2;
          // delta1 fully published -> chunked input gradient + forward product of the next step
          for (int kk = 0; kk < 8; ++kk) mbar_wait(wd_bar(smem, WdSmem::a_chunk + kk), a_par);
          a_par ^= 1;
          tcgen05_fence_after();
          WD_TR(0, 5 * 256);
          ring_wait(it);
          WD_TR(0, 6 * 256);
          gemm4_chunk(it, 0);
          if (NC > 1) { ring_wait(it + 1); gemm4_chunk(it + 1, 1); }
          for (int c = 0; c < NC; ++c) {
            xa_wait();                                                       // x'_c published, G[c & 1] drained
            WD_TR(0, 7 * 256 + c);
            if (!last) gemm1_chunk(it + c, c);
            xa_release();
            ring_release(it + c);
            WD_TR(0, 9 * 256 + c);
            if (c + 2 < NC) { ring_wait(it + c + 2); WD_TR(0, 8 * 256 + c); gemm4_chunk(it + c + 2, c + 2); WD_TR(0, 10 * 256 + c); }
          }
          it += NC;
          if (!last) if (leader) mma_commit(wd_bar(smem, WdSmem::acc_full));
        }
      }
      // the last commits must have landed in shared memory before the CTA may retire
      if (it > 0) mbar_wait(wd_bar(smem, WdSmem::ring_empty + (it - 1) % kWdStages), ((it - 1) / kWdStages) & 1);
    }
  } else if (P.n_peers > 0) {
    // ---- burst-end gather: warps 2 and 3 push every finished tile into the gathered buffers ----------------------
    // The epilogue's own store pattern (thread = row, 32 or 64 bytes per lane) is fine for the local L2 but makes one
    // NVLink packet per sector; these two warps re-read the tile from x_out (an L2 hit, contiguous [rows x d] floats) and
    // store it fully coalesced -- 512 contiguous bytes per warp instruction -- to the NVLS multicast address (one store
    // reaches every rank) or to each peer.  The push of a tile runs underneath the CTA's next tile; only the last one
    // is exposed.  (measured at 8 GPUs, 65 536 x 784 per GPU: stores from the epilogue 7.6 ms per step with peer stores,
    // 15.8 ms through multicast)
    const int pw = warp - 2;
    const int* cnt_p = reinterpret_cast<const int*>(smem + WdSmem::push_cnt);
    long long bulk_cnt = 0;   // pieces this warp has moved so far (slot and mbarrier phase of the next one)
    int want = 0;   // a counter, not an mbarrier phase: the pusher may fall more than one tile behind
    for (int tile = units->t_last; tile >= units->t_first; --tile) {
      if (mlp_unit_s1(units, tile, K) != K) continue;   // the burst of this tile ends on another CTA
      want += kWdEpiWarps;
      if (lane == 0) {
        int v;
        do {
          asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(cnt_p)) : "memory");
          if (v < want) __nanosleep(500);
        } while (v < want);
      }
      __syncwarp();
      const long long r0 = (long long)tile * kWdM;
      const long long r1 = (r0 + kWdM < P.n) ? r0 + kWdM : P.n;
      const long long e0 = r0 * P.d, cnt = (r1 - r0) * P.d;          // contiguous element range of the tile
      const float* src = P.x_out + e0;
      bool vec4 = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && (cnt % 4 == 0);
      for (int w = 0; w < P.n_peers; ++w) vec4 = vec4 && ((reinterpret_cast<uintptr_t>(P.peers[w] + P.peer_off + e0) & 15) == 0);
      if (vec4 && P.peer_bulk && !P.peer_mc) {
        if (lane == 0) bulk_cnt = wd_push_tile_bulk(P, smem, pw, src, e0, cnt, bulk_cnt);
      } else if (vec4) {
        const long long nv = cnt / 4;
        const float4* s4 = reinterpret_cast<const float4*>(src);
        for (long long i0 = 32 * pw; i0 < nv; i0 += 2 * 64) {       // 2 independent 16-byte loads in flight per lane
          float4 v[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const long long i = i0 + 64 * u + lane;
            if (i < nv) v[u] = __ldcg(s4 + i);
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const long long i = i0 + 64 * u + lane;
            if (i >= nv) continue;
            if (P.peer_mc) {
              mc_store4(P.peers[0] + P.peer_off + e0 + 4 * i, v[u].x, v[u].y, v[u].z, v[u].w);
            } else {
              for (int w = 0; w < P.n_peers; ++w) reinterpret_cast<float4*>(P.peers[w] + P.peer_off + e0)[i] = v[u];
            }
          }
        }
      } else {
        for (long long i = 32 * pw + lane; i < cnt; i += 64) {
          const float v = __ldcg(src + i);
          if (P.peer_mc) mc_store1(P.peers[0] + P.peer_off + e0 + i, v);
          else for (int w = 0; w < P.n_peers; ++w) P.peers[w][P.peer_off + e0 + i] = v;
        }
      }
    }
    if (P.peer_bulk && lane == 0) {   // every bulk store has completed before the CTA (and its staging slots) retire
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      __threadfence_system();
    }
  }
  } else {
    // ---- epilogue warps -------------------------------------------------------------------------------------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kWdEpiRegs));
    const int e = warp - kWdRoleWarps;
    const int q = warp & 3;                         // TMEM lane quarter this warp may access (hardware: warp % 4)
    const int cg = e >> 2;                          // column group 0..3
    const int row = 32 * q + lane;
    const int hcol = 32 * cg;                       // hidden columns [hcol, hcol + 32) in the 128-wide phases
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * q) << 16);
    const f32x2* b1 = reinterpret_cast<const f32x2*>(reinterpret_cast<const float*>(smem + WdSmem::b1) + hcol);
    const f32x2* b2 = reinterpret_cast<const f32x2*>(reinterpret_cast<const float*>(smem + WdSmem::b2) + hcol);
    const f32x2* w3 = reinterpret_cast<const f32x2*>(reinterpret_cast<const float*>(smem + WdSmem::w3) + hcol);
    const uint32_t a_hi = lane_addr + kWdTmA + hcol / 2, a_lo = lane_addr + kWdTmALo + hcol / 2;   // this thread's operand columns
    const uint32_t acc_bar = wd_bar(smem, WdSmem::acc_full);
    const uint32_t xa_full = wd_bar(smem, WdSmem::xa_full), xa_empty = wd_bar(smem, WdSmem::xa_empty);
    uint32_t xcnt = 0;  // running count of published state chunks: buffer xcnt & 1, use (xcnt >> 1) of that buffer
    const bool with_lo = P.passes == 3;
    // widest aligned access every pointer of this launch allows
    const uintptr_t ptr_bits = (uintptr_t)P.x_in | (uintptr_t)P.x_out | (uintptr_t)P.traj | (uintptr_t)P.x_out2;
    const int vec = (P.d % 8 == 0 && (ptr_bits & 31) == 0) ? 2 : ((P.d % 4 == 0 && (ptr_bits & 15) == 0) ? 1 : 0);
    uint32_t acc_par = 0, g_par = 0;  // g_par: bit b = parity of g_full[b]
    WD_TR_DECL((e == 0 || e == 4) && lane == 0);   // one warp of column groups 0 and 1
    const int tr_role = 1 + (e >> 2);
    (void)tr_role;

    for (int tile = units->t_last; tile >= units->t_first; --tile) {
      const long long grow = (long long)tile * kWdM + row;
      const bool rv = grow < P.n;
      const int s0 = mlp_unit_s0(units, tile), s1 = mlp_unit_s1(units, tile, K);
      // a unit that starts mid-burst continues the chain another CTA left in x_out
      if (s0 > 0) mlp_unit_acquire(P.sched, kWdEpiWarps);
      const float* x0src = (s0 == 0) ? P.x_in : P.x_out;
      // row of the launch's input that holds this chain (replay-buffer gather fused into the first load)
      const long long row0 = (s0 == 0 && P.row_index && rv) ? P.row_index[grow] : grow;
      // prologue: publish the unit's initial state chunk by chunk
      for (int c = 0; c < NC; ++c) {
        const int col0 = c * kWdChunk + 16 * cg;
        float v[16];
        wd_load_x16(x0src, row0, col0, P.d, rv, vec, v);
        const uint32_t xb = xcnt & 1;
        mbar_wait(xa_empty + 8 * xb, ((xcnt >> 1) & 1) ^ 1);
        ++xcnt;
        const uint32_t xa_hi = lane_addr + kWdTmXa + 64 * xb + 8 * cg;
        if (col0 < P.d) wd_put16(xa_hi, xa_hi + 32, v, with_lo);
        wd_publish(xa_full + 8 * xb);
      }
      int until_keep = P.thin - ((P.step_base + s0) % P.thin), kept = (P.step_base + s0) / P.thin;
      // only the stepping counter lives in registers; everything else of the stream is read from the parameters
      unsigned long long ctr_base = P.rng.ctr_base + (unsigned long long)s0 * P.rng.ctr_step;

      for (int k = s0; k < s1; ++k) {
        const int ti = k & tab.mask;
        const float h = tab.h[ti], c1 = tab.c1[ti], c2 = tab.c2[ti];
        const float* xsrc = (k == 0) ? P.x_in : P.x_out;
        const long long xrow = (k == 0 && P.row_index && rv) ? P.row_index[grow] : grow;
        const bool final_x2 = P.x_out2 && (k == K - 1);
        // E1: z1 -> h1 (A of GEMM2); act'(z1) -> TMEM [256, 384)
        WD_TR(tr_role, 1 * 256);
        mbar_wait(acc_bar, acc_par); acc_par ^= 1;
        tcgen05_fence_after();
        WD_TR(tr_role, 2 * 256);
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          f32x2 v[8], sd[8];
          tmem_ld16p(lane_addr + kWdTmZ1 + hcol + 16 * blk, v);
#pragma unroll
          for (int i = 0; i < 8; ++i) act2<ACT>(add2(v[i], b1[8 * blk + i]), v[i], sd[i]);
          tmem_st16p(lane_addr + kWdTmZ1 + hcol + 16 * blk, sd);   // act'(z1) replaces z1 (each thread rewrites what it read)
          wd_put16p(a_hi + 8 * blk, a_lo + 8 * blk, v, with_lo);
          wd_publish(wd_bar(smem, WdSmem::a_chunk + 2 * cg + blk));
        }
        // E2: z2 -> delta2 = w3 * act'(z2) (A of GEMM3)
        WD_TR(tr_role, 3 * 256);
        mbar_wait(acc_bar, acc_par); acc_par ^= 1;
        tcgen05_fence_after();
        WD_TR(tr_role, 4 * 256);
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          f32x2 v[8];
          tmem_ld16p(lane_addr + kWdTmZ2 + hcol + 16 * blk, v);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            f32x2 hh, dh;
            act2<ACT>(add2(v[i], b2[8 * blk + i]), hh, dh);
            v[i] = mul2(dh, w3[8 * blk + i]);
          }
          wd_put16p(a_hi + 8 * blk, a_lo + 8 * blk, v, with_lo);
          wd_publish(wd_bar(smem, WdSmem::a_chunk + 2 * cg + blk));
        }
        // E3: t -> delta1 = t * act'(z1) (A of every GEMM4 chunk)
        WD_TR(tr_role, 5 * 256);
        mbar_wait(acc_bar, acc_par); acc_par ^= 1;
        tcgen05_fence_after();
        WD_TR(tr_role, 6 * 256);
        tmem_st_wait();
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          f32x2 v[8], sd[8];
          tmem_ld16p_nowait(lane_addr + kWdTmT + hcol + 16 * blk, v);
          tmem_ld16p(lane_addr + kWdTmZ1 + hcol + 16 * blk, sd);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = mul2(v[i], sd[i]);
          wd_put16p(a_hi + 8 * blk, a_lo + 8 * blk, v, with_lo);
          wd_publish(wd_bar(smem, WdSmem::a_chunk + 2 * cg + blk));
        }
        // E4: per 64-column chunk: G_c -> Langevin update of the state chunk -> global + A operand of GEMM1'_c
        bool keep_now = false;
        if (P.traj && --until_keep == 0) { until_keep = P.thin; keep_now = kept < P.n_kept; ++kept; }
        for (int c = 0; c < NC; ++c) {
          const int col0 = c * kWdChunk + 16 * cg;
          const bool active = col0 < P.d;            // warp-uniform
          // the state chunk is requested first (an L2 hit from the previous step's store); the noise draw below
          // covers its latency
          float xc[16];
          WD_TR(tr_role, 7 * 256 + c);
          if (active) wd_load_x16(xsrc, xrow, col0, P.d, rv, vec, xc);
          f32x2 eps[8];
          if (active) {
            const long long li0 = grow * P.d + col0;
            if (P.rng.mode == 2 && P.d % 4 == 0 && col0 + 16 <= P.d) {
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) {
                const uint64_t qi = (uint64_t)(li0 + 4 * q4) >> 2;
                const uint4 w = philox4x32_10((uint32_t)qi, (uint32_t)(qi >> 32), (uint32_t)ctr_base,
                                              (uint32_t)(ctr_base >> 32), P.keys);
                normal4_fast_packed(w, eps[2 * q4], eps[2 * q4 + 1]);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float ev[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  const int col = 2 * i + u;
                  const bool in = rv && (col0 + col) < P.d;
                  ev[u] = 0.0f;
                  if (in) ev[u] = (P.rng.mode == 0) ? P.noise[(long long)k * (P.n * P.d) + li0 + col]
                                                    : normal_for_element_call(P.rng.k0, P.rng.k1, ctr_base, P.rng.T, P.rng.mode, (uint64_t)(li0 + col));
                }
                eps[i] = pack2(ev[0], ev[1]);
              }
            }
          }
          WD_TR(tr_role, 8 * 256 + c);
          mbar_wait(wd_bar(smem, WdSmem::g_full + (c & 1)), (g_par >> (c & 1)) & 1);
          g_par ^= 1u << (c & 1);
          tcgen05_fence_after();
          WD_TR(tr_role, 9 * 256 + c);
          f32x2 X[8];
          if (active) {
            f32x2 g[8];
            tmem_ld16p(lane_addr + kWdTmG + 64 * (c & 1) + 16 * cg, g);
            const float c12 = c1 * c2;
            // x' = (x - h g) + c2 c1 eps (base_integrator.py:728-729) on packed pairs; the fused roundings are within this
            // kernel's 2e-5 class
#pragma unroll
            for (int i = 0; i < 8; ++i) X[i] = fma2(eps[i], c12, fma2(g[i], -h, pack2(xc[2 * i], xc[2 * i + 1])));
            if (P.has_clamp) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float xa_, xb_;
                unpack2(X[i], xa_, xb_);
                X[i] = pack2(clamp_torch(xa_, P.clamp_lo, P.clamp_hi), clamp_torch(xb_, P.clamp_lo, P.clamp_hi));
              }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) unpack2(X[i], xc[2 * i], xc[2 * i + 1]);
            if (col0 + 16 > P.d) {  // ragged last block: columns beyond the state stay exactly zero
#pragma unroll
              for (int i = 0; i < 16; ++i) xc[i] = (col0 + i) < P.d ? xc[i] : 0.0f;
#pragma unroll
              for (int i = 0; i < 8; ++i) X[i] = pack2(xc[2 * i], xc[2 * i + 1]);
            }
          }
          // publish the new chunk to the tensor core first (GEMM1' of the next step is on the critical path), then
          // let the global stores drain underneath the next chunk's noise draw
          const uint32_t xb = xcnt & 1;
          WD_TR(tr_role, 10 * 256 + c);
          mbar_wait(xa_empty + 8 * xb, ((xcnt >> 1) & 1) ^ 1);
          ++xcnt;
          WD_TR(tr_role, 11 * 256 + c);
          if (active) {
            const uint32_t xa_hi = lane_addr + kWdTmXa + 64 * xb + 8 * cg;
            wd_put16p(xa_hi, xa_hi + 32, X, with_lo);
          }
          WD_TR(tr_role, 12 * 256 + c);
          wd_publish(xa_full + 8 * xb);
          WD_TR(tr_role, 13 * 256 + c);
          if (active) {
            wd_store_x16(P.x_out, grow * P.d, col0, P.d, rv, vec, xc);
            if (final_x2) wd_store_x16(P.x_out2, grow * P.d, col0, P.d, rv, vec, xc);
            if (keep_now) wd_store_x16(P.traj, (grow * P.n_kept + (kept - 1)) * P.d, col0, P.d, rv, vec, xc);
          }
        }
        ctr_base += P.rng.ctr_step;
      }
      if (P.n_peers > 0 && s1 == K) {   // the tile's final state is in x_out: hand it to the pusher warps
        __syncwarp();
        if (lane == 0) {
          __threadfence_block();
          asm volatile("red.release.cta.shared.add.s32 [%0], 1;" ::"r"(smem_u32(smem + WdSmem::push_cnt)) : "memory");
        }
      }
      if (s1 < K) mlp_unit_release(P.sched);  // the rest of this tile's burst runs on the next CTA
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem, 512);
  }
}

// ---- E(x) and grad E(x) for wide MLP energies (diagnostics / ebm_energy_f32 / ebm_gradient_f32) --------------------
// Plain fp32 FFMA utility kernel, not on the burst path: one warp owns kWdURows rows at a time, the rows are staged in
// shared memory, lane l owns hidden units l, l+32, l+64, l+96 in the forward products and state columns l + 32 m in the
// input-gradient product (coalesced W1 reads).  exact expf/tanhf activations as in ebm_mlp.cu.
constexpr int kWdURows = 4;
constexpr int kWdUWarps = 4;

template <int ACT>
__device__ __forceinline__ void wd_act_exact(float z, float& h, float& dh) {
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

struct WdUtilParams {
  const float* W1; const float* b1; const float* W2; const float* b2; const float* w3; const float* b3;
  int d, h1, h2, dpad;
  const float* x;
  float* energy;
  float* grad;
  long long n;
};

template <int ACT>
__global__ void __launch_bounds__(32 * kWdUWarps) mlp_wide_energy_grad_kernel(const WdUtilParams P) {
  extern __shared__ __align__(16) float wdu_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* xs = wdu_smem + (size_t)warp * (kWdURows * (P.dpad + 2 * kWdH));
  float* hs = xs + kWdURows * P.dpad;
  float* ds = hs + kWdURows * kWdH;
  const long long groups = (P.n + kWdURows - 1) / kWdURows;
  for (long long grp = (long long)blockIdx.x * kWdUWarps + warp; grp < groups; grp += (long long)gridDim.x * kWdUWarps) {
    const long long row0 = grp * kWdURows;
    for (int r = 0; r < kWdURows; ++r)
      for (int k = lane; k < P.dpad; k += 32)
        xs[r * P.dpad + k] = (row0 + r < P.n && k < P.d) ? P.x[(row0 + r) * P.d + k] : 0.0f;
    __syncwarp();
    float acc[kWdURows][4], dh1[kWdURows][4];
    // layer 1
#pragma unroll
    for (int r = 0; r < kWdURows; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.0f;
    for (int k = 0; k < P.d; ++k) {
      float w[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) { const int j = lane + 32 * v; w[v] = j < P.h1 ? P.W1[(long long)j * P.d + k] : 0.0f; }
#pragma unroll
      for (int r = 0; r < kWdURows; ++r) {
        const float a = xs[r * P.dpad + k];
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[r][v] = fmaf(a, w[v], acc[r][v]);
      }
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int j = lane + 32 * v;
      const float b = j < P.h1 ? P.b1[j] : 0.0f;
#pragma unroll
      for (int r = 0; r < kWdURows; ++r) {
        float h, dh;
        wd_act_exact<ACT>(acc[r][v] + b, h, dh);
        hs[r * kWdH + j] = j < P.h1 ? h : 0.0f;
        dh1[r][v] = j < P.h1 ? dh : 0.0f;
      }
    }
    __syncwarp();
    // layer 2 + energy
#pragma unroll
    for (int r = 0; r < kWdURows; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.0f;
    for (int i = 0; i < P.h1; ++i) {
      float w[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) { const int j = lane + 32 * v; w[v] = j < P.h2 ? P.W2[j * P.h1 + i] : 0.0f; }
#pragma unroll
      for (int r = 0; r < kWdURows; ++r) {
        const float a = hs[r * kWdH + i];
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[r][v] = fmaf(a, w[v], acc[r][v]);
      }
    }
    float en[kWdURows];
#pragma unroll
    for (int r = 0; r < kWdURows; ++r) en[r] = 0.0f;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int j = lane + 32 * v;
      const float b = j < P.h2 ? P.b2[j] : 0.0f, w3 = j < P.h2 ? P.w3[j] : 0.0f;
#pragma unroll
      for (int r = 0; r < kWdURows; ++r) {
        float h, dh;
        wd_act_exact<ACT>(acc[r][v] + b, h, dh);
        en[r] = fmaf(w3, h, en[r]);
        ds[r * kWdH + j] = w3 * dh;
      }
    }
    if (P.energy) {
#pragma unroll
      for (int r = 0; r < kWdURows; ++r) {
        float e = en[r];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
        if (lane == 0 && row0 + r < P.n) P.energy[row0 + r] = e + P.b3[0];
      }
    }
    if (P.grad) {
      __syncwarp();
      // t = W2^T delta2 ; delta1 = t * act'(z1)
#pragma unroll
      for (int r = 0; r < kWdURows; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.0f;
      for (int j = 0; j < P.h2; ++j) {
        float w[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) { const int i = lane + 32 * v; w[v] = i < P.h1 ? P.W2[j * P.h1 + i] : 0.0f; }
#pragma unroll
        for (int r = 0; r < kWdURows; ++r) {
          const float a = ds[r * kWdH + j];
#pragma unroll
          for (int v = 0; v < 4; ++v) acc[r][v] = fmaf(a, w[v], acc[r][v]);
        }
      }
      __syncwarp();
#pragma unroll
      for (int v = 0; v < 4; ++v)
#pragma unroll
        for (int r = 0; r < kWdURows; ++r) hs[r * kWdH + lane + 32 * v] = acc[r][v] * dh1[r][v];
      __syncwarp();
      // g = W1^T delta1, 128 state columns per pass
      for (int kb = 0; kb < P.d; kb += 128) {
#pragma unroll
        for (int r = 0; r < kWdURows; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.0f;
        for (int j = 0; j < P.h1; ++j) {
          float w[4];
#pragma unroll
          for (int v = 0; v < 4; ++v) { const int k = kb + lane + 32 * v; w[v] = k < P.d ? P.W1[(long long)j * P.d + k] : 0.0f; }
#pragma unroll
          for (int r = 0; r < kWdURows; ++r) {
            const float a = hs[r * kWdH + j];
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[r][v] = fmaf(a, w[v], acc[r][v]);
          }
        }
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int k = kb + lane + 32 * v;
#pragma unroll
          for (int r = 0; r < kWdURows; ++r)
            if (k < P.d && row0 + r < P.n) P.grad[(row0 + r) * P.d + k] = acc[r][v];
        }
      }
    }
    __syncwarp();
  }
}

bool mlp_wide_supported(const EbmEnergyDesc* e);

int mlp_wide_energy_grad_dispatch(const EbmEnergyDesc* e, const float* x, int64_t n, float* energy, float* grad,
                                  cudaStream_t st) {
  if (!mlp_wide_supported(e)) {
    set_error("MLP energy %d->%d->%d->1 is not supported (dim <= %d, hidden <= %d)", e->dim, e->hidden1, e->hidden2,
              kWdMaxDim, kWdH);
    return EBM_ERR_UNSUPPORTED;
  }
  const DeviceInfo& di = device_info(current_device());
  WdUtilParams P;
  P.W1 = e->buf[0]; P.b1 = e->buf[1]; P.W2 = e->buf[2]; P.b2 = e->buf[3]; P.w3 = e->buf[4]; P.b3 = e->buf[5];
  P.d = e->dim; P.h1 = e->hidden1; P.h2 = e->hidden2;
  P.dpad = (e->dim + 3) & ~3;
  P.x = x; P.energy = energy; P.grad = grad; P.n = n;
  const size_t smem = (size_t)kWdUWarps * kWdURows * (P.dpad + 2 * kWdH) * sizeof(float);
  long long groups = (n + kWdURows - 1) / kWdURows;
  long long ctas = (groups + kWdUWarps - 1) / kWdUWarps;
  const long long cap = (long long)di.sm_count * 2;
  const int grid = (int)(ctas < cap ? ctas : cap);
#define CALL(A)                                                                                          \
  {                                                                                                      \
    auto kern = mlp_wide_energy_grad_kernel<A>;                                                          \
    EBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
    kern<<<grid, 32 * kWdUWarps, smem, st>>>(P);                                                         \
  }
  switch (e->activation) {
    case EBM_ACT_SILU: CALL(EBM_ACT_SILU); break;
    case EBM_ACT_TANH: CALL(EBM_ACT_TANH); break;
    case EBM_ACT_RELU: CALL(EBM_ACT_RELU); break;
    default: CALL(EBM_ACT_SOFTPLUS); break;
  }
#undef CALL
  return launch_status("mlp_wide_energy_grad_kernel");
}

static size_t mlp_wide_weight_bytes(const EbmEnergyDesc* e) {
  const int nc = (e->dim + kWdChunk - 1) / kWdChunk;
  return (size_t)(nc + 2) * kWdStageBytes;
}
size_t mlp_wide_workspace_bytes(const EbmEnergyDesc* e) { return mlp_wide_weight_bytes(e) + kMlpFlagBytes; }


bool mlp_wide_supported(const EbmEnergyDesc* e) {
  return e->dim <= kWdMaxDim && e->hidden1 <= kWdH && e->hidden2 <= kWdH;
}

int langevin_mlp_wide_dispatch(const LangevinCall& c, int passes) {
  const EbmEnergyDesc* e = c.e;
  if (!mlp_wide_supported(e)) {
    set_error("wide MLP kernel supports dim <= %d and hidden widths <= %d (got %d-%d-%d)", kWdMaxDim, kWdH, e->dim,
              e->hidden1, e->hidden2);
    return EBM_ERR_UNSUPPORTED;
  }
  if (!e->buf[6]) {
    set_error("MLP energies with dim > 128 need a device workspace of ebm_workspace_bytes() bytes in buf[6]");
    return EBM_ERR_INVALID;
  }
  if (((uintptr_t)e->buf[6] & 127) != 0) { set_error("MLP workspace must be 128-byte aligned"); return EBM_ERR_INVALID; }
  const DeviceInfo& di = device_info(current_device());
  const long long numel = (long long)c.n * e->dim;
  WdParams P;
  memset(&P, 0, sizeof(P));
  P.ws = reinterpret_cast<const uint8_t*>(e->buf[6]);
  P.b1 = e->buf[1]; P.b2 = e->buf[3]; P.w3 = e->buf[4];
  P.d = e->dim; P.h1 = e->hidden1; P.h2 = e->hidden2;
  P.nc = (e->dim + kWdChunk - 1) / kWdChunk;
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
  // the weights may have changed since the last burst (training loop): re-split them on the caller's stream
  {
    const int items = (P.nc * kWdH * (kWdChunk / 8)) + kWdH * (kWdH / 8);
    mlp_wide_prep_kernel<<<(items + 255) / 256, 256, 0, c.st>>>(e->buf[0], e->buf[2], P.d, P.h1, P.h2, P.nc,
                                                                 reinterpret_cast<uint8_t*>(const_cast<float*>(e->buf[6])));
    int rc = launch_status("mlp_wide_prep_kernel");
    if (rc) return rc;
  }
  const long long tiles = (c.n + kWdM - 1) / kWdM;
  const int sms = (e->sm_margin > 0 && e->sm_margin < di.sm_count) ? di.sm_count - e->sm_margin : di.sm_count;
  const int grid = (int)(tiles < sms ? tiles : sms);
  int* flags = reinterpret_cast<int*>(const_cast<float*>(e->buf[6])) + mlp_wide_weight_bytes(e) / 4;
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
      P.peer_bulk = c.peer_mc ? 0 : wide_push_bulk();
      P.peer_off = c.peer_row_offset * e->dim;
      for (int w = 0; w < c.n_peers; ++w) P.peers[w] = c.peers[w];
    }
    P.n_steps = chunk;
    P.noise = c.noise ? c.noise + (long long)done * numel : nullptr;
    P.rng.ctr_base = c.offset / 4 + (unsigned long long)done * P.rng.ctr_step;
    P.step_base = done;
    int rc0 = mlp_schedule_setup(P.sched, tiles, chunk, grid, flags, c.st);
    if (rc0) return rc0;
    // the pusher's staging slots are the tail of the layout: only launches with a bulk gather ask for them
    const int smem_bytes = (P.n_peers > 0 && P.peer_bulk) ? WdSmem::total : WdSmem::push_stage;
#define CALL(A)                                                                                            \
  {                                                                                                        \
    auto kern = langevin_mlp_wide_kernel<A>;                                                               \
    EBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WdSmem::total));      \
    EBM_CUDA(mlp_launch_persistent(kern, grid, kWdThreads, smem_bytes, c.st, P, tab, tiles, chunk, grid)); \
  }
    switch (e->activation) {
      case EBM_ACT_SILU: CALL(EBM_ACT_SILU); break;
      case EBM_ACT_TANH: CALL(EBM_ACT_TANH); break;
      case EBM_ACT_RELU: CALL(EBM_ACT_RELU); break;
      default: CALL(EBM_ACT_SOFTPLUS); break;
    }
#undef CALL
    int rc = launch_status("langevin_mlp_wide_kernel");
    if (rc) return rc;
    done += chunk;
    src = c.x_out;
  }
  return 0;
}

}  // namespace ebm

#ifdef EBM_WD_TRACE
extern "C" int ebm_debug_wd_trace(unsigned long long* out, int n) {
  const size_t bytes = sizeof(unsigned long long) * (size_t)(n < 4 * ebm::kWdTraceLen ? n : 4 * ebm::kWdTraceLen);
  return (int)cudaMemcpyFromSymbol(out, ebm::g_wd_trace, bytes);
}
#endif
