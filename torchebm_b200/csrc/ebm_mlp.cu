// MLP-energy path, fp32 FFMA version (the parity anchor for the tensor-core version).
//
// Energy: E(x) = w3 . act(W2 act(W1 x + b1) + b2) + b3   (user Sequential energies,
// examples/20-training/01-mcmc-losses/01-cd-k/main.py:20-30); gradient
//   dE/dx = W1^T (act'(z1) * (W2^T (act'(z2) * w3)))       (what autograd of base_model.py:84-127 yields).
//
// One CTA = 8 warps = a tile of 64 chains; each warp owns 8 chains for the whole K-step burst.
// W1 and W2 are staged once per launch into shared memory in their natural [out, in] layout with a
// 16-byte-chunk XOR swizzle (chunk ^= (row / 4) & 7), so the same copy serves the forward products
// (lane owns 4 output rows, vector along `in`) and the backward products (lane owns 4 `in` columns,
// row fixed) without bank conflicts.  Activations travel between layers through two warp-private
// shared buffers; chain state x, act'(z1) and the accumulators stay in registers.  Nothing but the
// optional trajectory touches HBM between steps.
#include "api_common.cuh"

namespace ebm {

constexpr int kMlpMax = 128;   // padded width of every layer (D, H1, H2 <= 128 in this version)
constexpr int kMlpRows = 8;    // chains per warp
constexpr int kMlpWarps = 8;
constexpr int kMlpTile = kMlpRows * kMlpWarps;

struct MlpParams {
  const float* W1; const float* b1; const float* W2; const float* b2; const float* w3; const float* b3;
  int d, h1, h2;
  const float* x_in;
  float* x_out;
  const float* noise;
  float* traj;
  float* energy;  // energy/grad kernel only
  float* grad;
  long long n;
  int n_steps, thin, n_kept, thin_start, kept_base, has_clamp;
  float clamp_lo, clamp_hi;
  RowRng rng;
};

template <int ACT>
__device__ __forceinline__ void act_fwd(float z, float& h, float& dh) {
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
  } else {  // softplus, torch threshold 20
    h = z > 20.0f ? z : log1pf(expf(z));
    dh = 1.0f / (1.0f + expf(-z));
  }
}

__device__ __forceinline__ int swz(int row, int col) {  // float index inside a [128][128] swizzled matrix
  return row * kMlpMax + ((((col >> 2) ^ ((row >> 2) & 7)) << 2) | (col & 3));
}

__device__ void stage_matrix(float* dst, const float* __restrict__ src, int rows, int cols) {
  for (int i = threadIdx.x; i < kMlpMax * kMlpMax; i += blockDim.x) {
    const int r = i / kMlpMax, c = i - r * kMlpMax;
    dst[swz(r, c)] = (r < rows && c < cols) ? src[r * cols + c] : 0.0f;
  }
}
__device__ void stage_vector(float* dst, const float* __restrict__ src, int n) {
  for (int i = threadIdx.x; i < kMlpMax; i += blockDim.x) dst[i] = i < n ? src[i] : 0.0f;
}

// out[r][v] = sum_k A[r][k] * W[4*lane + v][k],  k < k4 (multiple of 4)
__device__ __forceinline__ void gemm_fwd(const float* __restrict__ A, const float* __restrict__ W, int k4, int lane,
                                         float (&acc)[kMlpRows][4]) {
#pragma unroll
  for (int r = 0; r < kMlpRows; ++r) { acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.0f; }
  const float* wrow = W + (4 * lane) * kMlpMax;
  const int sw = lane & 7;
#pragma unroll 2
  for (int k = 0; k < k4; k += 4) {
    const int chunk = (((k >> 2) ^ sw) << 2);
    float4 w[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) w[v] = *reinterpret_cast<const float4*>(wrow + v * kMlpMax + chunk);
#pragma unroll
    for (int r = 0; r < kMlpRows; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(A + r * kMlpMax + k);
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        acc[r][v] = fmaf(a.x, w[v].x, acc[r][v]);
        acc[r][v] = fmaf(a.y, w[v].y, acc[r][v]);
        acc[r][v] = fmaf(a.z, w[v].z, acc[r][v]);
        acc[r][v] = fmaf(a.w, w[v].w, acc[r][v]);
      }
    }
  }
}

// out[r][v] = sum_o A[r][o] * W[o][4*lane + v],  o < o4 (multiple of 4)
__device__ __forceinline__ void gemm_bwd(const float* __restrict__ A, const float* __restrict__ W, int o4, int lane,
                                         float (&acc)[kMlpRows][4]) {
#pragma unroll
  for (int r = 0; r < kMlpRows; ++r) { acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.0f; }
#pragma unroll 2
  for (int o = 0; o < o4; o += 4) {
    const int chunk = ((lane ^ ((o >> 2) & 7)) << 2);  // rows o..o+3 share (o >> 2)
    float4 w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = *reinterpret_cast<const float4*>(W + (o + j) * kMlpMax + chunk);
#pragma unroll
    for (int r = 0; r < kMlpRows; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(A + r * kMlpMax + o);
      acc[r][0] = fmaf(a.x, w[0].x, acc[r][0]); acc[r][1] = fmaf(a.x, w[0].y, acc[r][1]);
      acc[r][2] = fmaf(a.x, w[0].z, acc[r][2]); acc[r][3] = fmaf(a.x, w[0].w, acc[r][3]);
      acc[r][0] = fmaf(a.y, w[1].x, acc[r][0]); acc[r][1] = fmaf(a.y, w[1].y, acc[r][1]);
      acc[r][2] = fmaf(a.y, w[1].z, acc[r][2]); acc[r][3] = fmaf(a.y, w[1].w, acc[r][3]);
      acc[r][0] = fmaf(a.z, w[2].x, acc[r][0]); acc[r][1] = fmaf(a.z, w[2].y, acc[r][1]);
      acc[r][2] = fmaf(a.z, w[2].z, acc[r][2]); acc[r][3] = fmaf(a.z, w[2].w, acc[r][3]);
      acc[r][0] = fmaf(a.w, w[3].x, acc[r][0]); acc[r][1] = fmaf(a.w, w[3].y, acc[r][1]);
      acc[r][2] = fmaf(a.w, w[3].z, acc[r][2]); acc[r][3] = fmaf(a.w, w[3].w, acc[r][3]);
    }
  }
}

__device__ __forceinline__ void store_rows(float* buf, int lane, const float (&v)[kMlpRows][4]) {
#pragma unroll
  for (int r = 0; r < kMlpRows; ++r)
    *reinterpret_cast<float4*>(buf + r * kMlpMax + 4 * lane) = make_float4(v[r][0], v[r][1], v[r][2], v[r][3]);
}

struct MlpSmem {
  float* W1; float* W2; float* b1; float* b2; float* w3; float* bufA; float* bufB;
};
__device__ __forceinline__ MlpSmem carve(float* sm, int warp) {
  MlpSmem s;
  s.W1 = sm;
  s.W2 = s.W1 + kMlpMax * kMlpMax;
  s.b1 = s.W2 + kMlpMax * kMlpMax;
  s.b2 = s.b1 + kMlpMax;
  s.w3 = s.b2 + kMlpMax;
  s.bufA = s.w3 + kMlpMax + warp * kMlpRows * kMlpMax;
  s.bufB = s.w3 + kMlpMax + kMlpTile * kMlpMax + warp * kMlpRows * kMlpMax;
  return s;
}
constexpr size_t kMlpSmemBytes = (2 * kMlpMax * kMlpMax + 3 * kMlpMax + 2 * kMlpTile * kMlpMax) * sizeof(float);

__device__ __forceinline__ int round4(int v) { return (v + 3) & ~3; }

// forward + input-backward for the warp's 8 rows.  x (registers, lane owns columns 4*lane..+3) -> g.
// Returns per-row energies in e_out (valid on all lanes) when want_e.
template <int ACT>
__device__ __forceinline__ void mlp_grad_rows(const MlpSmem& s, const MlpParams& P, int lane,
                                              const float (&x)[kMlpRows][4], float (&g)[kMlpRows][4], bool want_e,
                                              float (&e_out)[kMlpRows]) {
  float acc[kMlpRows][4], s1[kMlpRows][4];
  store_rows(s.bufA, lane, x);
  __syncwarp();
  // layer 1 forward: z1 = x W1^T + b1
  gemm_fwd(s.bufA, s.W1, round4(P.d), lane, acc);
  {
    const float4 b = *reinterpret_cast<const float4*>(s.b1 + 4 * lane);
    const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int r = 0; r < kMlpRows; ++r)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        float h;
        act_fwd<ACT>(acc[r][v] + bb[v], h, s1[r][v]);
        acc[r][v] = h;
      }
  }
  store_rows(s.bufB, lane, acc);
  __syncwarp();
  // layer 2 forward: z2 = h1 W2^T + b2 ; delta2 = w3 * act'(z2)
  gemm_fwd(s.bufB, s.W2, round4(P.h1), lane, acc);
  {
    const float4 b = *reinterpret_cast<const float4*>(s.b2 + 4 * lane);
    const float4 w = *reinterpret_cast<const float4*>(s.w3 + 4 * lane);
    const float bb[4] = {b.x, b.y, b.z, b.w};
    const float ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int r = 0; r < kMlpRows; ++r) {
      float er = 0.0f;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        float h, dh;
        act_fwd<ACT>(acc[r][v] + bb[v], h, dh);
        er = fmaf(ww[v], h, er);
        acc[r][v] = ww[v] * dh;
      }
      if (want_e) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) er += __shfl_xor_sync(0xffffffffu, er, o);
        e_out[r] = er + P.b3[0];
      }
    }
  }
  store_rows(s.bufA, lane, acc);
  __syncwarp();
  // layer 2 backward: delta1 = (delta2 W2) * act'(z1)
  gemm_bwd(s.bufA, s.W2, round4(P.h2), lane, acc);
#pragma unroll
  for (int r = 0; r < kMlpRows; ++r)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[r][v] *= s1[r][v];
  store_rows(s.bufB, lane, acc);
  __syncwarp();
  // layer 1 backward: g = delta1 W1
  gemm_bwd(s.bufB, s.W1, round4(P.h1), lane, g);
  __syncwarp();
}

__device__ __forceinline__ void stage_all(const MlpSmem& s, const MlpParams& P) {
  stage_matrix(s.W1, P.W1, P.h1, P.d);
  stage_matrix(s.W2, P.W2, P.h2, P.h1);
  stage_vector(s.b1, P.b1, P.h1);
  stage_vector(s.b2, P.b2, P.h2);
  stage_vector(s.w3, P.w3, P.h2);
  __syncthreads();
}

template <int ACT>
__global__ void __launch_bounds__(kMlpWarps * 32, 1) langevin_mlp_kernel(const __grid_constant__ MlpParams P,
                                                                         const __grid_constant__ StepTable tab) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const MlpSmem s = carve(smem, warp);
  {
    MlpSmem s0 = carve(smem, 0);
    stage_all(s0, P);
  }
  const long long n_tiles = (P.n + kMlpTile - 1) / kMlpTile;
  const long long numel = P.n * P.d;
  const bool vec_ok = (P.d % 4 == 0) && ((reinterpret_cast<uintptr_t>(P.x_in) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(P.x_out) & 15) == 0);
  const bool quad_rng = (P.d % 4 == 0);
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long row0 = tile * kMlpTile + warp * kMlpRows;
    float x[kMlpRows][4], g[kMlpRows][4], e_unused[kMlpRows];
#pragma unroll
    for (int r = 0; r < kMlpRows; ++r) {
      const long long row = row0 + r;
#pragma unroll
      for (int v = 0; v < 4; ++v) x[r][v] = 0.0f;
      if (row < P.n) {
        if (vec_ok) {
          if (4 * lane < P.d) {
            const float4 t = *reinterpret_cast<const float4*>(P.x_in + row * P.d + 4 * lane);
            x[r][0] = t.x; x[r][1] = t.y; x[r][2] = t.z; x[r][3] = t.w;
          }
        } else {
#pragma unroll
          for (int v = 0; v < 4; ++v)
            if (4 * lane + v < P.d) x[r][v] = P.x_in[row * P.d + 4 * lane + v];
        }
      }
    }
    int until_keep = P.thin_start, kept = P.kept_base;
    RngStream rs;
    rs.k0 = P.rng.k0; rs.k1 = P.rng.k1; rs.T = P.rng.T; rs.mode = P.rng.mode; rs.ctr_base = P.rng.ctr_base;
    for (int k = 0; k < P.n_steps; ++k) {
      const int ti = k & tab.mask;
      const float h = tab.h[ti], c1 = tab.c1[ti], c2 = tab.c2[ti];
      mlp_grad_rows<ACT>(s, P, lane, x, g, false, e_unused);
#pragma unroll
      for (int r = 0; r < kMlpRows; ++r) {
        const long long row = row0 + r;
        const bool rv = row < P.n;
        float eps[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        const long long li0 = row * P.d + 4 * lane;
        if (rv && 4 * lane < P.d) {
          if (P.rng.mode == 0) {
#pragma unroll
            for (int v = 0; v < 4; ++v)
              if (4 * lane + v < P.d) eps[v] = P.noise[(long long)k * numel + li0 + v];
          } else if (P.rng.mode == 2 && quad_rng) {
            const uint64_t q = (uint64_t)li0 >> 2;
            const uint4 w = philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)rs.ctr_base,
                                          (uint32_t)(rs.ctr_base >> 32), rs.k0, rs.k1);
            const float4 nn = normal4(w);
            eps[0] = nn.x; eps[1] = nn.y; eps[2] = nn.z; eps[3] = nn.w;
          } else {
#pragma unroll
            for (int v = 0; v < 4; ++v)
              if (4 * lane + v < P.d) eps[v] = normal_for_element(rs, (uint64_t)(li0 + v));
          }
        }
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float x1 = __fsub_rn(x[r][v], __fmul_rn(h, g[r][v]));
          float xn = __fadd_rn(x1, __fmul_rn(c2, __fmul_rn(eps[v], c1)));
          if (P.has_clamp) xn = clamp_torch(xn, P.clamp_lo, P.clamp_hi);
          x[r][v] = (rv && 4 * lane + v < P.d) ? xn : 0.0f;
        }
      }
      rs.ctr_base += P.rng.ctr_step;
      if (P.traj && --until_keep == 0) {
        until_keep = P.thin;
        if (kept < P.n_kept) {
#pragma unroll
          for (int r = 0; r < kMlpRows; ++r) {
            const long long row = row0 + r;
#pragma unroll
            for (int v = 0; v < 4; ++v)
              if (row < P.n && 4 * lane + v < P.d) P.traj[(row * P.n_kept + kept) * P.d + 4 * lane + v] = x[r][v];
          }
        }
        ++kept;
      }
    }
#pragma unroll
    for (int r = 0; r < kMlpRows; ++r) {
      const long long row = row0 + r;
      if (row >= P.n) continue;
      if (vec_ok) {
        if (4 * lane < P.d)
          *reinterpret_cast<float4*>(P.x_out + row * P.d + 4 * lane) = make_float4(x[r][0], x[r][1], x[r][2], x[r][3]);
      } else {
#pragma unroll
        for (int v = 0; v < 4; ++v)
          if (4 * lane + v < P.d) P.x_out[row * P.d + 4 * lane + v] = x[r][v];
      }
    }
  }
}


// ---- HMC on MLP energies (hmc.py:244-312 + leapfrog.py:160-185) -----------------------------------------------------
// Same warp-owns-8-chains layout as the Langevin kernel above: per proposal the momentum draw, H0, L leapfrog steps
// (one fused forward + input-backward per step: the force at the bottom of step l is the force at the top of step
// l+1, and the forward pass of the last one is E(x')), H1, the Metropolis test and the select all happen in the warp;
// x, p and the carried force live in registers.  The pre-proposal state is parked in x_out (rows are warp-private), and
// the force at a restored state is recomputed at the top of the next proposal when any of the warp's rows rejected.
__device__ __forceinline__ void mlp_force(float (&g)[kMlpRows][4]) {
#pragma unroll
  for (int r = 0; r < kMlpRows; ++r)
#pragma unroll
    for (int v = 0; v < 4; ++v) g[r][v] = clamp_torch(-g[r][v], -kSafeClamp, kSafeClamp);   // safe=True, hmc.py:258-265
}

template <int ACT>
__global__ void __launch_bounds__(kMlpWarps * 32, 1) hmc_mlp_kernel(const __grid_constant__ MlpParams M,
                                                                    const __grid_constant__ HmcParams P,
                                                                    const __grid_constant__ HStepTable tab) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const MlpSmem s = carve(smem, warp);
  {
    MlpSmem s0 = carve(smem, 0);
    stage_all(s0, M);
  }
  const long long n_tiles = (P.n + kMlpTile - 1) / kMlpTile;
  const long long numel = P.n * P.d;
  const bool quad_rng = (P.d % 4 == 0);
  float minv[4], msqrt[4], mraw[4];
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const int col = 4 * lane + v;
    float mv = 1.0f;
    if (P.mass.kind == 2 && col < P.d) mv = P.mass.vec[col];
    mraw[v] = mv;
    minv[v] = fmaxf(mv, 1e-10f);   // torch.clamp(mass, min=1e-10), leapfrog.py:174
    msqrt[v] = sqrtf(mv);          // torch.sqrt(mass), hmc.py:129
  }
  auto kinetic = [&](const float (&p)[kMlpRows][4], float (&k_out)[kMlpRows]) {   // hmc.py:148-159
#pragma unroll
    for (int r = 0; r < kMlpRows; ++r) {
      float acc = 0.0f;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const float sq = __fmul_rn(p[r][v], p[r][v]);
        acc += (P.mass.kind == 2) ? __fdiv_rn(sq, mraw[v]) : sq;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      acc = __fmul_rn(0.5f, acc);
      if (P.mass.kind == 1) acc = __fdiv_rn(acc, P.mass.scalar);
      k_out[r] = clamp_torch(acc, 0.0f, 1e10f);
    }
  };

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long row0 = tile * kMlpTile + warp * kMlpRows;
    float x[kMlpRows][4], f[kMlpRows][4], p[kMlpRows][4], e_cur[kMlpRows], e_new[kMlpRows];
#pragma unroll
    for (int r = 0; r < kMlpRows; ++r) {
      const long long row = row0 + r;
#pragma unroll
      for (int v = 0; v < 4; ++v) x[r][v] = (row < P.n && 4 * lane + v < P.d) ? P.x_in[row * P.d + 4 * lane + v] : 0.0f;
    }
    RngStream rp, ru;
    rp.k0 = P.rng_p.k0; rp.k1 = P.rng_p.k1; rp.T = P.rng_p.T; rp.mode = P.rng_p.mode; rp.ctr_base = P.rng_p.ctr_base;
    ru.k0 = P.rng_u.k0; ru.k1 = P.rng_u.k1; ru.T = P.rng_u.T; ru.mode = P.rng_u.mode; ru.ctr_base = P.rng_u.ctr_base;
    int until_keep = P.thin_start, kept = P.kept_base;
    bool need_force = true;

    for (int i = 0; i < P.n_prop; ++i) {
      const float h = tab.h[i & tab.mask];
      const float half_h = __fmul_rn(0.5f, h);
      if (need_force) {   // E(x) and the force at the chain state (first proposal, or a row of this warp was restored)
        mlp_grad_rows<ACT>(s, M, lane, x, f, true, e_cur);
        mlp_force(f);
        need_force = false;
      }
      // park the pre-proposal state and draw the momentum (hmc.py:245 / :92-134)
#pragma unroll
      for (int r = 0; r < kMlpRows; ++r) {
        const long long row = row0 + r;
        const bool rv = row < P.n;
        const long long li0 = row * P.d + 4 * lane;
        float eps[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (rv && 4 * lane < P.d) {
#pragma unroll
          for (int v = 0; v < 4; ++v)
            if (4 * lane + v < P.d) P.x_out[li0 + v] = x[r][v];
          if (P.rng_p.mode == 0) {
#pragma unroll
            for (int v = 0; v < 4; ++v)
              if (4 * lane + v < P.d) eps[v] = P.noise_p[(long long)i * numel + li0 + v];
          } else if (P.rng_p.mode == 2 && quad_rng) {
            const uint64_t q = (uint64_t)li0 >> 2;
            const uint4 w = philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)rp.ctr_base,
                                          (uint32_t)(rp.ctr_base >> 32), rp.k0, rp.k1);
            const float4 nn = normal4(w);
            eps[0] = nn.x; eps[1] = nn.y; eps[2] = nn.z; eps[3] = nn.w;
          } else {
#pragma unroll
            for (int v = 0; v < 4; ++v)
              if (4 * lane + v < P.d) eps[v] = normal_for_element(rp, (uint64_t)(li0 + v));
          }
        }
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          if (P.mass.kind == 1) eps[v] = __fmul_rn(eps[v], P.mass.sqrt_scalar);   // hmc.py:124
          else if (P.mass.kind == 2) eps[v] = __fmul_rn(eps[v], msqrt[v]);         // hmc.py:133
          p[r][v] = eps[v];
        }
      }
      float k0[kMlpRows], h0[kMlpRows];
      kinetic(p, k0);
#pragma unroll
      for (int r = 0; r < kMlpRows; ++r) h0[r] = __fadd_rn(clamp_torch(e_cur[r], -1e10f, 1e10f), k0[r]);   // hmc.py:247-256
      // leapfrog (leapfrog.py:160-185), safe mode
      for (int l = 0; l < P.n_leapfrog; ++l) {
#pragma unroll
        for (int r = 0; r < kMlpRows; ++r)
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            p[r][v] = __fadd_rn(p[r][v], __fmul_rn(half_h, f[r][v]));   // p_half
            float dx = __fmul_rn(h, p[r][v]);
            if (P.mass.kind == 1) dx = __fdiv_rn(dx, P.mass.safe_scalar);
            else if (P.mass.kind == 2) dx = __fdiv_rn(dx, minv[v]);
            x[r][v] = (4 * lane + v < P.d) ? __fadd_rn(x[r][v], dx) : 0.0f;
          }
        mlp_grad_rows<ACT>(s, M, lane, x, f, true, e_new);
        mlp_force(f);
        bool dirty = false;
#pragma unroll
        for (int r = 0; r < kMlpRows; ++r)
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            p[r][v] = __fadd_rn(p[r][v], __fmul_rn(half_h, f[r][v]));
            const float xs = nan_to_num0(x[r][v]), ps = nan_to_num0(p[r][v]);
            dirty |= (__float_as_uint(xs) != __float_as_uint(x[r][v]));
            x[r][v] = xs;
            p[r][v] = ps;
          }
        if (__any_sync(0xffffffffu, dirty)) {   // the reference recomputes the force at the sanitised state
          mlp_grad_rows<ACT>(s, M, lane, x, f, true, e_new);
          mlp_force(f);
        }
      }
      // H1, Metropolis test (hmc.py:268-292)
      float k1[kMlpRows];
      kinetic(p, k1);
      int n_acc = 0;
      bool any_reject = false;
#pragma unroll
      for (int r = 0; r < kMlpRows; ++r) {
        const long long row = row0 + r;
        const bool rv = row < P.n;
        const float h1 = __fadd_rn(clamp_torch(e_new[r], -1e10f, 1e10f), k1[r]);
        const float dh = clamp_torch(__fsub_rn(h0[r], h1), -50.0f, 50.0f);
        float a = expf(dh);
        a = (a != a) ? a : fminf(a, 1.0f);   // clamp_(max=1.0) keeps NaN
        float u = 0.0f;
        if (rv) u = (P.rng_u.mode == 0) ? P.noise_u[(long long)i * P.n + row] : uniform_for_element(ru, (uint64_t)row);
        const bool accepted = u < a;
        if (accepted) {
          e_cur[r] = e_new[r];
          n_acc += rv ? 1 : 0;
        } else {
          any_reject = true;
#pragma unroll
          for (int v = 0; v < 4; ++v)
            x[r][v] = (rv && 4 * lane + v < P.d) ? P.x_out[row * P.d + 4 * lane + v] : 0.0f;
        }
      }
      need_force = any_reject;   // warp-uniform: every lane saw the same per-row decisions
      if (P.accept_count && lane == 0 && n_acc > 0) atomicAdd(P.accept_count + P.prop_base + i, n_acc);
      rp.ctr_base += P.rng_p.ctr_step;
      ru.ctr_base += P.rng_u.ctr_step;
      if (P.traj && --until_keep == 0) {
        until_keep = P.thin;
        if (kept < P.n_kept) {
#pragma unroll
          for (int r = 0; r < kMlpRows; ++r) {
            const long long row = row0 + r;
#pragma unroll
            for (int v = 0; v < 4; ++v)
              if (row < P.n && 4 * lane + v < P.d) P.traj[(row * P.n_kept + kept) * P.d + 4 * lane + v] = x[r][v];
          }
        }
        ++kept;
      }
    }
#pragma unroll
    for (int r = 0; r < kMlpRows; ++r) {
      const long long row = row0 + r;
      if (row >= P.n) continue;
#pragma unroll
      for (int v = 0; v < 4; ++v)
        if (4 * lane + v < P.d) P.x_out[row * P.d + 4 * lane + v] = x[r][v];
      if (P.energy_out && lane == 0) P.energy_out[row] = clamp_torch(e_cur[r], -1e10f, 1e10f);
    }
  }
}

template <int ACT>
__global__ void __launch_bounds__(kMlpWarps * 32, 1) mlp_energy_grad_kernel(const __grid_constant__ MlpParams P) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const MlpSmem s = carve(smem, warp);
  {
    MlpSmem s0 = carve(smem, 0);
    stage_all(s0, P);
  }
  const long long n_tiles = (P.n + kMlpTile - 1) / kMlpTile;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long row0 = tile * kMlpTile + warp * kMlpRows;
    float x[kMlpRows][4], g[kMlpRows][4], e[kMlpRows];
#pragma unroll
    for (int r = 0; r < kMlpRows; ++r) {
      const long long row = row0 + r;
#pragma unroll
      for (int v = 0; v < 4; ++v) x[r][v] = (row < P.n && 4 * lane + v < P.d) ? P.x_in[row * P.d + 4 * lane + v] : 0.0f;
    }
    mlp_grad_rows<ACT>(s, P, lane, x, g, P.energy != nullptr, e);
#pragma unroll
    for (int r = 0; r < kMlpRows; ++r) {
      const long long row = row0 + r;
      if (row >= P.n) continue;
      if (P.grad) {
#pragma unroll
        for (int v = 0; v < 4; ++v)
          if (4 * lane + v < P.d) P.grad[row * P.d + 4 * lane + v] = g[r][v];
      }
      if (P.energy && lane == 0) P.energy[row] = e[r];
    }
  }
}

static int fill_mlp(const EbmEnergyDesc* e, MlpParams& P) {
  if (e->hidden3 > 0) {
    set_error("three-hidden-layer MLP energies have no kernel for this operation (Langevin bursts with precision bf16x3 / bf16, "
              "energy and gradient evaluation only)");
    return EBM_ERR_UNSUPPORTED;
  }
  if (e->dim > kMlpMax || e->hidden1 > kMlpMax || e->hidden2 > kMlpMax) {
    set_error("MLP energy %d->%d->%d->1: widths above %d are not supported by this build", e->dim, e->hidden1,
              e->hidden2, kMlpMax);
    return EBM_ERR_UNSUPPORTED;
  }
  if (e->activation < EBM_ACT_SILU || e->activation > EBM_ACT_SOFTPLUS) { set_error("bad activation"); return EBM_ERR_INVALID; }
  memset(&P, 0, sizeof(P));
  P.W1 = e->buf[0]; P.b1 = e->buf[1]; P.W2 = e->buf[2]; P.b2 = e->buf[3]; P.w3 = e->buf[4]; P.b3 = e->buf[5];
  P.d = e->dim; P.h1 = e->hidden1; P.h2 = e->hidden2;
  return 0;
}

static int mlp_grid(const DeviceInfo& di, long long n) {
  long long tiles = (n + kMlpTile - 1) / kMlpTile;
  if (tiles > di.sm_count) tiles = di.sm_count;
  return (int)(tiles < 1 ? 1 : tiles);
}

#define EBM_ACT_DISPATCH(act, CALL)                        \
  switch (act) {                                           \
    case EBM_ACT_SILU: CALL(EBM_ACT_SILU); break;          \
    case EBM_ACT_TANH: CALL(EBM_ACT_TANH); break;          \
    case EBM_ACT_RELU: CALL(EBM_ACT_RELU); break;          \
    default: CALL(EBM_ACT_SOFTPLUS); break;                \
  }


// one chunk of proposals of ebm_hmc_burst_f32 for an MLP energy (called from the chunk loop in ebm_hmc.cu)
int hmc_mlp_launch(const EbmEnergyDesc* e, const HmcParams& P, const HStepTable& tab, cudaStream_t st) {
  MlpParams M;
  int rc = fill_mlp(e, M);
  if (rc) return rc;
  const DeviceInfo& di = device_info(current_device());
#define CALL(A)                                                                                              \
  {                                                                                                          \
    auto kern = hmc_mlp_kernel<A>;                                                                           \
    EBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMlpSmemBytes));   \
    kern<<<mlp_grid(di, P.n), kMlpWarps * 32, kMlpSmemBytes, st>>>(M, P, tab);                               \
  }
  EBM_ACT_DISPATCH(e->activation, CALL);
#undef CALL
  return launch_status("hmc_mlp_kernel");
}

int mlp_wide_energy_grad_dispatch(const EbmEnergyDesc* e, const float* x, int64_t n, float* energy, float* grad,
                                  cudaStream_t st);  // ebm_mlp_wide.cu
int mlp_deep_energy_grad_dispatch(const EbmEnergyDesc* e, const float* x, int64_t n, float* energy, float* grad,
                                  cudaStream_t st);  // ebm_mlp_deep.cu

int mlp_energy_grad_dispatch(const EbmEnergyDesc* e, const float* x, int64_t n, float* energy, float* grad,
                             cudaStream_t st) {
  if (e->hidden3 > 0) return mlp_deep_energy_grad_dispatch(e, x, n, energy, grad, st);
  if (e->dim > kMlpMax) return mlp_wide_energy_grad_dispatch(e, x, n, energy, grad, st);
  MlpParams P;
  int rc = fill_mlp(e, P);
  if (rc) return rc;
  const DeviceInfo& di = device_info(current_device());
  P.x_in = x; P.n = n; P.energy = energy; P.grad = grad;
#define CALL(A)                                                                                              \
  {                                                                                                          \
    auto kern = mlp_energy_grad_kernel<A>;                                                                   \
    EBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMlpSmemBytes));   \
    kern<<<mlp_grid(di, n), kMlpWarps * 32, kMlpSmemBytes, st>>>(P);                                         \
  }
  EBM_ACT_DISPATCH(e->activation, CALL);
#undef CALL
  return launch_status("mlp_energy_grad_kernel");
}

int langevin_mlp_tc_dispatch(const LangevinCall& c, int passes);    // ebm_mlp_tc.cu
int langevin_mlp_wide_dispatch(const LangevinCall& c, int passes);  // ebm_mlp_wide.cu
int langevin_mlp_deep_dispatch(const LangevinCall& c, int passes);  // ebm_mlp_deep.cu

int langevin_mlp_dispatch(const LangevinCall& c) {
  if (c.e->hidden3 > 0) {   // three hidden layers: tensor-core kernel with every A operand in tensor memory
    if (c.e->precision == EBM_MLP_BF16X3) return langevin_mlp_deep_dispatch(c, 3);
    if (c.e->precision == EBM_MLP_BF16) return langevin_mlp_deep_dispatch(c, 1);
    set_error("three-hidden-layer MLP energies run on the tensor-core kernel only (precision bf16x3 or bf16)");
    return EBM_ERR_UNSUPPORTED;
  }
  if (c.e->dim > kMlpMax) {  // state wider than one tile: streamed-operand tensor-core kernel
    if (c.e->precision == EBM_MLP_BF16X3) return langevin_mlp_wide_dispatch(c, 3);
    if (c.e->precision == EBM_MLP_BF16) return langevin_mlp_wide_dispatch(c, 1);
    set_error("MLP energies with dim > %d run on the tensor-core kernel only (precision bf16x3 or bf16)", kMlpMax);
    return EBM_ERR_UNSUPPORTED;
  }
  if (c.e->precision == EBM_MLP_BF16X3) return langevin_mlp_tc_dispatch(c, 3);
  if (c.e->precision == EBM_MLP_BF16) return langevin_mlp_tc_dispatch(c, 1);
  if (c.e->precision != EBM_MLP_FP32) { set_error("bad MLP precision %d", c.e->precision); return EBM_ERR_INVALID; }
  MlpParams P;
  int rc = fill_mlp(c.e, P);
  if (rc) return rc;
  const DeviceInfo& di = device_info(current_device());
  const long long numel = (long long)c.n * c.e->dim;
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
    P.n_steps = chunk;
    P.noise = c.noise ? c.noise + (long long)done * numel : nullptr;
    P.rng.ctr_base = c.offset / 4 + (unsigned long long)done * P.rng.ctr_step;
    P.thin_start = c.thin - (done % c.thin);
    P.kept_base = done / c.thin;
#define CALL(A)                                                                                              \
  {                                                                                                          \
    auto kern = langevin_mlp_kernel<A>;                                                                      \
    EBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMlpSmemBytes));   \
    kern<<<mlp_grid(di, c.n), kMlpWarps * 32, kMlpSmemBytes, c.st>>>(P, tab);                                \
  }
    EBM_ACT_DISPATCH(c.e->activation, CALL);
#undef CALL
    rc = launch_status("langevin_mlp_kernel");
    if (rc) return rc;
    done += chunk;
    src = c.x_out;
  }
  return 0;
}

}  // namespace ebm
