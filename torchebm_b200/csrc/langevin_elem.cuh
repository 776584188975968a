// K-step Langevin burst for elementwise-separable energies (DoubleWell, Harmonic, Rastrigin).
//
// Replaces the per-step loop of LangevinDynamics.sample (samplers/langevin_dynamics.py:157-185) +
// BaseSDERungeKuttaIntegrator.step (core/base_integrator.py:673-731).  Each thread owns one Philox
// block = 4 state elements, loads them once, runs all K steps in registers and stores once, so HBM
// sees 8 bytes per element per BURST; per step the only traffic is the optional trajectory write.
//
// Element ownership follows the RNG layout so that every Philox block is fully used:
//   TORCH   : thread (t, j) owns li = t + T*(4j + ii), ii = 0..3  (torch's grid-stride quad)
//   NATIVE / INJECTED : thread q owns li = 4q + ii                 (one float4)
#pragma once
#include "energies.cuh"
#include "rng.cuh"

namespace ebm {

constexpr int kSchedChunk = 64;
constexpr int kMaxPeers = 16;

struct StepTable {           // per-step coefficients, fp32 as torch rounds the Python doubles
  float h[kSchedChunk];      // step_size
  float c1[kSchedChunk];     // step_size ** 0.5
  float c2[kSchedChunk];     // (2 * noise_scale**2) ** 0.5
  int mask;                  // 0: constant schedule (entry 0), ~0: per-step entries
};

struct LangevinElemParams {
  const float* x_in;
  float* x_out;
  const float* noise;  // INJECTED: [n_steps, numel]
  float* traj;         // [n, n_kept, d] or null
  long long numel;
  int d;
  int n_steps;
  int thin;
  int n_kept;
  int thin_start;  // steps until the first kept sample of this launch (chunked schedules)
  int kept_base;   // index of the first kept sample of this launch
  int has_clamp;
  float clamp_lo, clamp_hi;
  // rng
  uint32_t k0, k1;
  unsigned long long ctr_base;  // TORCH: off/4 ; NATIVE: off/4
  unsigned long long ctr_step;  // TORCH: inc/4 per step ; NATIVE: 1 per step
  unsigned long long T;         // TORCH layout threads
  unsigned long long n_quads;   // number of owning threads
  unsigned long long quad_base; // first owning thread of this launch (host-pipelined bursts launch quad ranges)
  unsigned long long quad_end;  // one past the last owning thread of this launch
  // burst-end gather fused into the final store: the state is also written at element offset peer_off of every
  // peer-mapped gathered buffer (NVLink stores; ebm_langevin_burst_gather_f32)
  int n_peers;
  long long peer_off;
  float* peers[kMaxPeers];
};

// HEUN: the two-stage tableau of integrators/heun.py through the reference's generic RK path
// (core/base_integrator.py:300-347,387-397): k1 = f(x), k2 = f(x + h k1), x1 = x + h (k1/2 + k2/2), then the same noise.
// The Heun instantiations always compile the clamp in (bounds = -inf / +inf when the sampler has none).
template <class EnergyT, int RNG, bool TRAJ, bool CLAMP, bool HEUN = false>
__global__ void __launch_bounds__(256) langevin_elem_kernel(const __grid_constant__ LangevinElemParams P,
                                                            const EnergyT en,
                                                            const __grid_constant__ StepTable tab) {
  const unsigned long long gid = P.quad_base + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= P.quad_end) return;

  long long idx[4];
  uint32_t c2w, c3w;             // fixed counter words
  unsigned long long ctr;        // stepping counter words (lo, hi)
  if (RNG == 1) {
    const unsigned long long j = gid / P.T;
    const unsigned long long t = gid - j * P.T;
#pragma unroll
    for (int i = 0; i < 4; ++i) idx[i] = (long long)(t + P.T * (4ull * j + i));
    ctr = P.ctr_base + j;
    c2w = (uint32_t)t;
    c3w = (uint32_t)(t >> 32);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) idx[i] = (long long)(4ull * gid + i);
    ctr = P.ctr_base;
    c2w = (uint32_t)gid;          // NATIVE: ctr = (lo(q), hi(q), lo(step), hi(step))
    c3w = (uint32_t)(gid >> 32);
  }

  float x[4];
  bool ok[4];
  const bool vec = (RNG != 1) && (idx[3] < P.numel) && ((reinterpret_cast<uintptr_t>(P.x_in) & 15) == 0);
  if (vec) {
    const float4 v = *reinterpret_cast<const float4*>(P.x_in + idx[0]);
    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    ok[0] = ok[1] = ok[2] = ok[3] = true;
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ok[i] = idx[i] < P.numel;
      x[i] = ok[i] ? P.x_in[idx[i]] : 0.0f;
    }
  }

  long long tbase[4];
  if (TRAJ) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long r = idx[i] / P.d;
      const long long c = idx[i] - r * P.d;
      tbase[i] = r * (long long)P.n_kept * P.d + c;
    }
  }
  int until_keep = P.thin_start;
  int kept = P.kept_base;

  for (int k = 0; k < P.n_steps; ++k) {
    const int ti = k & tab.mask;
    const float h = tab.h[ti], c1 = tab.c1[ti], c2 = tab.c2[ti];
    float e[4];
    if (RNG == 0) {
      const float* nz = P.noise + (long long)k * P.numel;
#pragma unroll
      for (int i = 0; i < 4; ++i) e[i] = ok[i] ? nz[idx[i]] : 0.0f;
    } else {
      uint4 w;
      if (RNG == 1) w = philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), c2w, c3w, P.k0, P.k1);
      else          w = philox4x32_10(c2w, c3w, (uint32_t)ctr, (uint32_t)(ctr >> 32), P.k0, P.k1);
      ctr += P.ctr_step;
      const float4 nrm = normal4(w);
      e[0] = nrm.x; e[1] = nrm.y; e[2] = nrm.z; e[3] = nrm.w;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float g = en.grad(x[i]);
      if (HEUN) {
        const float g2 = en.grad(__fsub_rn(x[i], __fmul_rn(h, g)));
        g = __fadd_rn(__fmul_rn(0.5f, g), __fmul_rn(0.5f, g2));
      }
      // x1 = x + h*(1.0*(-g)) ; dw = eps*c1 ; x' = x1 + c2*dw   (base_integrator.py:387-397,728-729)
      const float x1 = __fsub_rn(x[i], __fmul_rn(h, g));
      float xn = __fadd_rn(x1, __fmul_rn(c2, __fmul_rn(e[i], c1)));
      if (CLAMP) xn = clamp_torch(xn, P.clamp_lo, P.clamp_hi);
      x[i] = xn;
    }
    if (TRAJ) {
      if (--until_keep == 0) {
        until_keep = P.thin;
        if (kept < P.n_kept) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (ok[i]) P.traj[tbase[i] + (long long)kept * P.d] = x[i];
        }
        ++kept;
      }
    }
  }

  if (vec && ((reinterpret_cast<uintptr_t>(P.x_out) & 15) == 0)) {
    *reinterpret_cast<float4*>(P.x_out + idx[0]) = make_float4(x[0], x[1], x[2], x[3]);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (ok[i]) P.x_out[idx[i]] = x[i];
  }
  for (int w = 0; w < P.n_peers; ++w) {  // one store per rank of the box, over NVLink for the remote ones
    float* dst = P.peers[w] + P.peer_off;
    if (vec && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
      *reinterpret_cast<float4*>(dst + idx[0]) = make_float4(x[0], x[1], x[2], x[3]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (ok[i]) dst[idx[i]] = x[i];
    }
  }
}

// ---- noise-free descent bursts (GradientDescentSampler / NesterovSampler, samplers/gradient_descent.py:123-138,
// 258-276) for the elementwise energies: one thread owns 4 consecutive elements for all K steps.  The reference's
// `torch.sub(x, grad, alpha=eta)`, `torch.add(x, v, alpha=mu)` and `v.mul_(mu).sub_(grad, alpha=eta)` are single-rounding
// fused multiply-adds in ATen (CPU and CUDA), so the updates here are explicit fmaf.
struct DescentElemParams {
  const float* x_in;
  float* x_out;
  float* traj;        // [n, n_kept, d] or null
  long long numel;
  int d, n_steps, thin, n_kept, thin_start, kept_base;
  int nesterov;       // 0: gradient descent, 1: Nesterov momentum
  float mu;
  const float* v_in;  // Nesterov: velocity carried between launches of one burst (null = zeros)
  float* v_out;       // Nesterov: velocity after this launch (null = not needed)
};

template <class EnergyT, bool NESTEROV, bool TRAJ>
__global__ void __launch_bounds__(256) descent_elem_kernel(const __grid_constant__ DescentElemParams P, const EnergyT en,
                                                           const __grid_constant__ StepTable tab) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i0 = 4 * q;
  if (i0 >= P.numel) return;
  float x[4], v[4];
  bool ok[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ok[i] = i0 + i < P.numel;
    x[i] = ok[i] ? P.x_in[i0 + i] : 0.0f;
    v[i] = (NESTEROV && ok[i] && P.v_in) ? P.v_in[i0 + i] : 0.0f;
  }
  long long tbase[4];
  if (TRAJ) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long r = (i0 + i) / P.d;
      tbase[i] = r * (long long)P.n_kept * P.d + ((i0 + i) - r * P.d);
    }
  }
  int until_keep = P.thin_start, kept = P.kept_base;
  for (int k = 0; k < P.n_steps; ++k) {
    const float eta = tab.h[k & tab.mask];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (NESTEROV) {
        const float look = __fmaf_rn(P.mu, v[i], x[i]);          // torch.add(x, v, alpha=mu)
        const float g = en.grad(look);
        v[i] = __fmaf_rn(-eta, g, __fmul_rn(v[i], P.mu));       // v.mul_(mu).sub_(grad, alpha=eta)
        x[i] = __fadd_rn(x[i], v[i]);
      } else {
        x[i] = __fmaf_rn(-eta, en.grad(x[i]), x[i]);             // torch.sub(x, grad, alpha=eta)
      }
    }
    if (TRAJ && --until_keep == 0) {
      until_keep = P.thin;
      if (kept < P.n_kept) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (ok[i]) P.traj[tbase[i] + (long long)kept * P.d] = x[i];
      }
      ++kept;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (!ok[i]) continue;
    P.x_out[i0 + i] = x[i];
    if (NESTEROV && P.v_out) P.v_out[i0 + i] = v[i];
  }
}

}  // namespace ebm
