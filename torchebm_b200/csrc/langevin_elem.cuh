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
#include "diag.cuh"
#include "energies.cuh"
#include "rng.cuh"

namespace ebm {

constexpr int kSchedChunk = 64;
constexpr int kMaxPeers = 16;

// Stores to an NVLS multicast address (NVSwitch replicates the write into every rank's copy of the symmetric buffer: one
// store instead of one per peer).  Multicast addresses may only be accessed with multimem.* instructions.
__device__ __forceinline__ void mc_store4(float* p, float a, float b, float c, float d) {
  asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void mc_store1(float* p, float a) {
  asm volatile("multimem.st.weak.global.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}

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
  double* diag_ws;     // in-burst diagnostics workspace [n_kept, diag_slot(d)] or null (diag.cuh); TRAJ instantiations only
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
  PhiloxKeys keys;              // the ten Philox round keys of (k0, k1)
  unsigned long long ctr_base;  // TORCH: off/4 ; NATIVE: off/4
  unsigned long long ctr_step;  // TORCH: inc/4 per step ; NATIVE: 1 per step
  unsigned long long T;         // TORCH layout threads
  unsigned long long n_quads;   // number of owning threads
  unsigned long long quad_base; // first owning thread of this launch (host-pipelined bursts launch quad ranges)
  unsigned long long quad_end;  // one past the last owning thread of this launch
  // burst-end gather fused into the final store: the state is also written at element offset peer_off of every
  // peer-mapped gathered buffer (NVLink stores; ebm_langevin_burst_gather_f32)
  int n_peers;
  int peer_mc;         // 1: peers[0] is an NVLS multicast address of the gathered buffers (n_peers == 1)
  long long peer_off;
  float* peers[kMaxPeers];
};

// HEUN: the two-stage tableau of integrators/heun.py through the reference's generic RK path
// (core/base_integrator.py:300-347,387-397): k1 = f(x), k2 = f(x + h k1), x1 = x + h (k1/2 + k2/2), then the same noise.
// The Heun instantiations always compile the clamp in (bounds = -inf / +inf when the sampler has none).
//
// The transform of the Philox words into normals runs on packed fp32x2 registers (rng.cuh), the update on scalar
// ones; rounding order is the reference's x' = fl(fl(x - fl(h g)) + fl(c2 fl(eps c1))) (base_integrator.py:387-397,728-729).
template <class EnergyT, int RNG, bool TRAJ, bool CLAMP, bool HEUN = false>
__global__ void __launch_bounds__(256) langevin_elem_kernel(const __grid_constant__ LangevinElemParams P,
                                                            const EnergyT en,
                                                            const __grid_constant__ StepTable tab) {
  const unsigned long long gid = P.quad_base + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = gid < P.quad_end;
  if (!TRAJ && !live) return;    // (the keeping instantiations reduce over whole warps: idle threads stay, owning nothing)

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
  const bool vec = live && (RNG != 1) && (idx[3] < P.numel) && ((reinterpret_cast<uintptr_t>(P.x_in) & 15) == 0);
  if (vec) {
    const float4 v = *reinterpret_cast<const float4*>(P.x_in + idx[0]);
    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    ok[0] = ok[1] = ok[2] = ok[3] = true;
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ok[i] = live && idx[i] < P.numel;
      x[i] = ok[i] ? P.x_in[idx[i]] : 0.0f;
    }
  }
  f32x2 X01 = pack2(x[0], x[1]), X23 = pack2(x[2], x[3]);

  long long tbase[4];
  int tcol[4];
  if (TRAJ) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long r = idx[i] / P.d;
      const long long c = idx[i] - r * P.d;
      tbase[i] = r * (long long)P.n_kept * P.d + c;
      tcol[i] = (int)c;
    }
  }
  int until_keep = P.thin_start;
  int kept = P.kept_base;

  // One element pair: gradient, drift step, noise, clamp.
  // Pipe facts (tools/pipe_probe.cu on B200, cycles per warp instruction per sub-partition): IMAD.WIDE 4.2, FFMA / FMUL
  // 1.5, FFMA2 / FMUL2 2.1 (= 1.04 per element), LOP3 2.1 on another pipe, MUFU 8.  IMAD.WIDE, scalar and packed fp32
  // all execute on the one FMA pipe, which is what this kernel is bound by: Philox's 20 IMAD.WIDE per thread-step are 83
  // of its ~195 cycles.  Packing therefore buys pipe time only through the 1.04 vs 1.5 per element.
  // TORCH / INJECTED streams: the reference's rounding order, every product and sum rounded separately (scalar ops).
  // NATIVE stream: no bit-parity contract -- contracted packed form (5 packed instructions per pair instead of 18 scalar).
  auto update2 = [&](f32x2 X, f32x2 E, float h, float c1s, float c2) -> f32x2 {
    if (RNG == 2 && !HEUN) {
      const f32x2 G = en.grad2_fast(X);
      f32x2 XN = fma2(E, __fmul_rn(c1s, c2), fma2(G, -h, X));
      if (CLAMP) {
        float a, b;
        unpack2(XN, a, b);
        XN = pack2(clamp_torch(a, P.clamp_lo, P.clamp_hi), clamp_torch(b, P.clamp_lo, P.clamp_hi));
      }
      return XN;
    }
    float xv[2], ev[2];
    unpack2(X, xv[0], xv[1]);
    unpack2(E, ev[0], ev[1]);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float g = en.grad(xv[u]);
      if (HEUN) {
        const float g2 = en.grad(__fsub_rn(xv[u], __fmul_rn(h, g)));
        g = __fadd_rn(__fmul_rn(0.5f, g), __fmul_rn(0.5f, g2));
      }
      // x1 = x + h*(1.0*(-g)) ; dw = eps*c1 ; x' = x1 + c2*dw   (base_integrator.py:387-397,728-729)
      const float x1 = __fsub_rn(xv[u], __fmul_rn(h, g));
      float xn = __fadd_rn(x1, __fmul_rn(c2, __fmul_rn(ev[u], c1s)));
      if (CLAMP) xn = clamp_torch(xn, P.clamp_lo, P.clamp_hi);
      xv[u] = xn;
    }
    return pack2(xv[0], xv[1]);
  };

  auto step = [&](int k, float h, float c1, float c2) {
    f32x2 E01, E23;
    float c1s = c1;
    if (RNG == 0) {
      const float* nz = P.noise + (long long)k * P.numel;
      float e[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) e[i] = ok[i] ? nz[idx[i]] : 0.0f;
      E01 = pack2(e[0], e[1]);
      E23 = pack2(e[2], e[3]);
    } else if (RNG == 1) {
      const uint4 w = philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), c2w, c3w, P.keys);
      ctr += P.ctr_step;
      neg_normal4_packed(w, E01, E23);     // bit-identical to torch's curand_normal4, negated
      c1s = -c1;
    } else {
      const uint4 w = philox4x32_10(c2w, c3w, (uint32_t)ctr, (uint32_t)(ctr >> 32), P.keys);
      ctr += P.ctr_step;
      normal4_fast_packed(w, E01, E23);    // the native stream pins the Philox words, not the transform's last bits
    }
    X01 = update2(X01, E01, h, c1s, c2);
    X23 = update2(X23, E23, h, c1s, c2);
    if (TRAJ) {
      if (--until_keep == 0) {
        until_keep = P.thin;
        if (kept < P.n_kept) {
          unpack2(X01, x[0], x[1]);
          unpack2(X23, x[2], x[3]);
          if (P.traj) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (ok[i]) P.traj[tbase[i] + (long long)kept * P.d] = x[i];
          }
          if (P.diag_ws) {   // column sums and sums of squares, sum of the energy terms (diag.cuh)
            double* slot = P.diag_ws + (long long)kept * diag_slot(P.d);
            double es = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (ok[i]) {
                const double xv = (double)x[i];
                atomicAdd(slot + tcol[i], xv);
                atomicAdd(slot + P.d + tcol[i], xv * xv);
                es += (double)en.term(x[i]);
              }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) es += __shfl_xor_sync(0xffffffffu, es, o);
            if ((threadIdx.x & 31) == 0) atomicAdd(slot + 2 * P.d, es);
          }
        }
        ++kept;
      }
    }
  };

  if (tab.mask == 0) {   // constant schedule: coefficients stay in registers
    const float h = tab.h[0], c1 = tab.c1[0], c2 = tab.c2[0];
    for (int k = 0; k < P.n_steps; ++k) step(k, h, c1, c2);
  } else {
    for (int k = 0; k < P.n_steps; ++k) step(k, tab.h[k], tab.c1[k], tab.c2[k]);
  }
  unpack2(X01, x[0], x[1]);
  unpack2(X23, x[2], x[3]);

  if (vec && ((reinterpret_cast<uintptr_t>(P.x_out) & 15) == 0)) {
    *reinterpret_cast<float4*>(P.x_out + idx[0]) = make_float4(x[0], x[1], x[2], x[3]);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (ok[i]) P.x_out[idx[i]] = x[i];
  }
  for (int w = 0; w < P.n_peers; ++w) {  // one store per rank of the box, over NVLink for the remote ones
    float* dst = P.peers[w] + P.peer_off;
    const bool v4 = vec && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
    if (P.peer_mc) {                     // one multicast store reaches every rank's copy
      if (v4) mc_store4(dst + idx[0], x[0], x[1], x[2], x[3]);
      else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (ok[i]) mc_store1(dst + idx[i], x[i]);
      }
    } else if (v4) {
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
