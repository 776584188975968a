// C-ABI: library bookkeeping, RNG test hook, energy/gradient, Euler-Maruyama step, Langevin bursts
// for the analytic energies, persistent-CD buffer gather/scatter.  See include/ebm_b200.h.
#include <stdarg.h>

#include <mutex>

#include "api_common.cuh"
#include "mlp_schedule.cuh"

namespace ebm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}

const DeviceInfo& device_info(int device) {
  static DeviceInfo info[64];
  static bool have[64] = {false};
  static std::mutex mu;
  if (device < 0 || device >= 64) device = 0;
  std::lock_guard<std::mutex> lock(mu);
  if (!have[device]) {
    DeviceInfo di{148, 2048, 227 * 1024};
    cudaDeviceGetAttribute(&di.sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&di.max_threads_per_sm, cudaDevAttrMaxThreadsPerMultiProcessor, device);
    cudaDeviceGetAttribute(&di.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    info[device] = di;
    have[device] = true;
  }
  return info[device];
}

// ---- RNG test hook ---------------------------------------------------------------------------
__global__ void rng_fill_kernel(float* out, long long numel, RngStream s, int kind) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long li = (long long)blockIdx.x * blockDim.x + threadIdx.x; li < numel; li += stride)
    out[li] = kind == 0 ? normal_for_element(s, (uint64_t)li) : uniform_for_element(s, (uint64_t)li);
}

// ---- Euler-Maruyama step with an opaque drift ------------------------------------------------
__global__ void em_step_kernel(const float* __restrict__ x, const float* __restrict__ drift,
                               const float* __restrict__ noise, float* __restrict__ out, long long numel, float h,
                               float c1, float c2, int has_noise) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += stride) {
    float v = __fadd_rn(x[i], __fmul_rn(h, drift[i]));
    if (has_noise) v = __fadd_rn(v, __fmul_rn(c2, __fmul_rn(noise[i], c1)));
    out[i] = v;
  }
}

// ---- persistent-CD buffer ----------------------------------------------------------------------
__global__ void pcd_gather_kernel(const float* __restrict__ buffer, long long row_elems, const long long* __restrict__ idx,
                                  long long batch, float* __restrict__ out) {
  const long long total = batch * row_elems;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = i / row_elems, c = i - r * row_elems;
    out[i] = buffer[idx[r] * row_elems + c];
  }
}
__global__ void pcd_noise_kernel(float* __restrict__ out, long long row_elems, const long long* __restrict__ rows,
                                 const float* __restrict__ noise, long long n_noise) {
  const long long total = n_noise * row_elems;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long j = i / row_elems, c = i - j * row_elems;
    float* p = out + rows[j] * row_elems + c;
    *p = __fadd_rn(*p, __fmul_rn(noise[i], 0.01f));  // base_loss.py:322-331
  }
}
__global__ void pcd_scatter_kernel(float* __restrict__ buffer, long long buffer_rows, long long row_elems, long long ptr,
                                   const float* __restrict__ samples, long long batch) {
  const long long total = batch * row_elems;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = i / row_elems, c = i - r * row_elems;
    long long dst = ptr + r;
    if (dst >= buffer_rows) dst -= buffer_rows;
    buffer[dst * row_elems + c] = samples[i];
  }
}

static inline int flat_grid(const DeviceInfo& di, long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = (long long)di.sm_count * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ---- energy / gradient dispatch ---------------------------------------------------------------
template <class RowE>
static int launch_energy_grad(const RowE& en, const EbmEnergyDesc* e, const float* x, int64_t n, float* energy,
                              float* grad, cudaStream_t st) {
  const DeviceInfo& di = device_info(current_device());
#define CALL(G, E)                                                                                  \
  {                                                                                                 \
    int ss;                                                                                         \
    const size_t smem = row_smem_bytes(e, G, ss);                                                   \
    if (smem > (size_t)di.max_smem_optin) { set_error("energy parameters need %zu B of shared memory", smem); return EBM_ERR_UNSUPPORTED; } \
    auto kern = row_energy_grad_kernel<RowE, G, E>;                                                 \
    if (smem > 48 * 1024) EBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<row_grid(di, n, G, 8), kRowThreads, smem, st>>>(en, x, n, e->dim, energy, grad, ss);     \
  }
  EBM_ROW_DISPATCH(e->dim, CALL);
#undef CALL
  return launch_status("row_energy_grad_kernel");
}

// ---- Langevin dispatch ------------------------------------------------------------------------

template <class ElemE>
static int launch_langevin_elem(const ElemE& en, const LangevinCall& c) {
  const DeviceInfo& di = device_info(current_device());
  const long long numel = (long long)c.n * c.e->dim;
  LangevinElemParams P;
  memset(&P, 0, sizeof(P));
  P.numel = numel;
  P.d = c.e->dim;
  P.thin = c.thin;
  P.n_kept = c.n_steps / c.thin;
  P.has_clamp = c.clamp != nullptr;
  if (c.clamp) { P.clamp_lo = c.clamp[0]; P.clamp_hi = c.clamp[1]; }
  P.traj = c.traj;
  P.diag_ws = c.diag_ws;
  const bool keep = c.traj || c.diag_ws;   // the keeping instantiations: trajectory and / or in-burst diagnostics
  if (c.scheme == 1 && !c.clamp) { P.clamp_lo = -INFINITY; P.clamp_hi = INFINITY; }   // the Heun kernel always clamps
  if (c.rng_mode == EBM_RNG_TORCH) {
    P.T = torch_threads(di, numel);
    const unsigned long long J = (((unsigned long long)numel + P.T - 1) / P.T + 3) / 4;
    P.n_quads = P.T * J;
    philox_expand_keys(P.keys, (uint32_t)c.seed, (uint32_t)(c.seed >> 32));
    P.ctr_step = torch_offset_increment(di, numel) / 4;
  } else {
    P.T = 1;
    P.n_quads = ((unsigned long long)numel + 3) / 4;
    philox_expand_keys(P.keys, (uint32_t)c.seed ^ kNativeTag0, (uint32_t)(c.seed >> 32) ^ kNativeTag1);
    P.ctr_step = 1;
  }
  P.quad_base = 0;
  P.quad_end = P.n_quads;
  if (c.quad_end > c.quad_begin) {
    P.quad_base = c.quad_begin;
    P.quad_end = c.quad_end < P.n_quads ? c.quad_end : P.n_quads;
    if (P.quad_base >= P.quad_end) return 0;
  }
  const unsigned long long blocks = (P.quad_end - P.quad_base + 255) / 256;
  if (blocks > 0x7fffffffull) { set_error("too many elements"); return EBM_ERR_UNSUPPORTED; }

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
    P.n_peers = 0;
    if (c.n_peers > 0 && done + chunk == c.n_steps) {  // the last launch of the burst also feeds the gathered buffers
      P.n_peers = c.n_peers;
      P.peer_mc = c.peer_mc;
      P.peer_off = c.peer_row_offset * c.e->dim;
      for (int w = 0; w < c.n_peers; ++w) P.peers[w] = c.peers[w];
    }
    P.noise = c.noise ? c.noise + (long long)done * numel : nullptr;
    P.ctr_base = c.offset / 4 + (unsigned long long)done * P.ctr_step;
    P.thin_start = c.thin - (done % c.thin);
    P.kept_base = done / c.thin;
#define LAUNCH(RNG)                                                                                          \
  if (keep) {                                                                                                \
    if (c.clamp) langevin_elem_kernel<ElemE, RNG, true, true><<<(unsigned)blocks, 256, 0, c.st>>>(P, en, tab);   \
    else         langevin_elem_kernel<ElemE, RNG, true, false><<<(unsigned)blocks, 256, 0, c.st>>>(P, en, tab);  \
  } else {                                                                                                   \
    if (c.clamp) langevin_elem_kernel<ElemE, RNG, false, true><<<(unsigned)blocks, 256, 0, c.st>>>(P, en, tab);  \
    else         langevin_elem_kernel<ElemE, RNG, false, false><<<(unsigned)blocks, 256, 0, c.st>>>(P, en, tab); \
  }
#define LAUNCH_HEUN(RNG)                                                                                     \
  if (keep)   langevin_elem_kernel<ElemE, RNG, true, true, true><<<(unsigned)blocks, 256, 0, c.st>>>(P, en, tab);     \
  else        langevin_elem_kernel<ElemE, RNG, false, true, true><<<(unsigned)blocks, 256, 0, c.st>>>(P, en, tab);
    if (c.scheme == 1) {
      if (c.rng_mode == EBM_RNG_INJECTED) { LAUNCH_HEUN(0) }
      else if (c.rng_mode == EBM_RNG_TORCH) { LAUNCH_HEUN(1) }
      else { LAUNCH_HEUN(2) }
    }
    else if (c.rng_mode == EBM_RNG_INJECTED) { LAUNCH(0) }
    else if (c.rng_mode == EBM_RNG_TORCH) { LAUNCH(1) }
    else { LAUNCH(2) }
#undef LAUNCH_HEUN
#undef LAUNCH
    int rc = launch_status("langevin_elem_kernel");
    if (rc) return rc;
    done += chunk;
    src = c.x_out;
  }
  return 0;
}

template <class RowE>
static int launch_langevin_row(const RowE& en, const LangevinCall& c) {
  const DeviceInfo& di = device_info(current_device());
  const long long numel = (long long)c.n * c.e->dim;
  LangevinRowParams P;
  memset(&P, 0, sizeof(P));
  P.n = c.n;
  P.d = c.e->dim;
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
#define CALL(G, E)                                                                                  \
  {                                                                                                 \
    int ss;                                                                                         \
    const size_t smem = row_smem_bytes(c.e, G, ss);                                                 \
    if (smem > (size_t)di.max_smem_optin) { set_error("energy parameters need %zu B of shared memory", smem); return EBM_ERR_UNSUPPORTED; } \
    P.scratch_stride = ss;                                                                          \
    auto kern = langevin_row_kernel<RowE, G, E>;                                                    \
    if (smem > 48 * 1024) EBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<row_grid(di, c.n, G, 4), kRowThreads, smem, c.st>>>(en, P, tab);                         \
  }
    EBM_ROW_DISPATCH(c.e->dim, CALL);
#undef CALL
    int rc = launch_status("langevin_row_kernel");
    if (rc) return rc;
    done += chunk;
    src = c.x_out;
  }
  return 0;
}

int langevin_mlp_dispatch(const LangevinCall& c);  // ebm_mlp.cu

static int langevin_dispatch(const LangevinCall& c) {
  if (c.scheme == 1 && c.e->kind != EBM_ENERGY_DOUBLE_WELL && c.e->kind != EBM_ENERGY_HARMONIC &&
      c.e->kind != EBM_ENERGY_RASTRIGIN) {
    set_error("the Heun burst is fused for the elementwise energies only (kind %d)", c.e->kind);
    return EBM_ERR_UNSUPPORTED;
  }
  switch (c.e->kind) {
    case EBM_ENERGY_DOUBLE_WELL: return launch_langevin_elem(make_dw(c.e), c);
    case EBM_ENERGY_HARMONIC: return launch_langevin_elem(make_harm(c.e), c);
    case EBM_ENERGY_RASTRIGIN: return launch_langevin_elem(make_rast(c.e), c);
    case EBM_ENERGY_GAUSSIAN: return launch_langevin_row(make_gauss(c.e), c);
    case EBM_ENERGY_MOG: return launch_langevin_row(make_mog(c.e), c);
    case EBM_ENERGY_MLP: return langevin_mlp_dispatch(c);
  }
  return EBM_ERR_INVALID;
}

template <class ElemE>
static int launch_descent_elem(const ElemE& en, const EbmEnergyDesc* e, const float* x_in, float* x_out, int64_t n,
                               int32_t n_steps, const double* hs, int32_t schedule_len, double momentum, float* velocity,
                               float* traj, int32_t thin, cudaStream_t st) {
  DescentElemParams P;
  memset(&P, 0, sizeof(P));
  P.numel = (long long)n * e->dim;
  P.d = e->dim;
  P.thin = thin;
  P.n_kept = n_steps / thin;
  P.traj = traj;
  P.nesterov = momentum >= 0.0 ? 1 : 0;
  P.mu = (float)(momentum >= 0.0 ? momentum : 0.0);
  const unsigned long long blocks = ((unsigned long long)(P.numel + 3) / 4 + 255) / 256;
  if (blocks > 0x7fffffffull) { set_error("too many elements"); return EBM_ERR_UNSUPPORTED; }
  const bool uniform = schedule_len == 1;
  int done = 0;
  const float* src = x_in;
  while (done < n_steps) {
    const int chunk = uniform ? n_steps : ((n_steps - done < kSchedChunk) ? (n_steps - done) : kSchedChunk);
    StepTable tab;
    memset(&tab, 0, sizeof(tab));
    if (uniform) { tab.h[0] = (float)hs[0]; tab.mask = 0; }
    else { for (int i = 0; i < chunk; ++i) tab.h[i] = (float)hs[done + i]; tab.mask = ~0; }
    P.x_in = src;
    P.x_out = x_out;
    P.n_steps = chunk;
    P.thin_start = thin - (done % thin);
    P.kept_base = done / thin;
    const bool more = done + chunk < n_steps;
    P.v_in = done > 0 ? velocity : nullptr;            // the burst starts from v = 0 (gradient_descent.py:238)
    P.v_out = (more || velocity) ? velocity : nullptr;
    if (P.nesterov && more && !velocity) { set_error("a scheduled Nesterov burst longer than %d steps needs a velocity buffer", kSchedChunk); return EBM_ERR_INVALID; }
    if (P.nesterov) {
      if (traj) descent_elem_kernel<ElemE, true, true><<<(unsigned)blocks, 256, 0, st>>>(P, en, tab);
      else      descent_elem_kernel<ElemE, true, false><<<(unsigned)blocks, 256, 0, st>>>(P, en, tab);
    } else {
      if (traj) descent_elem_kernel<ElemE, false, true><<<(unsigned)blocks, 256, 0, st>>>(P, en, tab);
      else      descent_elem_kernel<ElemE, false, false><<<(unsigned)blocks, 256, 0, st>>>(P, en, tab);
    }
    int rc = launch_status("descent_elem_kernel");
    if (rc) return rc;
    done += chunk;
    src = x_out;
  }
  return 0;
}

int mlp_energy_grad_dispatch(const EbmEnergyDesc* e, const float* x, int64_t n, float* energy, float* grad,
                             cudaStream_t st);  // ebm_mlp.cu

// SM-driven gather push: every thread streams 16-byte pieces of this rank's shard into all gathered buffers (local
// and peer-mapped).  Launched with a handful of CTAs next to a persistent burst that leaves as many SMs free
// (EbmEnergyDesc.sm_margin): NVLink takes stores from SMs at several times the rate of one copy engine.
struct PeerPushParams {
  const float4* src;
  long long n_vec;       // float4 count (tail handled by the last thread below)
  const float* src_tail;
  int n_tail;            // 0..3 trailing floats
  long long dst_off;     // element offset of this rank's shard inside every gathered buffer
  int world;
  float* peers[kMaxPeers];
};
__global__ void __launch_bounds__(1024) peer_push_kernel(const __grid_constant__ PeerPushParams P) {
  // 8 independent 16-byte loads in flight per thread (a few CTAs must cover the HBM latency on their own), then the
  // stores of each piece to every destination
  constexpr int U = 8;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long base = (long long)blockIdx.x * blockDim.x + threadIdx.x; base < P.n_vec; base += stride * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = base + u * stride;
      if (i < P.n_vec) v[u] = __ldcs(P.src + i);
    }
    for (int w = 0; w < P.world; ++w) {
      float4* dst = reinterpret_cast<float4*>(P.peers[w] + P.dst_off);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long i = base + u * stride;
        if (i < P.n_vec) dst[i] = v[u];
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < P.n_tail) {
    const float v = P.src_tail[threadIdx.x];
    for (int w = 0; w < P.world; ++w) P.peers[w][P.dst_off + 4 * P.n_vec + threadIdx.x] = v;
  }
}

// three non-blocking streams per device for the host-buffer entry point (module cache, created on first use);
// NULL when they cannot be created -- the caller then runs the single-stream path
static cudaStream_t* host_pipe_streams(int device) {
  static std::mutex mu;
  static cudaStream_t streams[64][3];
  static int state[64];  // 0 = not tried, 1 = ready, -1 = failed
  if (device < 0 || device >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (state[device] == 0) {
    state[device] = 1;
    for (int i = 0; i < 3; ++i)
      if (cudaStreamCreateWithFlags(&streams[device][i], cudaStreamNonBlocking) != cudaSuccess) { state[device] = -1; break; }
  }
  return state[device] == 1 ? streams[device] : nullptr;
}
size_t mlp_wide_workspace_bytes(const EbmEnergyDesc* e);  // ebm_mlp_wide.cu

// ---- diagnostics: column statistics of a state tensor, finalisation (diag.cuh) ------------------------------------
// CTA = 256 threads = 256 / cw row lanes x cw columns (cw = min(d, 256) rounded to a power of two would waste lanes;
// simply thread t <-> column (t % cw), row lane (t / cw)); a CTA walks its slab of rows, keeps fp64 partial sums in
// registers and issues one atomic per (column, row lane) at the end.
__global__ void __launch_bounds__(256) diag_accumulate_kernel(double* __restrict__ slot, const float* __restrict__ x,
                                                              const float* __restrict__ energy, long long n, int d,
                                                              long long rows_per_cta) {
  const int cw = d < 256 ? d : 256;
  const int lanes = 256 / cw;               // row lanes of this CTA
  const int col0 = threadIdx.x % cw, rl = threadIdx.x / cw;
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = r0 + rows_per_cta < n ? r0 + rows_per_cta : n;
  if (rl < lanes) {
    for (int c = col0; c < d; c += cw) {
      double s = 0.0, q = 0.0;
      for (long long r = r0 + rl; r < r1; r += lanes) {
        const double v = (double)x[r * d + c];
        s += v;
        q += v * v;
      }
      atomicAdd(slot + c, s);
      atomicAdd(slot + d + c, q);
    }
  }
  if (energy) {
    double es = 0.0;
    for (long long r = r0 + threadIdx.x; r < r1; r += 256) es += (double)energy[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) es += __shfl_xor_sync(0xffffffffu, es, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(slot + 2 * d, es);
  }
}

__global__ void diag_finalize_kernel(const double* __restrict__ ws, int n_kept, int d, double inv_n, long long n,
                                     float e_scale, float e_shift, float* __restrict__ mean, float* __restrict__ var,
                                     float* __restrict__ energy, const int32_t* __restrict__ accept_count, int thin,
                                     float* __restrict__ accept_rate) {
  const long long total = (long long)n_kept * d;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long j = i / d;
    const int c = (int)(i - j * d);
    const double* slot = ws + j * diag_slot(d);
    const double m = slot[c] * inv_n;
    mean[i] = (float)m;
    float v = 0.0f;                                        // one chain: the reference zeroes the variance
    if (n > 1) {
      v = (float)(slot[d + c] * inv_n - m * m);            // x.var(dim=0, unbiased=False)
      v = (v != v) ? v : fminf(fmaxf(v, 1e-10f), 1e10f);   // .clamp_(min=1e-10, max=1e10)
    }
    var[i] = v;
    if (c == 0) {
      energy[j] = __fadd_rn(__fmul_rn(e_scale, (float)(slot[2 * d] * inv_n)), e_shift);
      if (accept_rate) accept_rate[j] = (float)accept_count[(j + 1) * thin - 1] / (float)n;
    }
  }
}

int diag_accumulate(double* ws_slot, const float* x, const float* energy, int64_t n, int d, cudaStream_t st) {
  const DeviceInfo& di = device_info(current_device());
  long long ctas = (long long)di.sm_count * 4;
  const long long min_rows = 64;
  if (ctas * min_rows > n) ctas = (n + min_rows - 1) / min_rows;
  const long long rows_per_cta = (n + ctas - 1) / ctas;
  ctas = (n + rows_per_cta - 1) / rows_per_cta;
  diag_accumulate_kernel<<<(unsigned)ctas, 256, 0, st>>>(ws_slot, x, energy, n, d, rows_per_cta);
  return launch_status("diag_accumulate_kernel");
}

int diag_finalize(const double* ws, int n_kept, int d, int64_t n, float e_scale, float e_shift, float* mean, float* var,
                  float* energy, const int32_t* accept_count, int thin, float* accept_rate, cudaStream_t st) {
  const DeviceInfo& di = device_info(current_device());
  diag_finalize_kernel<<<flat_grid(di, (long long)n_kept * d, 256), 256, 0, st>>>(
      ws, n_kept, d, 1.0 / (double)n, n, e_scale, e_shift, mean, var, energy, accept_count, thin, accept_rate);
  return launch_status("diag_finalize_kernel");
}

}  // namespace ebm

using namespace ebm;

extern "C" {

int ebm_abi_version(void) { return EBM_ABI_VERSION; }
const char* ebm_last_error(void) { return g_err; }

int ebm_device_sm_count(int device) { return device_info(device).sm_count; }
int64_t ebm_torch_rng_threads(int device, int64_t numel) { return (int64_t)torch_threads(device_info(device), numel); }
int64_t ebm_torch_rng_offset_increment(int device, int64_t numel) {
  return (int64_t)torch_offset_increment(device_info(device), numel);
}

int ebm_rng_fill_f32(float* out, int64_t numel, int32_t rng_mode, int32_t kind, uint64_t seed, uint64_t offset,
                     void* stream) {
  EBM_CHECK_ARG(out && numel > 0, "out must be non-null and numel positive");
  EBM_CHECK_ARG(rng_mode == EBM_RNG_TORCH || rng_mode == EBM_RNG_NATIVE, "rng_mode must be TORCH or NATIVE");
  EBM_CHECK_ARG(offset % 4 == 0, "offset must be a multiple of 4");
  const DeviceInfo& di = device_info(current_device());
  RngStream s;
  s.mode = rng_mode;
  s.ctr_base = offset / 4;
  if (rng_mode == EBM_RNG_TORCH) { s.T = torch_threads(di, numel); s.k0 = (uint32_t)seed; s.k1 = (uint32_t)(seed >> 32); }
  else { s.T = 1; s.k0 = (uint32_t)seed ^ kNativeTag0; s.k1 = (uint32_t)(seed >> 32) ^ kNativeTag1; }
  rng_fill_kernel<<<flat_grid(di, numel, 256), 256, 0, (cudaStream_t)stream>>>(out, numel, s, kind);
  return launch_status("rng_fill_kernel");
}

int64_t ebm_workspace_bytes(const EbmEnergyDesc* e) {
  if (!e || e->kind != EBM_ENERGY_MLP) return 0;
  if (e->dim <= 128) return kMlpFlagBytes;  // hand-over flags of the balanced tile-step split (optional: buf[6] may be NULL)
  return (int64_t)mlp_wide_workspace_bytes(e);
}

int ebm_energy_f32(const EbmEnergyDesc* e, const float* x, int64_t n, float* energy, void* stream) {
  int rc = validate_desc(e);
  if (rc) return rc;
  EBM_CHECK_ARG(x && energy && n > 0, "x/energy must be non-null and n positive");
  cudaStream_t st = (cudaStream_t)stream;
  switch (e->kind) {
    case EBM_ENERGY_DOUBLE_WELL: return launch_energy_grad(ElemRow<DoubleWellE>{make_dw(e)}, e, x, n, energy, nullptr, st);
    case EBM_ENERGY_HARMONIC: return launch_energy_grad(ElemRow<HarmonicE>{make_harm(e)}, e, x, n, energy, nullptr, st);
    case EBM_ENERGY_RASTRIGIN: return launch_energy_grad(ElemRow<RastriginE>{make_rast(e)}, e, x, n, energy, nullptr, st);
    case EBM_ENERGY_GAUSSIAN: return launch_energy_grad(make_gauss(e), e, x, n, energy, nullptr, st);
    case EBM_ENERGY_MOG: return launch_energy_grad(make_mog(e), e, x, n, energy, nullptr, st);
    case EBM_ENERGY_MLP: return mlp_energy_grad_dispatch(e, x, n, energy, nullptr, st);
  }
  return EBM_ERR_INVALID;
}

int ebm_gradient_f32(const EbmEnergyDesc* e, const float* x, int64_t n, float* grad, void* stream) {
  int rc = validate_desc(e);
  if (rc) return rc;
  EBM_CHECK_ARG(x && grad && n > 0, "x/grad must be non-null and n positive");
  cudaStream_t st = (cudaStream_t)stream;
  switch (e->kind) {
    case EBM_ENERGY_DOUBLE_WELL: return launch_energy_grad(ElemRow<DoubleWellE>{make_dw(e)}, e, x, n, nullptr, grad, st);
    case EBM_ENERGY_HARMONIC: return launch_energy_grad(ElemRow<HarmonicE>{make_harm(e)}, e, x, n, nullptr, grad, st);
    case EBM_ENERGY_RASTRIGIN: return launch_energy_grad(ElemRow<RastriginE>{make_rast(e)}, e, x, n, nullptr, grad, st);
    case EBM_ENERGY_GAUSSIAN: return launch_energy_grad(make_gauss(e), e, x, n, nullptr, grad, st);
    case EBM_ENERGY_MOG: return launch_energy_grad(make_mog(e), e, x, n, nullptr, grad, st);
    case EBM_ENERGY_MLP: return mlp_energy_grad_dispatch(e, x, n, nullptr, grad, st);
  }
  return EBM_ERR_INVALID;
}

int ebm_euler_maruyama_step_f32(const float* x, const float* drift, const float* noise, float* out, int64_t numel,
                                double step_size, double noise_scale, void* stream) {
  EBM_CHECK_ARG(x && drift && out && numel > 0, "x/drift/out must be non-null and numel positive");
  const bool has_noise = noise_scale >= 0.0;
  EBM_CHECK_ARG(!has_noise || noise, "noise must be given when noise_scale >= 0");
  const DeviceInfo& di = device_info(current_device());
  StepTable t;
  fill_step(t, 0, step_size, has_noise ? noise_scale : 0.0);
  em_step_kernel<<<flat_grid(di, numel, 256), 256, 0, (cudaStream_t)stream>>>(x, drift, noise, out, numel, t.h[0],
                                                                              t.c1[0], t.c2[0], has_noise ? 1 : 0);
  return launch_status("em_step_kernel");
}

int ebm_langevin_burst_f32(const EbmEnergyDesc* e, const float* x_in, float* x_out, int64_t n, int32_t n_steps,
                           const double* step_size_host, const double* noise_scale_host, int32_t schedule_len,
                           const float* clamp_lo_hi_host, int32_t rng_mode, uint64_t seed, uint64_t offset,
                           const float* noise, float* traj, int32_t thin, void* stream) {
  int rc = validate_desc(e);
  if (rc) return rc;
  EBM_CHECK_ARG(x_in && x_out && n > 0, "x_in/x_out must be non-null and n positive");
  EBM_CHECK_ARG(n_steps > 0, "n_steps must be positive");
  EBM_CHECK_ARG(step_size_host && noise_scale_host, "schedules must be non-null");
  EBM_CHECK_ARG(schedule_len == 1 || schedule_len == n_steps, "schedule_len must be 1 or n_steps");
  EBM_CHECK_ARG(thin >= 1, "thin must be >= 1");
  EBM_CHECK_ARG(rng_mode >= EBM_RNG_INJECTED && rng_mode <= EBM_RNG_NATIVE, "bad rng_mode");
  EBM_CHECK_ARG(rng_mode != EBM_RNG_INJECTED || noise, "INJECTED rng needs a noise array");
  EBM_CHECK_ARG(offset % 4 == 0, "offset must be a multiple of 4");
  LangevinCall c{e, x_in, x_out, n, n_steps, step_size_host, noise_scale_host, schedule_len, clamp_lo_hi_host,
                 rng_mode, seed, offset, noise, traj, thin, (cudaStream_t)stream, nullptr, nullptr, 0, 0, nullptr, 0, 0, 0};
  return langevin_dispatch(c);
}

}  // extern "C"

// burst kernels with a peer-store epilogue: the elementwise kernel and the two tensor-core MLP kernels
static bool burst_stores_to_peers(const EbmEnergyDesc* e) {
  return e->kind == EBM_ENERGY_DOUBLE_WELL || e->kind == EBM_ENERGY_HARMONIC || e->kind == EBM_ENERGY_RASTRIGIN ||
         (e->kind == EBM_ENERGY_MLP && e->hidden3 == 0 && (e->precision == EBM_MLP_BF16X3 || e->precision == EBM_MLP_BF16));
}

// every other burst kernel: copy-engine pushes of the finished shard into every gathered buffer
static int push_to_peers(const float* src, size_t numel, float* const* peers, int world, long long elem_off, cudaStream_t st) {
  for (int w = 0; w < world; ++w)
    EBM_CUDA(cudaMemcpyAsync(peers[w] + elem_off, src, numel * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

extern "C" {

int ebm_langevin_burst_gather_f32(const EbmEnergyDesc* e, const float* x_in, float* x_out, int64_t n, int32_t n_steps,
                                  const double* step_size_host, const double* noise_scale_host, int32_t schedule_len,
                                  const float* clamp_lo_hi_host, int32_t rng_mode, uint64_t seed, uint64_t offset,
                                  float* const* peer_out_host, int32_t world, int64_t row_offset, void* stream) {
  int rc = validate_desc(e);
  if (rc) return rc;
  EBM_CHECK_ARG(x_in && x_out && n > 0, "x_in/x_out must be non-null and n positive");
  EBM_CHECK_ARG(n_steps > 0, "n_steps must be positive");
  EBM_CHECK_ARG(step_size_host && noise_scale_host, "schedules must be non-null");
  EBM_CHECK_ARG(schedule_len == 1 || schedule_len == n_steps, "schedule_len must be 1 or n_steps");
  EBM_CHECK_ARG(rng_mode == EBM_RNG_TORCH || rng_mode == EBM_RNG_NATIVE, "the gathering burst draws its own noise");
  EBM_CHECK_ARG(offset % 4 == 0, "offset must be a multiple of 4");
  const bool multicast = world < 0;   // -W: W unicast pointers followed by the NVLS multicast pointer
  if (multicast) world = -world;
  EBM_CHECK_ARG(peer_out_host && world >= 1 && world <= kMaxPeers, "peer_out_host must hold 1..16 pointers");
  EBM_CHECK_ARG(row_offset >= 0, "row_offset must be non-negative");
  for (int w = 0; w < world + (multicast ? 1 : 0); ++w) EBM_CHECK_ARG(peer_out_host[w], "null peer pointer");
  const bool fused = burst_stores_to_peers(e);
  LangevinCall c{e, x_in, x_out, n, n_steps, step_size_host, noise_scale_host, schedule_len, clamp_lo_hi_host,
                 rng_mode, seed, offset, nullptr, nullptr, 1, (cudaStream_t)stream, nullptr, nullptr, 0, 0,
                 fused ? peer_out_host : nullptr, fused ? world : 0, row_offset, 0};
  // multicast only where it wins (measured at 8 GPUs): the wide MLP kernel's coalesced pusher.  The elementwise kernel's
  // final store is already one coalesced store per peer, and the native-stream C2 run was 3 % faster with those.
  if (fused && multicast && e->kind == EBM_ENERGY_MLP && e->dim > 128 && !wide_push_bulk()) { c.peers = peer_out_host + world; c.n_peers = 1; c.peer_mc = 1; }
  rc = langevin_dispatch(c);
  if (rc || fused) return rc;
  return push_to_peers(x_out, (size_t)n * e->dim, peer_out_host, world, row_offset * e->dim, (cudaStream_t)stream);
}

int ebm_langevin_heun_burst_f32(const EbmEnergyDesc* e, const float* x_in, float* x_out, int64_t n, int32_t n_steps,
                                const double* step_size_host, const double* noise_scale_host, int32_t schedule_len,
                                const float* clamp_lo_hi_host, int32_t rng_mode, uint64_t seed, uint64_t offset,
                                const float* noise, float* traj, int32_t thin, void* stream) {
  int rc = validate_desc(e);
  if (rc) return rc;
  EBM_CHECK_ARG(x_in && x_out && n > 0, "x_in/x_out must be non-null and n positive");
  EBM_CHECK_ARG(n_steps > 0, "n_steps must be positive");
  EBM_CHECK_ARG(step_size_host && noise_scale_host, "schedules must be non-null");
  EBM_CHECK_ARG(schedule_len == 1 || schedule_len == n_steps, "schedule_len must be 1 or n_steps");
  EBM_CHECK_ARG(thin >= 1, "thin must be >= 1");
  EBM_CHECK_ARG(rng_mode >= EBM_RNG_INJECTED && rng_mode <= EBM_RNG_NATIVE, "bad rng_mode");
  EBM_CHECK_ARG(rng_mode != EBM_RNG_INJECTED || noise, "INJECTED rng needs a noise array");
  EBM_CHECK_ARG(offset % 4 == 0, "offset must be a multiple of 4");
  LangevinCall c{e, x_in, x_out, n, n_steps, step_size_host, noise_scale_host, schedule_len, clamp_lo_hi_host,
                 rng_mode, seed, offset, noise, traj, thin, (cudaStream_t)stream, nullptr, nullptr, 0, 0, nullptr, 0, 0, 1};
  return langevin_dispatch(c);
}

int ebm_peer_push_f32(const float* src, int64_t numel, float* const* peer_out_host, int32_t world, int64_t elem_offset,
                      int32_t max_ctas, void* stream) {
  EBM_CHECK_ARG(src && numel > 0 && peer_out_host, "src/peer_out_host must be non-null and numel positive");
  EBM_CHECK_ARG(world >= 1 && world <= kMaxPeers, "world must be 1..16");
  EBM_CHECK_ARG(elem_offset >= 0 && max_ctas >= 1, "elem_offset must be non-negative and max_ctas positive");
  PeerPushParams P;
  memset(&P, 0, sizeof(P));
  bool aligned = ((uintptr_t)src & 15) == 0 && (elem_offset % 4) == 0;
  for (int w = 0; w < world; ++w) {
    EBM_CHECK_ARG(peer_out_host[w], "null peer pointer");
    aligned = aligned && ((uintptr_t)peer_out_host[w] & 15) == 0;
    P.peers[w] = peer_out_host[w];
  }
  EBM_CHECK_ARG(aligned, "src, the gathered buffers and elem_offset must be 16-byte aligned");
  P.src = reinterpret_cast<const float4*>(src);
  P.n_vec = numel / 4;
  P.src_tail = src + 4 * P.n_vec;
  P.n_tail = (int)(numel - 4 * P.n_vec);
  P.dst_off = elem_offset;
  P.world = world;
  peer_push_kernel<<<max_ctas, 1024, 0, (cudaStream_t)stream>>>(P);
  return launch_status("peer_push_kernel");
}

int ebm_pcd_langevin_fused(const EbmEnergyDesc* e) {
  // the tensor-core MLP kernels read their start rows through an index and can write the final state twice
  return e && e->kind == EBM_ENERGY_MLP && (e->precision == EBM_MLP_BF16X3 || e->precision == EBM_MLP_BF16) ? 1 : 0;
}

}  // extern "C"

// the persistent-CD burst; peers != NULL: the negatives also land in every rank's gathered buffer (from inside the
// burst kernel's final store when it has a peer-store epilogue, copy-engine pushes otherwise)
static int pcd_langevin_burst_impl(const EbmEnergyDesc* e, float* buffer, int64_t buffer_rows, const int64_t* idx,
                                   int64_t ptr, float* x_out, float* scratch, int64_t n, int32_t n_steps,
                                   const double* step_size_host, const double* noise_scale_host, int32_t schedule_len,
                                   const float* clamp_lo_hi_host, int32_t rng_mode, uint64_t seed, uint64_t offset,
                                   const int64_t* noise_rows, const float* noise, int64_t n_noise, float* energy_out,
                                   int64_t* new_ptr_host, float* const* peers, int32_t world, int64_t row_offset,
                                   void* stream) {
  int rc = validate_desc(e);
  if (rc) return rc;
  float* const* mc_ptr = nullptr;   // world = -W: W unicast pointers followed by the NVLS multicast pointer
  if (peers) {
    if (world < 0) { world = -world; mc_ptr = peers + world; }
    EBM_CHECK_ARG(world >= 1 && world <= kMaxPeers, "peer_out_host must hold 1..16 pointers");
    EBM_CHECK_ARG(row_offset >= 0, "row_offset must be non-negative");
    for (int w = 0; w < world + (mc_ptr ? 1 : 0); ++w) EBM_CHECK_ARG(peers[w], "null peer pointer");
  } else {
    world = 0;
  }
  EBM_CHECK_ARG(buffer && x_out && buffer_rows > 0 && n > 0, "buffer/x_out must be non-null, sizes positive");
  EBM_CHECK_ARG(idx || n <= buffer_rows, "idx == NULL (chain i starts from row i) needs n <= buffer_rows");
  EBM_CHECK_ARG(ptr >= 0 && ptr < buffer_rows, "ptr out of range");
  EBM_CHECK_ARG(n_steps > 0, "n_steps must be positive");
  EBM_CHECK_ARG(step_size_host && noise_scale_host, "schedules must be non-null");
  EBM_CHECK_ARG(schedule_len == 1 || schedule_len == n_steps, "schedule_len must be 1 or n_steps");
  EBM_CHECK_ARG(rng_mode == EBM_RNG_TORCH || rng_mode == EBM_RNG_NATIVE, "the fused PCD burst draws its own noise");
  EBM_CHECK_ARG(offset % 4 == 0, "offset must be a multiple of 4");
  EBM_CHECK_ARG(n_noise == 0 || (noise_rows && noise), "noise_rows/noise must be given when n_noise > 0");
  const int64_t row_elems = e->dim;
  // The one-call form needs (a) a burst kernel that can write its final state to two destinations and (b) chain i
  // reading and writing row i of the buffer, which the caller states by passing idx == NULL (the reference's
  // stratified draw with stride 1, core/base_loss.py:307-312, when the batch is the whole buffer the write-back
  // replaces every row with ptr = 0, :409-413).  Exploration noise (:317-332) is then added in place to the few
  // noised rows first: every row is overwritten by the write-back anyway.  An explicit idx -- any permutation or
  // draw with replacement -- always goes gather -> burst -> FIFO scatter.
  const bool one_call = ebm_pcd_langevin_fused(e) != 0 && idx == nullptr && n == buffer_rows;
  if (one_call) {
    if (n_noise > 0) {
      const DeviceInfo& di = device_info(current_device());
      pcd_noise_kernel<<<flat_grid(di, n_noise * row_elems, 256), 256, 0, (cudaStream_t)stream>>>(
          buffer, row_elems, (const long long*)noise_rows, noise, n_noise);
      rc = launch_status("pcd_noise_kernel");
      if (rc) return rc;
    }
    LangevinCall c{e, buffer, x_out, n, n_steps, step_size_host, noise_scale_host, schedule_len, clamp_lo_hi_host,
                   rng_mode, seed, offset, nullptr, nullptr, 1, (cudaStream_t)stream, nullptr, buffer, 0, 0, peers, world,
                   row_offset, 0, nullptr};
    if (mc_ptr && e->dim > 128 && !wide_push_bulk()) { c.peers = mc_ptr; c.n_peers = 1; c.peer_mc = 1; }   // (the wide kernel's 16-byte-store pusher only)
    rc = langevin_dispatch(c);
    if (rc) return rc;
    if (world > 0 && !burst_stores_to_peers(e)) {   // (the three-hidden-layer kernel has no peer-store epilogue)
      rc = push_to_peers(x_out, (size_t)n * row_elems, peers, world, row_offset * row_elems, (cudaStream_t)stream);
      if (rc) return rc;
    }
    if (new_ptr_host) *new_ptr_host = 0;
  } else if (ebm_pcd_langevin_fused(e) != 0 && idx != nullptr && n_noise == 0) {
    // the burst kernel reads its start rows through idx; the FIFO write-back stays a separate launch (a destination
    // row may be another chain's source row)
    LangevinCall c{e, buffer, x_out, n, n_steps, step_size_host, noise_scale_host, schedule_len, clamp_lo_hi_host,
                   rng_mode, seed, offset, nullptr, nullptr, 1, (cudaStream_t)stream, (const long long*)idx, nullptr, 0, 0,
                   peers, world, row_offset, 0, nullptr};
    if (mc_ptr && e->dim > 128 && !wide_push_bulk()) { c.peers = mc_ptr; c.n_peers = 1; c.peer_mc = 1; }   // (the wide kernel's 16-byte-store pusher only)
    rc = langevin_dispatch(c);
    if (rc) return rc;
    if (world > 0 && !burst_stores_to_peers(e)) {   // (the three-hidden-layer kernel has no peer-store epilogue)
      rc = push_to_peers(x_out, (size_t)n * row_elems, peers, world, row_offset * row_elems, (cudaStream_t)stream);
      if (rc) return rc;
    }
    rc = ebm_pcd_scatter_f32(buffer, buffer_rows, row_elems, ptr, x_out, n, new_ptr_host, stream);
    if (rc) return rc;
  } else {   // gather (+ noise) -> burst -> FIFO scatter as separate launches
    EBM_CHECK_ARG(scratch, "scratch [n, dim] is required unless the burst kernel reads the buffer itself");
    if (idx) {
      rc = ebm_pcd_gather_f32(buffer, buffer_rows, row_elems, idx, n, scratch, noise_rows, noise, n_noise, stream);
    } else {
      EBM_CUDA(cudaMemcpyAsync(scratch, buffer, (size_t)n * row_elems * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
      if (n_noise > 0) {
        const DeviceInfo& di = device_info(current_device());
        pcd_noise_kernel<<<flat_grid(di, n_noise * row_elems, 256), 256, 0, (cudaStream_t)stream>>>(
            scratch, row_elems, (const long long*)noise_rows, noise, n_noise);
        rc = launch_status("pcd_noise_kernel");
      }
    }
    if (rc) return rc;
    rc = ebm_langevin_burst_f32(e, scratch, x_out, n, n_steps, step_size_host, noise_scale_host, schedule_len,
                                clamp_lo_hi_host, rng_mode, seed, offset, nullptr, nullptr, 1, stream);
    if (rc) return rc;
    rc = ebm_pcd_scatter_f32(buffer, buffer_rows, row_elems, ptr, x_out, n, new_ptr_host, stream);
    if (rc) return rc;
    if (world > 0) {
      rc = push_to_peers(x_out, (size_t)n * row_elems, peers, world, row_offset * row_elems, (cudaStream_t)stream);
      if (rc) return rc;
    }
  }
  if (energy_out) return ebm_energy_f32(e, x_out, n, energy_out, stream);   // E(x-) of the negatives
  return 0;
}

extern "C" {

int ebm_pcd_langevin_burst_f32(const EbmEnergyDesc* e, float* buffer, int64_t buffer_rows, const int64_t* idx,
                               int64_t ptr, float* x_out, float* scratch, int64_t n, int32_t n_steps,
                               const double* step_size_host, const double* noise_scale_host, int32_t schedule_len,
                               const float* clamp_lo_hi_host, int32_t rng_mode, uint64_t seed, uint64_t offset,
                               const int64_t* noise_rows, const float* noise, int64_t n_noise, float* energy_out,
                               int64_t* new_ptr_host, void* stream) {
  return pcd_langevin_burst_impl(e, buffer, buffer_rows, idx, ptr, x_out, scratch, n, n_steps, step_size_host,
                                 noise_scale_host, schedule_len, clamp_lo_hi_host, rng_mode, seed, offset, noise_rows, noise,
                                 n_noise, energy_out, new_ptr_host, nullptr, 0, 0, stream);
}

int ebm_pcd_langevin_burst_gather_f32(const EbmEnergyDesc* e, float* buffer, int64_t buffer_rows, const int64_t* idx,
                                      int64_t ptr, float* x_out, float* scratch, int64_t n, int32_t n_steps,
                                      const double* step_size_host, const double* noise_scale_host, int32_t schedule_len,
                                      const float* clamp_lo_hi_host, int32_t rng_mode, uint64_t seed, uint64_t offset,
                                      const int64_t* noise_rows, const float* noise, int64_t n_noise, float* energy_out,
                                      int64_t* new_ptr_host, float* const* peer_out_host, int32_t world,
                                      int64_t row_offset, void* stream) {
  EBM_CHECK_ARG(peer_out_host, "peer_out_host must be non-null");
  return pcd_langevin_burst_impl(e, buffer, buffer_rows, idx, ptr, x_out, scratch, n, n_steps, step_size_host,
                                 noise_scale_host, schedule_len, clamp_lo_hi_host, rng_mode, seed, offset, noise_rows, noise,
                                 n_noise, energy_out, new_ptr_host, peer_out_host, world, row_offset, stream);
}

int ebm_langevin_burst_host_f32(const EbmEnergyDesc* e, const float* x_in_host, float* x_out_host, float* scratch_dev,
                                int64_t n, int32_t n_steps, double step_size, double noise_scale, int32_t rng_mode,
                                uint64_t seed, uint64_t offset, void* stream) {
  int rc = validate_desc(e);
  if (rc) return rc;
  EBM_CHECK_ARG(x_in_host && x_out_host && scratch_dev && n > 0, "buffers must be non-null and n positive");
  EBM_CHECK_ARG(rng_mode == EBM_RNG_TORCH || rng_mode == EBM_RNG_NATIVE, "host entry point draws its own noise");
  EBM_CHECK_ARG(n_steps > 0, "n_steps must be positive");
  EBM_CHECK_ARG(offset % 4 == 0, "offset must be a multiple of 4");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t bytes = (size_t)n * e->dim * sizeof(float);
  // Elementwise energies: the burst is launched one resident wave of owning threads at a time (in both RNG layouts
  // a wave owns a contiguous block of 4 * wave elements), each wave on one of three internal streams between the
  // upload of its block and the download of its result, so the PCIe copies of block c+1 / c-1 run under the kernel
  // of block c.  The Philox addressing is per element, so the result is identical to the single-launch burst.
  if (e->kind == EBM_ENERGY_DOUBLE_WELL || e->kind == EBM_ENERGY_HARMONIC || e->kind == EBM_ENERGY_RASTRIGIN) {
    const int dev = current_device();
    const DeviceInfo& di = device_info(dev);
    const unsigned long long numel = (unsigned long long)n * e->dim;
    const unsigned long long wave = 256ull * di.sm_count * (di.max_threads_per_sm / 256);
    const bool layout_ok = rng_mode == EBM_RNG_NATIVE || torch_threads(di, (int64_t)numel) == wave;
    const unsigned long long n_chunks = (numel + 4 * wave - 1) / (4 * wave);
    cudaStream_t* pipe = layout_ok && n_chunks >= 2 ? host_pipe_streams(dev) : nullptr;
    if (pipe) {
      EBM_CUDA(cudaStreamSynchronize(st));
      for (unsigned long long ck = 0; ck < n_chunks; ++ck) {
        cudaStream_t s = pipe[ck % 3];
        const unsigned long long e0 = 4 * wave * ck;
        const unsigned long long e1 = (e0 + 4 * wave < numel) ? e0 + 4 * wave : numel;
        EBM_CUDA(cudaMemcpyAsync(scratch_dev + e0, x_in_host + e0, (e1 - e0) * sizeof(float), cudaMemcpyHostToDevice, s));
        LangevinCall c{e, scratch_dev, scratch_dev, n, n_steps, &step_size, &noise_scale, 1, nullptr, rng_mode, seed, offset,
                       nullptr, nullptr, 1, s, nullptr, nullptr, wave * ck, wave * (ck + 1), nullptr, 0, 0, 0};
        rc = langevin_dispatch(c);
        if (rc) return rc;
        EBM_CUDA(cudaMemcpyAsync(x_out_host + e0, scratch_dev + e0, (e1 - e0) * sizeof(float), cudaMemcpyDeviceToHost, s));
      }
      for (int i = 0; i < 3; ++i) EBM_CUDA(cudaStreamSynchronize(pipe[i]));
      return 0;
    }
  }
  EBM_CUDA(cudaMemcpyAsync(scratch_dev, x_in_host, bytes, cudaMemcpyHostToDevice, st));
  rc = ebm_langevin_burst_f32(e, scratch_dev, scratch_dev, n, n_steps, &step_size, &noise_scale, 1, nullptr, rng_mode,
                              seed, offset, nullptr, nullptr, 1, stream);
  if (rc) return rc;
  EBM_CUDA(cudaMemcpyAsync(x_out_host, scratch_dev, bytes, cudaMemcpyDeviceToHost, st));
  EBM_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int ebm_descent_burst_f32(const EbmEnergyDesc* e, const float* x_in, float* x_out, int64_t n, int32_t n_steps,
                          const double* step_size_host, int32_t schedule_len, double momentum, float* velocity, float* traj,
                          int32_t thin, void* stream) {
  int rc = validate_desc(e);
  if (rc) return rc;
  EBM_CHECK_ARG(x_in && x_out && n > 0, "x_in/x_out must be non-null and n positive");
  EBM_CHECK_ARG(n_steps > 0, "n_steps must be positive");
  EBM_CHECK_ARG(step_size_host, "step sizes must be non-null");
  EBM_CHECK_ARG(schedule_len == 1 || schedule_len == n_steps, "schedule_len must be 1 or n_steps");
  EBM_CHECK_ARG(thin >= 1, "thin must be >= 1");
  EBM_CHECK_ARG(momentum < 1.0, "momentum must be in [0, 1) (negative = plain gradient descent)");
  cudaStream_t st = (cudaStream_t)stream;
  switch (e->kind) {
    case EBM_ENERGY_DOUBLE_WELL: return launch_descent_elem(make_dw(e), e, x_in, x_out, n, n_steps, step_size_host, schedule_len, momentum, velocity, traj, thin, st);
    case EBM_ENERGY_HARMONIC: return launch_descent_elem(make_harm(e), e, x_in, x_out, n, n_steps, step_size_host, schedule_len, momentum, velocity, traj, thin, st);
    case EBM_ENERGY_RASTRIGIN: return launch_descent_elem(make_rast(e), e, x_in, x_out, n, n_steps, step_size_host, schedule_len, momentum, velocity, traj, thin, st);
    default:
      set_error("descent bursts are fused for the elementwise energies only (kind %d)", e->kind);
      return EBM_ERR_UNSUPPORTED;
  }
}

int ebm_pcd_gather_f32(const float* buffer, int64_t buffer_rows, int64_t row_elems, const int64_t* idx, int64_t batch,
                       float* out, const int64_t* noise_rows, const float* noise, int64_t n_noise, void* stream) {
  EBM_CHECK_ARG(buffer && idx && out, "buffer/idx/out must be non-null");
  EBM_CHECK_ARG(buffer_rows > 0 && row_elems > 0 && batch > 0, "sizes must be positive");
  EBM_CHECK_ARG(n_noise == 0 || (noise_rows && noise), "noise_rows/noise must be given when n_noise > 0");
  const DeviceInfo& di = device_info(current_device());
  cudaStream_t st = (cudaStream_t)stream;
  pcd_gather_kernel<<<flat_grid(di, batch * row_elems, 256), 256, 0, st>>>(buffer, row_elems, (const long long*)idx, batch, out);
  int rc = launch_status("pcd_gather_kernel");
  if (rc) return rc;
  if (n_noise > 0) {
    pcd_noise_kernel<<<flat_grid(di, n_noise * row_elems, 256), 256, 0, st>>>(out, row_elems, (const long long*)noise_rows, noise, n_noise);
    rc = launch_status("pcd_noise_kernel");
  }
  return rc;
}

int ebm_pcd_scatter_f32(float* buffer, int64_t buffer_rows, int64_t row_elems, int64_t ptr, const float* samples,
                        int64_t batch, int64_t* new_ptr_host, void* stream) {
  EBM_CHECK_ARG(buffer && samples, "buffer/samples must be non-null");
  EBM_CHECK_ARG(buffer_rows > 0 && row_elems > 0 && batch > 0, "sizes must be positive");
  EBM_CHECK_ARG(ptr >= 0 && ptr < buffer_rows, "ptr out of range");
  const DeviceInfo& di = device_info(current_device());
  cudaStream_t st = (cudaStream_t)stream;
  int64_t new_ptr;
  if (batch >= buffer_rows) {  // base_loss.py:409-413: keep the latest buffer_rows samples, ptr = 0
    const float* src = samples + (batch - buffer_rows) * row_elems;
    pcd_scatter_kernel<<<flat_grid(di, buffer_rows * row_elems, 256), 256, 0, st>>>(buffer, buffer_rows, row_elems, 0, src, buffer_rows);
    new_ptr = 0;
  } else {
    pcd_scatter_kernel<<<flat_grid(di, batch * row_elems, 256), 256, 0, st>>>(buffer, buffer_rows, row_elems, ptr, samples, batch);
    new_ptr = (ptr + batch) % buffer_rows;
  }
  if (new_ptr_host) *new_ptr_host = new_ptr;
  return launch_status("pcd_scatter_kernel");
}

int ebm_langevin_burst_diag_f32(const EbmEnergyDesc* e, const float* x_in, float* x_out, int64_t n, int32_t n_steps,
                                const double* step_size_host, const double* noise_scale_host, int32_t schedule_len,
                                const float* clamp_lo_hi_host, int32_t rng_mode, uint64_t seed, uint64_t offset,
                                const float* noise, float* traj, int32_t thin, int32_t heun, double* diag_ws,
                                float* scratch, float* diag_mean, float* diag_var, float* diag_energy, void* stream) {
  int rc = validate_desc(e);
  if (rc) return rc;
  EBM_CHECK_ARG(x_in && x_out && n > 0, "x_in/x_out must be non-null and n positive");
  EBM_CHECK_ARG(n_steps > 0, "n_steps must be positive");
  EBM_CHECK_ARG(step_size_host && noise_scale_host, "schedules must be non-null");
  EBM_CHECK_ARG(schedule_len == 1 || schedule_len == n_steps, "schedule_len must be 1 or n_steps");
  EBM_CHECK_ARG(thin >= 1 && n_steps / thin >= 1, "thin must be >= 1 and keep at least one sample");
  EBM_CHECK_ARG(rng_mode >= EBM_RNG_INJECTED && rng_mode <= EBM_RNG_NATIVE, "bad rng_mode");
  EBM_CHECK_ARG(rng_mode != EBM_RNG_INJECTED || noise, "INJECTED rng needs a noise array");
  EBM_CHECK_ARG(offset % 4 == 0, "offset must be a multiple of 4");
  EBM_CHECK_ARG(diag_ws && diag_mean && diag_var && diag_energy, "diagnostic outputs and workspace must be non-null");
  cudaStream_t st = (cudaStream_t)stream;
  const int d = e->dim;
  const int n_kept = n_steps / thin;
  EBM_CUDA(cudaMemsetAsync(diag_ws, 0, (size_t)n_kept * diag_slot(d) * sizeof(double), st));
  const bool elem = e->kind == EBM_ENERGY_DOUBLE_WELL || e->kind == EBM_ENERGY_HARMONIC || e->kind == EBM_ENERGY_RASTRIGIN;
  if (elem) {   // one burst launch: the kernel accumulates the statistics of every kept sample itself
    LangevinCall c{e, x_in, x_out, n, n_steps, step_size_host, noise_scale_host, schedule_len, clamp_lo_hi_host,
                   rng_mode, seed, offset, noise, traj, thin, st, nullptr, nullptr, 0, 0, nullptr, 0, 0, heun ? 1 : 0, diag_ws};
    rc = langevin_dispatch(c);
    if (rc) return rc;
    float es = 1.0f, eb = 0.0f;   // energy = finish(sum of terms): h * s, half_k * s, a*D + s
    if (e->kind == EBM_ENERGY_DOUBLE_WELL) es = e->p[0];
    else if (e->kind == EBM_ENERGY_HARMONIC) es = e->p[0];
    else eb = e->p[2];
    return diag_finalize(diag_ws, n_kept, d, n, es, eb, diag_mean, diag_var, diag_energy, nullptr, thin, nullptr, st);
  }
  // other energies: one sub-burst per kept sample, each followed by the energy and column-statistics kernels
  EBM_CHECK_ARG(scratch, "scratch [n] is required for this energy");
  EBM_CHECK_ARG(!heun, "the Heun burst is fused for the elementwise energies only");
  const DeviceInfo& di = device_info(current_device());
  const long long numel = (long long)n * d;
  const uint64_t per_step = rng_mode == EBM_RNG_TORCH ? torch_offset_increment(di, numel) : (rng_mode == EBM_RNG_NATIVE ? 4 : 0);
  const float* src = x_in;
  int done = 0;
  for (int j = 0; j <= n_kept; ++j) {
    const int len = j < n_kept ? thin : n_steps - done;   // the tail after the last kept sample
    if (len <= 0) break;
    LangevinCall c{e, src, x_out, n, len, schedule_len == 1 ? step_size_host : step_size_host + done,
                   schedule_len == 1 ? noise_scale_host : noise_scale_host + done, schedule_len == 1 ? 1 : len,
                   clamp_lo_hi_host, rng_mode, seed, offset + (uint64_t)done * per_step,
                   noise ? noise + (long long)done * numel : nullptr, nullptr, 1, st, nullptr, nullptr, 0, 0, nullptr, 0, 0, 0,
                   nullptr};
    rc = langevin_dispatch(c);
    if (rc) return rc;
    done += len;
    src = x_out;
    if (j < n_kept) {
      rc = ebm_energy_f32(e, x_out, n, scratch, stream);
      if (rc) return rc;
      rc = diag_accumulate(diag_ws + (long long)j * diag_slot(d), x_out, scratch, n, d, st);
      if (rc) return rc;
      if (traj)   // traj[:, j, :] = x
        EBM_CUDA(cudaMemcpy2DAsync(traj + (long long)j * d, (size_t)n_kept * d * sizeof(float), x_out, (size_t)d * sizeof(float),
                                   (size_t)d * sizeof(float), (size_t)n, cudaMemcpyDeviceToDevice, st));
    }
  }
  return diag_finalize(diag_ws, n_kept, d, n, 1.0f, 0.0f, diag_mean, diag_var, diag_energy, nullptr, thin, nullptr, st);
}

}  // extern "C"
