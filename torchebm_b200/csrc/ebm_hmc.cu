// C-ABI: leapfrog integration and the fused HMC proposal loop.  See include/ebm_b200.h.
#include "api_common.cuh"

namespace ebm {

static int make_mass(int32_t mass_kind, double mass_scalar, const float* mass_vec, MassSpec& ms) {
  memset(&ms, 0, sizeof(ms));
  ms.kind = mass_kind;
  ms.safe_scalar = 1.0f;
  ms.scalar = 1.0f;
  ms.sqrt_scalar = 1.0f;
  if (mass_kind == EBM_MASS_SCALAR) {
    ms.safe_scalar = (float)(mass_scalar > 1e-10 ? mass_scalar : 1e-10);  // leapfrog.py:170
    ms.scalar = (float)mass_scalar;                                      // hmc.py:151
    ms.sqrt_scalar = (float)sqrt(mass_scalar);                           // hmc.py:123-124
  } else if (mass_kind == EBM_MASS_VECTOR) {
    if (!mass_vec) { set_error("mass_vec must be given for EBM_MASS_VECTOR"); return EBM_ERR_INVALID; }
    ms.vec = mass_vec;
  } else if (mass_kind != EBM_MASS_NONE) {
    set_error("bad mass_kind %d", mass_kind);
    return EBM_ERR_INVALID;
  }
  return 0;
}

template <class RowE>
static int launch_leapfrog(const RowE& en, const EbmEnergyDesc* e, LeapfrogParams& P, cudaStream_t st) {
  const DeviceInfo& di = device_info(current_device());
#define CALL(G, E)                                                                                  \
  {                                                                                                 \
    int ss;                                                                                         \
    const size_t smem = row_smem_bytes(e, G, ss);                                                   \
    if (smem > (size_t)di.max_smem_optin) { set_error("energy parameters need %zu B of shared memory", smem); return EBM_ERR_UNSUPPORTED; } \
    P.scratch_stride = ss;                                                                          \
    auto kern = leapfrog_kernel<RowE, G, E>;                                                        \
    if (smem > 48 * 1024) EBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<row_grid(di, P.n, G, 4), kRowThreads, smem, st>>>(en, P);                                \
  }
  EBM_ROW_DISPATCH(e->dim, CALL);
#undef CALL
  return launch_status("leapfrog_kernel");
}

// diagnostics of the kept proposals (ebm_hmc_burst_diag_f32): launches end on kept proposals, whose state statistics
// and energies are accumulated right after the launch
struct HmcDiag {
  double* ws;        // [n_kept, diag_slot(d)], zeroed
  float* scratch;    // [n] energies of the state after a kept proposal
};

struct HmcCall {
  const EbmEnergyDesc* e;
  const double* hs;
  int32_t schedule_len;
  int32_t n_proposals;
  uint64_t offset;
  uint64_t inc_p, inc_u;  // generator offset consumed per proposal by the momentum / uniform draws
  cudaStream_t st;
  const HmcDiag* diag;
};

// the proposal loop is launched in chunks of at most kSchedChunk proposals (per-proposal step-size table); `launch`
// runs one chunk
template <class Launch>
static int hmc_chunks(const HmcCall& c, HmcParams& P, Launch launch) {
  const bool uniform = c.schedule_len == 1;
  const long long numel = P.n * P.d;
  const float* noise_p = P.noise_p;
  const float* noise_u = P.noise_u;
  const float* x_src = P.x_in;
  float* energy_out = P.energy_out;
  int done = 0;
  while (done < c.n_proposals) {
    int chunk = uniform ? c.n_proposals : ((c.n_proposals - done < kSchedChunk) ? (c.n_proposals - done) : kSchedChunk);
    if (c.diag) {   // end the launch on the next kept proposal (or on the end of the call behind the last kept one)
      const int to_keep = P.thin - (done % P.thin);
      if (done / P.thin < P.n_kept && to_keep < chunk) chunk = to_keep;
      if (chunk > kSchedChunk) chunk = kSchedChunk;
    }
    HStepTable tab;
    memset(&tab, 0, sizeof(tab));
    if (uniform) { tab.h[0] = (float)c.hs[0]; tab.mask = 0; }
    else { for (int i = 0; i < chunk; ++i) tab.h[i] = (float)c.hs[done + i]; tab.mask = ~0; }
    P.x_in = x_src;
    P.n_prop = chunk;
    P.prop_base = done;
    P.noise_p = noise_p ? noise_p + (long long)done * numel : nullptr;
    P.noise_u = noise_u ? noise_u + (long long)done * P.n : nullptr;
    P.thin_start = P.thin - (done % P.thin);
    P.kept_base = done / P.thin;
    P.energy_out = (done + chunk >= c.n_proposals) ? energy_out : nullptr;
    const bool kept_here = c.diag && ((done + chunk) % P.thin == 0) && ((done + chunk) / P.thin <= P.n_kept);
    if (kept_here) P.energy_out = c.diag->scratch;
    if (P.rng_p.mode == EBM_RNG_TORCH) {
      P.rng_p.ctr_base = (c.offset + (uint64_t)done * (c.inc_p + c.inc_u)) / 4;
      P.rng_u.ctr_base = (c.offset + (uint64_t)done * (c.inc_p + c.inc_u) + c.inc_p) / 4;
    } else {
      P.rng_p.ctr_base = c.offset / 4 + 2ull * done;
      P.rng_u.ctr_base = c.offset / 4 + 2ull * done + 1;
    }
    int rc = launch(P, tab);
    if (rc) return rc;
    done += chunk;
    x_src = P.x_out;
    if (kept_here) {
      rc = diag_accumulate(c.diag->ws + (long long)(done / P.thin - 1) * diag_slot(P.d), P.x_out, c.diag->scratch, P.n, P.d, c.st);
      if (rc) return rc;
      if (energy_out && done >= c.n_proposals)
        if (cudaMemcpyAsync(energy_out, c.diag->scratch, (size_t)P.n * sizeof(float), cudaMemcpyDeviceToDevice, c.st) != cudaSuccess)
          return (int)cudaGetLastError();
    }
  }
  return 0;
}

template <class RowE>
static int launch_hmc(const RowE& en, const HmcCall& c, HmcParams& P) {
  const DeviceInfo& di = device_info(current_device());
  return hmc_chunks(c, P, [&](HmcParams& Pc, const HStepTable& tab) -> int {
#define CALL(G, E)                                                                                  \
  {                                                                                                 \
    int ss;                                                                                         \
    const size_t smem = row_smem_bytes(c.e, G, ss);                                                 \
    if (smem > (size_t)di.max_smem_optin) { set_error("energy parameters need %zu B of shared memory", smem); return EBM_ERR_UNSUPPORTED; } \
    Pc.scratch_stride = ss;                                                                         \
    auto kern = hmc_kernel<RowE, G, E>;                                                             \
    if (smem > 48 * 1024) EBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<row_grid(di, Pc.n, G, 4), kRowThreads, smem, c.st>>>(en, Pc, tab);                       \
  }
    EBM_ROW_DISPATCH(c.e->dim, CALL);
#undef CALL
    return launch_status("hmc_kernel");
  });
}

int hmc_mlp_launch(const EbmEnergyDesc* e, const HmcParams& P, const HStepTable& tab, cudaStream_t st);  // ebm_mlp.cu
int hmc_mlp_tc_launch(const EbmEnergyDesc* e, const HmcParams& P, const HStepTable& tab, int passes, cudaStream_t st);  // ebm_mlp_tc.cu

}  // namespace ebm

using namespace ebm;

extern "C" {

int ebm_leapfrog_f32(const EbmEnergyDesc* e, const float* x_in, const float* p_in, float* x_out, float* p_out,
                     int64_t n, int32_t n_steps, double step_size, int32_t mass_kind, double mass_scalar,
                     const float* mass_vec, int32_t safe, void* stream) {
  int rc = validate_desc(e);
  if (rc) return rc;
  EBM_CHECK_ARG(x_in && p_in && x_out && p_out && n > 0, "state pointers must be non-null and n positive");
  EBM_CHECK_ARG(n_steps > 0, "n_steps must be positive");
  LeapfrogParams P;
  memset(&P, 0, sizeof(P));
  P.x_in = x_in; P.p_in = p_in; P.x_out = x_out; P.p_out = p_out;
  P.n = n; P.d = e->dim; P.n_steps = n_steps; P.safe = safe;
  P.h = (float)step_size;  // base_integrator.py:870-871: 0-d tensor of the state dtype
  rc = make_mass(mass_kind, mass_scalar, mass_vec, P.mass);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  switch (e->kind) {
    case EBM_ENERGY_DOUBLE_WELL: return launch_leapfrog(ElemRow<DoubleWellE>{make_dw(e)}, e, P, st);
    case EBM_ENERGY_HARMONIC: return launch_leapfrog(ElemRow<HarmonicE>{make_harm(e)}, e, P, st);
    case EBM_ENERGY_RASTRIGIN: return launch_leapfrog(ElemRow<RastriginE>{make_rast(e)}, e, P, st);
    case EBM_ENERGY_GAUSSIAN: return launch_leapfrog(make_gauss(e), e, P, st);
    case EBM_ENERGY_MOG: return launch_leapfrog(make_mog(e), e, P, st);
    default: set_error("leapfrog: energy kind %d has no fused kernel", e->kind); return EBM_ERR_UNSUPPORTED;
  }
}

static int hmc_burst_impl(const EbmEnergyDesc* e, const float* x_in, float* x_out, int64_t n, int32_t n_proposals,
                          int32_t n_leapfrog, const double* step_size_host, int32_t schedule_len, int32_t mass_kind,
                          double mass_scalar, const float* mass_vec, int32_t rng_mode, uint64_t seed, uint64_t offset,
                          const float* noise_p, const float* noise_u, float* traj, int32_t thin, int32_t* accept_count,
                          float* energy_out, void* stream, const HmcDiag* diag) {
  int rc = validate_desc(e);
  if (rc) return rc;
  EBM_CHECK_ARG(x_in && x_out && n > 0, "x_in/x_out must be non-null and n positive");
  EBM_CHECK_ARG(n_proposals > 0 && n_leapfrog > 0, "n_proposals and n_leapfrog must be positive");
  EBM_CHECK_ARG(step_size_host, "step_size_host must be non-null");
  EBM_CHECK_ARG(schedule_len == 1 || schedule_len == n_proposals, "schedule_len must be 1 or n_proposals");
  EBM_CHECK_ARG(thin >= 1, "thin must be >= 1");
  EBM_CHECK_ARG(rng_mode >= EBM_RNG_INJECTED && rng_mode <= EBM_RNG_NATIVE, "bad rng_mode");
  EBM_CHECK_ARG(rng_mode != EBM_RNG_INJECTED || (noise_p && noise_u), "INJECTED rng needs noise_p and noise_u");
  EBM_CHECK_ARG(offset % 4 == 0, "offset must be a multiple of 4");
  const DeviceInfo& di = device_info(current_device());
  HmcParams P;
  memset(&P, 0, sizeof(P));
  P.x_in = x_in; P.x_out = x_out; P.noise_p = noise_p; P.noise_u = noise_u; P.traj = traj;
  P.accept_count = accept_count; P.energy_out = energy_out;
  P.n = n; P.d = e->dim; P.n_leapfrog = n_leapfrog; P.thin = thin; P.n_kept = n_proposals / thin;
  rc = make_mass(mass_kind, mass_scalar, mass_vec, P.mass);
  if (rc) return rc;
  HmcCall c{e, step_size_host, schedule_len, n_proposals, offset, 0, 0, (cudaStream_t)stream, diag};
  P.rng_p.mode = P.rng_u.mode = rng_mode;
  if (rng_mode == EBM_RNG_TORCH) {
    const long long numel = (long long)n * e->dim;
    c.inc_p = torch_offset_increment(di, numel);
    c.inc_u = torch_offset_increment(di, n);
    P.rng_p.T = torch_threads(di, numel);
    P.rng_u.T = torch_threads(di, n);
    P.rng_p.k0 = P.rng_u.k0 = (uint32_t)seed;
    P.rng_p.k1 = P.rng_u.k1 = (uint32_t)(seed >> 32);
    P.rng_p.ctr_step = P.rng_u.ctr_step = (c.inc_p + c.inc_u) / 4;
  } else {
    P.rng_p.T = P.rng_u.T = 1;
    P.rng_p.k0 = P.rng_u.k0 = (uint32_t)seed ^ kNativeTag0;
    P.rng_p.k1 = P.rng_u.k1 = (uint32_t)(seed >> 32) ^ kNativeTag1;
    P.rng_p.ctr_step = P.rng_u.ctr_step = 2;
  }
  switch (e->kind) {
    case EBM_ENERGY_DOUBLE_WELL: return launch_hmc(ElemRow<DoubleWellE>{make_dw(e)}, c, P);
    case EBM_ENERGY_HARMONIC: return launch_hmc(ElemRow<HarmonicE>{make_harm(e)}, c, P);
    case EBM_ENERGY_RASTRIGIN: return launch_hmc(ElemRow<RastriginE>{make_rast(e)}, c, P);
    case EBM_ENERGY_GAUSSIAN: return launch_hmc(make_gauss(e), c, P);
    case EBM_ENERGY_MOG: return launch_hmc(make_mog(e), c, P);
    case EBM_ENERGY_MLP:
      if (e->dim > 128 || e->hidden1 > 128 || e->hidden2 > 128 || e->hidden3 > 0) {
        set_error("hmc: MLP energies wider than 128 or with three hidden layers have no fused HMC kernel");
        return EBM_ERR_UNSUPPORTED;
      }
      // bf16x3 (the default precision): tensor-core kernel, energies good to ~2e-5 relative; fp32: FFMA kernel; the
      // single-pass bf16 mode is refused here because accept decisions compare energies
      if (e->precision == EBM_MLP_BF16X3)
        return hmc_chunks(c, P, [&](HmcParams& Pc, const HStepTable& tab) -> int { return hmc_mlp_tc_launch(e, Pc, tab, 3, c.st); });
      if (e->precision == EBM_MLP_BF16) { set_error("hmc: precision bf16 (single pass) is not offered for HMC; use bf16x3 or fp32"); return EBM_ERR_UNSUPPORTED; }
      return hmc_chunks(c, P, [&](HmcParams& Pc, const HStepTable& tab) -> int { return hmc_mlp_launch(e, Pc, tab, c.st); });
    default: set_error("hmc: energy kind %d has no fused kernel", e->kind); return EBM_ERR_UNSUPPORTED;
  }
}

int ebm_hmc_burst_f32(const EbmEnergyDesc* e, const float* x_in, float* x_out, int64_t n, int32_t n_proposals,
                      int32_t n_leapfrog, const double* step_size_host, int32_t schedule_len, int32_t mass_kind,
                      double mass_scalar, const float* mass_vec, int32_t rng_mode, uint64_t seed, uint64_t offset,
                      const float* noise_p, const float* noise_u, float* traj, int32_t thin, int32_t* accept_count,
                      float* energy_out, void* stream) {
  return hmc_burst_impl(e, x_in, x_out, n, n_proposals, n_leapfrog, step_size_host, schedule_len, mass_kind, mass_scalar,
                        mass_vec, rng_mode, seed, offset, noise_p, noise_u, traj, thin, accept_count, energy_out, stream,
                        nullptr);
}

int ebm_hmc_burst_diag_f32(const EbmEnergyDesc* e, const float* x_in, float* x_out, int64_t n, int32_t n_proposals,
                           int32_t n_leapfrog, const double* step_size_host, int32_t schedule_len, int32_t mass_kind,
                           double mass_scalar, const float* mass_vec, int32_t rng_mode, uint64_t seed, uint64_t offset,
                           const float* noise_p, const float* noise_u, float* traj, int32_t thin, double* diag_ws,
                           float* scratch, int32_t* accept_count, float* diag_mean, float* diag_var, float* diag_energy,
                           float* diag_accept, void* stream) {
  EBM_CHECK_ARG(e && e->dim > 0, "null energy descriptor");
  EBM_CHECK_ARG(thin >= 1 && n_proposals / thin >= 1, "thin must be >= 1 and keep at least one sample");
  EBM_CHECK_ARG(diag_ws && scratch && accept_count && diag_mean && diag_var && diag_energy && diag_accept,
                "diagnostic outputs and scratch buffers must be non-null");
  cudaStream_t st = (cudaStream_t)stream;
  const int n_kept = n_proposals / thin;
  EBM_CUDA(cudaMemsetAsync(diag_ws, 0, (size_t)n_kept * diag_slot(e->dim) * sizeof(double), st));
  EBM_CUDA(cudaMemsetAsync(accept_count, 0, (size_t)n_proposals * sizeof(int32_t), st));
  HmcDiag dg{diag_ws, scratch};
  int rc = hmc_burst_impl(e, x_in, x_out, n, n_proposals, n_leapfrog, step_size_host, schedule_len, mass_kind, mass_scalar,
                          mass_vec, rng_mode, seed, offset, noise_p, noise_u, traj, thin, accept_count, nullptr, stream, &dg);
  if (rc) return rc;
  return diag_finalize(diag_ws, n_kept, e->dim, n, 1.0f, 0.0f, diag_mean, diag_var, diag_energy, accept_count, thin,
                       diag_accept, st);
}

}  // extern "C"
