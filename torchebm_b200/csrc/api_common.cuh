// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/ebm_b200.h"
#include "row_kernels.cuh"

namespace ebm {

void set_error(const char* fmt, ...);

struct DeviceInfo {
  int sm_count;
  int max_threads_per_sm;
  int max_smem_optin;
};
const DeviceInfo& device_info(int device);
int current_device();

#define EBM_CHECK_ARG(cond, msg)               \
  do {                                         \
    if (!(cond)) {                             \
      ::ebm::set_error("%s: %s", __func__, msg); \
      return EBM_ERR_INVALID;                  \
    }                                          \
  } while (0)

#define EBM_CUDA(call)                                                         \
  do {                                                                         \
    cudaError_t err__ = (call);                                                \
    if (err__ != cudaSuccess) {                                                \
      ::ebm::set_error("%s: %s -> %s", __func__, #call, cudaGetErrorString(err__)); \
      return (int)err__;                                                       \
    }                                                                          \
  } while (0)

inline int launch_status(const char* what) {
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(err));
    return (int)err;
  }
  return 0;
}

// torch's calc_execution_policy (DistributionTemplates.h:50-62)
inline uint64_t torch_threads(const DeviceInfo& di, int64_t numel) {
  uint64_t grid = ((uint64_t)numel + 255) / 256;
  const uint64_t cap = (uint64_t)di.sm_count * (uint64_t)(di.max_threads_per_sm / 256);
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  return grid * 256;
}
inline uint64_t torch_offset_increment(const DeviceInfo& di, int64_t numel) {
  const uint64_t T = torch_threads(di, numel);
  return (((uint64_t)numel - 1) / (T * 4) + 1) * 4;
}

// fp32 coefficients exactly as torch rounds the reference's Python doubles (base_integrator.py:728-729)
inline void fill_step(StepTable& t, int i, double h, double ns) {
  t.h[i] = (float)h;
  t.c1[i] = (float)pow(h, 0.5);
  t.c2[i] = (float)pow(2.0 * (ns * ns), 0.5);
}

// (G, EPT) of the row layout for a row length d; returns false when d is not supported
inline bool row_config(int d, int& G, int& EPT) {
  if (d <= 0) return false;
  if (d <= 2) { G = 2; EPT = 1; }
  else if (d <= 4) { G = 4; EPT = 1; }
  else if (d <= 8) { G = 8; EPT = 1; }
  else if (d <= 16) { G = 16; EPT = 1; }
  else if (d <= 32) { G = 32; EPT = 1; }
  else if (d <= 64) { G = 32; EPT = 2; }
  else if (d <= 128) { G = 32; EPT = 4; }
  else if (d <= 256) { G = 32; EPT = 8; }
  else if (d <= 1024) { G = 32; EPT = 32; }
  else return false;
  return true;
}

#define EBM_ROW_DISPATCH(d, CALL)                                  \
  do {                                                             \
    int G__, E__;                                                  \
    if (!::ebm::row_config(d, G__, E__)) {                         \
      ::ebm::set_error("row length %d not supported (max 1024)", d); \
      return EBM_ERR_UNSUPPORTED;                                  \
    }                                                              \
    if (E__ == 1) {                                                \
      switch (G__) {                                               \
        case 2: CALL(2, 1); break;                                 \
        case 4: CALL(4, 1); break;                                 \
        case 8: CALL(8, 1); break;                                 \
        case 16: CALL(16, 1); break;                               \
        default: CALL(32, 1); break;                               \
      }                                                            \
    } else if (E__ == 2) { CALL(32, 2); }                          \
    else if (E__ == 4) { CALL(32, 4); }                            \
    else if (E__ == 8) { CALL(32, 8); }                            \
    else { CALL(32, 32); }                                         \
  } while (0)

inline DoubleWellE make_dw(const EbmEnergyDesc* e) { return DoubleWellE{e->p[0], e->p[1]}; }
inline HarmonicE make_harm(const EbmEnergyDesc* e) { return HarmonicE{e->p[0]}; }
inline RastriginE make_rast(const EbmEnergyDesc* e) { return RastriginE{e->p[0], e->p[1], e->p[2]}; }
inline GaussianRow make_gauss(const EbmEnergyDesc* e) { return GaussianRow{e->buf[0], e->buf[1], e->dim}; }
inline MogRow make_mog(const EbmEnergyDesc* e) { return MogRow{e->buf[0], e->buf[1], e->buf[2], e->dim, e->n_components}; }

inline int validate_desc(const EbmEnergyDesc* e) {
  if (!e) { set_error("null energy descriptor"); return EBM_ERR_INVALID; }
  if (e->dim <= 0) { set_error("energy dim must be positive"); return EBM_ERR_INVALID; }
  switch (e->kind) {
    case EBM_ENERGY_DOUBLE_WELL: case EBM_ENERGY_HARMONIC: case EBM_ENERGY_RASTRIGIN: return 0;
    case EBM_ENERGY_GAUSSIAN:
      if (!e->buf[0] || !e->buf[1]) { set_error("gaussian needs mean and cov_inv"); return EBM_ERR_INVALID; }
      return 0;
    case EBM_ENERGY_MOG:
      if (!e->buf[0] || !e->buf[1] || !e->buf[2] || e->n_components <= 0) { set_error("mog needs means, sigmas, weights"); return EBM_ERR_INVALID; }
      return 0;
    case EBM_ENERGY_MLP:
      for (int i = 0; i < 6; ++i) if (!e->buf[i]) { set_error("mlp needs W1,b1,W2,b2,w3,b3"); return EBM_ERR_INVALID; }
      if (e->hidden1 <= 0 || e->hidden2 <= 0 || e->hidden3 < 0) { set_error("mlp hidden sizes must be positive"); return EBM_ERR_INVALID; }
      if (e->hidden3 > 0 && (!e->buf[7] || !e->buf[8])) { set_error("three-hidden-layer mlp needs W3 (buf[7]) and b3 (buf[8])"); return EBM_ERR_INVALID; }
      return 0;
    default: set_error("unknown energy kind %d", e->kind); return EBM_ERR_INVALID;
  }
}

// shared memory (bytes) a row kernel needs for energy `e` with layout group size G
inline size_t row_smem_bytes(const EbmEnergyDesc* e, int G, int& scratch_stride) {
  size_t staged = 0;
  scratch_stride = 0;
  if (e->kind == EBM_ENERGY_GAUSSIAN) { staged = (size_t)e->dim * e->dim + e->dim; scratch_stride = e->dim; }
  else if (e->kind == EBM_ENERGY_MOG) { staged = (size_t)e->n_components * e->dim + 2 * (size_t)e->n_components; scratch_stride = e->n_components; }
  return (staged + (size_t)(kRowThreads / G) * scratch_stride) * sizeof(float);
}

struct LangevinCall {
  const EbmEnergyDesc* e;
  const float* x_in;
  float* x_out;
  int64_t n;
  int32_t n_steps;
  const double* hs;
  const double* nss;
  int32_t schedule_len;
  const float* clamp;
  int32_t rng_mode;
  uint64_t seed, offset;
  const float* noise;
  float* traj;
  int32_t thin;
  cudaStream_t st;
  // persistent-CD fusion (ebm_pcd_langevin_burst_f32; MLP kernels only, NULL otherwise):
  const long long* row_index;  // chain i starts from row row_index[i] of x_in (the replay buffer)
  float* x_out2;               // the final state is also written here, row i -> row i (FIFO write-back when S == B)
  // elementwise kernels only: restrict the launch to owning threads [quad_begin, quad_end) of the burst (0, 0 = all);
  // the host-buffer entry point pipelines copies against such partial launches
  unsigned long long quad_begin, quad_end;
  // burst-end gather (ebm_langevin_burst_gather_f32): peer-mapped gathered buffers, this rank's first row in them
  float* const* peers;
  int n_peers;
  long long peer_row_offset;
  int scheme;   // 0 = Euler-Maruyama, 1 = Heun (elementwise energies only)
  // in-burst diagnostics (elementwise kernels only): fp64 workspace [n_steps / thin, diag_slot(dim)], see diag.cuh
  double* diag_ws;
  // burst-end gather through NVLS: peers[0] is a multicast address of the gathered buffers and n_peers == 1
  int peer_mc;
};

// Form of the wide MLP kernel's burst-end gather (ebm_mlp_wide.cu): bulk copies x_out -> shared memory -> every peer
// mapping (default; measured best at 2 and at 8 GPUs), or with EBM_B200_PUSH_BULK=0 16-byte stores -- through the NVLS
// multicast address when the caller passes one, else one per peer.
inline int wide_push_bulk() {
  static const int v = [] { const char* s = getenv("EBM_B200_PUSH_BULK"); return (s && s[0] == '0') ? 0 : 1; }();
  return v;
}

// diagnostics helpers shared by the Langevin and HMC entry points (ebm_core.cu)
// ws slot += column sums / sums of squares of x[n, d] and the sum of energy[n] (energy may be null)
int diag_accumulate(double* ws_slot, const float* x, const float* energy, int64_t n, int d, cudaStream_t st);
// ws -> mean / var [n_kept, d], energy [n_kept] = e_scale * (energy sum / n) + e_shift; accept (may be null):
// acceptance_rate[j] = accept_count[(j + 1) * thin - 1] / n
int diag_finalize(const double* ws, int n_kept, int d, int64_t n, float e_scale, float e_shift, float* mean, float* var,
                  float* energy, const int32_t* accept_count, int thin, float* accept_rate, cudaStream_t st);

inline int row_grid(const DeviceInfo& di, long long n, int G, int ctas_per_sm) {
  const long long rows_per_cta = kRowThreads / G;
  long long grid = (n + rows_per_cta - 1) / rows_per_cta;
  const long long cap = (long long)di.sm_count * ctas_per_sm;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  return (int)grid;
}

}  // namespace ebm
