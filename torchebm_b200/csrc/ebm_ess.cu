// Effective sample size of sampler diagnostics chains on the device (sm_100a).
//
// Replaces `_ess_from_chain` of the reference's benchmark harness (benchmarks/registry.py:348-365), which moves the
// energy chain of `return_diagnostics=True` to the host, takes an FFT autocorrelation and walks it with one `.item()`
// per lag.  Same estimator -- autocovariances c_k = sum_t x_t x_{t+k} of the centred chain, initial positive sequence
// (lags are summed until the first negative one), tau = 1 + 2 sum c_k / c_0, ESS = n / max(tau, 1) -- with the
// autocovariances as direct fp64 sums instead of an fp32 FFT (identical in exact arithmetic; the chains of interest are
// a few thousand kept samples, and the walk usually stops after a few lags).
// One CTA per chain; the centred chain sits in shared memory; 256 lags are evaluated per round (thread = lag, the reads
// of x[t + lag] are conflict-free, x[t] is a broadcast) and thread 0 walks the round's lags in order.
#include "api_common.cuh"

namespace ebm {

constexpr int kEssThreads = 256;

__device__ __forceinline__ double ess_block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < kEssThreads / 32; ++w) s += red[w];
  return s;
}

template <bool IN_SMEM>
__global__ void __launch_bounds__(kEssThreads) ess_kernel(const float* __restrict__ chains, long long n, float* __restrict__ ess_out) {
  extern __shared__ __align__(16) uint8_t ess_smem[];
  double* red = reinterpret_cast<double*>(ess_smem);                   // [8] block reduction
  double* lags = red + kEssThreads / 32;                               // [256] autocovariances of the current round
  float* xs = reinterpret_cast<float*>(lags + kEssThreads);            // [n] centred chain (IN_SMEM)
  const float* chain = chains + (long long)blockIdx.x * n;
  const int tid = threadIdx.x;
  if (n < 2) {                                                         // registry.py:351-352
    if (tid == 0) ess_out[blockIdx.x] = (float)n;
    return;
  }
  double s = 0.0;
  for (long long t = tid; t < n; t += kEssThreads) s += (double)chain[t];
  const float mean = (float)(ess_block_sum(s, red) / (double)n);       // :353 (fp32 mean, fp32 centring)
  double c0p = 0.0;
  for (long long t = tid; t < n; t += kEssThreads) {
    const float x = chain[t] - mean;
    if (IN_SMEM) xs[t] = x;
    c0p += (double)x * (double)x;
  }
  const double c0 = ess_block_sum(c0p, red);                           // (the barriers inside also publish xs)
  if (c0 == 0.0) {                                                     // :356-357
    if (tid == 0) ess_out[blockIdx.x] = (float)n;
    return;
  }
  __shared__ int done;
  __shared__ double total_sh;
  if (tid == 0) { done = 0; total_sh = 0.0; }
  __syncthreads();
  for (long long lag0 = 1; lag0 < n; lag0 += kEssThreads) {
    const long long lag = lag0 + tid;
    double ck = 0.0;
    if (lag < n) {
      const long long m = n - lag;
      if (IN_SMEM) {
        for (long long t = 0; t < m; ++t) ck += (double)xs[t] * (double)xs[t + lag];
      } else {
        for (long long t = 0; t < m; ++t) ck += (double)(chain[t] - mean) * (double)(chain[t + lag] - mean);
      }
    }
    lags[tid] = ck;
    __syncthreads();
    if (tid == 0) {                                                    // :360-363, in lag order
      double total = total_sh;
      const long long cnt = (n - lag0 < kEssThreads) ? (n - lag0) : kEssThreads;
      for (long long i = 0; i < cnt; ++i) {
        if (lags[i] < 0.0) { done = 1; break; }
        total += lags[i] / c0;
      }
      total_sh = total;
    }
    __syncthreads();
    if (done) break;
  }
  if (tid == 0) {
    const double tau = 1.0 + 2.0 * total_sh;                           // :364
    ess_out[blockIdx.x] = (float)((double)n / (tau > 1.0 ? tau : 1.0));  // :365
  }
}

}  // namespace ebm

using namespace ebm;

extern "C" int ebm_ess_f32(const float* chains, int64_t n_chains, int64_t n, float* ess_out, void* stream) {
  EBM_CHECK_ARG(chains && ess_out, "chains/ess_out must be non-null");
  EBM_CHECK_ARG(n_chains > 0 && n > 0, "n_chains and n must be positive");
  EBM_CHECK_ARG(n_chains <= 0x7fffffff, "too many chains");
  const DeviceInfo& di = device_info(current_device());
  const size_t fixed = (kEssThreads / 32 + kEssThreads) * sizeof(double);
  const size_t need = fixed + (size_t)n * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (need <= (size_t)di.max_smem_optin) {
    if (need > 48 * 1024) EBM_CUDA(cudaFuncSetAttribute(ess_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    ess_kernel<true><<<(unsigned)n_chains, kEssThreads, need, st>>>(chains, n, ess_out);
  } else {
    ess_kernel<false><<<(unsigned)n_chains, kEssThreads, fixed, st>>>(chains, n, ess_out);
  }
  return launch_status("ess_kernel");
}
