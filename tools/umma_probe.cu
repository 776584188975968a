// Standalone probe for the tcgen05 building blocks used by the tensor-core MLP kernel:
// no-swizzle core-matrix shared-memory layout, K-major and MN-major B descriptors on the SAME
// physical weight copy, kind::f16 (bf16 in, fp32 accumulate in TMEM), tcgen05.commit -> mbarrier,
// tcgen05.ld/st.  Compares against a double-precision host product of the bf16-rounded inputs.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/umma_probe tools/umma_probe.cu && /tmp/umma_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../torchebm_b200/csrc/umma.cuh"

using namespace ebm::umma;

constexpr int M = 128, N = 128, K = 128;

__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                    float* __restrict__ D1, float* __restrict__ D2,
                                                    float* __restrict__ D3) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem;                    // [M x K] bf16, core-matrix layout, 32 KB
  uint8_t* sW = smem + M * K * 2;        // [N x K] bf16 (W[o][i]), 32 KB
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int idx = tid; idx < M * K; idx += blockDim.x) {
    const int r = idx / K, k = idx % K;
    *reinterpret_cast<__nv_bfloat16*>(sA + core_offset(r, k, M)) = __float2bfloat16_rn(A[idx]);
    *reinterpret_cast<__nv_bfloat16*>(sW + core_offset(r, k, N)) = __float2bfloat16_rn(W[idx]);
  }
  if (tid == 0) mbar_init(smem_u32(&bar), 1);
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_slot), 512);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_slot;

  if (tid == 0) {
    const uint32_t a0 = smem_u32(sA), w0 = smem_u32(sW);
    const uint32_t idesc_k = make_idesc_bf16(M, N, /*b_mn_major=*/false);
    const uint32_t idesc_mn = make_idesc_bf16(M, N, /*b_mn_major=*/true);
    // D1[m, o] = sum_i A[m, i] W[o, i]   (B K-major: rows o, contiguous i)
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t ad = make_smem_desc(a0 + ks * 2 * (M * 16), M * 16, 128);
      const uint64_t bd = make_smem_desc(w0 + ks * 2 * (N * 16), N * 16, 128);
      mma_bf16(tmem + 0, ad, bd, idesc_k, ks > 0);
    }
    // D2[m, i] = sum_o A[m, o] W[o, i]   (B MN-major on the same copy: MN = i, K = o)
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t ad = make_smem_desc(a0 + ks * 2 * (M * 16), M * 16, 128);
      const uint64_t bd = make_smem_desc(w0 + ks * 256, /*lbo (next 8 k)*/ 128, /*sbo (next 8 n)*/ N * 16);
      mma_bf16(tmem + 128, ad, bd, idesc_mn, ks > 0);
    }
    mma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tcgen05_fence_after();

  // each thread owns TMEM lane = row tid; read 16 columns at a time
  const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c = 0; c < N; c += 16) {
    float v[16];
    tmem_ld16(lane_addr + c, v);
    for (int j = 0; j < 16; ++j) D1[tid * N + c + j] = v[j];
    tmem_ld16(lane_addr + 128 + c, v);
    for (int j = 0; j < 16; ++j) D2[tid * N + c + j] = v[j];
    // st round trip: write 2*v into columns 256.. and read it back
    for (int j = 0; j < 16; ++j) v[j] *= 2.0f;
    tmem_st16(lane_addr + 256 + c, v);
  }
  tmem_st_wait();
  for (int c = 0; c < N; c += 16) {
    float v[16];
    tmem_ld16(lane_addr + 256 + c, v);
    for (int j = 0; j < 16; ++j) D3[tid * N + c + j] = v[j];
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

int main() {
  const int n = M * K;
  float *hA = (float*)malloc(n * 4), *hW = (float*)malloc(n * 4);
  srand(1);
  for (int i = 0; i < n; ++i) { hA[i] = (rand() / (float)RAND_MAX - 0.5f) * 2; hW[i] = (rand() / (float)RAND_MAX - 0.5f) * 2; }
  float *dA, *dW, *d1, *d2, *d3;
  cudaMalloc(&dA, n * 4); cudaMalloc(&dW, n * 4); cudaMalloc(&d1, n * 4); cudaMalloc(&d2, n * 4); cudaMalloc(&d3, n * 4);
  cudaMemcpy(dA, hA, n * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, hW, n * 4, cudaMemcpyHostToDevice);
  const int smem = 2 * M * K * 2;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<1, 128, smem>>>(dA, dW, d1, d2, d3);
  cudaError_t err = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(err));
  if (err != cudaSuccess) return 1;
  float *h1 = (float*)malloc(n * 4), *h2 = (float*)malloc(n * 4), *h3 = (float*)malloc(n * 4);
  cudaMemcpy(h1, d1, n * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h2, d2, n * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h3, d3, n * 4, cudaMemcpyDeviceToHost);
  double e1 = 0, e2 = 0, e3 = 0;
  for (int m = 0; m < M; ++m)
    for (int j = 0; j < N; ++j) {
      double r1 = 0, r2 = 0;
      for (int k = 0; k < K; ++k) {
        r1 += (double)bf16_round(hA[m * K + k]) * bf16_round(hW[j * K + k]);
        r2 += (double)bf16_round(hA[m * K + k]) * bf16_round(hW[k * K + j]);
      }
      e1 = fmax(e1, fabs(r1 - h1[m * N + j]));
      e2 = fmax(e2, fabs(r2 - h2[m * N + j]));
      e3 = fmax(e3, fabs(2.0 * h2[m * N + j] - h3[m * N + j]));
    }
  printf("max err K-major B: %.3e   MN-major B: %.3e   st/ld round trip: %.3e\n", e1, e2, e3);
  printf("%s\n", (e1 < 1e-3 && e2 < 1e-3 && e3 == 0) ? "PROBE OK" : "PROBE FAILED");
  return 0;
}
