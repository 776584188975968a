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

// A operand in tensor memory (row = lane, column j = K elements 2j, 2j+1 as packed bf16x2)
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  const uint32_t acc = accumulate ? 1u : 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                    float* __restrict__ D1, float* __restrict__ D2,
                                                    float* __restrict__ D3, float* __restrict__ D4,
                                                    float* __restrict__ D5, long long* __restrict__ timing) {
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
  // ---- A operand from tensor memory: pack row tid of A as bf16x2 words into columns [384, 448) -----------------
  for (int c = 0; c < K / 2; c += 16) {
    float w[16];
    for (int j = 0; j < 16; ++j) {
      const __nv_bfloat162 p2 = __floats2bfloat162_rn(A[tid * K + 2 * (c + j)], A[tid * K + 2 * (c + j) + 1]);
      w[j] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&p2));
    }
    tmem_st16(lane_addr + 384 + c, w);
  }
  tmem_st_wait();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (tid == 0) {
    const uint32_t w0 = smem_u32(sW);
    const uint32_t idesc_k = make_idesc_bf16(M, N, false), idesc_mn = make_idesc_bf16(M, N, true);
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t bd = make_smem_desc(w0 + ks * 2 * (N * 16), N * 16, 128);
      mma_bf16_ts(tmem + 0, tmem + 384 + 8 * ks, bd, idesc_k, ks > 0);
    }
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t bd = make_smem_desc(w0 + ks * 256, 128, N * 16);
      mma_bf16_ts(tmem + 128, tmem + 384 + 8 * ks, bd, idesc_mn, ks > 0);
    }
    mma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 1);
  tcgen05_fence_after();
  for (int c = 0; c < N; c += 16) {
    float v[16];
    tmem_ld16(lane_addr + c, v);
    for (int j = 0; j < 16; ++j) D4[tid * N + c + j] = v[j];
    tmem_ld16(lane_addr + 128 + c, v);
    for (int j = 0; j < 16; ++j) D5[tid * N + c + j] = v[j];
  }
  // ---- issue-to-completion time of 256 MMAs: operands from shared memory vs A from tensor memory, N = 128 and 64 ----
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t par = 0;
  for (int variant = 0; variant < 8; ++variant) {
    const bool ts = variant & 1;
    const int nn = (variant & 2) ? 64 : 128;
    const bool two_acc = variant & 4;   // alternate between two accumulators (no back-to-back dependency)
    long long t0 = 0;
    if (warp == 0) {   // converged warp, one elected lane issues (as the kernels do)
      const bool leader = elect_one();
      const uint32_t a0 = smem_u32(sA), w0 = smem_u32(sW);
      const uint32_t idesc = make_idesc_bf16(M, nn, true);
      const uint64_t ad0 = make_smem_desc(a0, M * 16, 128), bd0 = make_smem_desc(w0, 128, N * 16);
      t0 = clock64();
      for (int it = 0; it < 32; ++it) {
#pragma unroll
        for (int ks = 0; ks < K / 16; ++ks) {
          const uint32_t d = tmem + ((two_acc && (ks & 1)) ? 128 : 0);
          if (leader) {
            if (ts) mma_bf16_ts(d, tmem + 384 + 8 * ks, bd0 + (uint64_t)(ks * 256 >> 4), idesc, true);
            else mma_bf16(d, ad0 + (uint64_t)(ks * 2 * (M * 16) >> 4), bd0 + (uint64_t)(ks * 256 >> 4), idesc, true);
          }
        }
      }
      if (leader) mma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), par); par ^= 1;
    if (tid == 0) timing[variant] = clock64() - t0;
    __syncthreads();
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
  float *dA, *dW, *d1, *d2, *d3, *d4, *d5;
  long long* dT;
  cudaMalloc(&dA, n * 4); cudaMalloc(&dW, n * 4); cudaMalloc(&d1, n * 4); cudaMalloc(&d2, n * 4); cudaMalloc(&d3, n * 4);
  cudaMalloc(&d4, n * 4); cudaMalloc(&d5, n * 4); cudaMalloc(&dT, 128);
  cudaMemcpy(dA, hA, n * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, hW, n * 4, cudaMemcpyHostToDevice);
  const int smem = 2 * M * K * 2;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<1, 128, smem>>>(dA, dW, d1, d2, d3, d4, d5, dT);
  cudaError_t err = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(err));
  if (err != cudaSuccess) return 1;
  float *h1 = (float*)malloc(n * 4), *h2 = (float*)malloc(n * 4), *h3 = (float*)malloc(n * 4);
  cudaMemcpy(h1, d1, n * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h2, d2, n * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h3, d3, n * 4, cudaMemcpyDeviceToHost);
  float *h4 = (float*)malloc(n * 4), *h5 = (float*)malloc(n * 4);
  long long hT[8];
  cudaMemcpy(h4, d4, n * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h5, d5, n * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(hT, dT, 64, cudaMemcpyDeviceToHost);
  double e4 = 0, e5 = 0;
  for (int i = 0; i < n; ++i) { e4 = fmax(e4, fabs((double)h4[i] - h1[i])); e5 = fmax(e5, fabs((double)h5[i] - h2[i])); }
  printf("A from tensor memory vs A from shared memory: K-major B %.3e   MN-major B %.3e  (%s)\n", e4, e5,
         (e4 == 0 && e5 == 0) ? "TS OK" : "TS MISMATCH");
  printf("cycles per MMA (256 back to back, K=16, one accumulator):  N=128 SS %.1f TS %.1f   N=64 SS %.1f TS %.1f\n",
         hT[0] / 256.0, hT[1] / 256.0, hT[2] / 256.0, hT[3] / 256.0);
  printf("cycles per MMA (two accumulators alternating):             N=128 SS %.1f TS %.1f   N=64 SS %.1f TS %.1f\n",
         hT[4] / 256.0, hT[5] / 256.0, hT[6] / 256.0, hT[7] / 256.0);
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
