// Micro-benchmark of sm_100a issue / pipe throughput for the instruction classes the fused sampler kernels are made
// of (Philox = IMAD.WIDE + LOP3, Box-Muller / activations = FFMA(2) + MUFU, operand splitting = F2FP / PRMT / I2FP).
// Prints cycles per warp instruction per SM sub-partition at full occupancy (16 warps per sub-partition).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_probe tools/pipe_probe.cu && tools/pipe_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 4096
typedef unsigned long long u64;

#define OP_IMADW(i) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"((uint32_t)w[i]), "r"(0xD2511F53u));
#define OP_LOP3(i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(a[(i + 1) & 7]), "r"(k));
#define OP_FFMA(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(c0), "f"(c1));
#define OP_FMUL(i) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c0));
#define OP_FFMA2(i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pc0), "l"(pc1));
#define OP_FMUL2(i) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc0));
#define OP_EX2(i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
#define OP_RCP(i) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
#define OP_I2FP(i) asm volatile("{.reg .f32 t; cvt.rn.f32.u32 t, %0; mov.b32 %0, t;}" : "+r"(a[i]));
#define OP_F2FP(i) asm volatile("{.reg .f32 t, u; mov.b32 t, %0; mov.b32 u, %1; cvt.rn.bf16x2.f32 %0, t, u;}" : "+r"(a[i]) : "r"(a[(i + 1) & 7]));
#define OP_PRMT(i) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(a[i]) : "r"(a[(i + 1) & 7]));
#define OP_IADD(i) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(k));
#define OP_SHL(i) asm volatile("shl.b32 %0, %0, 16;" : "+r"(a[i]));

#define REP8(OP) OP(0) OP(1) OP(2) OP(3) OP(4) OP(5) OP(6) OP(7)

#define KERNEL(name, BODY, NINSTR)                                                                   \
  __global__ void __launch_bounds__(256) name(float* out, uint32_t k, float c0, float c1) {           \
    uint32_t a[8]; float f[8]; u64 w[8]; u64 p[8];                                                    \
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x + i; f[i] = 1.0f + i * 1e-3f; w[i] = threadIdx.x * 77 + i; \
      p[i] = ((u64)__float_as_uint(f[i]) << 32) | __float_as_uint(f[i]); }                            \
    u64 pc0 = ((u64)__float_as_uint(c0) << 32) | __float_as_uint(c0);                                 \
    u64 pc1 = ((u64)__float_as_uint(c1) << 32) | __float_as_uint(c1);                                 \
    _Pragma("unroll 1") for (int it = 0; it < ITERS; ++it) { BODY }                                                       \
    float s = 0; for (int i = 0; i < 8; ++i) s += f[i] + a[i] + (float)w[i] + (float)p[i];            \
    if (s == 123.456f) out[0] = s;                                                                    \
  }                                                                                                   \
  static const int name##_n = NINSTR;

KERNEL(k_imadw, REP8(OP_IMADW), 8)
KERNEL(k_lop3, REP8(OP_LOP3), 8)
KERNEL(k_iadd, REP8(OP_IADD), 8)
KERNEL(k_ffma, REP8(OP_FFMA), 8)
KERNEL(k_fmul, REP8(OP_FMUL), 8)
KERNEL(k_ffma2, REP8(OP_FFMA2), 8)
KERNEL(k_fmul2, REP8(OP_FMUL2), 8)
KERNEL(k_ex2, REP8(OP_EX2), 8)
KERNEL(k_rcp, REP8(OP_RCP), 8)
KERNEL(k_i2fp, REP8(OP_I2FP), 8)
KERNEL(k_f2fp, REP8(OP_F2FP), 8)
KERNEL(k_prmt, REP8(OP_PRMT), 8)
KERNEL(k_shl, REP8(OP_SHL), 8)
KERNEL(k_philox, REP8(OP_IMADW) REP8(OP_LOP3), 16)
KERNEL(k_imadw_ffma, REP8(OP_IMADW) REP8(OP_FFMA), 16)
KERNEL(k_imadw_ffma2, REP8(OP_IMADW) REP8(OP_FFMA2), 16)
KERNEL(k_ffma_ffma2, REP8(OP_FFMA) REP8(OP_FFMA2), 16)
KERNEL(k_ffma_lop3, REP8(OP_FFMA) REP8(OP_LOP3), 16)
KERNEL(k_ffma2_lop3, REP8(OP_FFMA2) REP8(OP_LOP3), 16)
KERNEL(k_ffma_ex2, REP8(OP_FFMA) OP_EX2(0) OP_EX2(1), 10)
KERNEL(k_imadw_lop3_ffma, REP8(OP_IMADW) REP8(OP_LOP3) REP8(OP_FFMA), 24)
KERNEL(k_imadw_lop3_ffma2, REP8(OP_IMADW) REP8(OP_LOP3) REP8(OP_FFMA2), 24)
KERNEL(k_imadw_lop3_ffma_x2, REP8(OP_IMADW) REP8(OP_LOP3) REP8(OP_FFMA) REP8(OP_FMUL), 32)
KERNEL(k_mix_c2, REP8(OP_IMADW) REP8(OP_LOP3) REP8(OP_FFMA) REP8(OP_FMUL) REP8(OP_FFMA2) OP_EX2(0) OP_EX2(1) OP_I2FP(2) OP_I2FP(3), 44)

template <class K>
static void run(const char* name, K kern, int n_instr, int sms) {
  float* out;
  cudaMalloc(&out, 4);
  const int blocks = sms * 8;   // 8 CTAs x 256 threads = 64 warps per SM = 16 per sub-partition
  kern<<<blocks, 256>>>(out, 1, 1.0000001f, 1e-9f);
  cudaDeviceSynchronize();
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  kern<<<blocks, 256>>>(out, 1, 1.0000001f, 1e-9f);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  int clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double cycles = ms * 1e-3 * clk_khz * 1e3;
  const double warp_instr_per_smsp = 16.0 * ITERS * n_instr;
  printf("%-24s %8.3f ms   %6.3f cycles per warp-instruction per sub-partition (at %d MHz nominal)\n", name, ms,
         cycles / warp_instr_per_smsp, clk_khz / 1000);
  cudaFree(out);
}

int main() {
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
#define RUN(k) run(#k, k, k##_n, sms)
  RUN(k_imadw); RUN(k_lop3); RUN(k_iadd); RUN(k_ffma); RUN(k_fmul); RUN(k_ffma2); RUN(k_fmul2); RUN(k_ex2); RUN(k_rcp);
  RUN(k_i2fp); RUN(k_f2fp); RUN(k_prmt); RUN(k_shl);
  RUN(k_philox); RUN(k_imadw_ffma); RUN(k_imadw_ffma2); RUN(k_ffma_ffma2); RUN(k_ffma_lop3); RUN(k_ffma2_lop3); RUN(k_ffma_ex2);
  RUN(k_imadw_lop3_ffma); RUN(k_imadw_lop3_ffma2); RUN(k_imadw_lop3_ffma_x2); RUN(k_mix_c2);
  return 0;
}
