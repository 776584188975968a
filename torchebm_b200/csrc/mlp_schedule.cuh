// Balanced work split of the persistent MLP Langevin kernels.
//
// A burst is `tiles` x `K` tile-steps (tile = 128 chains, one Langevin step).  Handing out whole tiles leaves
// ceil(tiles / SMs) rounds of work on some SMs and one round less on the others (65 536 chains = 512 tiles on 148 SMs:
// 4 rounds for 3.46 rounds of work).  Instead every CTA gets a contiguous range of the linearised (tile-major)
// tile-step sequence, equal to within one tile-step, so a tile's K steps may be split between CTA b-1 (its first
// steps) and CTA b (the rest).  Because tiles > CTAs whenever anything is split, a range is longer than K: a tile
// spans at most two CTAs.  Each CTA walks its range from its LAST tile to its FIRST, so the shared tile's head is
// the first thing CTA b-1 does and its tail the last thing CTA b does: the hand-over (chain state through x_out in
// global memory, one release/acquire flag per CTA) is never waited on for long and cannot deadlock -- the head unit
// of any CTA depends on nothing.  The RNG is counter based per (element, step), so results do not depend on the split.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace ebm {

constexpr int kMlpFlagBytes = 2048;  // one int per worker (a CTA, or one of the two tile pipelines of a CTA; <= 512)

struct MlpSchedule {
  long long quota;  // work quanta per CTA (floor)
  int rem;          // the first `rem` CTAs take one quantum more
  int gran;         // tile-steps per quantum: 1 = balanced split, K = whole tiles only (no hand-over, no flags needed)
  int* flags;       // [grid], zeroed before the launch; flags[b] counts epilogue warps of CTA b done with its head unit
};

// The CTA's range, reduced to four ints that live in shared memory (not in registers across the hot loops):
// tiles [t_first, t_last], first tile from step first_s0, last tile up to step last_s1.
struct MlpUnits {
  int t_first, t_last, first_s0, last_s1;
};
// `worker` = index of the entity that walks a range: the CTA, or (kernels that run two tile pipelines per CTA)
// 2 * blockIdx.x + pipeline; the host sets the schedule up for as many workers
__device__ __forceinline__ void mlp_units_compute(const MlpSchedule& s, int K, volatile MlpUnits* u, long long worker) {
  const long long b = worker;
  const long long lin_begin = (b * s.quota + (b < s.rem ? b : s.rem)) * s.gran;
  const long long lin_end = lin_begin + (s.quota + (b < s.rem ? 1 : 0)) * s.gran;
  const long long tf = lin_begin / K, tl = (lin_end - 1) / K;
  u->t_first = (int)tf;
  u->t_last = (int)tl;
  u->first_s0 = (int)(lin_begin - tf * K);
  u->last_s1 = (int)(lin_end - tl * K);
}
__device__ __forceinline__ void mlp_units_compute(const MlpSchedule& s, int K, volatile MlpUnits* u) {
  mlp_units_compute(s, K, u, blockIdx.x);
}
// steps [s0, s1) of `tile` that belong to this CTA
__device__ __forceinline__ int mlp_unit_s0(const volatile MlpUnits* u, int tile) { return tile == u->t_first ? u->first_s0 : 0; }
__device__ __forceinline__ int mlp_unit_s1(const volatile MlpUnits* u, int tile, int K) { return tile == u->t_last ? u->last_s1 : K; }

// called by every epilogue warp (all lanes) after its last global store of a head unit
__device__ __forceinline__ void mlp_unit_release(const MlpSchedule& s, int worker) {
  __threadfence();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) atomicAdd(s.flags + worker, 1);
}
__device__ __forceinline__ void mlp_unit_release(const MlpSchedule& s) { mlp_unit_release(s, blockIdx.x); }
// called by every epilogue warp (all lanes) before its first global load of a tail unit
__device__ __forceinline__ void mlp_unit_acquire(const MlpSchedule& s, int n_warps, int worker) {
  if ((threadIdx.x & 31) == 0) {
    const int* f = s.flags + (worker - 1);
    int v;
    do {
      asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if (v < n_warps) __nanosleep(200);
    } while (v < n_warps);
  }
  __syncwarp();
}

__device__ __forceinline__ void mlp_unit_acquire(const MlpSchedule& s, int n_warps) { mlp_unit_acquire(s, n_warps, blockIdx.x); }

// host: fill the schedule for `tiles` x `k_steps` on `grid` workers and zero the flags on the stream
inline int mlp_schedule_setup(MlpSchedule& s, long long tiles, int k_steps, int grid, int* flags, cudaStream_t st) {
  const long long total = tiles * (long long)k_steps;
  s.quota = total / grid;
  s.rem = (int)(total % grid);
  s.gran = 1;
  s.flags = flags;
  // (a tile can be shared by two workers whenever the ranges do not fall on tile boundaries)
  cudaError_t err = cudaMemsetAsync(flags, 0, kMlpFlagBytes, st);
  if (err != cudaSuccess) return (int)err;
  return 0;
}

// host: no flag memory available -> every CTA takes whole tiles only (contiguous blocks of tiles; no hand-over)
inline void mlp_schedule_whole_tiles(MlpSchedule& s, long long tiles, int k_steps, int grid) {
  s.quota = tiles / grid;
  s.rem = (int)(tiles % grid);
  s.gran = k_steps;
  s.flags = nullptr;
}


// host: launch a persistent kernel `kern(P, tab)`.  With the balanced split (P.sched.gran == 1) a worker spin-waits on
// its predecessor's flag, which is only safe when every CTA of the grid is resident at once -- not guaranteed when other
// kernels share the device (a gather on another stream, a second burst).  The launch is therefore cooperative: the
// driver co-schedules the whole grid or refuses, and a refusal falls back to whole tiles per worker (no hand-over).
template <class Kern, class Params, class Table>
inline cudaError_t mlp_launch_persistent(Kern kern, int grid, int threads, size_t smem, cudaStream_t st, Params& P,
                                         const Table& tab, long long tiles, int k_steps, int workers) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  const bool balanced = P.sched.gran == 1 && P.sched.flags != nullptr;
  cfg.numAttrs = balanced ? 1 : 0;
  cudaError_t err = cudaLaunchKernelEx(&cfg, kern, P, tab);
  if (err != cudaSuccess && balanced) {
    (void)cudaGetLastError();
    mlp_schedule_whole_tiles(P.sched, tiles, k_steps, workers);
    cfg.numAttrs = 0;
    err = cudaLaunchKernelEx(&cfg, kern, P, tab);
  }
  return err;
}

}  // namespace ebm
