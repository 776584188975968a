/*
 * ebm_b200.h -- C ABI of the B200-native MCMC negative-sampling hot path.
 *
 * Drop-in boundary for torchebm's sampler inner loops (reference = soran-ghaderi/torchebm @ a77aeee,
 * paths relative to its root).  All entry points:
 *   - take plain pointers and sizes (no torch types); `*_dev` / unmarked pointers are DEVICE pointers,
 *     `const double* ..._host` are HOST arrays that are consumed before the call returns;
 *   - are asynchronous on `stream` (a cudaStream_t passed as void*), allocate nothing, keep no state
 *     besides a per-device property cache, and are safe to call concurrently on different streams;
 *   - return 0 on success, a positive cudaError_t, or a negative EBM_ERR_* code.  `ebm_last_error()`
 *     returns a thread-local description of the last failure.
 * State is fp32, row-major contiguous [n, dim].
 */
#ifndef EBM_B200_H
#define EBM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EBM_ABI_VERSION 9

#define EBM_ERR_INVALID     (-1) /* bad argument (null pointer, non-positive size, ...) */
#define EBM_ERR_UNSUPPORTED (-2) /* valid request this build has no kernel for (e.g. dim too large) */

/* Energy kinds: torchebm/core/base_model.py (DoubleWell :130-148, Gaussian :151-210, Harmonic :213-229,
 * Rastrigin :297-316); MLP = user Sequential(Linear, act, Linear, act, Linear(.,1)) energies
 * (examples/20-training/01-mcmc-losses/01-cd-k/main.py:20-30); MoG is not in the reference (north_star). */
enum {
  EBM_ENERGY_DOUBLE_WELL = 0,
  EBM_ENERGY_HARMONIC    = 1,
  EBM_ENERGY_RASTRIGIN   = 2,
  EBM_ENERGY_GAUSSIAN    = 3,
  EBM_ENERGY_MOG         = 4,
  EBM_ENERGY_MLP         = 5
};

enum { EBM_ACT_SILU = 0, EBM_ACT_TANH = 1, EBM_ACT_RELU = 2, EBM_ACT_SOFTPLUS = 3 };

/* Where the Gaussian noise / uniforms come from.
 *   INJECTED: read from caller-provided device arrays (parity with the CPU oracle).
 *   TORCH   : Philox4x32-10 laid out exactly like torch's CUDA `randn_like` / `normal_` / `rand`
 *             (ATen/native/cuda/DistributionTemplates.h:50-91), so a burst consumes the generator
 *             stream the reference sampler would: same seed + offset => same chains.
 *   NATIVE  : Philox4x32-10, one block per aligned quad of consecutive elements (cheapest). */
enum { EBM_RNG_INJECTED = 0, EBM_RNG_TORCH = 1, EBM_RNG_NATIVE = 2 };

enum { EBM_MASS_NONE = 0, EBM_MASS_SCALAR = 1, EBM_MASS_VECTOR = 2 };

/* Arithmetic of the MLP-energy products.
 *   FP32  : fp32 FFMA on the CUDA cores (summation order differs from cuBLAS, ~1e-6 relative).
 *   BF16X3: tcgen05 tensor cores, every operand split into bf16 hi + lo, hi*hi + lo*hi + hi*lo accumulated in fp32
 *           tensor memory (~2e-5 relative); the default of the Python samplers.  Widths: D <= 128 keeps the chain
 *           on chip for the whole burst; 128 < D <= 4096 (e.g. 784-128-128-1) streams the state and W1 per step.
 *   BF16  : tcgen05 tensor cores, single bf16 pass (~4e-3 relative). */
enum { EBM_MLP_FP32 = 0, EBM_MLP_BF16X3 = 1, EBM_MLP_BF16 = 2 };

typedef struct EbmEnergyDesc {
  int32_t kind;         /* EBM_ENERGY_* */
  int32_t dim;          /* D */
  int32_t n_components; /* MoG: K */
  int32_t hidden1;      /* MLP: H1 */
  int32_t hidden2;      /* MLP: H2 */
  int32_t activation;   /* MLP: EBM_ACT_* */
  int32_t precision;    /* MLP: EBM_MLP_* (Langevin burst only; energy/gradient evaluation is always FP32) */
  int32_t sm_margin;    /* MLP: SMs the persistent burst kernels leave free (0 = use all).  A burst that runs next to
                           a collective on another stream needs >= 1: its CTAs fill every SM they get, and a barrier or
                           copy kernel of the collective would otherwise wait for the whole burst. */
  int32_t hidden3;      /* MLP: H3 > 0 = three hidden layers, E = w4 . act(W3 act(W2 act(W1 x + b1) + b2) + b3) + b4
                           (benchmarks/distributed_fsdp2.py:43-53); 0 = two hidden layers */
  int32_t reserved;     /* 0 */
  /* scalar parameters (fp32, already rounded the way torch rounds the Python doubles):
   *   DoubleWell: p[0] = barrier_height, p[1] = b*b
   *   Harmonic  : p[0] = 0.5*k
   *   Rastrigin : p[0] = a, p[1] = 2*pi, p[2] = a*D                                  */
  float p[4];
  /* device buffers:
   *   Gaussian: buf[0] = mean[D], buf[1] = cov_inv[D,D]
   *   MoG     : buf[0] = means[K,D], buf[1] = sigmas[K], buf[2] = weights[K]
   *   MLP     : buf[0] = W1[H1,D], buf[1] = b1[H1], buf[2] = W2[H2,H1], buf[3] = b2[H2],
   *             buf[4] = w3[H2], buf[5] = b3[1]            (torch [out,in] layout);
   *             three hidden layers (hidden3 > 0): buf[4] = w4[H3], buf[5] = b4[1] (the output layer),
   *             buf[7] = W3[H3,H2], buf[8] = b3[H3]; every width <= 128; the Langevin burst (tensor cores, precision
   *             BF16X3 / BF16), ebm_energy_f32, ebm_gradient_f32 and the persistent-CD entry points take it, the
   *             fused HMC kernels do not (EBM_ERR_UNSUPPORTED);
   *             buf[6] = scratch workspace of ebm_workspace_bytes() bytes, 128-byte aligned, used by the
   *             Langevin burst: hand-over flags of the balanced (tile, step-range) work split and, when D > 128,
   *             the per-call bf16 hi/lo re-split of the weights.  Required when D > 128; optional (NULL = whole
   *             tiles per SM, no balancing) when D <= 128.  One workspace must not be shared by bursts running
   *             concurrently on different streams.   */
  const float* buf[10];
} EbmEnergyDesc;

int         ebm_abi_version(void);
const char* ebm_last_error(void);

/* Number of SMs and T = 256 * min(SMs * maxThreadsPerSM/256, ceil(numel/256)) of torch's distribution
 * kernels on `device`; the generator offset a torch-layout draw of `numel` elements consumes. */
int     ebm_device_sm_count(int device);
int64_t ebm_torch_rng_threads(int device, int64_t numel);
int64_t ebm_torch_rng_offset_increment(int device, int64_t numel);

/* Bytes of device scratch the Langevin burst needs in e->buf[6] (0 when the energy needs none). */
int64_t ebm_workspace_bytes(const EbmEnergyDesc* e);

/* E(x) -> energy[n]; replaces `model(x)` (BaseModel.forward, base_model.py:49-60). */
int ebm_energy_f32(const EbmEnergyDesc* e, const float* x, int64_t n, float* energy, void* stream);

/* grad_x E(x) -> grad[n, dim]; replaces `BaseModel.gradient` (base_model.py:62-127). */
int ebm_gradient_f32(const EbmEnergyDesc* e, const float* x, int64_t n, float* grad, void* stream);

/* One Euler-Maruyama update with a caller-supplied drift (integrator-level boundary):
 * out = (x + h*drift) + c2*(noise*c1), c1 = (float)h^0.5, c2 = (float)(2 ns^2)^0.5.
 * Replaces BaseSDERungeKuttaIntegrator.step for the EM tableau (core/base_integrator.py:673-731).
 * `noise` may be NULL when noise_scale < 0 (pure ODE step). */
int ebm_euler_maruyama_step_f32(const float* x, const float* drift, const float* noise, float* out,
                                int64_t numel, double step_size, double noise_scale, void* stream);

/* K-step Langevin burst; replaces the loop of LangevinDynamics.sample
 * (samplers/langevin_dynamics.py:157-185) including gradient, EM update, noise draw, optional clamp
 * and trajectory thinning, with the chain state resident in registers between steps.
 *   step_size_host / noise_scale_host: `schedule_len` doubles each; schedule_len is 1 (constant) or n_steps.
 *   clamp_lo_hi_host: NULL or 2 floats.
 *   rng_mode INJECTED: `noise` = [n_steps, n, dim]; otherwise (seed, offset) address the Philox stream;
 *     TORCH consumes n_steps * ebm_torch_rng_offset_increment(n*dim), NATIVE consumes 4 * n_steps.
 *   traj: NULL or [n, n_steps/thin, dim]; x_out may alias x_in. */
int ebm_langevin_burst_f32(const EbmEnergyDesc* e, const float* x_in, float* x_out, int64_t n,
                           int32_t n_steps, const double* step_size_host, const double* noise_scale_host,
                           int32_t schedule_len, const float* clamp_lo_hi_host, int32_t rng_mode,
                           uint64_t seed, uint64_t offset, const float* noise, float* traj, int32_t thin,
                           void* stream);

/* Same burst with the Heun (improved Euler) SDE scheme: LangevinDynamics(integrator="heun"), i.e. the generic
 * Runge-Kutta path of core/base_integrator.py:300-347,387-397,673-731 with the tableau of integrators/heun.py
 * (two gradient evaluations per step, same additive noise and generator consumption as Euler-Maruyama).  Fused for
 * the elementwise energies; other energies return EBM_ERR_UNSUPPORTED. */
int ebm_langevin_heun_burst_f32(const EbmEnergyDesc* e, const float* x_in, float* x_out, int64_t n, int32_t n_steps,
                                const double* step_size_host, const double* noise_scale_host, int32_t schedule_len,
                                const float* clamp_lo_hi_host, int32_t rng_mode, uint64_t seed, uint64_t offset,
                                const float* noise, float* traj, int32_t thin, void* stream);

/* Noise-free descent burst; replaces the loops of GradientDescentSampler.sample (x <- x - eta * grad E(x),
 * samplers/gradient_descent.py:123-138) and NesterovSampler.sample (lookahead gradient + momentum, :258-276) for the
 * elementwise energies (others: EBM_ERR_UNSUPPORTED).  momentum < 0 selects plain gradient descent, 0 <= momentum < 1
 * Nesterov (the velocity starts at 0).  velocity: NULL or [n, dim] scratch that receives the final velocity (required
 * for per-step schedules longer than 64 steps).  traj / thin / schedule_len as in ebm_langevin_burst_f32. */
int ebm_descent_burst_f32(const EbmEnergyDesc* e, const float* x_in, float* x_out, int64_t n, int32_t n_steps,
                          const double* step_size_host, int32_t schedule_len, double momentum, float* velocity, float* traj,
                          int32_t thin, void* stream);

/* Same burst on one rank of a box whose chains are sharded over `world` GPUs, with the burst-end all-gather (the one
 * collective of the sharded path) fused into the kernel's final store: besides x_out[n, dim] the final state is written
 * at rows [row_offset, row_offset + n) of EVERY rank's gathered buffer.  peer_out_host[w] (host array of `world` device
 * pointers, world <= 16) is rank w's gathered [n_total, dim] buffer as mapped into THIS process (peer / symmetric
 * memory; the own rank's entry is its local buffer).  Remote stores travel over NVLink from inside the burst kernel's
 * final state store (elementwise energies; MLP energies on the tensor-core kernels, i.e. ebm_pcd_langevin_fused(e) != 0)
 * or as copy-engine pushes (other energies).  The caller runs a cross-rank barrier on the stream afterwards, before any
 * rank reads its gathered buffer.  Replaces the all_gather of the negatives (utils/distributed.py:43-70).
 * NVLS: pass world = -W and W + 1 pointers, the last one the MULTICAST address of the gathered buffers (symmetric
 * memory's multicast mapping): kernels with a peer-store epilogue then issue ONE multimem.st per 16 bytes, which NVSwitch
 * replicates into all W copies, instead of W stores; the others keep pushing to the W unicast pointers.
 * MLP energies with dim > 128 (the streamed-state kernel): two otherwise idle warps per SM move every finished tile
 * with 8 KB bulk copies, x_out -> shared memory -> one bulk store per unicast pointer (measured best at 2 and 8 GPUs), and
 * ignore the multicast pointer; the environment variable EBM_B200_PUSH_BULK=0 selects 16-byte stores instead
 * (through the multicast address when one was passed). */
int ebm_langevin_burst_gather_f32(const EbmEnergyDesc* e, const float* x_in, float* x_out, int64_t n, int32_t n_steps,
                                  const double* step_size_host, const double* noise_scale_host, int32_t schedule_len,
                                  const float* clamp_lo_hi_host, int32_t rng_mode, uint64_t seed, uint64_t offset,
                                  float* const* peer_out_host, int32_t world, int64_t row_offset, void* stream);

/* Burst-end gather for bursts whose kernel has no peer-store epilogue: store src[numel] at element offset `elem_offset`
 * of every rank's gathered buffer (peer_out_host as in ebm_langevin_burst_gather_f32) with at most `max_ctas` CTAs of
 * 1024 threads, so that it can run on the SMs a persistent burst leaves free (EbmEnergyDesc.sm_margin) on another
 * stream.  Pointers and offset 16-byte aligned.  A cross-rank barrier must follow. */
int ebm_peer_push_f32(const float* src, int64_t numel, float* const* peer_out_host, int32_t world, int64_t elem_offset,
                      int32_t max_ctas, void* stream);

/* Same burst with HOST buffers: copies x_in_host -> device scratch, runs the burst, copies the result
 * back into x_out_host and synchronises `stream`.  `scratch_dev` must hold n*dim floats.
 * Energy parameter buffers in `e` stay device pointers.  Used for the end-to-end measurement. */
int ebm_langevin_burst_host_f32(const EbmEnergyDesc* e, const float* x_in_host, float* x_out_host,
                                float* scratch_dev, int64_t n, int32_t n_steps, double step_size,
                                double noise_scale, int32_t rng_mode, uint64_t seed, uint64_t offset,
                                void* stream);

/* L leapfrog steps; replaces LeapfrogIntegrator.integrate (integrators/leapfrog.py:116-187) for a
 * recognised energy (drift = -grad E).  safe != 0 applies the +-1e6 force clamp and NaN->0 / inf->FLT_MAX
 * sanitising of core/base_integrator.py:875-889.  mass_kind/mass_scalar/mass_vec follow leapfrog.py:167-177. */
int ebm_leapfrog_f32(const EbmEnergyDesc* e, const float* x_in, const float* p_in, float* x_out,
                     float* p_out, int64_t n, int32_t n_steps, double step_size, int32_t mass_kind,
                     double mass_scalar, const float* mass_vec, int32_t safe, void* stream);

/* n_proposals HMC proposals; replaces the loop of HamiltonianMonteCarlo.sample (samplers/hmc.py:244-312):
 * momentum draw, H0, L leapfrog steps (safe mode), H1, Metropolis accept/reject.
 *   step_size_host: schedule_len (1 or n_proposals) doubles.
 *   INJECTED: noise_p = [n_proposals, n, dim] standard normals, noise_u = [n_proposals, n] uniforms.
 *   TORCH consumes per proposal increment(n*dim) + increment(n); NATIVE consumes 8 per proposal.
 *   traj: NULL or [n, n_proposals/thin, dim].
 *   accept_count: NULL or int32[n_proposals], += number of accepted chains per proposal (zero it first).
 *   energy_out: NULL or [n], energy of the final state (clamped to +-1e10 like hmc.py:247). */
int ebm_hmc_burst_f32(const EbmEnergyDesc* e, const float* x_in, float* x_out, int64_t n,
                      int32_t n_proposals, int32_t n_leapfrog, const double* step_size_host,
                      int32_t schedule_len, int32_t mass_kind, double mass_scalar, const float* mass_vec,
                      int32_t rng_mode, uint64_t seed, uint64_t offset, const float* noise_p,
                      const float* noise_u, float* traj, int32_t thin, int32_t* accept_count,
                      float* energy_out, void* stream);

/* Persistent-CD replay buffer (core/base_loss.py:266-337, :390-426), device side.
 * gather : out[i, :] = buffer[idx[i], :] (+ 0.01 * noise[j, :] for i = noise_rows[j], j < n_noise).
 * scatter: FIFO write of `samples` at row `ptr` with wraparound; when batch >= buffer_rows the last
 *          buffer_rows samples overwrite the whole buffer.  Returns the new ptr through *new_ptr_host. */
int ebm_pcd_gather_f32(const float* buffer, int64_t buffer_rows, int64_t row_elems, const int64_t* idx,
                       int64_t batch, float* out, const int64_t* noise_rows, const float* noise,
                       int64_t n_noise, void* stream);
int ebm_pcd_scatter_f32(float* buffer, int64_t buffer_rows, int64_t row_elems, int64_t ptr,
                        const float* samples, int64_t batch, int64_t* new_ptr_host, void* stream);

/* Persistent-CD negative sampling in one call (the sampling half of ContrastiveDivergence.forward,
 * losses/contrastive_divergence.py:127-139): start points buffer[idx[i], :] (get_start_points, core/base_loss.py:293-314)
 * plus the exploration noise 0.01 * noise[j, :] on chain noise_rows[j] (:317-332; n_noise may be 0), n_steps Langevin
 * steps into x_out[n, dim], FIFO write-back of x_out into the buffer at `ptr` (update_buffer, :390-426), and, when
 * energy_out != NULL, E(x-) of the negatives into energy_out[n].
 *   idx == NULL states that chain i starts from row i (the reference's stratified draw at stride 1).  With
 *   n == buffer_rows the write-back then replaces the whole buffer, and for energies with ebm_pcd_langevin_fused(e) != 0
 *   the burst kernel reads its start rows from the buffer and writes its final state to both destinations: no gather or
 *   scatter pass; the noise is added in place to the noised rows first.  Any explicit idx runs gather -> burst ->
 *   scatter and needs `scratch` [n, dim].  Row length of the buffer = e->dim.  *new_ptr_host receives the new ptr. */
int ebm_pcd_langevin_fused(const EbmEnergyDesc* e);
int ebm_pcd_langevin_burst_f32(const EbmEnergyDesc* e, float* buffer, int64_t buffer_rows, const int64_t* idx,
                               int64_t ptr, float* x_out, float* scratch, int64_t n, int32_t n_steps,
                               const double* step_size_host, const double* noise_scale_host, int32_t schedule_len,
                               const float* clamp_lo_hi_host, int32_t rng_mode, uint64_t seed, uint64_t offset,
                               const int64_t* noise_rows, const float* noise, int64_t n_noise, float* energy_out,
                               int64_t* new_ptr_host, void* stream);
/* The same call on one rank of a chain-sharded box (BASELINE config 5): the negatives also land at rows
 * [row_offset, row_offset + n) of every rank's gathered buffer, from inside the burst kernel's final store where it has
 * a peer-store epilogue.  peer_out_host / world / row_offset and the barrier rule as in ebm_langevin_burst_gather_f32. */
int ebm_pcd_langevin_burst_gather_f32(const EbmEnergyDesc* e, float* buffer, int64_t buffer_rows, const int64_t* idx,
                                      int64_t ptr, float* x_out, float* scratch, int64_t n, int32_t n_steps,
                                      const double* step_size_host, const double* noise_scale_host, int32_t schedule_len,
                                      const float* clamp_lo_hi_host, int32_t rng_mode, uint64_t seed, uint64_t offset,
                                      const int64_t* noise_rows, const float* noise, int64_t n_noise, float* energy_out,
                                      int64_t* new_ptr_host, float* const* peer_out_host, int32_t world,
                                      int64_t row_offset, void* stream);

/* Bursts that also produce the reference's per-kept-sample diagnostics (return_diagnostics=True:
 * samplers/langevin_dynamics.py:170-185, samplers/hmc.py:294-310) in the same call:
 *   diag_mean / diag_var [n_kept, dim]: batch mean and biased variance (clamped to [1e-10, 1e10]; 0 for one chain),
 *   diag_energy [n_kept]: batch mean of E(x) (HMC: of the energies clamped to +-1e10),
 *   diag_accept [n_kept] (HMC): acceptance fraction of the kept proposal;  n_kept = n_steps / thin >= 1.
 * diag_ws: n_kept * (2*dim + 2) doubles of device scratch (sums are accumulated in fp64); scratch: n floats;
 * accept_count (HMC): n_proposals int32 of device scratch.  Elementwise energies accumulate inside the Langevin burst
 * kernel (one launch + finalize); every other case runs one sub-burst per kept sample followed by the energy and
 * column-statistics kernels, still inside this one call.  Other arguments as in the plain bursts; heun != 0 selects
 * the Heun scheme. */
int ebm_langevin_burst_diag_f32(const EbmEnergyDesc* e, const float* x_in, float* x_out, int64_t n, int32_t n_steps,
                                const double* step_size_host, const double* noise_scale_host, int32_t schedule_len,
                                const float* clamp_lo_hi_host, int32_t rng_mode, uint64_t seed, uint64_t offset,
                                const float* noise, float* traj, int32_t thin, int32_t heun, double* diag_ws,
                                float* scratch, float* diag_mean, float* diag_var, float* diag_energy, void* stream);
int ebm_hmc_burst_diag_f32(const EbmEnergyDesc* e, const float* x_in, float* x_out, int64_t n, int32_t n_proposals,
                           int32_t n_leapfrog, const double* step_size_host, int32_t schedule_len, int32_t mass_kind,
                           double mass_scalar, const float* mass_vec, int32_t rng_mode, uint64_t seed, uint64_t offset,
                           const float* noise_p, const float* noise_u, float* traj, int32_t thin, double* diag_ws,
                           float* scratch, int32_t* accept_count, float* diag_mean, float* diag_var, float* diag_energy,
                           float* diag_accept, void* stream);

/* Effective sample size of n_chains chains of n samples each (chains[n_chains, n], contiguous): replaces
 * `_ess_from_chain` of the reference's benchmark harness (benchmarks/registry.py:348-365), which the harness applies to
 * the energy chain of a sampler's diagnostics (:766-770).  Same estimator (centred chain, autocovariances, lags summed up
 * to the first negative one, ESS = n / max(1 + 2 sum rho_k, 1); n for chains shorter than 2 or of zero variance); the
 * autocovariances are direct fp64 sums instead of an fp32 FFT.  ess_out[n_chains]. */
int ebm_ess_f32(const float* chains, int64_t n_chains, int64_t n, float* ess_out, void* stream);

/* Fill out[numel] with the TORCH- or NATIVE-layout normal (kind 0) / uniform (kind 1) stream at
 * (seed, offset): test hook that exposes exactly what the fused kernels draw. */
int ebm_rng_fill_f32(float* out, int64_t numel, int32_t rng_mode, int32_t kind, uint64_t seed,
                     uint64_t offset, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EBM_B200_H */
