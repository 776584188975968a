// In-burst diagnostics (samplers/langevin_dynamics.py:170-185, samplers/hmc.py:294-310): per kept sample the batch mean
// and (biased, clamped) variance of every state column and the batch-mean energy.  Sums are accumulated in fp64
// (column sums over up to millions of chains; var = E[x^2] - E[x]^2 needs the headroom) in a workspace of
// kDiagSlot(d) doubles per kept sample, and turned into the fp32 outputs by one finalize launch at the end.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ebm {

// workspace layout per kept sample: sum[d], sumsq[d], energy sum, spare
__host__ __device__ inline long long diag_slot(int d) { return 2ll * d + 2; }

struct DiagAccum {
  double* ws;   // [n_kept, diag_slot(d)] zeroed before the burst, or null
  int d;
};

}  // namespace ebm
