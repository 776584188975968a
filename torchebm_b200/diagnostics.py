"""Sampler quality diagnostics computed on the device.

`ess_from_chain` / `ess_from_diagnostics` replace the effective-sample-size step of the reference's benchmark harness
(benchmarks/registry.py:348-365 and :760-770): there the energy chain of `return_diagnostics=True` goes to the host,
through an FFT autocorrelation and a Python walk with one `.item()` per lag; here the chain stays where the burst left
it and one kernel (`ebm_ess_f32`) returns the number.
"""

from __future__ import annotations

import torch

from . import ops


def ess_from_chain(chain: torch.Tensor) -> torch.Tensor:
    """ESS of a 1-D chain (registry.py:348-365), or of every row of `[n_chains, n]`; a CUDA float32 tensor, no sync."""
    return ops.ess(chain.to(torch.float32))


def ess_from_diagnostics(diagnostics) -> torch.Tensor:
    """ESS of the energy chain of a sampler's diagnostics (registry.py:766-770): the `"energy"` entry of the dict the
    samplers return with `return_diagnostics=True` (samplers/langevin_dynamics.py:170-185, one batch-mean energy per kept
    sample), or, for the stacked tensor layout `[n_kept, 3 or 4, ...]` the harness indexes, `diagnostics[:, 2, 0, 0]`."""
    if isinstance(diagnostics, dict):
        chain = diagnostics["energy"]
    else:
        if diagnostics.dim() < 2 or diagnostics.shape[1] < 3:
            raise ValueError("diagnostics must be the sampler's dict or a [n_kept, >= 3, ...] tensor")
        chain = diagnostics[:, 2]
    while chain.dim() > 1:
        chain = chain[:, 0]
    return ess_from_chain(chain.contiguous())
