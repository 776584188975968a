"""Effective sample size of a 1-D chain: CPU restatement of `_ess_from_chain`, benchmarks/registry.py:348-365 of the
reference (FFT autocorrelation, initial positive sequence: lags are summed until the first negative autocorrelation).

TEST INFRASTRUCTURE ONLY -- nothing under torchebm_b200/ imports this module.  Pinned against
tests/golden/ess_chains.npz, produced by the unmodified reference function (tests/golden/make_golden_ess.py)."""

import numpy as np
import torch


def ess_from_chain(chain: torch.Tensor) -> float:
    """registry.py:348-365, operation for operation (fp32 rfft of the centred chain zero-padded to 2n)."""
    n = len(chain)
    if n < 2:                                   # :351-352
        return float(n)
    x = chain - chain.mean()                    # :353
    fft_x = torch.fft.rfft(x, n=2 * n)          # :354
    acf = torch.fft.irfft(fft_x * fft_x.conj(), n=2 * n)[:n]   # :355
    if acf[0].item() == 0:                      # :356-357
        return float(n)
    acf = acf / acf[0]                          # :358
    total = 0.0
    for i in range(1, n):                       # :360-363
        if acf[i].item() < 0:
            break
        total += acf[i].item()
    tau = 1.0 + 2.0 * total                     # :364
    return n / max(tau, 1.0)                    # :365


def ess_direct(chain) -> float:
    """The same estimator with the autocovariances as direct fp64 sums (what the CUDA kernel computes): identical in
    exact arithmetic, used to bound the fp32-FFT rounding of the reference form in the tests."""
    x = np.asarray(chain, dtype=np.float64)
    n = x.shape[0]
    if n < 2:
        return float(n)
    x = x - x.mean()
    c0 = float(np.dot(x, x))
    if c0 == 0.0:
        return float(n)
    total = 0.0
    for k in range(1, n):
        ck = float(np.dot(x[: n - k], x[k:]))
        if ck < 0:
            break
        total += ck / c0
    return n / max(1.0 + 2.0 * total, 1.0)
