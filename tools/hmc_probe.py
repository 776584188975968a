"""Timing probe for the fused HMC kernels on an MLP energy (not a bench): per-proposal time vs L and RNG layout."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchebm_b200 as te
from torchebm_b200 import _lib, ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
n, d = 148 * 128, 128
x = torch.randn(n, d, device=dev)
out = torch.empty_like(x)
for prec in ("bf16x3", "fp32"):
    model = te.MLPEnergy(dim=d, hidden=128, activation="silu", precision=prec).to(dev)
    desc = te.energy_descriptor(model, d, dev)
    for mode in ("torch", "native"):
        for L in (5, 10, 20):
            run = lambda: ops.hmc_burst(desc, x, 2, L, [0.05], rng_mode=_lib.RNG_MODES[mode], seed=1, offset=0, out=out)
            for _ in range(2):
                run()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                run()
            b.record()
            torch.cuda.synchronize()
            us = a.elapsed_time(b) / 3 / 2 * 1e3
            print(f"{prec:7s} {mode:7s} L={L:2d}  {us:8.1f} us per proposal (one tile per SM)  {us / (L + 1):6.2f} us per evaluation")
