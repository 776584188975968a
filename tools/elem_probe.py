"""Timing probe for the elementwise Langevin kernel: shard sizes of C2 (65536 / N chains), both RNG layouts, with and
without a workspace (not a bench; the balanced elementwise variant it compared was rejected)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchebm_b200 as te
from torchebm_b200 import _lib, ops

dev = torch.device("cuda:0")
for mode in ("torch", "native"):
    for n in (65536, 32768, 16384, 8192):
        desc = te.energy_descriptor(te.DoubleWellModel(2.0, 1.0), 128, dev)
        x = torch.randn(n, 128, device=dev).clamp_(-3, 3)
        out = torch.empty_like(x)
        for ws in (True, False):
            if not ws:
                desc.c.buf[6] = None
            run = lambda: ops.langevin_burst(desc, x, 500, [0.01], [1.0], rng_mode=_lib.RNG_MODES[mode], seed=1, offset=0, out=out)
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                run()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 5
            print(f"{mode:7s} n={n:6d} {'workspace' if ws else 'no-ws    '} {ms:.3f} ms  {n * 500 / ms * 1e3:.3e} chain-steps/s")
