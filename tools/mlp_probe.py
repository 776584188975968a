"""Timing probe for the MLP Langevin kernels: tile-step time versus number of busy SMs / tiles (not a bench)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchebm_b200 as te
from torchebm_b200 import _lib, ops

dev = torch.device("cuda:0")
d = int(sys.argv[1]) if len(sys.argv) > 1 else 784
k = int(sys.argv[2]) if len(sys.argv) > 2 else 20
torch.manual_seed(0)
model = te.MLPEnergy(dim=d, hidden=128, activation="silu").to(dev)
desc = te.energy_descriptor(model, d, dev)
for tiles in [int(a) for a in (sys.argv[3] if len(sys.argv) > 3 else "37,74,148,296,444,512,592").split(",")]:
    n = tiles * 128
    x = torch.randn(n, d, device=dev).clamp_(-3, 3)
    out = torch.empty_like(x)
    for _ in range(3):
        ops.langevin_burst(desc, x, k, [0.01], [1.0], rng_mode=_lib.RNG_NATIVE, seed=1, offset=0, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    reps = 5
    for _ in range(reps):
        ops.langevin_burst(desc, x, k, [0.01], [1.0], rng_mode=_lib.RNG_NATIVE, seed=1, offset=0, out=out)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    rounds = max(1.0, tiles / 148)
    print(f"d={d} k={k} tiles={tiles} ms={ms:.3f} us_per_tile_step_per_SM={ms * 1e3 / (rounds * k):.2f} chain-steps/s={n * k / ms * 1e3:.3e}")
