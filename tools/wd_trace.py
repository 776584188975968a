"""Timeline of CTA 0 of langevin_mlp_wide_kernel (tuning only; needs the trace build:
    python -m torchebm_b200.build --variant trace EBM_WD_TRACE
    EBM_B200_LIB=torchebm_b200/lib/libebm_b200_trace.so python tools/wd_trace.py [d] [k] [tiles] > gpurun_out/wd_trace.txt
Prints, per role, the records (cycle relative to the first record, tag kind, index)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchebm_b200 as te
from torchebm_b200 import _lib, ops

dev = torch.device("cuda:0")
d = int(sys.argv[1]) if len(sys.argv) > 1 else 784
k = int(sys.argv[2]) if len(sys.argv) > 2 else 20
tiles = int(sys.argv[3]) if len(sys.argv) > 3 else 512
torch.manual_seed(0)
model = te.MLPEnergy(dim=d, hidden=128, activation="silu").to(dev)
desc = te.energy_descriptor(model, d, dev)
n = tiles * 128
x = torch.randn(n, d, device=dev).clamp_(-3, 3)
out = torch.empty_like(x)
for _ in range(3):
    ops.langevin_burst(desc, x, k, [0.01], [1.0], rng_mode=_lib.RNG_NATIVE, seed=1, offset=0, out=out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
ops.langevin_burst(desc, x, k, [0.01], [1.0], rng_mode=_lib.RNG_NATIVE, seed=1, offset=0, out=out)
b.record()
torch.cuda.synchronize()
print(f"# d={d} k={k} tiles={tiles} ms={a.elapsed_time(b):.3f}")
lib = _lib.load()
L = 8192
buf = (C.c_uint64 * (4 * L))()
fn = lib.ebm_debug_wd_trace
fn.argtypes = [C.POINTER(C.c_uint64), C.c_int]
rc = fn(buf, 4 * L)
assert rc == 0, rc
recs = []
for role in range(4):
    for i in range(L):
        v = buf[role * L + i]
        if v == 0:
            break
        recs.append((v >> 16, role, (v & 0xffff) >> 8, v & 0xff))
recs.sort()
t0 = recs[0][0]
names = {0: {1: "step", 2: "W2rdy", 3: "G2iss", 4: "G3iss", 5: "d1all", 6: "W1_0", 7: "xa_c", 8: "ring_c+2", 9: "G1_issued", 10: "G4_issued"},
         1: {1: "E1wait", 2: "E1go", 3: "E2wait", 4: "E2go", 5: "E3wait", 6: "E3go", 7: "chunk", 8: "noise_done", 9: "G_rdy", 10: "upd_done", 11: "xa_free", 12: "sts_done", 13: "published"},
         3: {1: "load"}}
names[2] = names[1]
lim = int(os.environ.get("WD_TRACE_MAX", "1500"))
for t, role, kind, idx in recs[:lim]:
    print(f"{t - t0:9d} {'MMA EP0 EP1 PRD'.split()[role]} {names[role].get(kind, kind)} {idx}")
