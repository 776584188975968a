"""2+-rank probe: does torch symmetric memory (peer-mapped buffers + device barrier) work on this box?  (not a test)"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, d = 1024, 128
t = symm_mem.empty((world * n, d), dtype=torch.float32, device=torch.device("cuda", local))
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "multicast", hdl.has_multicast_support(local and 0 or 0, local) if False else None, flush=True)
try:
    print(rank, "multicast_ptr", hex(hdl.multicast_ptr), flush=True)
except Exception as e:  # noqa: BLE001
    print(rank, "no multicast:", repr(e)[:100], flush=True)
t.zero_()
hdl.barrier(channel=0)
mine = torch.full((n, d), float(rank + 1), device=t.device)
for peer in range(world):
    buf = hdl.get_buffer(peer, (world * n, d), torch.float32)
    buf[rank * n:(rank + 1) * n].copy_(mine)
hdl.barrier(channel=0)
torch.cuda.synchronize()
want = torch.arange(1, world + 1, device=t.device, dtype=torch.float32).repeat_interleave(n)
print(rank, "gathered ok:", bool((t[:, 0] == want).all()), flush=True)
dist.destroy_process_group()
