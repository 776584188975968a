"""Chain-axis sharding across the GPUs of one box and the burst-end gather.

Chains are independent for every in-scope energy, so a K-step burst needs no communication: rank r owns
the contiguous block of rows [r*N/W, (r+1)*N/W) and a decorrelated Philox stream (`base_seed + rank`, the
convention of tests/distributed/test_generator_ranks.py:38-51 in the reference).  The only collective is
one all-gather of the `[N/W, D]` shards at the end of a burst, and only when the caller needs the global
tensor (the reference keeps PCD buffers rank-local, core/base_loss.py:131-134).  Helper semantics follow
torchebm/utils/distributed.py:30-125 (identity when not distributed).
"""

from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized()


def get_rank(group=None) -> int:
    return dist.get_rank(group) if is_distributed() else 0


def get_world_size(group=None) -> int:
    return dist.get_world_size(group) if is_distributed() else 1


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows [lo, hi) of rank `rank`; requires n % world == 0 (equal per-rank batches, like the reference)."""
    if n % world != 0:
        raise ValueError(f"n_chains ({n}) must be divisible by the world size ({world})")
    per = n // world
    return rank * per, (rank + 1) * per


def shard_chains(x: torch.Tensor, group=None) -> torch.Tensor:
    lo, hi = shard_bounds(x.shape[0], get_rank(group), get_world_size(group))
    return x[lo:hi]


def all_gather_cat(x: torch.Tensor, group=None, dim: int = 0) -> torch.Tensor:
    world = get_world_size(group)
    if world == 1:
        return x
    x = x.detach().contiguous()
    if dim == 0:
        out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x, group=group)
        return out
    parts = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(parts, x, group=group)
    return torch.cat(parts, dim=dim)


def gather_chains(x_local: torch.Tensor, group=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Burst-end all-gather of the per-rank `[N/W, D]` shards into `[N, D]` (rank order)."""
    world = get_world_size(group)
    if world == 1:
        return x_local
    x_local = x_local.contiguous()
    if out is None:
        out = torch.empty((world * x_local.shape[0],) + tuple(x_local.shape[1:]), dtype=x_local.dtype,
                          device=x_local.device)
    dist.all_gather_into_tensor(out, x_local, group=group)
    return out


def broadcast_tensor(t: torch.Tensor, src: int = 0, group=None) -> torch.Tensor:
    if get_world_size(group) == 1:
        return t
    device = t.device
    backend = dist.get_backend(group)
    work = t.cuda() if (backend == "nccl" and not t.is_cuda) else t.clone()
    dist.broadcast(work, src=src, group=group)
    return work.to(device)


def rank_generator(base_seed: int, device, group=None) -> torch.Generator:
    """Per-rank generator `base_seed + rank` (docs/developer_guide/distributed.md:80-81 in the reference)."""
    g = torch.Generator(device=device)
    g.manual_seed(int(base_seed) + get_rank(group))
    return g


class PeerGatherBuffer:
    """The gathered `[n_total, dim]` chain tensor of a sharded burst, allocated in symmetric (peer-mapped) memory so that
    every rank's burst kernel can store its shard straight into every other rank's copy over NVLink
    (`ops.langevin_burst_gather`): the burst-end all-gather without a separate collective launch.  Where the fabric offers
    NVLS, the kernels store through the buffer's multicast address instead: one store per 16 bytes, replicated by NVSwitch
    (except the streamed-state MLP kernel, whose pusher warps move finished tiles with bulk copies to every peer mapping).

    Collective: every rank of `group` must construct it (rendezvous) with the same shape.  `barrier()` is a device-side
    cross-rank barrier on the current stream: call it after the burst before reading `tensor`, and again before the
    next burst overwrites the buffers if a reader may still be using them."""

    def __init__(self, n_total: int, dim: int, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        if not is_distributed():
            raise RuntimeError("PeerGatherBuffer needs an initialised process group")
        self.group = group if group is not None else dist.group.WORLD
        self.tensor = symm_mem.empty((n_total, dim), dtype=torch.float32, device=torch.device(device))
        self.handle = symm_mem.rendezvous(self.tensor, self.group)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        # NVLS: one multicast address that NVSwitch replicates into every rank's copy (None when the box has none, or
        # with EBM_B200_NO_MULTICAST=1 for A/B measurements); burst kernels then store once instead of once per peer
        self.mc_ptr = None
        if not os.environ.get("EBM_B200_NO_MULTICAST"):
            try:
                mc = int(self.handle.multicast_ptr)
                self.mc_ptr = mc if mc != 0 else None
            except Exception:  # noqa: BLE001  (no multicast support in this torch build / on this fabric)
                self.mc_ptr = None
        self.rank = int(self.handle.rank)
        self.world = int(self.handle.world_size)
        if n_total % self.world != 0:
            raise ValueError(f"n_total ({n_total}) must be divisible by the world size ({self.world})")
        self.rows_per_rank = n_total // self.world

    def barrier(self) -> None:
        self.handle.barrier(channel=0)

    def push(self, x_local: torch.Tensor) -> None:
        """Copy-engine gather: store this rank's finished shard into every rank's gathered tensor with peer-to-peer
        DMA copies on the current stream (no SM is used, so it overlaps a burst running on another stream), then the
        cross-rank barrier.  For bursts whose kernel has no peer-store epilogue (MLP energies)."""
        if x_local.shape[0] != self.rows_per_rank:
            raise ValueError("every rank must hold n_total / world chains")
        lo = self.rank * self.rows_per_rank
        shape, dtype = tuple(self.tensor.shape), self.tensor.dtype
        # one stream: with a stream per destination the 8-GPU gather went from 7.7 to 6.1 ms per step (several copy
        # engines at once) but the 2-GPU step degraded from 3.35 to 24.6 ms, and at 8 GPUs NCCL on spare SMs is faster
        # anyway, so the simple form is kept
        for w in range(self.world):
            dst = self.tensor if w == self.rank else self.handle.get_buffer(w, shape, dtype)
            dst[lo:lo + self.rows_per_rank].copy_(x_local, non_blocking=True)
        self.barrier()

    def push_sm(self, x_local: torch.Tensor, max_ctas: int) -> None:
        """SM-driven gather: a small kernel (`max_ctas` CTAs, to fit the SMs a persistent burst leaves free through
        `sm_margin`) stores this rank's shard into every rank's gathered tensor over NVLink, then the barrier."""
        from . import ops

        if x_local.shape[0] != self.rows_per_rank:
            raise ValueError("every rank must hold n_total / world chains")
        ops.peer_push(x_local.contiguous(), self.ptrs, self.rank * self.rows_per_rank * self.tensor.shape[1], max_ctas)
        self.barrier()

    def burst(self, desc, x_local: torch.Tensor, n_steps: int, step_sizes, noise_scales, **kw) -> torch.Tensor:
        """Run this rank's burst and push the result into every rank's gathered tensor; returns the local result.
        The gathered tensor is complete on all ranks after the barrier this method issues."""
        from . import ops

        if x_local.shape[0] != self.rows_per_rank:
            raise ValueError("every rank must hold n_total / world chains")
        out = ops.langevin_burst_gather(desc, x_local, n_steps, step_sizes, noise_scales, self.ptrs,
                                        self.rank * self.rows_per_rank, multicast_ptr=self.mc_ptr, **kw)
        self.barrier()
        return out
