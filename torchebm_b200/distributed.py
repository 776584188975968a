"""Chain-axis sharding across the GPUs of one box and the burst-end gather.

Chains are independent for every in-scope energy, so a K-step burst needs no communication: rank r owns
the contiguous block of rows [r*N/W, (r+1)*N/W) and a decorrelated Philox stream (`base_seed + rank`, the
convention of tests/distributed/test_generator_ranks.py:38-51 in the reference).  The only collective is
one all-gather of the `[N/W, D]` shards at the end of a burst, and only when the caller needs the global
tensor (the reference keeps PCD buffers rank-local, core/base_loss.py:131-134).  Helper semantics follow
torchebm/utils/distributed.py:30-125 (identity when not distributed).
"""

from __future__ import annotations

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
