"""N > 1 host logic on CPU: two real gloo processes (no mocks), like the reference's
tests/distributed/dist_harness.py.  Covers the chain-axis shard/gather and the persistent-CD cross-rank mix
(tests/distributed/test_pcd_buffer_ranks.py:60-105 of the reference: exact partition, rank-0 seed fixes the
permutation, FIFO pointer untouched)."""

import os
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

WORLD = 2


def _worker(rank, world, store_path, out_dir):
    os.environ["CUDA_VISIBLE_DEVICES"] = ""
    store = dist.FileStore(store_path, world)
    dist.init_process_group("gloo", store=store, rank=rank, world_size=world)
    import torchebm_b200 as te
    from torchebm_b200 import distributed as D

    res = {}
    # shard / gather round trip
    full = torch.arange(8 * 3, dtype=torch.float32).reshape(8, 3)
    local = D.shard_chains(full)
    res["bounds"] = D.shard_bounds(8, rank, world)
    res["gathered_equal"] = torch.equal(D.gather_chains(local * 1.0), full)
    res["gathered_cat_equal"] = torch.equal(D.all_gather_cat(local), full)
    try:
        D.shard_bounds(7, rank, world)
        res["uneven_raises"] = False
    except ValueError:
        res["uneven_raises"] = True
    # per-rank generators decorrelate, shared seed reproduces
    res["rank_seed"] = D.rank_generator(100, "cpu").initial_seed()
    # persistent-CD mix: pooled chains are re-dealt with none lost or duplicated
    m = te.DoubleWellModel()
    cd = te.ContrastiveDivergence(m, te.LangevinDynamics(m), k_steps=1, persistent=True, buffer_size=6, init_steps=0)
    cd.initialize_buffer((2,))
    cd.replay_buffer.copy_(torch.arange(12, dtype=torch.float32).reshape(6, 2) + 100 * rank)
    cd._buffer_ptr_int = 3
    cd.buffer_ptr.fill_(3)
    gen = torch.Generator().manual_seed(1234 if rank == 0 else 999)  # only rank 0's seed may matter
    cd.mix_buffer_across_ranks(generator=gen)
    res["mixed"] = cd.replay_buffer.clone()
    res["ptr"] = (cd._buffer_ptr_int, int(cd.buffer_ptr))
    torch.save(res, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_shard_gather_and_pcd_mix():
    with tempfile.TemporaryDirectory() as tmp:
        store = os.path.join(tmp, "store")
        mp.start_processes(_worker, args=(WORLD, store, tmp), nprocs=WORLD, join=True, start_method="spawn")
        r = [torch.load(os.path.join(tmp, f"rank{i}.pt")) for i in range(WORLD)]
    assert r[0]["bounds"] == (0, 4) and r[1]["bounds"] == (4, 8)
    for x in r:
        assert x["gathered_equal"] and x["gathered_cat_equal"] and x["uneven_raises"]
        assert x["ptr"] == (3, 3)
    assert r[0]["rank_seed"] == 100 and r[1]["rank_seed"] == 101
    pooled = torch.cat([torch.arange(12, dtype=torch.float32).reshape(6, 2) + 100 * k for k in range(WORLD)])
    mixed = torch.cat([r[0]["mixed"], r[1]["mixed"]])
    perm = torch.randperm(12, generator=torch.Generator().manual_seed(1234))
    assert torch.equal(mixed, pooled[perm])  # exact partition under rank 0's permutation
    assert sorted(mixed[:, 0].tolist()) == sorted(pooled[:, 0].tolist())
