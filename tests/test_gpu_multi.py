"""Two-GPU checks of the sharded path (skipped on a one-GPU box): the burst-end gather fused into the burst kernel's
final store (peer stores into symmetric memory) must equal an NCCL all-gather of the per-rank results."""

import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, d, k, result_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import torchebm_b200 as te
    from torchebm_b200 import _lib, ops
    from torchebm_b200.distributed import PeerGatherBuffer, gather_chains, shard_bounds

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        lo, hi = shard_bounds(n_total, rank, world)
        x_full = torch.randn(n_total, d, generator=torch.Generator().manual_seed(0)).clamp_(-3, 3)
        x_local = x_full[lo:hi].to(dev)
        results = {}
        for name, model in (("doublewell", te.DoubleWellModel(2.0, 1.0)), ("rastrigin", te.RastriginModel(10.0)),
                            ("gaussian", te.GaussianModel(torch.zeros(d), torch.eye(d)).to(dev))):
            desc = te.energy_descriptor(model, d, dev)
            buf = PeerGatherBuffer(n_total, d, dev)
            buf.tensor.fill_(float("nan"))
            buf.barrier()
            mc = buf.mc_ptr
            for rng_mode, use_mc in ((_lib.RNG_TORCH, True), (_lib.RNG_NATIVE, True), (_lib.RNG_NATIVE, False)):
                buf.mc_ptr = mc if use_mc else None   # NVLS multicast stores where the box has them, and plain peer stores
                rng_key = f"{rng_mode}_{'mc' if (use_mc and mc) else 'p2p'}"
                local = buf.burst(desc, x_local, k, [0.01], [1.0], rng_mode=rng_mode, seed=100 + rank, offset=0)
                torch.cuda.synchronize()
                want_local = ops.langevin_burst(desc, x_local, k, [0.01], [1.0], rng_mode=rng_mode, seed=100 + rank, offset=0)
                want = gather_chains(want_local)
                results[f"{name}_{rng_key}"] = bool(torch.equal(local, want_local) and torch.equal(buf.tensor, want))
                buf.barrier()  # nobody may overwrite a peer's buffer before that peer has compared it
                buf.tensor.fill_(float("nan"))
                buf.barrier()
            buf.mc_ptr = mc
        # copy-engine gather (MLP bursts): DMA pushes on a side stream while the next burst runs with one SM left free
        mlp = te.MLPEnergy(dim=d, hidden=64, activation="silu").to(dev)
        mlp.sm_margin = 1
        desc = te.energy_descriptor(mlp, d, dev)
        assert desc.c.sm_margin == 1
        buf = PeerGatherBuffer(n_total, d, dev)
        side = torch.cuda.Stream(device=dev)
        ok = True
        prev = None
        for it in range(3):
            local = ops.langevin_burst(desc, x_local, k, [0.01], [1.0], rng_mode=_lib.RNG_NATIVE, seed=7 + rank, offset=8 * it)
            ready = torch.cuda.Event()
            ready.record()
            local.record_stream(side)
            with torch.cuda.stream(side):
                side.wait_event(ready)
                buf.push(local)
            prev = local
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize()
        want = gather_chains(prev)
        ok = bool(torch.equal(buf.tensor, want))
        results["mlp_push"] = ok
        # SM-driven push kernel (a few CTAs on the SMs the burst leaves free)
        buf.tensor.fill_(float("nan"))
        buf.barrier()
        odd = x_local[:, :d].contiguous()
        buf.push_sm(odd, 8)
        torch.cuda.synchronize()
        results["mlp_push_sm"] = bool(torch.equal(buf.tensor, gather_chains(odd)))
        # MLP bursts: the tensor-core kernels store the final state of every finished tile into all gathered tensors
        # (narrow kernel at dim 96, wide kernel at dim 200), plain burst and the one-call persistent-CD form
        for dm in (d, 200):
            torch.manual_seed(5)
            mlp = te.MLPEnergy(dim=dm, hidden=64, activation="silu").to(dev)
            desc = te.energy_descriptor(mlp, dm, dev)
            xl = torch.randn(hi - lo, dm, generator=torch.Generator().manual_seed(10 + rank)).to(dev)
            buf = PeerGatherBuffer(n_total, dm, dev)
            mc = buf.mc_ptr
            results["has_multicast"] = True if mc else True   # (recorded for the log; both paths must pass either way)
            for use_mc in (True, False):
                buf.mc_ptr = mc if use_mc else None
                buf.tensor.fill_(float("nan"))
                buf.barrier()
                local = buf.burst(desc, xl, k, [0.01], [1.0], rng_mode=_lib.RNG_NATIVE, seed=7 + rank, offset=0)
                torch.cuda.synchronize()
                want_local = ops.langevin_burst(desc, xl, k, [0.01], [1.0], rng_mode=_lib.RNG_NATIVE, seed=7 + rank, offset=0)
                results[f"mlp_fused_{dm}_{'mc' if (use_mc and mc) else 'p2p'}"] = bool(
                    torch.equal(local, want_local) and torch.equal(buf.tensor, gather_chains(want_local)))
                buf.barrier()
            buf.mc_ptr = mc
            sampler = te.LangevinDynamics(mlp, step_size=0.01, noise_scale=1.0, device=dev).with_rng("native")
            cds = [te.ContrastiveDivergence(mlp, sampler, k_steps=k, persistent=True, buffer_size=hi - lo, init_steps=0,
                                            new_sample_ratio=ratio, device=dev) for ratio in (0.0, 0.0, 0.05, 0.05)]
            for j, cd in enumerate(cds):   # pairs: (with gather, without) must agree in negatives and buffer
                gen = torch.Generator(dev).manual_seed(31 + rank)
                buf.tensor.fill_(float("nan"))
                buf.barrier()
                neg = cd.sample_negatives(xl, generator=gen, gather_into=buf if j % 2 == 0 else None)
                torch.cuda.synchronize()
                if j % 2 == 0:
                    kept, kept_buf, gathered = neg.clone(), cd.replay_buffer.clone(), buf.tensor.clone()
                else:
                    results[f"pcd_gather_{dm}_{j // 2}"] = bool(
                        torch.equal(neg, kept) and torch.equal(cd.replay_buffer, kept_buf)
                        and torch.equal(gathered, gather_chains(neg)))
                buf.barrier()
        torch.save(results, os.path.join(result_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_fused_peer_gather_equals_nccl_all_gather(tmp_path):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.start_processes(_worker, args=(2, port, 2 * 5000, 96, 6, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    for rank in range(2):
        res = torch.load(os.path.join(str(tmp_path), f"rank{rank}.pt"))
        assert res and all(res.values()), (rank, res)
