"""Tensor-level wrappers over the C ABI (include/ebm_b200.h).  PyTorch is plumbing here: it owns the
device memory and the stream; every computation below happens in libebm_b200.so.

All functions require fp32 CUDA tensors and raise otherwise: there is no CPU or eager fallback.
"""

from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from .core import EnergyDescriptor


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _req(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: torchebm_b200 has no CPU path")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _bind_workspace(desc: EnergyDescriptor, device) -> None:
    """The MLP workspace (hand-over flags of the balanced work split; for wide states the re-split weights) must not be
    shared by bursts that run concurrently, so it is keyed by (device, stream) and resolved HERE, for the stream the
    launch goes to -- a descriptor built once and launched on two streams gets two workspaces.  A descriptor whose
    workspace pointer was cleared (whole-tile scheduling, tests) stays that way."""
    if desc.kind != "mlp" or not desc.c.buf[6]:
        return
    from .core import _mlp_workspace

    nbytes = int(_lib.load().ebm_workspace_bytes(C.byref(desc.c)))
    if nbytes <= 0:
        return
    ws = _mlp_workspace(device, nbytes)
    if desc.c.buf[6] != ws.data_ptr():
        desc.c.buf[6] = ws.data_ptr()
        desc.keep.append(ws)


def device_index(device) -> int:
    device = torch.device(device)
    return torch.cuda.current_device() if device.index is None else device.index


def torch_offset_increment(device, numel: int) -> int:
    return int(_lib.load().ebm_torch_rng_offset_increment(device_index(device), int(numel)))


def rng_consumed_langevin(device, numel: int, n_steps: int, rng_mode: int) -> int:
    if rng_mode == _lib.RNG_TORCH:
        return n_steps * torch_offset_increment(device, numel)
    if rng_mode == _lib.RNG_NATIVE:
        return 4 * n_steps
    return 0


def rng_consumed_hmc(device, n: int, d: int, n_proposals: int, rng_mode: int) -> int:
    if rng_mode == _lib.RNG_TORCH:
        return n_proposals * (torch_offset_increment(device, n * d) + torch_offset_increment(device, n))
    if rng_mode == _lib.RNG_NATIVE:
        return 8 * n_proposals
    return 0


def energy(desc: EnergyDescriptor, x: torch.Tensor) -> torch.Tensor:
    x = _req(x, "x")
    out = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().ebm_energy_f32(C.byref(desc.c), x.data_ptr(), x.shape[0], out.data_ptr(), _stream(x.device))
    _lib.check(rc, "ebm_energy_f32")
    return out


def gradient(desc: EnergyDescriptor, x: torch.Tensor) -> torch.Tensor:
    x = _req(x, "x")
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = _lib.load().ebm_gradient_f32(C.byref(desc.c), x.data_ptr(), x.shape[0], out.data_ptr(), _stream(x.device))
    _lib.check(rc, "ebm_gradient_f32")
    return out


def euler_maruyama_step(x: torch.Tensor, drift: torch.Tensor, noise: Optional[torch.Tensor], step_size: float,
                        noise_scale: Optional[float]) -> torch.Tensor:
    x, drift = _req(x, "x"), _req(drift, "drift")
    if noise is not None:
        noise = _req(noise, "noise")
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = _lib.load().ebm_euler_maruyama_step_f32(
            x.data_ptr(), drift.data_ptr(), _ptr(noise), out.data_ptr(), x.numel(), float(step_size),
            -1.0 if noise_scale is None else float(noise_scale), _stream(x.device))
    _lib.check(rc, "ebm_euler_maruyama_step_f32")
    return out


def langevin_burst(desc: EnergyDescriptor, x: torch.Tensor, n_steps: int, step_sizes: Sequence[float],
                   noise_scales: Sequence[float], *, clamp: Optional[Tuple[float, float]] = None,
                   rng_mode: int = _lib.RNG_TORCH, seed: int = 0, offset: int = 0,
                   noise: Optional[torch.Tensor] = None, traj: Optional[torch.Tensor] = None, thin: int = 1,
                   out: Optional[torch.Tensor] = None, scheme: str = "euler_maruyama",
                   diag: Optional[dict] = None) -> torch.Tensor:
    """K-step burst (`scheme`: "euler_maruyama" or, for the elementwise energies, "heun").  `step_sizes` / `noise_scales`
    have length 1 (constant) or n_steps.  Returns the final state (a new tensor unless `out` is given; `out` may be `x`
    for an in-place burst).  `diag`: dict of preallocated "mean" [n_kept, d], "var" [n_kept, d], "energy" [n_kept]
    (langevin_dynamics.py:170-185) filled by the same call."""
    x = _req(x, "x")
    if out is None:
        out = torch.empty_like(x)
    if noise is not None:
        noise = _req(noise, "noise")
    assert len(step_sizes) == len(noise_scales) and len(step_sizes) in (1, n_steps)
    hs, ns = _lib.doubles(list(step_sizes)), _lib.doubles(list(noise_scales))
    cl = (C.c_float * 2)(clamp[0], clamp[1]) if clamp is not None else None
    heun = scheme != "euler_maruyama"
    _bind_workspace(desc, x.device)
    with torch.cuda.device(x.device):
        if diag is not None and n_steps // thin > 0:
            ws = torch.empty((n_steps // thin) * (2 * desc.dim + 2), dtype=torch.float64, device=x.device)
            scratch = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
            rc = _lib.load().ebm_langevin_burst_diag_f32(
                C.byref(desc.c), x.data_ptr(), out.data_ptr(), x.shape[0], int(n_steps), hs, ns, len(step_sizes), cl,
                int(rng_mode), int(seed), int(offset), _ptr(noise), _ptr(traj), int(thin), 1 if heun else 0, ws.data_ptr(),
                scratch.data_ptr(), diag["mean"].data_ptr(), diag["var"].data_ptr(), diag["energy"].data_ptr(),
                _stream(x.device))
            _lib.check(rc, "ebm_langevin_burst_diag_f32")
            return out
        fn = _lib.load().ebm_langevin_heun_burst_f32 if heun else _lib.load().ebm_langevin_burst_f32
        rc = fn(C.byref(desc.c), x.data_ptr(), out.data_ptr(), x.shape[0], int(n_steps), hs, ns, len(step_sizes), cl,
                int(rng_mode), int(seed), int(offset), _ptr(noise), _ptr(traj), int(thin), _stream(x.device))
    _lib.check(rc, "ebm_langevin_heun_burst_f32" if heun else "ebm_langevin_burst_f32")
    return out


def langevin_burst_gather(desc: EnergyDescriptor, x: torch.Tensor, n_steps: int, step_sizes: Sequence[float],
                          noise_scales: Sequence[float], peer_ptrs: Sequence[int], row_offset: int, *,
                          clamp: Optional[Tuple[float, float]] = None, rng_mode: int = _lib.RNG_TORCH, seed: int = 0,
                          offset: int = 0, out: Optional[torch.Tensor] = None,
                          multicast_ptr: Optional[int] = None) -> torch.Tensor:
    """K-step burst whose final state also lands at rows [row_offset, row_offset + n) of every rank's gathered buffer
    (`peer_ptrs[w]` = rank w's buffer as mapped into this process; `multicast_ptr` = the buffers' NVLS multicast address,
    used instead by the kernels that store from their epilogue).  A cross-rank barrier must follow on the stream."""
    x = _req(x, "x")
    if out is None:
        out = torch.empty_like(x)
    assert len(step_sizes) == len(noise_scales) and len(step_sizes) in (1, n_steps)
    hs, ns = _lib.doubles(list(step_sizes)), _lib.doubles(list(noise_scales))
    cl = (C.c_float * 2)(clamp[0], clamp[1]) if clamp is not None else None
    peers = (C.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
    _bind_workspace(desc, x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().ebm_langevin_burst_gather_f32(
            C.byref(desc.c), x.data_ptr(), out.data_ptr(), x.shape[0], int(n_steps), hs, ns, len(step_sizes), cl,
            int(rng_mode), int(seed), int(offset), peers, len(peer_ptrs), int(row_offset), _stream(x.device))
    _lib.check(rc, "ebm_langevin_burst_gather_f32")
    return out


def _peer_array(peer_ptrs: Sequence[int], multicast_ptr: Optional[int]):
    """(pointer array, world) of the gather entry points: world = -W announces W unicast pointers followed by the NVLS
    multicast pointer (include/ebm_b200.h)."""
    ptrs = [int(p) for p in peer_ptrs]
    if multicast_ptr:
        return (C.c_void_p * (len(ptrs) + 1))(*ptrs, int(multicast_ptr)), -len(ptrs)
    return (C.c_void_p * len(ptrs))(*ptrs), len(ptrs)


def descent_burst(desc: EnergyDescriptor, x: torch.Tensor, n_steps: int, step_sizes: Sequence[float], *,
                  momentum: Optional[float] = None, traj: Optional[torch.Tensor] = None, thin: int = 1,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """K noise-free descent steps (momentum None: gradient descent; else Nesterov starting from zero velocity)."""
    x = _req(x, "x")
    if out is None:
        out = torch.empty_like(x)
    assert len(step_sizes) in (1, n_steps)
    hs = _lib.doubles(list(step_sizes))
    vel = torch.empty_like(x) if (momentum is not None and len(step_sizes) > 1 and n_steps > 64) else None
    with torch.cuda.device(x.device):
        rc = _lib.load().ebm_descent_burst_f32(C.byref(desc.c), x.data_ptr(), out.data_ptr(), x.shape[0], int(n_steps), hs,
                                               len(step_sizes), -1.0 if momentum is None else float(momentum), _ptr(vel),
                                               _ptr(traj), int(thin), _stream(x.device))
    _lib.check(rc, "ebm_descent_burst_f32")
    return out


def peer_push(x_local: torch.Tensor, peer_ptrs: Sequence[int], elem_offset: int, max_ctas: int) -> None:
    """SM-driven gather push of a finished shard into every rank's gathered buffer (at most `max_ctas` CTAs)."""
    x_local = _req(x_local, "x_local")
    peers = (C.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
    with torch.cuda.device(x_local.device):
        rc = _lib.load().ebm_peer_push_f32(x_local.data_ptr(), x_local.numel(), peers, len(peer_ptrs), int(elem_offset),
                                           int(max_ctas), _stream(x_local.device))
    _lib.check(rc, "ebm_peer_push_f32")


def leapfrog(desc: EnergyDescriptor, x: torch.Tensor, p: torch.Tensor, step_size: float, n_steps: int,
             mass=None, safe: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    x, p = _req(x, "x"), _req(p, "p")
    xo, po = torch.empty_like(x), torch.empty_like(p)
    kind, ms, mv = _mass_args(mass, x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().ebm_leapfrog_f32(C.byref(desc.c), x.data_ptr(), p.data_ptr(), xo.data_ptr(), po.data_ptr(),
                                          x.shape[0], int(n_steps), float(step_size), kind, ms, _ptr(mv),
                                          1 if safe else 0, _stream(x.device))
    _lib.check(rc, "ebm_leapfrog_f32")
    return xo, po


def _mass_args(mass, device):
    if mass is None:
        return _lib.MASS_NONE, 0.0, None
    if isinstance(mass, float):
        return _lib.MASS_SCALAR, float(mass), None
    mv = mass.detach().to(device=device, dtype=torch.float32).contiguous()
    return _lib.MASS_VECTOR, 0.0, mv


def hmc_burst(desc: EnergyDescriptor, x: torch.Tensor, n_proposals: int, n_leapfrog: int, step_sizes: Sequence[float],
              *, mass=None, rng_mode: int = _lib.RNG_TORCH, seed: int = 0, offset: int = 0,
              noise_p: Optional[torch.Tensor] = None, noise_u: Optional[torch.Tensor] = None,
              traj: Optional[torch.Tensor] = None, thin: int = 1, accept_count: Optional[torch.Tensor] = None,
              energy_out: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
              diag: Optional[dict] = None) -> torch.Tensor:
    """`diag`: dict of preallocated "mean" / "var" [n_kept, d], "energy" / "acceptance_rate" [n_kept] (hmc.py:294-310)
    filled by the same call."""
    x = _req(x, "x")
    if out is None:
        out = torch.empty_like(x)
    if noise_p is not None:
        noise_p, noise_u = _req(noise_p, "noise_p"), _req(noise_u, "noise_u")
    if accept_count is not None and accept_count.dtype != torch.int32:
        raise TypeError("accept_count must be int32")
    assert len(step_sizes) in (1, n_proposals)
    hs = _lib.doubles(list(step_sizes))
    kind, ms, mv = _mass_args(mass, x.device)
    _bind_workspace(desc, x.device)
    with torch.cuda.device(x.device):
        if diag is not None and n_proposals // thin > 0:
            n_kept = n_proposals // thin
            ws = torch.empty(n_kept * (2 * desc.dim + 2), dtype=torch.float64, device=x.device)
            scratch = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
            acc = torch.empty(n_proposals, dtype=torch.int32, device=x.device)
            rc = _lib.load().ebm_hmc_burst_diag_f32(
                C.byref(desc.c), x.data_ptr(), out.data_ptr(), x.shape[0], int(n_proposals), int(n_leapfrog), hs,
                len(step_sizes), kind, ms, _ptr(mv), int(rng_mode), int(seed), int(offset), _ptr(noise_p), _ptr(noise_u),
                _ptr(traj), int(thin), ws.data_ptr(), scratch.data_ptr(), acc.data_ptr(), diag["mean"].data_ptr(),
                diag["var"].data_ptr(), diag["energy"].data_ptr(), diag["acceptance_rate"].data_ptr(), _stream(x.device))
            _lib.check(rc, "ebm_hmc_burst_diag_f32")
            return out
        rc = _lib.load().ebm_hmc_burst_f32(
            C.byref(desc.c), x.data_ptr(), out.data_ptr(), x.shape[0], int(n_proposals), int(n_leapfrog), hs,
            len(step_sizes), kind, ms, _ptr(mv), int(rng_mode), int(seed), int(offset), _ptr(noise_p), _ptr(noise_u),
            _ptr(traj), int(thin), _ptr(accept_count), _ptr(energy_out), _stream(x.device))
    _lib.check(rc, "ebm_hmc_burst_f32")
    return out


def pcd_gather(buffer: torch.Tensor, idx: torch.Tensor, noise_rows: Optional[torch.Tensor] = None,
               noise: Optional[torch.Tensor] = None) -> torch.Tensor:
    buffer = _req(buffer, "buffer")
    if idx.dtype != torch.int64 or not idx.is_cuda:
        raise TypeError("idx must be a CUDA int64 tensor")
    idx = idx.contiguous()
    row_elems = buffer[0].numel()
    out = torch.empty((idx.shape[0],) + tuple(buffer.shape[1:]), dtype=torch.float32, device=buffer.device)
    n_noise = 0
    if noise_rows is not None:
        noise_rows = noise_rows.contiguous()
        noise = _req(noise, "noise")
        n_noise = noise_rows.shape[0]
    with torch.cuda.device(buffer.device):
        rc = _lib.load().ebm_pcd_gather_f32(buffer.data_ptr(), buffer.shape[0], row_elems, idx.data_ptr(), idx.shape[0],
                                            out.data_ptr(), _ptr(noise_rows), _ptr(noise), n_noise,
                                            _stream(buffer.device))
    _lib.check(rc, "ebm_pcd_gather_f32")
    return out


def pcd_scatter(buffer: torch.Tensor, ptr: int, samples: torch.Tensor) -> int:
    """FIFO write-back in place; returns the new pointer (host int, no sync)."""
    if not buffer.is_contiguous():
        raise ValueError("replay buffer must be contiguous")
    samples = _req(samples, "samples")
    row_elems = buffer[0].numel()
    new_ptr = C.c_int64(0)
    with torch.cuda.device(buffer.device):
        rc = _lib.load().ebm_pcd_scatter_f32(buffer.data_ptr(), buffer.shape[0], row_elems, int(ptr), samples.data_ptr(),
                                             samples.shape[0], C.byref(new_ptr), _stream(buffer.device))
    _lib.check(rc, "ebm_pcd_scatter_f32")
    return int(new_ptr.value)


def pcd_langevin_fused(desc: EnergyDescriptor) -> bool:
    return bool(_lib.load().ebm_pcd_langevin_fused(C.byref(desc.c)))


def pcd_langevin_burst(desc: EnergyDescriptor, buffer: torch.Tensor, idx: Optional[torch.Tensor], ptr: int, n_steps: int,
                       step_sizes: Sequence[float], noise_scales: Sequence[float], *,
                       clamp: Optional[Tuple[float, float]] = None, rng_mode: int = _lib.RNG_TORCH, seed: int = 0,
                       offset: int = 0, noise_rows: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None,
                       energy_out: Optional[torch.Tensor] = None, batch: Optional[int] = None,
                       peer_ptrs: Optional[Sequence[int]] = None, row_offset: int = 0,
                       multicast_ptr: Optional[int] = None) -> Tuple[torch.Tensor, int]:
    """Start points `buffer[idx]` (+ 0.01 * noise[j] on chain noise_rows[j]) -> K-step burst -> FIFO write-back into
    `buffer` (in place).  `idx=None`: chain i starts from row i (`batch` chains, default the whole buffer) -- the
    stride-1 case of core/base_loss.py:307-312.  Returns the negatives and the new FIFO pointer (host int, no sync);
    `energy_out[n]` receives E(negatives).  `peer_ptrs` (see `langevin_burst_gather`): the negatives also land at rows
    `[row_offset, row_offset + n)` of every rank's gathered buffer; a cross-rank barrier must follow on the stream."""
    if not buffer.is_contiguous() or buffer.ndim != 2:
        raise ValueError("replay buffer must be a contiguous [S, D] tensor")
    buffer = _req(buffer, "buffer")
    if idx is not None:
        if idx.dtype != torch.int64 or not idx.is_cuda:
            raise TypeError("idx must be a CUDA int64 tensor")
        idx = idx.contiguous()
        n = idx.shape[0]
    else:
        n = buffer.shape[0] if batch is None else int(batch)
    n_noise = 0
    if noise_rows is not None:
        if noise_rows.dtype != torch.int64 or not noise_rows.is_cuda:
            raise TypeError("noise_rows must be a CUDA int64 tensor")
        noise_rows = noise_rows.contiguous()
        noise = _req(noise, "noise")
        n_noise = noise_rows.shape[0]
    out = torch.empty((n, buffer.shape[1]), dtype=torch.float32, device=buffer.device)
    reads_buffer = pcd_langevin_fused(desc) and ((idx is None and n == buffer.shape[0]) or (idx is not None and n_noise == 0))
    scratch = None if reads_buffer else torch.empty_like(out)
    assert len(step_sizes) == len(noise_scales) and len(step_sizes) in (1, n_steps)
    hs, ns = _lib.doubles(list(step_sizes)), _lib.doubles(list(noise_scales))
    cl = (C.c_float * 2)(clamp[0], clamp[1]) if clamp is not None else None
    new_ptr = C.c_int64(0)
    args = (C.byref(desc.c), buffer.data_ptr(), buffer.shape[0], _ptr(idx), int(ptr), out.data_ptr(), _ptr(scratch), n,
            int(n_steps), hs, ns, len(step_sizes), cl, int(rng_mode), int(seed), int(offset), _ptr(noise_rows), _ptr(noise),
            n_noise, _ptr(energy_out), C.byref(new_ptr))
    _bind_workspace(desc, buffer.device)
    with torch.cuda.device(buffer.device):
        if peer_ptrs is None:
            rc = _lib.load().ebm_pcd_langevin_burst_f32(*args, _stream(buffer.device))
        else:
            peers, world = _peer_array(peer_ptrs, multicast_ptr)
            rc = _lib.load().ebm_pcd_langevin_burst_gather_f32(*args, peers, world, int(row_offset), _stream(buffer.device))
    _lib.check(rc, "ebm_pcd_langevin_burst_f32" if peer_ptrs is None else "ebm_pcd_langevin_burst_gather_f32")
    return out, int(new_ptr.value)


def ess(chains: torch.Tensor) -> torch.Tensor:
    """Effective sample size of every row of `chains[n_chains, n]` (or of one 1-D chain) on the device: the estimator of
    benchmarks/registry.py:348-365.  Returns a float32 tensor of shape `[n_chains]` (0-d for a 1-D chain); no sync."""
    chains = _req(chains, "chains")
    one = chains.ndim == 1
    if one:
        chains = chains.unsqueeze(0)
    if chains.ndim != 2 or chains.shape[1] < 1 or chains.shape[0] < 1:
        raise ValueError("chains must be a non-empty [n_chains, n] tensor or a 1-D chain")
    out = torch.empty(chains.shape[0], dtype=torch.float32, device=chains.device)
    with torch.cuda.device(chains.device):
        rc = _lib.load().ebm_ess_f32(chains.data_ptr(), chains.shape[0], chains.shape[1], out.data_ptr(), _stream(chains.device))
    _lib.check(rc, "ebm_ess_f32")
    return out[0] if one else out


def rng_fill(numel: int, device, rng_mode: int, kind: int, seed: int, offset: int) -> torch.Tensor:
    out = torch.empty(numel, dtype=torch.float32, device=device)
    with torch.cuda.device(out.device):
        rc = _lib.load().ebm_rng_fill_f32(out.data_ptr(), numel, rng_mode, kind, int(seed), int(offset), _stream(out.device))
    _lib.check(rc, "ebm_rng_fill_f32")
    return out


def langevin_burst_host(desc: EnergyDescriptor, x_host: torch.Tensor, out_host: torch.Tensor, scratch: torch.Tensor,
                        n_steps: int, step_size: float, noise_scale: float, rng_mode: int, seed: int, offset: int) -> None:
    """End-to-end entry: pinned host in -> burst -> pinned host out, synchronised on return."""
    assert not x_host.is_cuda and not out_host.is_cuda and scratch.is_cuda
    _bind_workspace(desc, scratch.device)
    with torch.cuda.device(scratch.device):
        rc = _lib.load().ebm_langevin_burst_host_f32(
            C.byref(desc.c), x_host.data_ptr(), out_host.data_ptr(), scratch.data_ptr(), x_host.shape[0], int(n_steps),
            float(step_size), float(noise_scale), int(rng_mode), int(seed), int(offset), _stream(scratch.device))
    _lib.check(rc, "ebm_langevin_burst_host_f32")
