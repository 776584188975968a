"""Persistent-CD replay buffer kept device-side, and the `ContrastiveDivergence` loss that drives it.

Mirrors torchebm/core/base_loss.py:116-529 (`BaseContrastiveDivergence`: lazy `[S, *shape]` buffer,
stratified start points, exploration noise, FIFO write-back, cross-rank mix, state_dict) and
torchebm/losses/contrastive_divergence.py:13-223 (the CD-k loss).  The index draws
(`randint` / `randperm`) stay PyTorch calls so the generator stream matches the reference draw for
draw; the row gather + exploration noise and the FIFO scatter are library
kernels working in place on the registered buffer, and the K-step negative chain is one fused burst.
The loss value itself (`compute_loss`) is ordinary autograd PyTorch: it is the training objective,
not part of the sampling path.
"""

from __future__ import annotations

import warnings
from typing import Optional, Tuple, Union

import torch

from . import ops
from .core import TorchEBMModule


class FusedPCDMixin:
    """Device-side replay-buffer logic shared by the standalone loss below and by the reference-derived one of dropin.py:
    library kernels for the row gather + exploration noise and for the FIFO write-back, and the one-call negative
    sampler.  Relies only on the attributes of core/base_loss.py:155-188 (`replay_buffer`, `buffer_ptr`,
    `_buffer_ptr_int`, `buffer_size`, `new_sample_ratio`, `k_steps`, `sampler`, `persistent`, `buffer_initialized`)."""

    # base_loss.py:266-337
    def get_start_points(self, x: torch.Tensor, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        x = x.to(device=self.device, dtype=self.dtype)
        batch_size = x.shape[0]
        if not self.persistent:
            return x.detach().clone()
        if not self.buffer_initialized:
            self.initialize_buffer(tuple(x.shape[1:]), generator=generator)
            if not self.buffer_initialized:
                raise RuntimeError("Buffer initialization failed.")
        if not self.replay_buffer.is_cuda:
            return self._get_start_points_unfused(x, generator)
        indices = self._draw_start_indices(batch_size, generator)
        noise_rows, noise = self._draw_exploration_noise(x, generator)
        return ops.pcd_gather(self.replay_buffer, indices, noise_rows, noise)

    def _draw_start_indices(self, batch_size: int, generator: Optional[torch.Generator]) -> torch.Tensor:
        """base_loss.py:293-312: stratified start rows (uniform with replacement when the buffer is the smaller)."""
        if self.buffer_size < batch_size:
            warnings.warn(
                f"Buffer size ({self.buffer_size}) is smaller than batch size ({batch_size}). Sampling with replacement.",
                UserWarning)
            return torch.randint(0, self.buffer_size, (batch_size,), device=self.device, generator=generator)
        stride = self.buffer_size // batch_size
        base = torch.arange(0, batch_size, device=self.device) * stride
        offset = torch.randint(0, stride, (batch_size,), device=self.device, generator=generator)
        return (base + offset) % self.buffer_size

    def _draw_exploration_noise(self, x: torch.Tensor, generator: Optional[torch.Generator]):
        """base_loss.py:317-332: the rows that get `+ 0.01 * randn` and that noise (same draws, same order)."""
        if self.new_sample_ratio <= 0.0:
            return None, None
        batch_size = x.shape[0]
        n_new = max(1, int(batch_size * self.new_sample_ratio))
        noise_rows = torch.randperm(batch_size, device=self.device, generator=generator)[:n_new]
        noise = torch.randn((n_new,) + tuple(x.shape[1:]), dtype=self.dtype, device=self.device, generator=generator)
        return noise_rows, noise

    def sample_negatives(self, x: torch.Tensor, model_kwargs: Optional[dict] = None,
                         generator: Optional[torch.Generator] = None, energy_out: Optional[torch.Tensor] = None,
                         gather_into=None) -> torch.Tensor:
        """The sampling half of `ContrastiveDivergence.forward` (contrastive_divergence.py:127-139): start points,
        `k_steps` of the sampler, buffer write-back.  Same draws from `generator`, same buffer and pointer state as
        `get_start_points` -> `sampler.sample` -> `update_buffer`; when the sampler offers `sample_from_buffer` and
        nothing stands in the way (2-D state, no conditioning) the three steps are ONE library call: exploration noise
        included, and with `buffer_size == batch` the burst kernel reads its start rows straight from the replay buffer
        and writes its final state straight back.  `energy_out[batch]` (optional) receives E(x-).  `gather_into` (a
        `distributed.PeerGatherBuffer`, chain-sharded runs): the negatives also land in every rank's gathered tensor --
        the all-gather of utils/distributed.py:43-70 -- stored by the burst kernel itself when the one-call path runs,
        pushed afterwards otherwise; complete on all ranks after the barrier this method issues."""
        pred = self._sample_negatives(x, model_kwargs, generator, energy_out, gather_into)
        if gather_into is not None:
            if not getattr(self, "_gathered_in_burst", False):
                gather_into.push(pred)       # (push ends with the barrier)
            else:
                gather_into.barrier()
        return pred

    def _sample_negatives(self, x, model_kwargs, generator, energy_out, gather_into) -> torch.Tensor:
        self._gathered_in_burst = False
        fused = getattr(self.sampler, "sample_from_buffer", None)
        if self.persistent and fused is not None and not model_kwargs and x.ndim == 2:
            x = x.to(device=self.device, dtype=self.dtype)
            if not self.buffer_initialized:
                self.initialize_buffer(tuple(x.shape[1:]), generator=generator)
                if not self.buffer_initialized:
                    raise RuntimeError("Buffer initialization failed.")
            if self.replay_buffer.is_cuda and self.replay_buffer.is_contiguous() and self.replay_buffer.dtype == torch.float32:
                batch = x.shape[0]
                indices = self._draw_start_indices(batch, generator)
                noise_rows, noise = self._draw_exploration_noise(x, generator)
                # buffer_size == batch: the stratified draw has stride 1, i.e. it is arange(batch) by construction
                # (base_loss.py:307-312); the library takes its no-gather form only when told so explicitly (idx None)
                identity = self.buffer_size == batch
                extra = {} if gather_into is None else {"gather_into": gather_into}
                res = fused(self.replay_buffer, None if identity else indices, self._buffer_ptr_int, self.k_steps,
                            noise_rows=noise_rows, noise=noise, energy_out=energy_out, generator=generator, **extra)
                if res is not None:
                    self._gathered_in_burst = gather_into is not None
                    pred, new_ptr = res
                    self._buffer_ptr_int = new_ptr
                    self.buffer_ptr.fill_(new_ptr)
                    return pred
                # nothing was consumed beyond the draws above: continue with the gathered start points
                start_points = ops.pcd_gather(self.replay_buffer, indices, noise_rows, noise)
                pred = self.sampler.sample(x=start_points, n_steps=self.k_steps, model_kwargs=model_kwargs, generator=generator)
                with torch.no_grad():
                    self.update_buffer(pred)
                if energy_out is not None:
                    with torch.no_grad():
                        energy_out.copy_(self.model(pred))
                return pred
        start_points = self.get_start_points(x, generator=generator)
        pred = self.sampler.sample(x=start_points, n_steps=self.k_steps, model_kwargs=model_kwargs, generator=generator)
        if self.persistent:
            with torch.no_grad():
                self.update_buffer(pred)
        if energy_out is not None:
            with torch.no_grad():
                energy_out.copy_(self.model(pred, **(model_kwargs or {})))
        return pred

    # base_loss.py:390-426
    def update_buffer(self, samples: torch.Tensor) -> None:
        if not self.persistent or not self.buffer_initialized:
            return
        if not self.replay_buffer.is_cuda:
            return self._update_buffer_unfused(samples)
        samples = samples.to(device=self.device, dtype=self.dtype).detach()
        new_ptr = ops.pcd_scatter(self.replay_buffer, self._buffer_ptr_int, samples)
        self._buffer_ptr_int = new_ptr
        self.buffer_ptr.fill_(new_ptr)

    # hooks for buffers that are not on a CUDA device: the standalone loss has no such path, dropin.py hands them to
    # the reference's own implementation
    def _get_start_points_unfused(self, x, generator):
        raise RuntimeError("the persistent-CD buffer must live on a CUDA device: torchebm_b200 has no CPU path")

    def _update_buffer_unfused(self, samples):
        raise RuntimeError("the persistent-CD buffer must live on a CUDA device: torchebm_b200 has no CPU path")


class BaseContrastiveDivergence(FusedPCDMixin, TorchEBMModule):
    def __init__(self, model, sampler, k_steps: int = 1, persistent: bool = False, buffer_size: int = 100,
                 new_sample_ratio: float = 0.0, init_steps: int = 0, dtype: torch.dtype = torch.float32,
                 device: Optional[Union[str, torch.device]] = None, *args, **kwargs):
        super().__init__(dtype=dtype, device=device, *args, **kwargs)
        self.model = model
        self.sampler = sampler
        self.k_steps = k_steps
        self.persistent = persistent
        self.buffer_size = buffer_size
        self.new_sample_ratio = new_sample_ratio
        self.init_steps = init_steps
        self.register_buffer("replay_buffer", None)
        self.register_buffer("buffer_ptr", torch.tensor(0, dtype=torch.long, device=device))
        self._buffer_ptr_int: int = 0
        self.buffer_initialized = False

    # base_loss.py:190-264
    def initialize_buffer(self, data_shape_no_batch: Tuple[int, ...], buffer_chunk_size: int = 1024,
                          init_noise_scale: float = 0.01, generator: Optional[torch.Generator] = None):
        if not self.persistent or self.buffer_initialized:
            return
        if self.buffer_size <= 0:
            raise ValueError(f"Replay buffer size must be positive, got {self.buffer_size}")
        shape = (self.buffer_size,) + tuple(data_shape_no_batch)
        self.replay_buffer = torch.randn(shape, dtype=self.dtype, device=self.device, generator=generator) * init_noise_scale
        if self.init_steps > 0:
            with torch.no_grad():
                chunk = min(self.buffer_size, buffer_chunk_size)
                for i in range(0, self.buffer_size, chunk):
                    end = min(i + chunk, self.buffer_size)
                    cur = self.replay_buffer[i:end].clone()
                    try:
                        upd = self.sampler.sample(x=cur, n_steps=self.init_steps, generator=generator).detach()
                        if upd.shape == cur.shape:
                            self.replay_buffer[i:end] = upd
                        else:
                            warnings.warn(f"Sampler output shape mismatch during buffer init for chunk {i}-{end}.")
                    except Exception as e:  # base_loss.py:254-257 keeps the noise for a failing chunk
                        warnings.warn(f"Error during buffer initialization sampling for chunk {i}-{end}: {e}.")
        self.buffer_ptr.zero_()
        self._buffer_ptr_int = 0
        self.buffer_initialized = True
        return self.replay_buffer

    # base_loss.py:428-481
    def mix_buffer_across_ranks(self, process_group=None, generator: Optional[torch.Generator] = None) -> None:
        if not self.persistent:
            raise RuntimeError("mix_buffer_across_ranks requires a persistent loss (persistent=True).")
        if not self.buffer_initialized:
            raise RuntimeError("The replay buffer is not initialized; run one training step or call initialize_buffer() first.")
        from .distributed import all_gather_cat, broadcast_tensor, get_rank, get_world_size

        if get_world_size(process_group) == 1:
            return
        gathered = all_gather_cat(self.replay_buffer, group=process_group)
        perm = torch.randperm(gathered.shape[0], generator=generator)
        perm = broadcast_tensor(perm, src=0, group=process_group)
        start = get_rank(process_group) * self.buffer_size
        idx = perm[start:start + self.buffer_size].to(gathered.device)
        if gathered.is_cuda:
            self.replay_buffer.copy_(ops.pcd_gather(gathered, idx))
        else:  # gloo/CPU process groups in the host-logic tests
            self.replay_buffer.copy_(gathered[idx])

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        self._buffer_ptr_int = int(self.buffer_ptr.item())


class ContrastiveDivergence(BaseContrastiveDivergence):
    """contrastive_divergence.py:13-223."""

    def __init__(self, model, sampler, k_steps=10, persistent=False, buffer_size=10000, init_steps=100,
                 new_sample_ratio=0.05, energy_reg_weight=0.001, add_noise_to_real=False, noise_scale=1e-4,
                 dtype=torch.float32, device=torch.device("cpu"), *args, **kwargs):
        super().__init__(model=model, sampler=sampler, k_steps=k_steps, persistent=persistent, buffer_size=buffer_size,
                         new_sample_ratio=new_sample_ratio, init_steps=init_steps, dtype=dtype, device=device,
                         *args, **kwargs)
        self.energy_reg_weight = energy_reg_weight
        self.add_noise_to_real = add_noise_to_real
        self.noise_scale = noise_scale

    def forward(self, x: torch.Tensor, *args, model_kwargs: Optional[dict] = None,
                generator: Optional[torch.Generator] = None, **kwargs):
        model_kwargs = self._prepare_model_kwargs(model_kwargs)
        pred_samples = self.sample_negatives(x, model_kwargs=model_kwargs, generator=generator)
        kwargs.setdefault("energy_reg_weight", self.energy_reg_weight)
        kwargs.setdefault("add_noise_to_real", self.add_noise_to_real)
        kwargs.setdefault("noise_scale", self.noise_scale)
        loss = self.compute_loss(x, pred_samples, *args, model_kwargs=model_kwargs, generator=generator, **kwargs)
        return loss, pred_samples

    def compute_loss(self, x, pred_x, *args, model_kwargs: Optional[dict] = None,
                     generator: Optional[torch.Generator] = None, **kwargs) -> torch.Tensor:
        x = x.to(self.device, self.dtype)
        pred_x = pred_x.to(self.device, self.dtype)
        mk = model_kwargs or {}
        with torch.set_grad_enabled(True):
            if kwargs.get("add_noise_to_real", self.add_noise_to_real):
                ns = kwargs.get("noise_scale", self.noise_scale)
                x_energy = self.model(x + ns * torch.randn_like(x, generator=generator), **mk)
            else:
                x_energy = self.model(x, **mk)
            pred_energy = self.model(pred_x, **mk)
        loss = torch.mean(x_energy) - torch.mean(pred_energy)
        reg = kwargs.get("energy_reg_weight", self.energy_reg_weight)
        if reg > 0:
            loss = loss + reg * (torch.mean(x_energy**2) + torch.mean(pred_energy**2))
        fallback = torch.tensor(0.1, device=loss.device, dtype=loss.dtype)
        return torch.where(torch.isfinite(loss), loss, fallback)
