"""Oracle for the persistent-CD replay buffer.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows torchebm/core/base_loss.py:190-264 (`initialize_buffer`), :266-337
(`get_start_points`), :390-426 (`update_buffer`), :428-481 (`mix_buffer_across_ranks`)
and the caller torchebm/losses/contrastive_divergence.py:127-139.
"""

from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch


class ReplayBuffer:
    def __init__(self, buffer_size: int, new_sample_ratio: float = 0.0, init_steps: int = 0):
        self.buffer_size = buffer_size
        self.new_sample_ratio = new_sample_ratio
        self.init_steps = init_steps
        self.buffer: Optional[torch.Tensor] = None
        self.ptr = 0

    # base_loss.py:190-264
    def initialize(self, data_shape: Tuple[int, ...], device, generator=None, sampler: Optional[Callable] = None,
                   chunk: int = 1024, init_noise_scale: float = 0.01):
        if self.buffer_size <= 0:
            raise ValueError(f"Replay buffer size must be positive, got {self.buffer_size}")
        shape = (self.buffer_size,) + tuple(data_shape)
        self.buffer = torch.randn(shape, dtype=torch.float32, device=device, generator=generator) * init_noise_scale
        if self.init_steps > 0 and sampler is not None:
            cs = min(self.buffer_size, chunk)
            for i in range(0, self.buffer_size, cs):
                end = min(i + cs, self.buffer_size)
                cur = self.buffer[i:end].clone()
                self.buffer[i:end] = sampler(cur, self.init_steps, generator)
        self.ptr = 0

    # base_loss.py:288-312
    def start_indices(self, batch_size: int, device, generator=None) -> torch.Tensor:
        if self.buffer_size < batch_size:
            return torch.randint(0, self.buffer_size, (batch_size,), device=device, generator=generator)
        stride = self.buffer_size // batch_size
        base = torch.arange(0, batch_size, device=device) * stride
        offset = torch.randint(0, stride, (batch_size,), device=device, generator=generator)
        return (base + offset) % self.buffer_size

    # base_loss.py:314-332
    def get_start_points(self, batch_size: int, generator=None, indices: Optional[torch.Tensor] = None,
                         noise_rows: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None):
        dev = self.buffer.device
        if indices is None:
            indices = self.start_indices(batch_size, dev, generator)
        start = self.buffer[indices]
        if self.new_sample_ratio > 0.0:
            n_new = max(1, int(batch_size * self.new_sample_ratio))
            if noise_rows is None:
                noise_rows = torch.randperm(batch_size, device=dev, generator=generator)[:n_new]
            if noise is None:
                noise = torch.randn_like(start[noise_rows], generator=generator)
            start[noise_rows] = start[noise_rows] + noise * 0.01
        return start

    # base_loss.py:390-426
    def update(self, samples: torch.Tensor):
        b = samples.shape[0]
        s = self.buffer_size
        if b >= s:
            self.buffer[:] = samples[-s:]
            self.ptr = 0
            return
        end = (self.ptr + b) % s
        if end > self.ptr:
            self.buffer[self.ptr:end] = samples
        else:
            first = s - self.ptr
            self.buffer[self.ptr:] = samples[:first]
            self.buffer[:end] = samples[first:]
        self.ptr = end


def mix(gathered: torch.Tensor, perm: torch.Tensor, rank: int, buffer_size: int) -> torch.Tensor:
    """base_loss.py:477-481: this rank's shard of the permuted pooled chains."""
    start = rank * buffer_size
    return gathered[perm[start:start + buffer_size]]
