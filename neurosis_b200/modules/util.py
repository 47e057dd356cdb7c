"""Shared helpers of the drop-in modules (layout views, zero-init, sinusoidal embedding)."""
from __future__ import annotations

import torch
from torch import Tensor, nn

from .. import ops


def zero_module(module: nn.Module) -> nn.Module:
    """zero all parameters (reference modules/diffusion/util.py:180-186)."""
    for p in module.parameters():
        nn.init.zeros_(p)
    return module


def as_nhwc(x: Tensor, cpad: int | None = None) -> Tensor:
    """(N,C,H,W) -> contiguous bf16 (N,H,W,C).  Zero-copy when x is already a channels-last bf16
    view (what every module of this package returns); otherwise one layout-conversion kernel."""
    if x.dim() != 4:
        raise ValueError(f"expected a 4-D NCHW tensor, got {tuple(x.shape)}")
    if x.dtype == torch.bfloat16 and (cpad is None or cpad == x.shape[1]):
        v = x.permute(0, 2, 3, 1)
        if v.is_contiguous():
            return v
    return ops.to_nhwc(x, cpad)


def from_nhwc(y: Tensor) -> Tensor:
    """(N,H,W,C) contiguous -> logical (N,C,H,W) channels-last view (no copy)."""
    return y.permute(0, 3, 1, 2)


def timestep_embedding(timesteps: Tensor, dim: int, max_period: int = 10000, repeat_only: bool = False) -> Tensor:
    """cos|sin embedding (reference modules/diffusion/util.py:152-177), bf16 output [N, dim]."""
    if repeat_only:
        raise NotImplementedError("repeat_only embeddings are not used by the training path")
    return ops.timestep_embedding(timesteps, dim, float(max_period))
