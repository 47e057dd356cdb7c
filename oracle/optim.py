"""ORACLE (test infrastructure): CPU fp32 restatement of the reference's optimizer-side updates.

Follows /root/reference/src/neurosis/optimizers/adafactor.py: `_get_lr` (:129-139), `_rms` (:147-149),
`_approx_sq_grad` (:151-157) and the body of `step` (:176-246); and /root/reference/src/neurosis/modules/ema.py:40-59
(`LitEma.forward`).  Pinned by tests/golden/reference_golden_next.npz (generated from the reference classes).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import Tensor


def adafactor_init_state(p: Tensor, beta1: Optional[float]) -> dict:
    st = {"step": 0, "RMS": 0}
    if beta1 is not None:
        st["exp_avg"] = torch.zeros_like(p)
    if p.dim() >= 2:
        st["exp_avg_sq_row"] = torch.zeros(p.shape[:-1])
        st["exp_avg_sq_col"] = torch.zeros(p.shape[:-2] + p.shape[-1:])
    else:
        st["exp_avg_sq"] = torch.zeros_like(p)
    return st


def adafactor_step(p: Tensor, grad: Tensor, st: dict, *, lr: Optional[float] = None, eps=(1e-30, 1e-3),
                   clip_threshold: float = 1.0, decay_rate: float = -0.8, beta1: Optional[float] = None,
                   weight_decay: float = 0.0, scale_parameter: bool = True, relative_step: bool = True,
                   warmup_init: bool = False) -> float:
    """one in-place update of fp32 `p` (adafactor.py:209-244); returns the step size used."""
    st["step"] += 1
    st["RMS"] = p.norm(2) / (p.numel() ** 0.5)
    rel = lr
    if relative_step:
        min_step = 1e-6 * st["step"] if warmup_init else 1e-2
        rel = min(min_step, 1.0 / math.sqrt(st["step"]))
    scale = max(eps[1], st["RMS"]) if scale_parameter else 1.0
    step_size = scale * rel
    beta2t = 1.0 - math.pow(st["step"], decay_rate)
    update = grad ** 2 + eps[0]
    if p.dim() >= 2:
        row, col = st["exp_avg_sq_row"], st["exp_avg_sq_col"]
        row.mul_(beta2t).add_(update.mean(dim=-1), alpha=1.0 - beta2t)
        col.mul_(beta2t).add_(update.mean(dim=-2), alpha=1.0 - beta2t)
        r_factor = (row / row.mean(dim=-1, keepdim=True)).rsqrt_().unsqueeze(-1)
        c_factor = col.unsqueeze(-2).rsqrt()
        update = torch.mul(r_factor, c_factor) * grad
    else:
        v = st["exp_avg_sq"]
        v.mul_(beta2t).add_(update, alpha=1.0 - beta2t)
        update = v.rsqrt() * grad
    update = update / ((update.norm(2) / (update.numel() ** 0.5)) / clip_threshold).clamp(min=1.0)
    update = update * step_size
    if beta1 is not None:
        st["exp_avg"].mul_(beta1).add_(update, alpha=1 - beta1)
        update = st["exp_avg"]
    if weight_decay != 0:
        p.add_(p, alpha=float(-weight_decay * step_size))
    p.add_(-update)
    return float(step_size)


def ema_decay(decay: float, num_updates: int) -> float:
    """decay used by the `num_updates`-th call (num_updates already incremented), ema.py:43-46 in fp32."""
    d = torch.tensor(decay, dtype=torch.float32)
    if num_updates >= 0:
        n = torch.tensor(num_updates, dtype=torch.int)
        d = min(d, (1 + n) / (10 + n))
    return float(d)


def ema_update(shadow: Tensor, p: Tensor, decay: float) -> None:
    """shadow.sub_((1 - decay) * (shadow - p)) (ema.py:48-57)."""
    one_minus_decay = 1.0 - torch.tensor(decay, dtype=torch.float32)
    shadow.sub_(one_minus_decay * (shadow - p))
