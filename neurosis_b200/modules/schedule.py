"""Sigma tables (discretizations) and per-step sigma draws of the diffusion objective.

Host-side, O(1000) scalar work done once (tables) or O(batch) per step (draws).  Bit-exact parity
with the reference is obtained by issuing the *same sequence of torch ops and dtypes*:

  reference modules/diffusion/discretization.py:17-36   `Discretization.__call__` — note that the
      `do_append_zero` ARGUMENT is ignored there (only the constructor flag is honoured); kept.
  :149-171  LegacyDDPMDiscretization: betas = linspace(sqrt(a), sqrt(b), T, f64)**2 (util.py:31),
      alphas_cumprod = cumprod(1 - betas, dtype=f32), sigma = sqrt((1-acp)/acp) in f32, flipped.
  modules/diffusion/sampling/sigma_generators.py:17-166 — generators; `DiscreteSigmaGenerator`
      does `clamp(t.long(), 0, num_idx-1)` on t in [0,1) (always index 0), reproduced literally.

Unlike the reference the tables carry no autograd graph (reference discretization.py:164-166
leaves `requires_grad_(True)` on them, which breaks the second backward — SURVEY.md §0.9).
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from math import log
from typing import Optional

import numpy as np
import torch
from torch import Tensor


def append_zero(x: Tensor) -> Tensor:
    return torch.cat([x, x.new_zeros([1])])


def append_dims(x: Tensor, ndim: int) -> Tensor:
    extra = ndim - x.ndim
    if extra < 0:
        raise ValueError(f"can't extend tensor from {x.ndim} to {ndim} dimensions!")
    return x[(...,) + (None,) * extra]


def make_beta_schedule(schedule: str, n_timestep: int, linear_start: float = 1e-4, linear_end: float = 2e-2,
                       cosine_s: float = 8e-3) -> Tensor:
    f64 = torch.float64
    if schedule == "linear":
        return torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=f64) ** 2
    if schedule == "cosine":
        ts = torch.arange(n_timestep + 1, dtype=f64) / n_timestep + cosine_s
        alphas = torch.cos(ts / (1 + cosine_s) * np.pi / 2).pow(2)
        alphas = alphas / alphas[0]
        return torch.clamp(1 - alphas[1:] / alphas[:-1], min=0, max=0.999)
    if schedule == "sqrt_linear":
        return torch.linspace(linear_start, linear_end, n_timestep, dtype=f64)
    if schedule == "sqrt":
        return torch.linspace(linear_start, linear_end, n_timestep, dtype=f64) ** 0.5
    raise ValueError(f"unknown schedule: {schedule}")


def spaced_steps(num_substeps: int, max_step: int) -> np.ndarray:
    return np.linspace(max_step - 1, 0, num_substeps, endpoint=False).astype(int)[::-1]


class Discretization(ABC):
    def __init__(self, do_append_zero: bool = True):
        self.do_append_zero = do_append_zero

    def __call__(self, n: int, do_append_zero: bool = True, device: str | torch.device = "cpu",
                 flip: bool = False) -> Tensor:
        sigmas = self.get_sigmas(n, device=device)
        if self.do_append_zero:  # sic: the call argument is not consulted (reference behaviour)
            sigmas = append_zero(sigmas)
        return sigmas.flip((0,)) if flip else sigmas

    @abstractmethod
    def get_sigmas(self, n: int, device: str | torch.device) -> Tensor:
        raise NotImplementedError


class LegacyDDPMDiscretization(Discretization):
    def __init__(self, linear_start: float = 0.00085, linear_end: float = 0.0120, num_timesteps: int = 1000):
        super().__init__()
        self.num_timesteps = num_timesteps
        self.alphas = 1.0 - make_beta_schedule("linear", num_timesteps, linear_start, linear_end)
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0, dtype=torch.float32)

    def get_sigmas(self, n: int, device: str | torch.device = "cpu") -> Tensor:
        if n < self.num_timesteps:
            acp = self.alphas_cumprod[spaced_steps(n, self.num_timesteps).copy()].clone()  # .copy(): the reference crashes here (negative numpy stride)
        elif n == self.num_timesteps:
            acp = self.alphas_cumprod.clone()
        else:
            raise ValueError(f"n ({n}) must be less than or equal to num_timesteps ({self.num_timesteps})")
        sigmas = ((1 - acp) / acp) ** 0.5
        return sigmas.flip(0).to(device, dtype=torch.float32)


class EDMcDiscretization(Discretization):
    def __init__(self, sigma_min: float = 0.001, sigma_max: float = 1000.0):
        super().__init__()
        self.sigma_min, self.sigma_max = sigma_min, sigma_max

    def get_sigmas(self, n: int, device: str | torch.device = "cpu") -> Tensor:
        return torch.linspace(log(self.sigma_min), log(self.sigma_max), n, dtype=torch.float32).exp().flip(0).to(device)


class EDMcSimpleDiscretization(Discretization):
    def __init__(self, sigma_min: float = 0.001, sigma_max: float = 1000.0, num_sigmas: int = 1000):
        super().__init__()
        self.sigma_min, self.sigma_max, self.num_sigmas = sigma_min, sigma_max, num_sigmas

    def get_sigmas(self, n: int, device: str | torch.device = "cpu") -> Tensor:
        table = torch.linspace(log(self.sigma_min), log(self.sigma_max), self.num_sigmas, dtype=torch.float32).exp()
        step = len(table) / n
        picked = [float(table[-(1 + int(i * step))]) for i in range(n)] + [0.0]
        return torch.tensor(picked).to(device)


class RectifiedFlowDiscretization(Discretization):
    def __init__(self, start_shift: float = 0.0, end_shift: float = 0.001, do_append_zero: bool = False):
        super().__init__(do_append_zero=do_append_zero)
        self.start_shift, self.end_shift = start_shift, end_shift

    def get_sigmas(self, n: int, device: str | torch.device = "cpu") -> Tensor:
        t = torch.linspace(self.start_shift, 1 - self.end_shift, n, dtype=torch.float64)
        return (t / (1.0 - t)).flip(0).to(device, dtype=torch.float32)


class RectifiedFlowComfyDiscretization(RectifiedFlowDiscretization):
    def get_sigmas(self, n: int, device: str | torch.device = "cpu") -> Tensor:
        t = torch.linspace(self.start_shift, 1 - self.end_shift, n, dtype=torch.float64)
        return t.flip(0).to(device, dtype=torch.float32)


class TanZeroSNRDiscretization(Discretization):
    def __init__(self, start_shift: float = 0.001, end_shift: float = 0.001, scale: float = 1.0):
        super().__init__()
        self.start_shift, self.end_shift, self.scale = start_shift, end_shift, scale

    def get_sigmas(self, n: int, device: str | torch.device = "cpu") -> Tensor:
        half_pi = torch.acos(torch.zeros(1, dtype=torch.float64))[0]
        grid = torch.linspace(self.start_shift, half_pi - self.end_shift, n, dtype=torch.float64)
        return torch.tan(grid).mul(self.scale).flip(0).to(device, dtype=torch.float32)


class EDMDiscretization(Discretization):
    def __init__(self, sigma_min: float = 0.002, sigma_max: float = 80.0, rho: float = 7.0):
        super().__init__()
        self.sigma_min, self.sigma_max, self.rho = sigma_min, sigma_max, rho

    def get_sigmas(self, n: int, device: str | torch.device = "cpu") -> Tensor:
        ramp = torch.linspace(0, 1, n, device=device, dtype=torch.float32)
        lo, hi = self.sigma_min ** (1 / self.rho), self.sigma_max ** (1 / self.rho)
        return (hi + ramp * (lo - hi)) ** self.rho


# ------------------------------------------------------------------------------------------------
# per-step sigma draws
# ------------------------------------------------------------------------------------------------
class SigmaGenerator(ABC):
    @abstractmethod
    def __call__(self, n_samples: int, t: Optional[Tensor] = None) -> Tensor:
        raise NotImplementedError


def _t64(n_samples: int, t: Optional[Tensor]) -> Tensor:
    return t.to(torch.float64) if t is not None else torch.rand((n_samples,), dtype=torch.float64)


class EDMSigmaGenerator(SigmaGenerator):
    def __init__(self, p_mean: float = -1.2, p_std: float = 1.2, scale: float = 2.0):
        self.p_mean, self.p_std, self.scale = p_mean, p_std, scale

    def __call__(self, n_samples: int, t: Optional[Tensor] = None) -> Tensor:
        t = t.to(torch.float32) if t is not None else torch.randn((n_samples,), dtype=torch.float32)
        return (self.p_mean + self.p_std * t).exp() * self.scale


class DiscreteSigmaGenerator(SigmaGenerator):
    def __init__(self, discretization: Discretization, num_idx: int = 1000, do_append_zero: bool = True,
                 flip: bool = True):
        self.num_idx = num_idx
        self.sigmas = discretization(num_idx, do_append_zero=do_append_zero, flip=flip)

    def idx_to_sigma(self, idx) -> Tensor:
        return self.sigmas[idx]

    def __call__(self, n_samples: int, t: Optional[Tensor] = None) -> Tensor:
        if t is not None:
            idx = torch.clamp(t.long(), 0, self.num_idx - 1)
        else:
            idx = torch.randint(0, self.num_idx, (n_samples,))
        return self.idx_to_sigma(idx)


class CosineScheduleSigmaGenerator(SigmaGenerator):
    def __init__(self, s: float = 0.008, sigma_data: float = 1.0):
        self.s = torch.tensor([s])
        self.sigma_data = sigma_data
        self.min_var = torch.cos(self.s / (1 + self.s) * torch.pi * 0.5) ** 2

    def __call__(self, n_samples: int, t: Optional[Tensor] = None, shift: int = 1, return_logSNR: bool = False):
        if t is None:
            t = (1 - torch.rand(n_samples)).add(0.001).clamp(0.001, 1.0)
        s, min_var = self.s.to(t.device), self.min_var.to(t.device)
        var = torch.cos((s + t) / (1 + s) * torch.pi * 0.5).clamp(0, 1) ** 2 / min_var
        var = 0.0001 + var * 0.9999
        log_snr = (var / (1 - var)).log()
        if shift != 1:
            log_snr += 2 * np.log(1 / shift)
        return log_snr if return_logSNR else torch.exp(-log_snr / 2) * self.sigma_data


class TanScheduleSigmaGenerator(SigmaGenerator):
    def __init__(self, start_shift: float = 0.001, end_shift: float = 0.001, scale: float = 1.0, clip: bool = True):
        self.start_shift, self.end_shift, self.scale, self.clip = start_shift, end_shift, scale, clip

    def __call__(self, n_samples: int, t: Optional[Tensor] = None) -> Tensor:
        quarter_turn = torch.acos(torch.zeros(1, dtype=torch.float64))
        angle = quarter_turn * _t64(n_samples, t)
        if self.clip:
            angle = angle.clip(torch.tensor([self.start_shift], dtype=torch.float64), quarter_turn - self.end_shift)
        return torch.tan(angle).mul(self.scale).to(torch.float32)


class RectifiedFlowSigmaGenerator(SigmaGenerator):
    def __init__(self, start_shift: float = 0.0, end_shift: float = 0.001, clip: bool = True):
        self.start_shift, self.end_shift, self.clip = start_shift, end_shift, clip

    def _t(self, n_samples: int, t: Optional[Tensor]) -> Tensor:
        t = _t64(n_samples, t)
        return t.clip(self.start_shift, 1 - self.end_shift) if self.clip else t

    def __call__(self, n_samples: int, t: Optional[Tensor] = None) -> Tensor:
        t = self._t(n_samples, t)
        return (t / (1 - t)).to(torch.float32)


class RectifiedFlowComfySigmaGenerator(RectifiedFlowSigmaGenerator):
    def __call__(self, n_samples: int, t: Optional[Tensor] = None) -> Tensor:
        return self._t(n_samples, t).to(torch.float32)
