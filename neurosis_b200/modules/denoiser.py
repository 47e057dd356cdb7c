"""Denoiser, sigma preconditioning (c_skip, c_out, c_in, c_noise) and loss weightings.

Drop-in for /root/reference/src/neurosis/modules/diffusion/{denoiser.py:14-97,
denoiser_preconditioning.py:8-105, denoiser_weighting.py:7-101}.  The O(batch) scalar formulas stay
torch expressions with the reference's exact dtypes (they decide the *integer* timestep index fed to
the UNet, which must be bit-exact: `sigma_to_idx` = argmin |sigma - table| over the 1001-entry fp32
table, first minimum wins).  The per-element work — `inputs * c_in` and
`net * c_out + inputs * c_skip` — runs in nk_lincomb_per_sample.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Tuple, Union

import torch
from torch import Tensor, nn

from .. import ops
from .schedule import Discretization, append_dims


# ---- preconditioning ---------------------------------------------------------------------------
class DenoiserPreconditioning(ABC):
    def __call__(self, sigma: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        return self.get_c_skip(sigma), self.get_c_out(sigma), self.get_c_in(sigma), self.get_c_noise(sigma)

    @abstractmethod
    def get_c_skip(self, sigma: Tensor) -> Tensor: ...

    @abstractmethod
    def get_c_out(self, sigma: Tensor) -> Tensor: ...

    @abstractmethod
    def get_c_in(self, sigma: Tensor) -> Tensor: ...

    @abstractmethod
    def get_c_noise(self, sigma: Tensor) -> Tensor: ...

    def get_snr(self, sigma: Tensor) -> Tensor:
        return 1 / sigma ** 2.0


class EpsPreconditioning(DenoiserPreconditioning):
    def get_c_skip(self, sigma):
        return torch.ones_like(sigma, device=sigma.device)

    def get_c_out(self, sigma):
        return -sigma

    def get_c_in(self, sigma):
        return 1.0 / (sigma ** 2.0 + 1.0) ** 0.5

    def get_c_noise(self, sigma):
        return sigma.clone()


class VPreconditioning(EpsPreconditioning):
    def get_c_skip(self, sigma):
        return 1.0 / (sigma ** 2 + 1.0)

    def get_c_out(self, sigma):
        return -sigma / (sigma ** 2 + 1.0) ** 0.5


class VPreconditioningWithEDMcNoise(VPreconditioning):
    def get_c_noise(self, sigma):
        return 0.25 * sigma.log()


class EDMPreconditioning(DenoiserPreconditioning):
    def __init__(self, sigma_data: float = 1.0):
        self.sigma_data = sigma_data

    def get_c_skip(self, sigma):
        return self.sigma_data ** 2 / (sigma ** 2 + self.sigma_data ** 2)

    def get_c_out(self, sigma):
        return sigma * self.sigma_data / (sigma ** 2 + self.sigma_data ** 2) ** 0.5

    def get_c_in(self, sigma):
        return 1 / (sigma ** 2 + self.sigma_data ** 2) ** 0.5

    def get_c_noise(self, sigma):
        return 0.25 * sigma.log()


class RectifiedFlowXLPreconditioning(DenoiserPreconditioning):
    def get_c_skip(self, sigma):
        return torch.ones_like(sigma, device=sigma.device)

    def get_c_out(self, sigma):
        return -sigma

    def get_c_in(self, sigma):
        s_t = 1.0 / (1.0 + sigma)
        noise_std = ((1.0 / (sigma + 1.0)) ** 2.0 + (sigma / (sigma + 1.0)) ** 2.0) ** 0.5
        return s_t / noise_std

    def get_c_noise(self, sigma):
        return 1000.0 * (sigma / (1 + sigma))


class RectifiedFlowComfyPreconditioning(RectifiedFlowXLPreconditioning):
    def get_c_in(self, sigma):
        return (sigma ** 2.0 + (1.0 - sigma) ** 2.0) ** -0.5

    def get_c_noise(self, sigma):
        return 1000.0 * sigma


# ---- loss weightings -----------------------------------------------------------------------------
class DenoiserWeighting(ABC):
    @abstractmethod
    def __call__(self, sigma: Tensor) -> Tensor: ...


class UnitWeighting(DenoiserWeighting):
    def __call__(self, sigma):
        return torch.ones_like(sigma, device=sigma.device)


class EpsWeighting(DenoiserWeighting):
    def __call__(self, sigma):
        return sigma ** -2.0


class EDMWeighting(DenoiserWeighting):
    def __init__(self, sigma_data: float = 1.0):
        self.sigma_data = sigma_data

    def __call__(self, sigma):
        return (sigma ** 2 + self.sigma_data ** 2) / (sigma * self.sigma_data) ** 2


def _logit_normal_pi(t: Tensor, logit: Tensor, m: float, s: float) -> Tensor:
    half_pi = torch.acos(torch.zeros(1, dtype=torch.float64))[0]
    return (1 / (s * (4.0 * half_pi) ** 0.5)) * (1 / (t * (1.0 - t))) * torch.exp(-0.5 * (logit - m) ** 2 / s ** 2)


class RectifiedFlowWeighting(DenoiserWeighting):
    def __init__(self, m: float = 0.0, s: float = 1.0):
        self.m, self.s = m, s

    def __call__(self, sigma):
        sigma = sigma.to(torch.float64)
        t = sigma / (1.0 + sigma)
        return (1 / (1 - t) ** 2) * _logit_normal_pi(t, torch.log(sigma), self.m, self.s)


class RectifiedFlowComfyWeighting(RectifiedFlowWeighting):
    def __call__(self, sigma):
        t = sigma.to(torch.float64)
        return (1 / (1 - t) ** 2) * _logit_normal_pi(t, torch.log(t / (1 - t)), self.m, self.s)


class MinSNRGammaModifier(DenoiserWeighting):
    def __init__(self, weighting: DenoiserWeighting, gamma: float = 5, v_pred: bool = False):
        self.weighting, self.gamma, self.v_pred = weighting, gamma, v_pred

    def __call__(self, sigma):
        snr = 1.0 / sigma ** 2
        capped = torch.min(snr, torch.full_like(snr, self.gamma))
        return self.weighting(sigma) * capped.div(snr + 1.0 if self.v_pred else snr)


# ---- denoisers -------------------------------------------------------------------------------------
class Denoiser(nn.Module):
    def __init__(self, preconditioning: DenoiserPreconditioning):
        super().__init__()
        self.preconditioning = preconditioning

    def possibly_quantize_sigma(self, sigma: Tensor) -> Tensor:
        return sigma

    def possibly_quantize_c_noise(self, c_noise: Tensor) -> Tensor:
        return c_noise

    def forward(self, network: nn.Module, inputs: Tensor, sigma: Tensor, cond: dict, output_mode: str = "D",
                **additional_model_inputs) -> Tensor:
        sigma = self.possibly_quantize_sigma(sigma)
        shape = sigma.shape
        c_skip, c_out, c_in, c_noise = self.preconditioning(append_dims(sigma, inputs.ndim))
        c_noise = self.possibly_quantize_c_noise(c_noise.reshape(shape))
        c_in, c_out, c_skip = (c.to(inputs.dtype).reshape(-1) for c in (c_in, c_out, c_skip))
        net_inputs = ops.lincomb_per_sample(inputs, c_in)  # inputs * c_in
        net_outputs = network(net_inputs, c_noise, cond, **additional_model_inputs)
        if output_mode == "F":
            return net_outputs
        return ops.denoise_combine(net_outputs, inputs, c_out, c_skip)  # F * c_out + inputs * c_skip


class DiscreteDenoiser(Denoiser):
    sigmas: Tensor
    log_sigmas: Tensor

    def __init__(self, preconditioning: DenoiserPreconditioning, num_idx: int, discretization: Discretization,
                 do_append_zero: bool = False, quantize_c_noise: bool = True, flip: bool = False):
        super().__init__(preconditioning)
        self.num_idx = num_idx
        self.quantize_c_noise = quantize_c_noise
        self.do_append_zero = do_append_zero
        self.flip = flip
        sigmas = discretization(self.num_idx, do_append_zero=self.do_append_zero, flip=self.flip)
        self.register_buffer("sigmas", sigmas, persistent=False)
        self.register_buffer("log_sigmas", sigmas.log(), persistent=False)

    def sigma_to_idx(self, sigma: Tensor) -> Tensor:
        dists = sigma - self.sigmas[:, None]
        return dists.abs().argmin(dim=0).view(sigma.shape)

    def idx_to_sigma(self, idx: Union[Tensor, int]) -> Tensor:
        return self.sigmas[idx]

    def possibly_quantize_sigma(self, sigma: Tensor) -> Tensor:
        return self.idx_to_sigma(self.sigma_to_idx(sigma))

    def possibly_quantize_c_noise(self, c_noise: Tensor) -> Tensor:
        return self.sigma_to_idx(c_noise) if self.quantize_c_noise else c_noise
