"""ORACLE (test infrastructure): CPU restatement of the diffusion objective around the UNet.

Follows /root/reference/src/neurosis/modules/diffusion/
  discretization.py:17-36,149-171 + util.py:22-46  LegacyDDPM sigma table (f64 betas, f32 cumprod)
  sampling/sigma_generators.py:38-57               DiscreteSigmaGenerator (clamp(t.long()) / randint)
  denoiser.py:28-57,83-97                          DiscreteDenoiser: nearest-sigma quantisation (argmin),
                                                   c_skip/c_out/c_in/c_noise, D = F*c_out + z*c_skip
  denoiser_preconditioning.py:33-44, denoiser_weighting.py:22-25   Eps preconditioning / weighting
  loss.py:105-155 + modules/losses/functions.py:81-94   noise mix and per-sample weighted MSE
and the engine glue models/diffusion.py:186-233 (x scale_factor, loss.mean()).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
from torch import Tensor


def ddpm_sigma_table(n: int = 1000, linear_start: float = 0.00085, linear_end: float = 0.0120,
                     num_timesteps: int = 1000, flip: bool = False) -> Tensor:
    """(n+1,) fp32: descending sigmas + trailing 0.0 (flip=False) or ascending with leading 0.0 (flip=True).
    The trailing zero is always present: Discretization.__call__ ignores its do_append_zero argument."""
    betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, num_timesteps, dtype=torch.float64) ** 2
    acp = torch.cumprod(1.0 - betas, dim=0, dtype=torch.float32)
    if n < num_timesteps:
        import numpy as np
        steps = np.linspace(num_timesteps - 1, 0, n, endpoint=False).astype(int)[::-1]
        acp = acp[steps.copy()]
    sig = (((1 - acp) / acp) ** 0.5).flip(0).to(torch.float32)
    sig = torch.cat([sig, sig.new_zeros([1])])
    return sig.flip((0,)) if flip else sig


def discrete_sigma_draw(table_flipped: Tensor, num_idx: int, n: int, t: Optional[Tensor]) -> Tensor:
    idx = torch.clamp(t.long(), 0, num_idx - 1) if t is not None else torch.randint(0, num_idx, (n,))
    return table_flipped[idx]


def sigma_to_idx(table: Tensor, sigma: Tensor) -> Tensor:
    return (sigma - table[:, None]).abs().argmin(dim=0).view(sigma.shape)


def eps_preconditioning(sigma: Tensor):
    """(c_skip, c_out, c_in, c_noise) of EpsPreconditioning."""
    return torch.ones_like(sigma), -sigma, 1.0 / (sigma ** 2.0 + 1.0) ** 0.5, sigma.clone()


def discrete_denoise(network: Callable, table: Tensor, z: Tensor, sigma: Tensor, cond: dict) -> Tensor:
    sigma = table[sigma_to_idx(table, sigma)]
    shape = sigma.shape
    sb = sigma[(...,) + (None,) * (z.ndim - sigma.ndim)]
    c_skip, c_out, c_in, c_noise = eps_preconditioning(sb)
    t_idx = sigma_to_idx(table, c_noise.reshape(shape))
    out = network(z * c_in.to(z.dtype), t_idx, cond)
    return out * c_out.to(z.dtype) + z * c_skip.to(z.dtype)


def diffusion_loss(network: Callable, table: Tensor, x: Tensor, cond: dict, sigmas: Tensor, noise: Tensor) -> Tensor:
    """edm objective with EpsWeighting: loss[b] = mean((D - x)^2) * sigma^-2 in fp32."""
    sigmas = sigmas.to(x)
    sb = sigmas[(...,) + (None,) * (x.ndim - 1)]
    z = x + sb * noise
    D = discrete_denoise(network, table, z, sigmas, cond)
    w = sigmas ** -2.0
    per = ((D.float() - x.float()) ** 2).flatten(1).mean(1)
    return per * w.float()


def rf_comfy_loss(network: Callable, x: Tensor, cond: dict, sigmas: Tensor, noise: Tensor) -> Tensor:
    """rectified-flow objective (loss.py:130-139) with a continuous `Denoiser` (denoiser.py:28-57, output mode "F"),
    RectifiedFlowComfyPreconditioning (denoiser_preconditioning.py:93-105: c_in = (s^2 + (1-s)^2)^-1/2, c_noise = 1000 s)
    and RectifiedFlowComfyWeighting (denoiser_weighting.py:58-75, float64): loss[b] = mean((F - noise)^2) * w(s)."""
    sigmas = sigmas.to(x)
    sb = sigmas[(...,) + (None,) * (x.ndim - 1)]
    z = (1.0 - sb) * x + sb * noise
    c_in = (sb ** 2.0 + (1.0 - sb) ** 2.0) ** -0.5
    c_noise = (1000.0 * sb).reshape(sigmas.shape)
    F_out = network(z * c_in.to(z.dtype), c_noise, cond)
    t = sigmas.to(torch.float64)
    half_pi = torch.acos(torch.zeros(1, dtype=torch.float64))[0]
    w = (1 / (1 - t) ** 2) * ((1 / (1.0 * (4.0 * half_pi) ** 0.5)) * (1 / (t * (1.0 - t)))
                              * torch.exp(-0.5 * (torch.log(t / (1 - t)) - 0.0) ** 2 / 1.0 ** 2))
    per = ((F_out.float() - noise.float()) ** 2).flatten(1).mean(1)
    return per * w.float()
