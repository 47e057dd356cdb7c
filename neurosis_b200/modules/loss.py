"""Diffusion loss, network wrappers and the loss-hook boundary.

Drop-in for /root/reference/src/neurosis/modules/diffusion/loss.py:20-157 (`DiffusionLoss`,
`StandardDiffusionLoss`), wrappers.py:7-40 (`IdentityWrapper`, `OpenAIWrapper`) and
modules/hooks/common.py:11-51 (`LossHook`).  Sampling order and dtypes of the reference are kept:
t ~ U[0,1) float64 from the *CPU* generator, sigma = generator(B, t).to(inputs), noise =
randn_like(inputs) from the device generator.  Element-wise work is done by nk_noise_mix,
nk_lincomb_per_sample and nk_weighted_mse_{fwd,bwd}.

`TagFrequencyHook` implements the per-sample tag-frequency loss scaling named by
configs/hook/loss_scale.example.yml.  The reference does NOT ship that hook (its class path
`neurosis.dataset.processing.TagFrequencyHook` does not exist in the tree), so its arithmetic is
defined here against the config's parameter names — parity unpinned, see DESIGN.md.
"""
from __future__ import annotations

import importlib
import random
from abc import ABC, abstractmethod
from collections import defaultdict
from typing import Callable, Optional

import torch
from torch import Tensor, nn

from .. import ops
from .denoiser import Denoiser, DenoiserWeighting
from .schedule import SigmaGenerator, append_dims


# batch key under which a `LossHook.pre_hook` leaves per-sample loss multipliers for the loss to consume: they are folded
# into the (B,) weight vector of the weighted-MSE / L1 reduction kernel instead of being a separate
# multiply on the reduced loss; APPLIED_KEY tells the hook's post-loss call that this already happened
SAMPLE_WEIGHT_KEY = "_nk_loss_sample_weights"
APPLIED_KEY = "_nk_loss_sample_weights_applied"


class DiffusionLoss(ABC, nn.Module):
    def __init__(self, noise_offset: float = 0.0, noise_offset_chance: float = 0.0, *args, **kwargs):
        super().__init__()
        self.noise_offset = min(max(noise_offset, 0.0), 1.0)
        self.noise_offset_chance = min(max(noise_offset_chance, 0.0), 1.0)

    def apply_noise_offset(self, noise: Tensor, inputs: Tensor) -> Tensor:
        if self.noise_offset <= 0:
            return noise
        if self.noise_offset_chance == 1.0 or random.random() < self.noise_offset_chance:
            offset = torch.randn(inputs.shape[:2] + (1,) * (inputs.ndim - 2)).to(noise)
            return noise + self.noise_offset * offset
        return noise

    def forward(self, network: nn.Module, denoiser: Denoiser, conditioner: Callable[[dict], dict], inputs: Tensor,
                batch: dict, return_dict: bool = False):
        cond = conditioner(batch)
        return self._forward(network, denoiser, cond, inputs, batch, return_dict)

    @abstractmethod
    def _forward(self, network, denoiser, cond, inputs, batch, return_dict=False): ...

    @abstractmethod
    def get_loss(self, outputs: Tensor, target: Tensor, w: Tensor): ...


class StandardDiffusionLoss(DiffusionLoss):
    def __init__(self, sigma_generator: SigmaGenerator, loss_weighting: DenoiserWeighting, loss_type: str = "l2",
                 snr_gamma: float = 0.0, noise_offset: float = 0.0, noise_offset_chance: float = 0.0,
                 input_keys: str | list[str] = [], objective_type: str = "edm"):
        super().__init__(noise_offset, noise_offset_chance)
        self.sigma_generator = sigma_generator
        self.loss_weighting = loss_weighting
        self.snr_gamma = snr_gamma
        self.objective_type = str(getattr(objective_type, "value", objective_type)).lower()
        kind = str(getattr(loss_type, "value", loss_type)).lower()
        if kind in ("l2", "mse"):
            self.loss_type = "l2"
        elif kind == "l1":
            self.loss_type = "l1"
        else:
            raise ValueError(f"Unknown loss type: '{loss_type}'")
        if self.objective_type not in ("edm", "rf"):
            raise ValueError(f"Unknown objective type: '{objective_type}'")
        self.input_keys = set(input_keys if isinstance(input_keys, list) else [input_keys])

    def draw_sigmas(self, n: int) -> Tensor:
        """host-side sigma draw exactly as `_forward` does it (CPU generator, float64 t)."""
        return self.sigma_generator(n, torch.rand((n,), dtype=torch.float64))

    def _forward(self, network, denoiser, cond, inputs, batch, return_dict=False, *,
                 t: Optional[Tensor] = None, noise: Optional[Tensor] = None, sigmas: Optional[Tensor] = None,
                 sample_weights: Optional[Tensor] = None):
        """`t` / `noise` overrides exist for parity tests (the reference draws both internally); `sigmas` (a device
        tensor filled from `draw_sigmas`) lets the whole step be captured in a CUDA graph without a host copy;
        `sample_weights` (B,) — or `batch[SAMPLE_WEIGHT_KEY]`, left there by a loss hook's `pre_hook` — are per-sample
        loss multipliers that enter the reduction kernel through its weight vector: loss[b] = sw[b] * w(sigma[b]) *
        mean((D - T)^2)."""
        extra = {k: batch[k] for k in batch if k in self.input_keys}
        n = inputs.shape[0]
        if sigmas is None:
            if t is None:
                t = torch.rand((n,), dtype=torch.float64)
            sigmas = self.sigma_generator(n, t).to(inputs)
        else:
            sigmas = sigmas.to(inputs)
        if noise is None:
            noise = torch.randn_like(inputs)
        noise = self.apply_noise_offset(noise, inputs)
        rf = self.objective_type == "rf"
        z_t = ops.noise_mix(inputs, noise, sigmas, rectified_flow=rf)
        weight = self.loss_weighting(sigmas)
        if sample_weights is None and isinstance(batch, dict) and SAMPLE_WEIGHT_KEY in batch:
            sample_weights = batch[SAMPLE_WEIGHT_KEY]
            batch[APPLIED_KEY] = True
        if sample_weights is not None:
            weight = weight.reshape(-1) * sample_weights.to(device=weight.device, dtype=weight.dtype).reshape(-1)
        if rf:
            out = denoiser(network, z_t, sigmas, cond, "F", **extra)
            loss = self.get_loss(out, noise, weight)
        else:
            out = denoiser(network, z_t, sigmas, cond, "D", **extra)
            loss = self.get_loss(out, inputs, weight)
        if return_dict:
            return loss, {"sigmas": sigmas, "t": t}
        return loss

    def get_loss(self, outputs: Tensor, target: Tensor, weight: Tensor) -> Tensor:
        if self.loss_type == "l1":
            return ops.weighted_l1(outputs, target, weight)
        return ops.weighted_mse(outputs, target, weight)


# ---- network wrappers ------------------------------------------------------------------------------
class IdentityWrapper(nn.Module):
    def __init__(self, diffusion_model: nn.Module, compile_model: bool = False, **kwargs):
        super().__init__()
        # torch.compile is deliberately not applied: the hot path is already hand-written kernels
        self.diffusion_model = diffusion_model

    def forward(self, *args, **kwargs):
        return self.diffusion_model(*args, **kwargs)


class OpenAIWrapper(IdentityWrapper):
    def forward(self, x: Tensor, t: Tensor, c: dict, **kwargs) -> Tensor:
        concat = c.get("concat", None)
        if concat is not None and concat.numel() > 0:
            x = torch.cat((x, concat.to(x)), dim=1)
        return self.diffusion_model(x, timesteps=t, context=c.get("crossattn", None), y=c.get("vector", None),
                                    **kwargs)


# ---- loss hooks --------------------------------------------------------------------------------------
class LossHook(ABC):
    """pre_hook(trainer, module, batch, idx) -> batch ; __call__(module, batch, loss[B], loss_dict) ->
    (loss[B], loss_dict), invoked between the loss and `.mean()` (reference models/diffusion.py:207-226)."""

    def __init__(self, name: Optional[str] = None, **kwargs):
        self.name = name or self.__class__.__name__

    def __call__(self, pl_module, batch: dict, loss: Tensor, loss_dict: dict = {}, **kwargs):
        return self.batch_hook(pl_module, batch, loss, loss_dict, **kwargs)

    def pre_hook(self, trainer, pl_module, batch, batch_idx):
        return batch

    @abstractmethod
    def batch_hook(self, pl_module, batch: dict, loss: Tensor, loss_dict: dict = {}, **kwargs): ...


class TagFreqScale:
    """piecewise-constant multiplier by running tag count: scales = [[count_threshold, multiplier], ...]."""

    def __init__(self, scales: list[list[float]]):
        self.scales = sorted((float(c), float(m)) for c, m in scales)

    def __call__(self, count: float) -> float:
        mult = self.scales[0][1]
        for thr, m in self.scales:
            if count > thr:
                mult = m
        return mult


class TagRewards:
    def __init__(self, **rewards: float):
        self.rewards = {k: float(v) for k, v in rewards.items()}

    def __call__(self, tag: str) -> float:
        return self.rewards.get(tag, 1.0)


class TagFrequencyHook(LossHook):
    """Per-sample loss scaling from running tag frequencies (host-side string work, one (B,) multiply on
    device).  For every caption: split on `tag_sep`, keep tags accepted by `check_fn`; each kept tag has a
    decayed running count (count <- beta*count + 1 when seen); its multiplier is
    freq_scale(count) * tag_rewards(tag); the sample multiplier is the mean over kept tags blended as
    1 + strength * alpha * (mean - 1), and loss[b] is multiplied by it."""

    def __init__(self, input_key: str = "caption", tag_sep: str = " ", check_fn: Optional[str | Callable] = None,
                 alpha: float = 0.2, beta: float = 0.99, strength: float = 1.0,
                 freq_scale: Optional[TagFreqScale] = None, tag_rewards: Optional[TagRewards] = None, **kwargs):
        super().__init__(**kwargs)
        self.input_key, self.tag_sep = input_key, tag_sep
        if isinstance(check_fn, str):
            mod, _, fn = check_fn.rpartition(".")
            check_fn = getattr(importlib.import_module(mod), fn)
        self.check_fn = check_fn
        self.alpha, self.beta, self.strength = alpha, beta, strength
        self.freq_scale = freq_scale or TagFreqScale([[-1, 1.0]])
        self.tag_rewards = tag_rewards or TagRewards()
        self.counts: dict[str, float] = defaultdict(float)

    def sample_weights(self, captions: list) -> list[float]:
        weights = []
        for cap in captions:
            if isinstance(cap, bytes):
                cap = cap.decode("utf-8", errors="ignore")
            tags = [t for t in str(cap).split(self.tag_sep) if t and (self.check_fn is None or self.check_fn(t))]
            if not tags:
                weights.append(1.0)
                continue
            mults = []
            for tag in tags:
                self.counts[tag] = self.beta * self.counts[tag] + 1.0
                mults.append(self.freq_scale(self.counts[tag]) * self.tag_rewards(tag))
            mean = sum(mults) / len(mults)
            weights.append(1.0 + self.strength * self.alpha * (mean - 1.0))
        return weights

    def pre_hook(self, trainer, pl_module, batch, batch_idx):
        """the multipliers depend only on the captions, so they are computed BEFORE the loss and handed to it through
        the batch: the loss folds them into the weight vector of its reduction kernel (no extra pass over the loss)."""
        if isinstance(batch, dict) and self.input_key in batch:
            batch[SAMPLE_WEIGHT_KEY] = torch.tensor(self.sample_weights(list(batch[self.input_key])), dtype=torch.float32)
            batch.pop(APPLIED_KEY, None)
        return batch

    def batch_hook(self, pl_module, batch: dict, loss: Tensor, loss_dict: dict = {}, **kwargs):
        loss_dict = dict(loss_dict)
        if batch.pop(APPLIED_KEY, False):  # the loss already carries the multipliers (pre_hook path)
            w = batch.pop(SAMPLE_WEIGHT_KEY)
            loss_dict[f"{self.name}/scale_mean"] = w.mean()
            return loss, loss_dict
        w = batch.pop(SAMPLE_WEIGHT_KEY, None)  # pre_hook ran but the loss did not consume them (foreign loss class)
        if w is None:
            w = torch.tensor(self.sample_weights(list(batch[self.input_key])), dtype=torch.float32)
        w = w.to(device=loss.device, dtype=loss.dtype)
        loss_dict[f"{self.name}/scale_mean"] = w.mean()
        return loss * w, loss_dict
