"""The diffusion training step: VAE latent encode -> loss -> hooks -> mean -> backward.

Host-side mirror of `DiffusionEngine.{encode_first_stage, forward, training_step}`
(/root/reference/src/neurosis/models/diffusion.py:186-233) without the Lightning dependency: the
30 lines of glue there contain no arithmetic beyond `* scale_factor` and `loss.mean()`.  Everything
heavy is delegated to the drop-in modules of `neurosis_b200.modules`.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
from torch import Tensor, nn

from .modules.loss import DiffusionLoss, LossHook, OpenAIWrapper
from .modules.denoiser import Denoiser


class DiffusionEngine(nn.Module):
    def __init__(self, model: nn.Module, denoiser: Denoiser, first_stage_model: Optional[nn.Module],
                 conditioner: nn.Module, loss_fn: DiffusionLoss, scale_factor: float = 1.0,
                 input_key: str = "image", vae_batch_size: Optional[int] = None,
                 forward_hooks: Sequence[LossHook] = (), use_ema: bool = False, ema_decay_rate: float = 0.9999,
                 ckpt_path: Optional[str] = None, **kwargs):
        super().__init__()
        self.model = model if isinstance(model, OpenAIWrapper) else OpenAIWrapper(model)
        self.denoiser = denoiser
        self.first_stage_model = first_stage_model
        if first_stage_model is not None:
            for p in first_stage_model.parameters():
                p.requires_grad_(False)
        self.conditioner = conditioner
        self.loss_fn = loss_fn
        self.scale_factor = scale_factor
        self.input_key = input_key
        self.vae_batch_size = vae_batch_size
        self.forward_hooks = list(forward_hooks)
        self.global_step = 0
        # EMA of the UNet wrapper (reference models/diffusion.py:92-98) on the multi-tensor kernel; the shadow buffers
        # appear under `model_ema.*` in the state dict as in the reference
        self.use_ema = use_ema
        if use_ema:
            from .optim import LitEma
            self.model_ema = LitEma(self.model, decay=ema_decay_rate)
        else:
            self.model_ema = None
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path)

    def init_from_ckpt(self, path) -> None:
        """reference models/diffusion.py:127-144 (strict=False, relocated VAE keys tolerated)."""
        from .checkpoint import init_from_ckpt
        self.last_ckpt_report = init_from_ckpt(self, path)

    def on_train_batch_end(self, *args, **kwargs) -> None:
        """EMA update after the optimizer step (reference models/diffusion.py:242-244)."""
        if self.use_ema:
            self.model_ema(self.model)

    def get_input(self, batch: dict) -> Tensor:
        return batch[self.input_key]

    @torch.no_grad()
    def encode_first_stage(self, x: Tensor) -> Tensor:
        enc = self.first_stage_model
        if self.vae_batch_size is None or x.shape[0] <= self.vae_batch_size:
            z = enc(x, regularize=True) if _takes_regularize(enc) else enc.encode(x)
        else:
            bs = self.vae_batch_size
            parts = [enc(x[i: i + bs], regularize=True) if _takes_regularize(enc) else enc.encode(x[i: i + bs])
                     for i in range(0, x.shape[0], bs)]
            z = torch.cat(parts, 0)
        return self.scale_factor * z

    def forward(self, x: Tensor, batch: dict, **kwargs):
        return self.loss_fn(self.model, self.denoiser, self.conditioner, x, batch, return_dict=True, **kwargs)

    def training_step(self, batch: dict, batch_idx: int = 0) -> Tensor:
        for hook in self.forward_hooks:
            batch = hook.pre_hook(None, self, batch, batch_idx)
        x = self.get_input(batch)
        if self.first_stage_model is not None:
            x = self.encode_first_stage(x)
        batch["global_step"] = self.global_step
        loss, loss_dict = self(x, batch)
        for hook in self.forward_hooks:
            loss, loss_dict = hook(self, batch, loss, loss_dict)
        self.last_loss_dict = loss_dict
        return loss.mean()


def _takes_regularize(enc: nn.Module) -> bool:
    return not hasattr(enc, "encode") or enc.__class__.__name__ == "Encoder"
