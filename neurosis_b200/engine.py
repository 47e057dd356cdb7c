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
                 ckpt_path: Optional[str] = None, optimizer=None, scheduler=None, **kwargs):
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
        self.optimizer, self.scheduler = optimizer, scheduler  # callables: params -> optimizer, optimizer -> scheduler
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

    def configure_optimizers(self):
        """parameter groups of the reference (models/diffusion.py:261-296): one "UNet" group plus one per trainable
        embedder, each with its optional `initial_lr` (`base_lr` attribute).  `self.optimizer` / `self.scheduler` are
        the callables the YAML's `optimizer:` / `scheduler:` nodes produce (default: the fused Adafactor with the
        example configs' arguments)."""
        groups = []
        unet = {"name": "UNet", "params": list(self.model.parameters())}
        if getattr(self.model, "base_lr", None) is not None:
            unet["initial_lr"] = self.model.base_lr
        groups.append(unet)
        for emb in getattr(self.conditioner, "embedders", []):
            if getattr(emb, "is_trainable", False):
                g = {"name": getattr(emb, "name", emb.__class__.__name__), "params": list(emb.parameters())}
                if getattr(emb, "base_lr", None) is not None:
                    g["initial_lr"] = emb.base_lr
                groups.append(g)
        if self.optimizer is not None:
            opt = self.optimizer(groups)
        else:
            from .optim import Adafactor
            opt = Adafactor(groups, scale_parameter=True, relative_step=True, warmup_init=True)
        if self.scheduler is not None:
            return {"optimizer": opt, "lr_scheduler": {"scheduler": self.scheduler(opt), "interval": "step"}}
        return opt

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
            # the reference ignores pre_hook's return value (models/diffusion.py:209-210, hooks mutate the batch in
            # place); a hook that returns a new dict is honoured, one that returns None keeps the batch
            new_batch = hook.pre_hook(None, self, batch, batch_idx)
            if new_batch is not None:
                batch = new_batch
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
