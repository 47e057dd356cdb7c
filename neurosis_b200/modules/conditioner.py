"""Conditioner output contract of the training step (boundary only).

The hot path consumes the *output* of `GeneralConditioner` (reference
modules/encoders/embedding.py:59-61,90-149): a dict {"crossattn": (B,77,ctx), "vector": (B,adm),
"concat": (B,C,H,W)} assembled from the embedders' outputs by rank (2 -> vector, 3 -> crossattn,
4/5 -> concat).  The frozen text encoders are out of scope; `IdentityEncoder` passes precomputed
embeddings through and `ConcatTimestepEmbedderND` (encoders/metadata.py:14-36) builds the SDXL
size/crop Fourier features — on the device through the sinusoidal-embedding kernel when its input is a
CUDA tensor (the vector path of the conditioner), with the reference's per-sample
UCG dropout and `force_zero_embeddings` handling in `GeneralConditioner.forward` (embedding.py:134-142).
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np
import torch
from torch import Tensor, nn

OUTPUT_DIM2KEYS = {2: "vector", 3: "crossattn", 4: "concat", 5: "concat"}
KEY2CATDIM = {"vector": 1, "crossattn": 2, "concat": 1}


class AbstractEmbModel(nn.Module):
    """embedder base of the reference (encoders/embedding.py:17-56): same constructor arguments in the same order
    (`name`, `input_key`, `ucg_rate`, `is_trainable`, `base_lr`); `input_keys` (multi-input embedders) is accepted as a
    keyword.  `base_lr` feeds `DiffusionEngine.configure_optimizers` (per-group `initial_lr`)."""

    def __init__(self, name: Optional[str] = None, input_key: Optional[str] = None, ucg_rate: Optional[float] = 0.0,
                 is_trainable: Optional[bool] = None, base_lr: Optional[float] = None,
                 input_keys: Optional[Sequence[str]] = None, **kwargs):
        super().__init__()
        self.name = name or str(self.__class__.__name__)
        self.input_key = input_key
        self.input_keys = list(input_keys) if input_keys is not None else None
        self.is_trainable = is_trainable or False
        self.ucg_rate = ucg_rate
        self.base_lr = base_lr

    def freeze(self) -> None:
        self.eval()
        self.requires_grad_(False)


class IdentityEncoder(AbstractEmbModel):
    def forward(self, x: Tensor) -> Tensor:
        return x


class ConcatTimestepEmbedderND(AbstractEmbModel):
    """(B, k) scalars -> (B, k*outdim) sinusoidal features, cos|sin per scalar (encoders/metadata.py:14-36 with
    `Timestep` = modules/diffusion/util.py:152-177).  CUDA input: one launch of the embedding kernel, bf16 output (the
    UNet's label_emb consumes bf16); CPU input (host-side shape checks, oracle): fp32 torch expression."""

    def __init__(self, outdim: int, **kwargs):
        super().__init__(**kwargs)
        self.outdim = outdim

    def forward(self, x) -> Tensor:
        if isinstance(x, list):
            x = torch.stack(x, dim=-1)
        if x.ndim == 1:
            x = x[:, None]
        if x.ndim != 2:
            raise ValueError(f"Expected 2D input, got {x.ndim}D")
        b, k = x.shape
        if x.is_cuda:
            from .. import ops
            return ops.timestep_embedding(x.reshape(-1).float().contiguous(), self.outdim).view(b, k * self.outdim)
        half = self.outdim // 2
        freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device=x.device) / half)
        args = x.reshape(-1)[:, None].float() * freqs[None]
        emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
        return emb.reshape(b, k * self.outdim)


class GeneralConditioner(nn.Module):
    def __init__(self, emb_models: Sequence[AbstractEmbModel]):
        super().__init__()
        if len(emb_models) == 0:
            raise ValueError("no embedders were added")
        for idx, emb in enumerate(emb_models):
            if getattr(emb, "input_key", None) is None and getattr(emb, "input_keys", None) is None:
                raise KeyError(f"need either 'input_key' or 'input_keys' for embedder #{idx} {type(emb).__name__}")
        self.embedders = nn.ModuleList(list(emb_models))
        self.rng = np.random.default_rng()

    def forward(self, batch: dict, force_zero_embeddings: Optional[list] = None) -> dict:
        out: dict[str, Tensor] = {}
        force_zero_embeddings = force_zero_embeddings or []
        for emb in self.embedders:
            with torch.set_grad_enabled(emb.is_trainable):
                if emb.input_key is not None:
                    inputs = batch[emb.input_key]
                    if isinstance(inputs, list) and isinstance(emb, ConcatTimestepEmbedderND) and not torch.is_tensor(
                            inputs[0]):
                        # list of (w, h) tuples from the aspect-bucket loader (embedding.py:108-113): ONE host->device
                        # copy of a (B, k) array instead of a Python-list tensor build per embedder
                        ref = batch.get("image")
                        inputs = torch.as_tensor(np.asarray(inputs, dtype=np.float32)).to(
                            ref.device if torch.is_tensor(ref) else "cpu", non_blocking=True)
                    if emb.ucg_rate > 0.0 and emb.input_key == "caption" and self.rng.random() < emb.ucg_rate:
                        inputs = [" "] * len(inputs)
                    res = emb(inputs)
                else:
                    res = emb(*[batch[k] for k in emb.input_keys])
            for r in (res if isinstance(res, (list, tuple)) else [res]):
                key = OUTPUT_DIM2KEYS[r.dim()]
                if emb.input_key is not None and emb.input_key in force_zero_embeddings:
                    r = torch.zeros_like(r)
                elif emb.ucg_rate > 0.0 and emb.input_key != "caption":
                    keep = torch.bernoulli(torch.full((r.shape[0],), 1.0 - emb.ucg_rate, device=r.device))
                    r = r.mul(keep.to(r.dtype).reshape((-1,) + (1,) * (r.dim() - 1)))  # O(batch) mask, per sample
                if key in out:
                    if out[key].dtype != r.dtype:  # bf16 Fourier features next to an fp32 pooled embedding
                        r = r.to(out[key].dtype)
                    out[key] = torch.cat((out[key], r), KEY2CATDIM[key])
                else:
                    out[key] = r
        return out
