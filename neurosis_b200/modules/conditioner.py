"""Conditioner output contract of the training step (boundary only).

The hot path consumes the *output* of `GeneralConditioner` (reference
modules/encoders/embedding.py:59-61,90-149): a dict {"crossattn": (B,77,ctx), "vector": (B,adm),
"concat": (B,C,H,W)} assembled from the embedders' outputs by rank (2 -> vector, 3 -> crossattn,
4/5 -> concat).  The frozen text encoders are out of scope; `IdentityEncoder` passes precomputed
embeddings through and `ConcatTimestepEmbedderND` (encoders/metadata.py:14-36) builds the SDXL
size/crop Fourier features.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch
from torch import Tensor, nn

OUTPUT_DIM2KEYS = {2: "vector", 3: "crossattn", 4: "concat", 5: "concat"}
KEY2CATDIM = {"vector": 1, "crossattn": 2, "concat": 1}


class AbstractEmbModel(nn.Module):
    def __init__(self, input_key: Optional[str] = None, input_keys: Optional[Sequence[str]] = None,
                 is_trainable: bool = False, ucg_rate: float = 0.0, **kwargs):
        super().__init__()
        self.input_key = input_key
        self.input_keys = list(input_keys) if input_keys is not None else None
        self.is_trainable = is_trainable
        self.ucg_rate = ucg_rate


class IdentityEncoder(AbstractEmbModel):
    def forward(self, x: Tensor) -> Tensor:
        return x


class ConcatTimestepEmbedderND(AbstractEmbModel):
    """(B, k) scalars -> (B, k*outdim) sinusoidal features, cos|sin per scalar, fp32."""

    def __init__(self, outdim: int, **kwargs):
        super().__init__(**kwargs)
        self.outdim = outdim

    def forward(self, x: Tensor) -> Tensor:
        if x.ndim == 1:
            x = x[:, None]
        b, k = x.shape
        half = self.outdim // 2
        freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device=x.device) / half)
        args = x.reshape(-1)[:, None].float() * freqs[None]
        emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
        return emb.reshape(b, k * self.outdim)


class GeneralConditioner(nn.Module):
    def __init__(self, emb_models: Sequence[AbstractEmbModel]):
        super().__init__()
        self.embedders = nn.ModuleList(list(emb_models))

    def forward(self, batch: dict, force_zero_embeddings: Optional[list] = None) -> dict:
        out: dict[str, Tensor] = {}
        for emb in self.embedders:
            with torch.set_grad_enabled(emb.is_trainable):
                if emb.input_key is not None:
                    res = emb(batch[emb.input_key])
                else:
                    res = emb(*[batch[k] for k in emb.input_keys])
            for r in (res if isinstance(res, (list, tuple)) else [res]):
                key = OUTPUT_DIM2KEYS[r.dim()]
                out[key] = torch.cat((out[key], r), KEY2CATDIM[key]) if key in out else r
        return out
