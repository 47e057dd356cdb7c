"""Transformer blocks of the UNet on the sm_100a kernels.

Drop-in for /root/reference/src/neurosis/modules/attention.py: same class names, constructor
signatures, attribute names (=> identical state-dict keys) and call conventions as
`GEGLU` (:50-57), `FeedForward` (:60-74), `MemoryEfficientCrossAttention` /
`TorchSDPCrossAttention` / `CrossAttention` (:187-417), `BasicTransformerBlock` (:420-511) and
`SpatialTransformer` (:567-667).  The math is executed by libnk_b200.so:

  LayerNorm          -> nk_layernorm_fwd/bwd
  to_q/k/v, to_out   -> nk_linear_* (tcgen05 GEMM; `+ x` residual fused in the to_out / FF epilogue)
  softmax(QK^T/sqrt d)V -> nk_attention_fwd (flash-style, TMEM) ; backward: batched tcgen05 GEMMs
  GEGLU              -> GEMM + nk_geglu_fwd/bwd (exact-erf GELU)
  GroupNorm(eps 1e-6)-> nk_groupnorm_fwd/bwd on NHWC

Activations are bf16 (the reference runs this stack under bf16 autocast); parameters stay fp32
nn.Parameters exactly as in the reference so its checkpoints load unchanged.
"""
from __future__ import annotations

import os

from typing import Any, Optional

import torch
from torch import Tensor, nn

from .. import ops
from .util import as_nhwc, from_nhwc, zero_module


FUSED_GEGLU_MIN_DIM = 1 << 30 if os.environ.get("NK_NO_FUSED_GEGLU") else int(os.environ.get("NK_FUSED_GEGLU_MIN_DIM", "1024"))


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x: Tensor) -> Tensor:
        return ops.geglu(ops.linear(x, self.proj.weight, self.proj.bias))


class _GeluLinear(nn.Sequential):
    """non-gated FeedForward input projection (nn.Sequential(nn.Linear, nn.GELU) in the reference)."""

    def __init__(self, dim: int, inner: int):
        super().__init__(nn.Linear(dim, inner), nn.GELU())

    def forward(self, x: Tensor) -> Tensor:
        raise NotImplementedError("glu=False feed-forward is not used by any reference config")


class FeedForward(nn.Module):
    def __init__(self, dim: int, dim_out: Optional[int] = None, mult: int = 4, glu: bool = False,
                 dropout: float = 0.0):
        super().__init__()
        inner_dim = int(dim * mult)
        dim_out = dim_out or dim
        project_in = GEGLU(dim, inner_dim) if glu else _GeluLinear(dim, inner_dim)
        if dropout:
            raise NotImplementedError("dropout > 0 is not supported (all reference configs use 0.0)")
        self.net = nn.Sequential(project_in, nn.Dropout(dropout), nn.Linear(inner_dim, dim_out))

    def forward(self, x: Tensor, residual: Optional[Tensor] = None) -> Tensor:
        out = self.net[2]
        proj = getattr(self.net[0], "proj", None)
        if proj is not None and proj.out_features % 32 == 0 and proj.in_features >= FUSED_GEGLU_MIN_DIM:
            # gate (forward) and its derivative (backward) live in the epilogues of the two GEMMs.  Only where the
            # reduction is long enough to hide the heavier epilogue behind the main loop: at dim 640 (K = 640, 10
            # k-iterations per tile) the fused kernels measured slower than GEMM + elementwise kernel
            # (profiles/r02_breakdown_v26.txt: 8.9 vs 7.3 ms forward, 6.9 vs 4.6 ms backward at 65536 tokens).
            return ops.feed_forward_geglu(x, proj.weight, proj.bias, out.weight, out.bias, residual)
        h = self.net[0](x)
        return ops.linear(h, out.weight, out.bias, residual)


class CrossAttention(nn.Module):
    """Self / cross attention.  State-dict keys: to_q.weight, to_k.weight, to_v.weight,
    to_out.0.weight, to_out.0.bias (reference attention.py:283-290)."""

    def __init__(self, query_dim: int, context_dim: Optional[int] = None, heads: int = 8, dim_head: int = 64,
                 dropout: float = 0.0, **kwargs: Any):
        super().__init__()
        inner_dim = dim_head * heads
        context_dim = context_dim or query_dim
        if dropout:
            raise NotImplementedError("attention dropout > 0 is not supported")
        self.heads = heads
        self.dim_head = dim_head
        self.scale = dim_head ** -0.5
        self.to_q = nn.Linear(query_dim, inner_dim, bias=False)
        self.to_k = nn.Linear(context_dim, inner_dim, bias=False)
        self.to_v = nn.Linear(context_dim, inner_dim, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, query_dim), nn.Identity())
        # self-attention with head_dim 64: one stacked q|k|v GEMM per direction (NK_NO_FUSED_QKV=1: A/B switch)
        self.fuse_qkv = not os.environ.get("NK_NO_FUSED_QKV")

    def forward(self, x: Tensor, context: Optional[Tensor] = None, mask: Optional[Tensor] = None,
                additional_tokens: Optional[Tensor] = None, n_times_crossframe_attn_in_self: int = 0,
                residual: Optional[Tensor] = None) -> Tensor:
        if mask is not None or additional_tokens is not None or n_times_crossframe_attn_in_self:
            raise NotImplementedError("mask / additional_tokens / cross-frame attention are not supported")
        b, n, _ = x.shape
        h, d = self.heads, self.dim_head
        out = self.to_out[0]
        if (context is None and d <= 64 and d % 8 == 0 and self.fuse_qkv and x.dtype == torch.bfloat16
                and self.to_k.weight.shape == self.to_q.weight.shape):
            # self-attention: q, k, v are one GEMM (forward, data gradient and weight gradient alike)
            o = ops.self_attention_qkv(x, self.to_q.weight, self.to_k.weight, self.to_v.weight, h, self.scale)
            return ops.linear(o, out.weight, out.bias, residual)
        context = x if context is None else ops.cast_bf16(context) if context.dtype != torch.bfloat16 else context
        q = ops.linear(x, self.to_q.weight).view(b, n, h, d)
        if (ops.FUSE_CROSS_KV and context is not x and context.dim() == 3 and context.is_contiguous()
                and self.to_k.weight.shape == self.to_v.weight.shape and d % 8 == 0):
            # the two context projections as one GEMM (ops.CrossAttentionKVFn); off by default, see ops.FUSE_CROSS_KV
            o = ops.cross_attention_kv(q, context, self.to_k.weight, self.to_v.weight, h, self.scale).view(b, n, h * d)
            return ops.linear(o, out.weight, out.bias, residual)
        k = ops.linear(context, self.to_k.weight).view(b, context.shape[1], h, d)
        v = ops.linear(context, self.to_v.weight).view(b, context.shape[1], h, d)
        o = ops.attention(q, k, v, self.scale).view(b, n, h * d)
        return ops.linear(o, out.weight, out.bias, residual)


# the reference's three flavours are numerically the same function; all map to the B200 kernel
MemoryEfficientCrossAttention = CrossAttention
TorchSDPCrossAttention = CrossAttention


class BasicTransformerBlock(nn.Module):
    ATTENTION_MODES = {
        "softmax": CrossAttention,
        "softmax-xformers": CrossAttention,
        "torch-sdp": CrossAttention,
        "b200": CrossAttention,
    }

    def __init__(self, dim: int, n_heads: int, d_head: int, dropout: float = 0.0, context_dim: Optional[int] = None,
                 gated_ff: bool = True, checkpoint: bool = True, disable_self_attn: bool = False,
                 attn_mode: str = "softmax", sdp_backend: Any = None):
        super().__init__()
        if attn_mode not in self.ATTENTION_MODES:
            raise ValueError(f"Unknown attention mode: {attn_mode}")
        attn_cls = self.ATTENTION_MODES[attn_mode]
        self.disable_self_attn = disable_self_attn
        self.attn1 = attn_cls(query_dim=dim, heads=n_heads, dim_head=d_head, dropout=dropout,
                              context_dim=context_dim if disable_self_attn else None, backend=sdp_backend)
        self.ff = FeedForward(dim, dropout=dropout, glu=gated_ff)
        self.attn2 = attn_cls(query_dim=dim, context_dim=context_dim, heads=n_heads, dim_head=d_head,
                              dropout=dropout, backend=sdp_backend)
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.norm3 = nn.LayerNorm(dim)
        # activation checkpointing is a memory policy of the reference (attention.py:482-485); 180 GB of
        # HBM make it unnecessary at the benchmarked batch sizes, so the flag is accepted and ignored.
        self.checkpoint = checkpoint

    def forward(self, x: Tensor, context: Optional[Tensor] = None, additional_tokens: Optional[Tensor] = None,
                n_times_crossframe_attn_in_self: int = 0) -> Tensor:
        return self._forward(x, context, additional_tokens, n_times_crossframe_attn_in_self)

    def _ln(self, norm: nn.LayerNorm, x: Tensor) -> Tensor:
        return ops.layer_norm(x, norm.weight, norm.bias, norm.eps)

    def _ln_res(self, norm: nn.LayerNorm, x: Tensor):
        return ops.layer_norm_residual(x, norm.weight, norm.bias, norm.eps)

    def _forward(self, x: Tensor, context: Optional[Tensor] = None, additional_tokens: Optional[Tensor] = None,
                 n_times_crossframe_attn_in_self: int = 0) -> Tensor:
        if additional_tokens is not None or n_times_crossframe_attn_in_self:
            raise NotImplementedError("additional_tokens / cross-frame attention are not supported")
        # x -> x + f(LN(x)), three times; `_ln_res` hands x back so that the gradient of the residual use is folded into
        # the LayerNorm-backward kernel instead of a separate autograd add
        x, h = self._ln_res(self.norm1, x)
        x = self.attn1(h, context=context if self.disable_self_attn else None, residual=x)
        x, h = self._ln_res(self.norm2, x)
        x = self.attn2(h, context=context, residual=x)
        x, h = self._ln_res(self.norm3, x)
        x = self.ff(h, residual=x)
        return x


class SpatialTransformer(nn.Module):
    """GroupNorm -> proj_in -> depth x BasicTransformerBlock -> proj_out (zero-init) -> + x_in
    (reference attention.py:567-667).  NHWC activations make `b c h w -> b (h w) c` a free view."""

    def __init__(self, in_channels: int, n_heads: int, d_head: int, depth: int = 1, dropout: float = 0.0,
                 context_dim: Optional[int | list[int]] = None, disable_self_attn: bool = False,
                 use_linear: bool = False, attn_type: str = "softmax", use_checkpoint: bool = True,
                 sdp_backend: Any = None):
        super().__init__()
        if context_dim is not None:
            if not isinstance(context_dim, (list, tuple)):
                context_dim = [context_dim]
            context_dim = list(context_dim)
            if len(context_dim) != depth:
                if not all(c == context_dim[0] for c in context_dim):
                    raise ValueError("need homogenous context_dim to match depth automatically")
                context_dim = [context_dim[0]] * depth
        else:
            context_dim = [None] * depth
        self.in_channels = in_channels
        self.norm = nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)
        inner_dim = n_heads * d_head
        if use_linear:
            self.proj_in = nn.Linear(in_channels, inner_dim)
        else:
            self.proj_in = nn.Conv2d(in_channels, inner_dim, kernel_size=1, stride=1, padding=0)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner_dim, n_heads, d_head, dropout=dropout, context_dim=context_dim[d],
                                  disable_self_attn=disable_self_attn, attn_mode=attn_type,
                                  checkpoint=use_checkpoint, sdp_backend=sdp_backend)
            for d in range(depth)
        ])
        if use_linear:
            self.proj_out = zero_module(nn.Linear(inner_dim, in_channels))
        else:
            self.proj_out = zero_module(nn.Conv2d(inner_dim, in_channels, kernel_size=1, stride=1, padding=0))
        self.use_linear = use_linear

    def forward(self, x: Tensor, context: Optional[Tensor | list] = None) -> Tensor:
        if not isinstance(context, list):
            context = [context]
        xn = as_nhwc(x)  # (B, H, W, C) bf16
        b, h, w, c = xn.shape
        t = ops.group_norm(xn, self.norm.weight, self.norm.bias, 32, self.norm.eps, silu=False)
        t = t.view(b, h * w, c)
        # a 1x1 convolution on NHWC is the same GEMM as the linear projection
        w_in = self.proj_in.weight if self.use_linear else self.proj_in.weight.view(self.proj_in.weight.shape[0], -1)
        t = ops.linear(t, w_in, self.proj_in.bias)
        for i, block in enumerate(self.transformer_blocks):
            t = block(t, context=context[i if len(context) > 1 else 0])
        w_out = self.proj_out.weight if self.use_linear else self.proj_out.weight.view(self.proj_out.weight.shape[0], -1)
        y = ops.linear(t, w_out, self.proj_out.bias, residual=xn.view(b, h * w, c))
        return from_nhwc(y.view(b, h, w, c))
