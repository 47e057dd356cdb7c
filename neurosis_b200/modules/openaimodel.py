"""UNetModel and its building blocks on the sm_100a kernels.

Drop-in for /root/reference/src/neurosis/modules/diffusion/openaimodel.py: `TimestepBlock` (:52-62),
`TimestepEmbedSequential` (:65-93), `Upsample` (:96-143), `Downsample` (:146-197), `ResBlock`
(:200-342), `Timestep` and `UNetModel` (:446-840) keep their constructor signatures, attribute
names and therefore state-dict keys (`input_blocks.4.0.out_layers.3.weight`, `out.2.weight`, ...).

Inside the network activations are bf16 NHWC; every module accepts / returns logical (N,C,H,W)
tensors (channels-last views between the modules of this package, so no copies are made).

Kernel mapping per ResBlock (reference _forward :315-342):
  GroupNorm+SiLU           -> nk_groupnorm_fwd(silu=1)
  conv3x3 + emb[:, :, None, None]   -> nk_conv2d_fwd with the embedding as per-image bias (fused)
  GroupNorm+SiLU, conv3x3 + skip(x) -> nk_conv2d_fwd with the skip tensor as fused residual
"""
from __future__ import annotations

from abc import abstractmethod
from typing import Optional, Union

import torch
from torch import Tensor, nn

from .. import ops
from .attention import SpatialTransformer
from .util import as_nhwc, from_nhwc, timestep_embedding, zero_module


def conv_nd(dims: int, *args, **kwargs) -> nn.Module:
    if dims != 2:
        raise NotImplementedError("only 2-D UNets are supported (all reference configs use dims=2)")
    return nn.Conv2d(*args, **kwargs)


def _silu_cached(emb: Tensor) -> Tensor:
    """SiLU(emb) is identical for every ResBlock of a forward pass: compute it once per tensor."""
    s = getattr(emb, "_nk_silu", None)
    if s is None:
        s = ops.silu(emb if emb.dtype == torch.bfloat16 else ops.cast_bf16(emb))
        try:
            emb._nk_silu = s
        except AttributeError:
            pass
    return s


def _conv3x3(conv: nn.Conv2d, x: Tensor, bias_img: Optional[Tensor] = None, residual: Optional[Tensor] = None) -> Tensor:
    if conv.kernel_size != (3, 3) or conv.stride != (1, 1) or conv.padding != (1, 1):
        raise NotImplementedError(f"unsupported conv geometry {conv}")
    return ops.conv2d(x, conv.weight, conv.bias, bias_img, residual)


def _conv1x1(conv: nn.Conv2d, x: Tensor) -> Tensor:
    n, h, w, c = x.shape
    y = ops.linear(x.view(n * h * w, c), conv.weight.view(conv.weight.shape[0], -1), conv.bias)
    return y.view(n, h, w, -1)


class TimestepBlock(nn.Module):
    @abstractmethod
    def forward(self, x: Tensor, emb: Tensor) -> Tensor:
        raise NotImplementedError


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    """passes `emb` to TimestepBlocks and `context` to SpatialTransformers (reference :65-93)."""

    def forward(self, x: Tensor, emb: Tensor, context: Optional[Tensor] = None, *args, **kwargs) -> Tensor:
        for layer in self:
            if isinstance(layer, TimestepBlock):
                x = layer(x, emb)
            elif isinstance(layer, SpatialTransformer):
                x = layer(x, context)
            elif isinstance(layer, nn.Conv2d):  # the stem conv of input_blocks[0]
                cin = layer.weight.shape[1]
                xn = as_nhwc(x, cpad=max(cin, 64))
                x = from_nhwc(_conv3x3(layer, xn))
            else:
                x = layer(x)
        return x


class Upsample(nn.Module):
    def __init__(self, channels: int, use_conv: bool, dims: int = 2, out_channels: Optional[int] = None,
                 padding: int = 1, third_up: bool = False, kernel_size: int = 3, scale_factor: int = 2):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.dims = dims
        if dims != 2 or scale_factor != 2:
            raise NotImplementedError("only 2-D nearest 2x upsampling is supported")
        if use_conv:
            self.conv = conv_nd(dims, self.channels, self.out_channels, kernel_size, padding=padding)

    def forward(self, x: Tensor) -> Tensor:
        xn = as_nhwc(x)
        assert xn.shape[-1] == self.channels
        y = ops.upsample2x(xn)
        if self.use_conv:
            y = _conv3x3(self.conv, y)
        return from_nhwc(y)


class Downsample(nn.Module):
    def __init__(self, channels: int, use_conv: bool, dims: int = 2, out_channels: Optional[int] = None,
                 padding: int = 1, third_down: bool = False):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.dims = dims
        if not use_conv or dims != 2 or padding != 1:
            raise NotImplementedError("only the learned 3x3 stride-2 downsampling is supported")
        self.op = conv_nd(dims, self.channels, self.out_channels, 3, stride=2, padding=padding)

    def forward(self, x: Tensor) -> Tensor:
        xn = as_nhwc(x)
        assert xn.shape[-1] == self.channels
        return from_nhwc(ops.conv2d_stride2(xn, self.op.weight, self.op.bias, asymmetric=False))


class ResBlock(TimestepBlock):
    def __init__(self, channels: int, emb_channels: int, dropout: float, out_channels: Optional[int] = None,
                 use_conv: bool = False, use_scale_shift_norm: bool = False, dims: int = 2,
                 use_checkpoint: bool = False, up: bool = False, down: bool = False, kernel_size: int = 3,
                 exchange_temb_dims: bool = False, skip_t_emb: bool = False):
        super().__init__()
        if up or down or use_scale_shift_norm or exchange_temb_dims or skip_t_emb or kernel_size != 3:
            raise NotImplementedError("resblock up/down, scale-shift norm and video options are not supported")
        if dropout:
            raise NotImplementedError("dropout > 0 is not supported (all reference configs use 0.0)")
        self.channels = channels
        self.emb_channels = emb_channels
        self.dropout = dropout
        self.out_channels = out_channels if out_channels is not None else channels
        self.use_conv = use_conv
        self.use_checkpoint = use_checkpoint  # accepted, not needed (see BasicTransformerBlock)
        self.use_scale_shift_norm = use_scale_shift_norm
        self.in_layers = nn.Sequential(
            nn.GroupNorm(32, channels), nn.SiLU(), conv_nd(dims, channels, self.out_channels, 3, padding=1))
        self.updown = False
        self.h_upd = self.x_upd = nn.Identity()
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, self.out_channels))
        self.out_layers = nn.Sequential(
            nn.GroupNorm(32, self.out_channels), nn.SiLU(), nn.Dropout(p=dropout),
            zero_module(conv_nd(dims, self.out_channels, self.out_channels, 3, padding=1)))
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        elif use_conv:
            self.skip_connection = conv_nd(dims, channels, self.out_channels, 3, padding=1)
        else:
            self.skip_connection = conv_nd(dims, channels, self.out_channels, 1)

    def forward(self, x: Tensor, emb: Tensor) -> Tensor:
        return self._forward(x, emb)

    def _forward(self, x: Tensor, emb: Tensor) -> Tensor:
        xn = as_nhwc(x)
        gn1, conv1 = self.in_layers[0], self.in_layers[2]
        gn2, conv2 = self.out_layers[0], self.out_layers[3]
        h = ops.group_norm(xn, gn1.weight, gn1.bias, gn1.num_groups, gn1.eps, silu=True)
        lin = self.emb_layers[1]
        emb_out = ops.linear(_silu_cached(emb), lin.weight, lin.bias, out_f32=True)  # (N, C_out) fp32
        h = _conv3x3(conv1, h, bias_img=emb_out)  # conv + bias + emb[:, :, None, None]
        h = ops.group_norm(h, gn2.weight, gn2.bias, gn2.num_groups, gn2.eps, silu=True)
        if isinstance(self.skip_connection, nn.Identity):
            skip = xn
        elif self.skip_connection.kernel_size == (1, 1):
            skip = _conv1x1(self.skip_connection, xn)
        else:
            skip = _conv3x3(self.skip_connection, xn)
        return from_nhwc(_conv3x3(conv2, h, residual=skip))  # skip(x) + h fused in the epilogue


class Timestep(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.dim = dim

    def forward(self, t: Tensor) -> Tensor:
        return timestep_embedding(t, self.dim)


class _EmbedMLP(nn.Sequential):
    """Linear -> SiLU -> Linear on (N, d) embeddings (time_embed / label_emb[0])."""

    def __init__(self, d_in: int, d_out: int):
        super().__init__(nn.Linear(d_in, d_out), nn.SiLU(), nn.Linear(d_out, d_out))

    def forward(self, x: Tensor) -> Tensor:
        x = x if x.dtype == torch.bfloat16 else ops.cast_bf16(x)
        h = ops.silu(ops.linear(x, self[0].weight, self[0].bias))
        return ops.linear(h, self[2].weight, self[2].bias)


class UNetModel(nn.Module):
    """The full UNet with attention and timestep embedding (reference openaimodel.py:446-840)."""

    def __init__(
        self,
        in_channels: int,
        model_channels: int,
        out_channels: int,
        num_res_blocks: int,
        attention_resolutions: int | list[int] | tuple[int, ...],
        dropout: float = 0.0,
        channel_mult: Union[list, tuple] = (1, 2, 4, 8),
        conv_resample: bool = True,
        dims: int = 2,
        num_classes: Optional[int | str] = None,
        use_checkpoint: bool = False,
        num_heads: int = -1,
        num_head_channels: int = -1,
        num_heads_upsample: int = -1,
        use_scale_shift_norm: bool = False,
        resblock_updown: bool = False,
        transformer_depth: int | list[int] = 1,
        context_dim: Optional[int] = None,
        disable_self_attentions: Optional[list[bool]] = None,
        num_attention_blocks: Optional[list[int]] = None,
        disable_middle_self_attn: bool = False,
        disable_middle_transformer: bool = False,
        use_linear_in_transformer: bool = False,
        spatial_transformer_attn_type: str = "softmax",
        adm_in_channels: Optional[int] = None,
    ):
        super().__init__()
        if num_heads_upsample == -1:
            num_heads_upsample = num_heads
        if num_heads == -1 and num_head_channels == -1:
            raise ValueError("Either num_heads or num_head_channels has to be set")
        if resblock_updown:
            raise NotImplementedError("resblock_updown is not used by the reference configs")
        if isinstance(attention_resolutions, int):
            attention_resolutions = [attention_resolutions]

        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        if isinstance(transformer_depth, int):
            transformer_depth = len(channel_mult) * [transformer_depth]
        transformer_depth_middle = transformer_depth[-1]
        if isinstance(num_res_blocks, int):
            self.num_res_blocks = len(channel_mult) * [num_res_blocks]
        else:
            if len(num_res_blocks) != len(channel_mult):
                raise ValueError("num_res_blocks must be an int or have one entry per channel_mult level")
            self.num_res_blocks = list(num_res_blocks)
        if disable_self_attentions is not None and len(disable_self_attentions) != len(channel_mult):
            raise ValueError("disable_self_attentions must have one entry per channel_mult level")
        if num_attention_blocks is not None and len(num_attention_blocks) != len(self.num_res_blocks):
            raise ValueError("num_attention_blocks must have one entry per level")

        self.attention_resolutions = attention_resolutions
        self.dropout = dropout
        self.channel_mult = channel_mult
        self.conv_resample = conv_resample
        self.num_classes = num_classes
        self.use_checkpoint = use_checkpoint
        self.num_heads = num_heads
        self.num_head_channels = num_head_channels
        self.num_heads_upsample = num_heads_upsample

        time_embed_dim = model_channels * 4
        self.time_embed = _EmbedMLP(model_channels, time_embed_dim)

        if self.num_classes is not None:
            if self.num_classes == "sequential":
                if adm_in_channels is None:
                    raise ValueError("adm_in_channels is required when num_classes == 'sequential'")
                self.label_emb = nn.Sequential(_EmbedMLP(adm_in_channels, time_embed_dim))
            elif self.num_classes == "timestep":
                self.label_emb = nn.Sequential(Timestep(model_channels), _EmbedMLP(model_channels, time_embed_dim))
            else:
                raise NotImplementedError(f"num_classes={self.num_classes!r} is not used by the reference configs")

        def heads_for(ch: int) -> tuple[int, int]:
            if num_head_channels == -1:
                return num_heads, ch // num_heads
            return ch // num_head_channels, num_head_channels

        def transformer(ch: int, depth: int, disabled_sa: bool) -> SpatialTransformer:
            nh, dh = heads_for(ch)
            return SpatialTransformer(ch, nh, dh, depth=depth, context_dim=context_dim, disable_self_attn=disabled_sa,
                                      use_linear=use_linear_in_transformer, attn_type=spatial_transformer_attn_type,
                                      use_checkpoint=use_checkpoint)

        def resblock(cin: int, cout: int) -> ResBlock:
            return ResBlock(cin, time_embed_dim, dropout, out_channels=cout, dims=dims, use_checkpoint=use_checkpoint,
                            use_scale_shift_norm=use_scale_shift_norm)

        self.input_blocks = nn.ModuleList(
            [TimestepEmbedSequential(conv_nd(dims, in_channels, model_channels, 3, padding=1))])
        self._feature_size = model_channels
        input_block_chans = [model_channels]
        ch = model_channels
        ds = 1
        for level, mult in enumerate(channel_mult):
            for nr in range(self.num_res_blocks[level]):
                layers: list[nn.Module] = [resblock(ch, mult * model_channels)]
                ch = mult * model_channels
                if ds in attention_resolutions:
                    disabled_sa = (disable_self_attentions[level]
                                   if (context_dim is not None and disable_self_attentions is not None) else False)
                    if num_attention_blocks is None or nr < num_attention_blocks[level]:
                        layers.append(transformer(ch, transformer_depth[level], disabled_sa))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                self._feature_size += ch
                input_block_chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(
                    TimestepEmbedSequential(Downsample(ch, conv_resample, dims=dims, out_channels=ch)))
                input_block_chans.append(ch)
                ds *= 2
                self._feature_size += ch

        self.middle_block = TimestepEmbedSequential(
            resblock(ch, ch),
            transformer(ch, transformer_depth_middle, disable_middle_self_attn)
            if not disable_middle_transformer else nn.Identity(),
            resblock(ch, ch),
        )
        self._feature_size += ch

        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(self.num_res_blocks[level] + 1):
                ich = input_block_chans.pop()
                layers = [resblock(ch + ich, model_channels * mult)]
                ch = model_channels * mult
                if ds in attention_resolutions:
                    disabled_sa = disable_self_attentions[level] if disable_self_attentions is not None else False
                    if num_attention_blocks is None or i < num_attention_blocks[level]:
                        layers.append(transformer(ch, transformer_depth[level], disabled_sa))
                if level and i == self.num_res_blocks[level]:
                    layers.append(Upsample(ch, conv_resample, dims=dims, out_channels=ch))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
                self._feature_size += ch

        self.out = nn.Sequential(
            nn.GroupNorm(32, ch), nn.SiLU(), zero_module(conv_nd(dims, model_channels, out_channels, 3, padding=1)))

    def forward(self, x: Tensor, timesteps: Optional[Tensor] = None, context: Optional[Tensor] = None,
                y: Optional[Tensor] = None, **kwargs) -> Tensor:
        if (y is not None) != (self.num_classes is not None):
            raise ValueError(f"y must be None for non-class-conditional models, got {y=}")
        if not x.is_cuda:
            raise RuntimeError("neurosis_b200.UNetModel runs on CUDA (sm_100a) only")
        t_emb = timestep_embedding(timesteps, self.model_channels)
        emb = self.time_embed(t_emb)
        if self.num_classes is not None:
            assert y.shape[0] == x.shape[0]
            emb = ops.add_bf16(emb, self.label_emb(y))
        if context is not None and context.dtype != torch.bfloat16:
            context = ops.cast_bf16(context)

        hs = []
        h = x
        for module in self.input_blocks:
            h = module(h, emb, context)
            hs.append(h)
        h = self.middle_block(h, emb, context)
        for module in self.output_blocks:
            h = from_nhwc(ops.cat(as_nhwc(h), as_nhwc(hs.pop())))
            h = module(h, emb, context)

        gn, conv = self.out[0], self.out[2]
        hn = ops.group_norm(as_nhwc(h), gn.weight, gn.bias, gn.num_groups, gn.eps, silu=True)
        out = _conv3x3(conv, hn)  # (N, H, W, 64) with out_channels valid
        return ops.from_nhwc_f32(out, self.out_channels).to(x.dtype if x.dtype.is_floating_point else torch.float32)
