"""KL-f8 VAE on the sm_100a kernels: the encoder (latent encode of the diffusion training step) and, for the VAE
training step, the `Decoder` (model.py:609-765), `Upsample` (:51-62), the sampling
posterior and `AutoencoderKL.forward/training_step` (models/autoencoder.py:203-293, 429-504).

Drop-in for /root/reference/src/neurosis/modules/diffusion/model.py: `Normalize` (layers.py:5-7),
`ResnetBlock` (:85-134), `AttnBlock` (:144-172, the "vanilla"/xformers semantics — NOT the buggy
`TorchSDPAttnBlock`), `Downsample` (:65-82, asymmetric (0,1,0,1) pad + 3x3 s2),
`Encoder` (:456-606) incl. the `standalone` quant_conv, and `DiagonalGaussianRegularizer` in mode
(sample=False) form (regularizers.py:23-41).  `AutoencoderKL.encode` mirrors
models/autoencoder.py:469-487.  Attribute names (=> state-dict keys such as
`down.0.block.0.norm1.weight`, `mid.attn_1.q.weight`) are unchanged.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
from torch import Tensor, nn

from .. import ops
from .util import as_nhwc, from_nhwc


def Normalize(in_channels: int, num_groups: int = 32) -> nn.GroupNorm:
    return nn.GroupNorm(num_groups=num_groups, num_channels=in_channels, eps=1e-6, affine=True)


def _gn(norm: nn.GroupNorm, x: Tensor, silu: bool) -> Tensor:
    return ops.group_norm(x, norm.weight, norm.bias, norm.num_groups, norm.eps, silu=silu)


def _conv1x1(conv: nn.Conv2d, x: Tensor, residual: Optional[Tensor] = None) -> Tensor:
    n, h, w, c = x.shape
    r = residual.reshape(n * h * w, -1) if residual is not None else None
    y = ops.linear(x.reshape(n * h * w, c), conv.weight.view(conv.weight.shape[0], -1), conv.bias, r)
    return y.view(n, h, w, -1)


class Downsample(nn.Module):
    def __init__(self, in_channels: int, with_conv: bool):
        super().__init__()
        if not with_conv:
            raise NotImplementedError("average-pool downsampling is not used by the reference configs")
        self.with_conv = with_conv
        self.padding = nn.ConstantPad2d((0, 1, 0, 1), 0)
        self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=2, padding=0)

    def forward(self, x: Tensor) -> Tensor:
        return from_nhwc(ops.conv2d_stride2(as_nhwc(x), self.conv.weight, self.conv.bias, asymmetric=True))


class Upsample(nn.Module):
    """nearest 2x + 3x3 conv (reference model.py:51-62)."""

    def __init__(self, in_channels: int, with_conv: bool):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1)

    def forward(self, x: Tensor) -> Tensor:
        y = ops.upsample2x(as_nhwc(x))
        if self.with_conv:
            y = ops.conv2d(y, self.conv.weight, self.conv.bias)
        return from_nhwc(y)


class ResnetBlock(nn.Module):
    def __init__(self, *, in_channels: int, out_channels: Optional[int] = None, conv_shortcut: bool = False,
                 dropout: float = 0.0, temb_channels: int = 512):
        super().__init__()
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels = in_channels, out_channels
        self.use_conv_shortcut = conv_shortcut
        if temb_channels > 0 or dropout > 0.0:
            raise NotImplementedError("the VAE path uses temb_channels=0 and dropout=0")
        self.norm1 = Normalize(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.norm2 = Normalize(out_channels)
        self.dropout = nn.Identity()
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        if in_channels != out_channels:
            if conv_shortcut:
                self.conv_shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
            else:
                self.nin_shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=1, padding=0)

    def forward(self, x: Tensor, temb: Optional[Tensor] = None) -> Tensor:
        xn = as_nhwc(x)
        h = ops.conv2d(_gn(self.norm1, xn, True), self.conv1.weight, self.conv1.bias)
        h = _gn(self.norm2, h, True)
        if self.in_channels != self.out_channels:
            if self.use_conv_shortcut:
                skip = ops.conv2d(xn, self.conv_shortcut.weight, self.conv_shortcut.bias)
            else:
                skip = _conv1x1(self.nin_shortcut, xn)
        else:
            skip = xn
        return from_nhwc(ops.conv2d(h, self.conv2.weight, self.conv2.bias, None, skip))


class AttnBlock(nn.Module):
    """single-head attention over all pixels, head_dim = channels (512): materialised tcgen05 path (score GEMM ->
    row softmax -> PV GEMM), a few images per call to bound the fp32 score matrix."""

    def __init__(self, in_channels: int):
        super().__init__()
        self.in_channels = in_channels
        self.norm = Normalize(in_channels)
        self.q = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.k = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.v = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.proj_out = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.max_images_per_call = 2  # bounds the fp32 score matrix (1 GiB per 1024^2 image)

    def forward(self, x: Tensor, **kwargs) -> Tensor:
        xn = as_nhwc(x)
        n, h, w, c = xn.shape
        t = _gn(self.norm, xn, False)
        q, k, v = (_conv1x1(m, t).view(n, h * w, 1, c) for m in (self.q, self.k, self.v))
        outs = []
        step = n if c in ops.FLASH_HEAD_DIMS else self.max_images_per_call
        for i in range(0, n, step):
            j = min(n, i + step)
            outs.append(ops.attention(q[i:j], k[i:j], v[i:j], c ** -0.5))
        o = outs[0] if len(outs) == 1 else torch.cat(outs, 0)
        return from_nhwc(_conv1x1(self.proj_out, o.view(n, h, w, c), residual=xn))


def make_attn(in_channels: int, attn_type: str = "vanilla") -> nn.Module:
    if attn_type in ("vanilla", "vanilla-xformers", "torch-sdp", "b200"):
        return AttnBlock(in_channels)
    if attn_type == "none":
        return nn.Identity()
    raise NotImplementedError(f"attention type {attn_type!r} is not supported")


class DiagonalGaussianRegularizer(nn.Module):
    """DiagonalGaussianRegularizer (regularizers.py:23-41): z = posterior.sample() or .mode(), log["kl_loss"] =
    sum(kl) / B.  Sampling draws eps with torch.randn on the moments' device (the reference draws on the CPU and
    copies, distributions.py:39-41 — a different RNG stream, same distribution); `eps` can be passed for parity."""

    def __init__(self, sample: bool = True):  # the reference's default (regularizers.py:24); the engine's latent
        super().__init__()                     # encode constructs it with sample=False (autoencoder.py:443)
        self.sample = sample

    def get_trainable_parameters(self):
        yield from ()

    def forward(self, z: Tensor, eps: Optional[Tensor] = None):
        if not self.sample:
            # mode = the mean half of the moments; the diffusion step discards the KL term the reference logs here
            # (SURVEY.md §8(a) row 3), so no kernel runs
            return torch.chunk(z, 2, dim=1)[0], {}
        if eps is None:
            b, c2 = z.shape[:2]
            eps = torch.randn((b, c2 // 2, *z.shape[2:]), dtype=torch.float32, device=z.device)
        out, kl = ops.diag_gaussian(z, eps)
        return out, {"kl_loss": kl.sum() / kl.shape[0]}


class Encoder(nn.Module):
    def __init__(self, *, ch: int, out_ch: int, ch_mult: Sequence[int] = (1, 2, 4, 8), num_res_blocks: int,
                 attn_resolutions: Sequence[int], dropout: float = 0.0, resamp_with_conv: bool = True,
                 in_channels: int, resolution: int, z_channels: int, double_z: bool = True,
                 use_linear_attn: bool = False, attn_type: str = "vanilla", embed_dim: int = 256,
                 standalone: bool = False, **kwargs):
        super().__init__()
        self.ch = ch
        self.temb_ch = 0
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        self.resolution = resolution
        self.in_channels = in_channels
        self.conv_in = nn.Conv2d(in_channels, ch, kernel_size=3, stride=1, padding=1)
        curr_res = resolution
        in_ch_mult = (1,) + tuple(ch_mult)
        self.in_ch_mult = in_ch_mult
        self.down = nn.ModuleList()
        block_in = ch
        for i_level in range(self.num_resolutions):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_in = ch * in_ch_mult[i_level]
            block_out = ch * ch_mult[i_level]
            for _ in range(num_res_blocks):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, temb_channels=0, dropout=dropout))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(make_attn(block_in, attn_type=attn_type))
            down = nn.Module()
            down.block = block
            down.attn = attn
            if i_level != self.num_resolutions - 1:
                down.downsample = Downsample(block_in, resamp_with_conv)
                curr_res = curr_res // 2
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=0, dropout=dropout)
        self.mid.attn_1 = make_attn(block_in, attn_type=attn_type)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=0, dropout=dropout)
        self.norm_out = Normalize(block_in)
        self.z_out = 2 * z_channels if double_z else z_channels
        self.conv_out = nn.Conv2d(block_in, self.z_out, kernel_size=3, stride=1, padding=1)
        self.regularizer = DiagonalGaussianRegularizer(sample=False)
        self.max_batch_size = None
        self.standalone = standalone
        if standalone:
            self.quant_conv = nn.Conv2d((1 + double_z) * z_channels, (1 + double_z) * embed_dim, 1)
        else:
            self.quant_conv = nn.Identity()

    def encode(self, x: Tensor) -> Tensor:
        """(N, 3, H, W) in [-1, 1] -> moments (N, H/8, W/8, 64-padded) NHWC bf16."""
        w_in = self.conv_in.weight
        if (x.dim() == 4 and x.shape[1] in (1, 3, 4) and x.dtype == torch.float32 and w_in.shape[0] >= 64
                and not (torch.is_grad_enabled() and x.requires_grad)):
            # RGB patches -> one K = 64 GEMM (the image takes no gradient; the weight gradient, needed by the VAE
            # training step only, is the transposed GEMM over the same patch matrix)
            if torch.is_grad_enabled() and w_in.requires_grad:
                h = ops.conv3x3_thin_input(x, w_in, self.conv_in.bias)
            else:
                h = ops.conv3x3_thin_input_fwd(x, w_in, self.conv_in.bias)
        else:
            h = ops.conv2d(as_nhwc(x, cpad=64), w_in, self.conv_in.bias)
        h = from_nhwc(h)
        for i_level in range(self.num_resolutions):
            for i_block in range(self.num_res_blocks):
                h = self.down[i_level].block[i_block](h, None)
                if len(self.down[i_level].attn) > 0:
                    h = self.down[i_level].attn[i_block](h)
            if i_level != self.num_resolutions - 1:
                h = self.down[i_level].downsample(h)
        h = self.mid.block_1(h, None)
        h = self.mid.attn_1(h)
        h = self.mid.block_2(h, None)
        hn = _gn(self.norm_out, as_nhwc(h), True)
        return ops.conv2d(hn, self.conv_out.weight, self.conv_out.bias)  # padded to 64 channels

    def _encode_quant(self, x: Tensor, quant_conv: Optional[nn.Module] = None) -> Tensor:
        z = self.encode(x)  # (N, h, w, 64) with z_out valid channels
        qc = quant_conv if quant_conv is not None else self.quant_conv
        if isinstance(qc, nn.Conv2d):
            n, h, w, cp = z.shape
            wq = torch.zeros((qc.weight.shape[0], cp), dtype=qc.weight.dtype, device=qc.weight.device)
            wq[:, : qc.weight.shape[1]] = qc.weight.detach().view(qc.weight.shape[0], -1)
            y = ops.linear_fwd(z.view(n * h * w, cp), ops.cast_bf16(wq), ops.f32_param(qc.bias), out_f32=True)
            return y.view(n, h, w, -1).permute(0, 3, 1, 2).contiguous()
        return ops.nhwc_to_nchw(z, self.z_out, out_f32=True)

    def moments(self, x: Tensor, quant_conv: Optional[nn.Module] = None) -> Tensor:
        """differentiable encode: (N, 3, H, W) -> (N, 2*embed, H/8, W/8) fp32 moments after quant_conv (the VAE
        training step; `_encode_quant` is the no-grad form used by the diffusion step)."""
        z = self.encode(x)
        qc = quant_conv if quant_conv is not None else self.quant_conv
        cz = self.z_out
        if isinstance(qc, nn.Conv2d):
            z = ops.conv1x1_thin(z, qc.weight, qc.bias)
            cz = qc.weight.shape[0]
        return ops.from_nhwc_f32(z, cz)

    def forward(self, x: Tensor, regularize: bool = False):
        if self.max_batch_size is None:
            z = self._encode_quant(x)
        else:
            bs = self.max_batch_size
            z = torch.cat([self._encode_quant(x[i: i + bs]) for i in range(0, x.shape[0], bs)], 0)
        if regularize:
            z, _ = self.regularizer(z)
        return z


class Decoder(nn.Module):
    """KL-f8 decoder (reference model.py:609-765): [post_quant_conv] -> conv_in -> mid (res, attn, res) -> per level
    (num_res_blocks + 1 ResnetBlocks [+ attn], nearest-2x + conv except at level 0) -> GroupNorm+SiLU -> conv_out.
    Input z (N, z_channels | embed_dim, h, w) fp32/bf16 NCHW; output (N, out_ch, 8h, 8w) fp32 NCHW."""

    def __init__(self, *, ch: int, out_ch: int, ch_mult: Sequence[int] = (1, 2, 4, 8), num_res_blocks: int,
                 attn_resolutions: Sequence[int], dropout: float = 0.0, resamp_with_conv: bool = True,
                 in_channels: int, resolution: int, z_channels: int, give_pre_end: bool = False,
                 tanh_out: bool = False, use_linear_attn: bool = False, attn_type: str = "vanilla",
                 embed_dim: int = 256, standalone: bool = False, **kwargs):
        super().__init__()
        if tanh_out or give_pre_end:
            raise NotImplementedError("tanh_out / give_pre_end are not used by the reference configs")
        self.ch = ch
        self.temb_ch = 0
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        self.resolution = resolution
        self.in_channels = in_channels
        self.out_ch = out_ch
        self.give_pre_end, self.tanh_out = give_pre_end, tanh_out
        block_in = ch * ch_mult[self.num_resolutions - 1]
        curr_res = resolution // 2 ** (self.num_resolutions - 1)
        self.z_shape = (1, z_channels, curr_res, curr_res)
        self.conv_in = nn.Conv2d(z_channels, block_in, kernel_size=3, stride=1, padding=1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=0, dropout=dropout)
        self.mid.attn_1 = make_attn(block_in, attn_type=attn_type)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=0, dropout=dropout)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_out = ch * ch_mult[i_level]
            for _ in range(num_res_blocks + 1):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, temb_channels=0, dropout=dropout))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(make_attn(block_in, attn_type=attn_type))
            up = nn.Module()
            up.block = block
            up.attn = attn
            if i_level != 0:
                up.upsample = Upsample(block_in, resamp_with_conv)
                curr_res = curr_res * 2
            self.up.insert(0, up)  # prepend: up[i] is resolution level i, as in the reference state dict
        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, out_ch, kernel_size=3, stride=1, padding=1)
        self.max_batch_size = None
        self.standalone = standalone
        self.post_quant_conv = nn.Conv2d(embed_dim, z_channels, 1) if standalone else nn.Identity()

    def get_last_layer(self, **kwargs) -> Tensor:
        return self.conv_out.weight

    def decode(self, z: Tensor, post_quant_conv: Optional[nn.Module] = None, **kwargs) -> Tensor:
        self.last_z_shape = z.shape
        h = ops.to_nhwc(z.float(), 64)  # thin latent, zero-padded to one 64-channel K slab
        pq = post_quant_conv if post_quant_conv is not None else self.post_quant_conv
        if isinstance(pq, nn.Conv2d):
            h = ops.conv1x1_thin(h, pq.weight, pq.bias)
        h = from_nhwc(ops.conv2d(h, self.conv_in.weight, self.conv_in.bias))
        h = self.mid.block_1(h, None)
        h = self.mid.attn_1(h)
        h = self.mid.block_2(h, None)
        for i_level in reversed(range(self.num_resolutions)):
            for i_block in range(self.num_res_blocks + 1):
                h = self.up[i_level].block[i_block](h, None)
                if len(self.up[i_level].attn) > 0:
                    h = self.up[i_level].attn[i_block](h)
            if i_level != 0:
                h = self.up[i_level].upsample(h)
        hn = _gn(self.norm_out, as_nhwc(h), True)
        y = ops.conv2d(hn, self.conv_out.weight, self.conv_out.bias)  # out_ch valid channels of 64
        return ops.from_nhwc_f32(y, self.out_ch)

    def forward(self, z: Tensor, cat_zero: bool = False, **kwargs):
        if self.max_batch_size is None:
            return self.decode(z, **kwargs)
        bs = self.max_batch_size
        dec = [self.decode(z[i: i + bs], **kwargs) for i in range(0, z.shape[0], bs)]
        return torch.cat(dec, 0) if cat_zero else dec  # the reference returns the list unless cat_zero (model.py:752-765)


class AutoencoderKL(nn.Module):
    """The reference AutoencoderKL / AutoencodingEngineLegacy (models/autoencoder.py:429-504) without Lightning:
    encoder -> quant_conv -> DiagonalGaussian regularizer -> post_quant_conv -> decoder, and the non-adversarial
    branch of `inner_training_step` (:203-246; optimizer_idx 0 with a simple loss).  State-dict keys: encoder.*,
    decoder.*, quant_conv.*, post_quant_conv.* as in the reference.  `encode` under no_grad with a mode regularizer is
    the latent encode of the diffusion step; `forward`/`training_step` are the VAE training step (config 5)."""

    def __init__(self, embed_dim: int, ddconfig: dict, loss: Optional[nn.Module] = None,
                 regularizer: Optional[nn.Module] = None, with_decoder: bool = True, input_key: str = "image",
                 **kwargs):
        super().__init__()
        cfg = dict(ddconfig)
        cfg.pop("standalone", None)
        cfg.pop("embed_dim", None)
        self.encoder = Encoder(**cfg, embed_dim=embed_dim, standalone=False)
        z_ch = cfg["z_channels"]
        dz = 1 + cfg.get("double_z", True)
        self.quant_conv = nn.Conv2d(dz * z_ch, dz * embed_dim, 1)
        if with_decoder:
            self.decoder = Decoder(**cfg, embed_dim=embed_dim, standalone=False)
            self.post_quant_conv = nn.Conv2d(embed_dim, z_ch, 1)
        else:
            self.decoder = None
        self.regularization = regularizer if regularizer is not None else DiagonalGaussianRegularizer(sample=False)
        self.loss = loss if loss is not None else nn.Identity()
        self.embed_dim = embed_dim
        self.input_key = input_key
        self.max_batch_size = kwargs.pop("max_batch_size", None)
        self.encoder.max_batch_size = self.max_batch_size
        if self.decoder is not None:
            self.decoder.max_batch_size = self.max_batch_size

    def get_input(self, batch: dict) -> Tensor:
        return batch[self.input_key]

    def get_last_layer(self) -> Tensor:
        return self.decoder.get_last_layer()

    def get_autoencoder_params(self, decoder_only: bool = False) -> list:
        params = list(self.decoder.parameters())
        if not decoder_only:
            params += list(self.encoder.parameters())
        return params

    def encode(self, x: Tensor, return_reg_log: bool = False, eps: Optional[Tensor] = None):
        if not torch.is_grad_enabled() and not getattr(self.regularization, "sample", False):
            z = self.encoder._encode_quant(x, self.quant_conv)  # same arithmetic as Encoder(standalone=True)
            z, reg_log = self.regularization(z)
            return (z, reg_log) if return_reg_log else z
        bs = self.max_batch_size or x.shape[0]
        parts = [self.encoder.moments(x[i: i + bs], self.quant_conv) for i in range(0, x.shape[0], bs)]
        m = parts[0] if len(parts) == 1 else torch.cat(parts, 0)
        z, reg_log = self.regularization(m, eps) if eps is not None else self.regularization(m)
        return (z, reg_log) if return_reg_log else z

    def decode(self, z: Tensor, **decoder_kwargs) -> Tensor:
        bs = self.max_batch_size or z.shape[0]
        parts = [self.decoder.decode(z[i: i + bs], self.post_quant_conv, **decoder_kwargs)
                 for i in range(0, z.shape[0], bs)]
        return parts[0] if len(parts) == 1 else torch.cat(parts, 0)

    def forward(self, x: Tensor, eps: Optional[Tensor] = None, **additional_decode_kwargs):
        z, reg_log = self.encode(x, return_reg_log=True, eps=eps)
        xrec = self.decode(z, **additional_decode_kwargs)
        return z, xrec, reg_log

    def inner_training_step(self, batch: dict, batch_idx: int = 0, optimizer_idx: int = 0) -> Tensor:
        """autoencoder branch (optimizer_idx 0) with a simple reconstruction loss `loss(x, xrec)`; the default is the
        plain L2 of SURVEY.md §8(d) config 5 through the reduction kernel.  The LPIPS/discriminator losses of the
        reference (modules/autoencoding/losses) are outside the hot-path scope."""
        if optimizer_idx != 0:
            raise ValueError(f"Unknown optimizer ID {optimizer_idx}")
        x = self.get_input(batch)
        z, xrec, reg_log = self(x, eps=batch.get("posterior_eps"))
        if isinstance(self.loss, nn.Identity):
            aeloss = ops.mse_loss(xrec, x)
        else:
            aeloss = self.loss(x, xrec)
            if isinstance(aeloss, tuple):
                aeloss = aeloss[0]
        self.last_log = {"train/loss/rec": aeloss.detach(), **reg_log}
        return aeloss

    def training_step(self, batch: dict, batch_idx: int = 0) -> Tensor:
        """returns the loss; the caller owns backward / optimizer (the reference uses Lightning manual optimisation,
        autoencoder.py:276-289)."""
        return self.inner_training_step(batch, batch_idx, 0)
