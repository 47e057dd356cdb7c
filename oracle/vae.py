"""ORACLE (test infrastructure): CPU fp32 restatement of the reference KL-f8 VAE encoder.

Follows /root/reference/src/neurosis/modules/diffusion/model.py: Encoder.encode/forward (:558-606),
ResnetBlock.forward (:113-134), AttnBlock ("vanilla", :144-172), Downsample (:65-82, pad (0,1,0,1) then
conv k3 s2 p0), Normalize = GroupNorm(32, C, eps=1e-6) (modules/layers.py:5-7) and
DiagonalGaussianRegularizer(sample=False) = first half of the channels (regularizers.py:31-41,
distributions.py:30-35,71-72).

For the VAE training step: Decoder.decode (:719-747), Upsample (:51-62), the sampling
posterior (distributions.py:29-51) and AutoencodingEngineLegacy.encode/decode + inner_training_step with a plain L2
reconstruction loss (models/autoencoder.py:203-246, 469-504).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import Tensor


def _gn(sd, p, x):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], 1e-6)


def _conv(sd, p, x, stride=1, padding=1):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)


def resnet_block(sd, p, x):
    h = _conv(sd, p + ".conv1", F.silu(_gn(sd, p + ".norm1", x)))
    h = _conv(sd, p + ".conv2", F.silu(_gn(sd, p + ".norm2", h)))
    if p + ".nin_shortcut.weight" in sd:
        x = _conv(sd, p + ".nin_shortcut", x, padding=0)
    return x + h


def attn_block(sd, p, x):
    b, c, h, w = x.shape
    t = _gn(sd, p + ".norm", x)
    q, k, v = (_conv(sd, f"{p}.{n}", t, padding=0).permute(0, 2, 3, 1).reshape(b, h * w, c) for n in "qkv")
    a = torch.softmax(q @ k.transpose(1, 2) * c ** -0.5, dim=-1) @ v
    a = a.reshape(b, h, w, c).permute(0, 3, 1, 2)
    return x + _conv(sd, p + ".proj_out", a, padding=0)


def vae_encode(sd: dict, cfg: dict, x: Tensor, regularize: bool = True) -> Tensor:
    """x (B,3,H,W) -> moments (B, 2*z, H/8, W/8) -> quant_conv (if present) -> mean half."""
    nres = len(cfg["ch_mult"])
    h = _conv(sd, "conv_in", x)
    for lvl in range(nres):
        for blk in range(cfg["num_res_blocks"]):
            h = resnet_block(sd, f"down.{lvl}.block.{blk}", h)
            if f"down.{lvl}.attn.{blk}.norm.weight" in sd:
                h = attn_block(sd, f"down.{lvl}.attn.{blk}", h)
        if lvl != nres - 1:
            h = _conv(sd, f"down.{lvl}.downsample.conv", F.pad(h, (0, 1, 0, 1)), stride=2, padding=0)
    h = resnet_block(sd, "mid.block_1", h)
    h = attn_block(sd, "mid.attn_1", h)
    h = resnet_block(sd, "mid.block_2", h)
    h = _conv(sd, "conv_out", F.silu(_gn(sd, "norm_out", h)))
    if "quant_conv.weight" in sd:
        h = _conv(sd, "quant_conv", h, padding=0)
    if regularize:
        h = torch.chunk(h, 2, dim=1)[0]
    return h


def vae_moments(sd: dict, cfg: dict, x: Tensor) -> Tensor:
    """Encoder.encode -> quant_conv: the (B, 2*embed, h, w) moments (differentiable)."""
    return vae_encode(sd, cfg, x, regularize=False)


def diag_gaussian(moments: Tensor, eps: Tensor | None):
    """DiagonalGaussianDistribution (distributions.py:29-51): returns (z, kl[B]); eps None -> mode."""
    mean, logvar = torch.chunk(moments, 2, dim=1)
    logvar = torch.clamp(logvar, -30.0, 20.0)
    std, var = torch.exp(0.5 * logvar), torch.exp(logvar)
    z = mean if eps is None else mean + std * eps
    kl = 0.5 * torch.sum(torch.pow(mean, 2) + var - 1.0 - logvar, dim=[1, 2, 3])
    return z, kl


def vae_decode(sd: dict, cfg: dict, z: Tensor) -> Tensor:
    """[post_quant_conv ->] Decoder.decode (model.py:719-747): z (B, embed|z_ch, h, w) -> (B, out_ch, 8h, 8w)."""
    nres = len(cfg["ch_mult"])
    h = z
    if "post_quant_conv.weight" in sd:
        h = _conv(sd, "post_quant_conv", h, padding=0)
    h = _conv(sd, "conv_in", h)
    h = resnet_block(sd, "mid.block_1", h)
    h = attn_block(sd, "mid.attn_1", h)
    h = resnet_block(sd, "mid.block_2", h)
    for lvl in reversed(range(nres)):
        for blk in range(cfg["num_res_blocks"] + 1):
            h = resnet_block(sd, f"up.{lvl}.block.{blk}", h)
            if f"up.{lvl}.attn.{blk}.norm.weight" in sd:
                h = attn_block(sd, f"up.{lvl}.attn.{blk}", h)
        if lvl != 0:
            h = _conv(sd, f"up.{lvl}.upsample.conv", F.interpolate(h, scale_factor=2.0, mode="nearest"))
    return _conv(sd, "conv_out", F.silu(_gn(sd, "norm_out", h)))


def vae_train_loss(enc_sd: dict, dec_sd: dict, cfg: dict, x: Tensor, eps: Tensor | None):
    """AutoencodingEngine.forward + the simple-loss branch of inner_training_step (autoencoder.py:198-231) with
    loss = F.mse_loss(xrec, x): returns (loss, z, xrec, kl_loss)."""
    z, kl = diag_gaussian(vae_moments(enc_sd, cfg, x), eps)
    xrec = vae_decode(dec_sd, cfg, z)
    return F.mse_loss(xrec, x), z, xrec, kl.sum() / kl.shape[0]


def vae_decoder_param_shapes(cfg: dict, embed_dim: int = 4, standalone: bool = True) -> dict[str, tuple]:
    ch, mult, nrb = cfg["ch"], list(cfg["ch_mult"]), cfg["num_res_blocks"]
    shapes: dict[str, tuple] = {}

    def conv(p, i, o, k):
        shapes[p + ".weight"] = (o, i, k, k)
        shapes[p + ".bias"] = (o,)

    def norm(p, c):
        shapes[p + ".weight"] = (c,)
        shapes[p + ".bias"] = (c,)

    def res(p, i, o):
        norm(p + ".norm1", i)
        conv(p + ".conv1", i, o, 3)
        norm(p + ".norm2", o)
        conv(p + ".conv2", o, o, 3)
        if i != o:
            conv(p + ".nin_shortcut", i, o, 1)

    def attn(p, c):
        norm(p + ".norm", c)
        for n in ("q", "k", "v", "proj_out"):
            conv(f"{p}.{n}", c, c, 1)

    bi = ch * mult[-1]
    res_now = cfg["resolution"] // 2 ** (len(mult) - 1)
    conv("conv_in", cfg["z_channels"], bi, 3)
    res("mid.block_1", bi, bi)
    attn("mid.attn_1", bi)
    res("mid.block_2", bi, bi)
    for lvl in reversed(range(len(mult))):
        bo = ch * mult[lvl]
        for blk in range(nrb + 1):
            res(f"up.{lvl}.block.{blk}", bi, bo)
            bi = bo
            if res_now in cfg["attn_resolutions"]:
                attn(f"up.{lvl}.attn.{blk}", bi)
        if lvl != 0:
            conv(f"up.{lvl}.upsample.conv", bi, bi, 3)
            res_now *= 2
    norm("norm_out", bi)
    conv("conv_out", bi, cfg["out_ch"], 3)
    if standalone:
        conv("post_quant_conv", embed_dim, cfg["z_channels"], 1)
    return shapes


def vae_param_shapes(cfg: dict, embed_dim: int = 4, standalone: bool = True) -> dict[str, tuple]:
    ch, mult, nrb = cfg["ch"], list(cfg["ch_mult"]), cfg["num_res_blocks"]
    shapes: dict[str, tuple] = {}

    def conv(p, i, o, k):
        shapes[p + ".weight"] = (o, i, k, k)
        shapes[p + ".bias"] = (o,)

    def norm(p, c):
        shapes[p + ".weight"] = (c,)
        shapes[p + ".bias"] = (c,)

    def res(p, i, o):
        norm(p + ".norm1", i)
        conv(p + ".conv1", i, o, 3)
        norm(p + ".norm2", o)
        conv(p + ".conv2", o, o, 3)
        if i != o:
            conv(p + ".nin_shortcut", i, o, 1)

    def attn(p, c):
        norm(p + ".norm", c)
        for n in ("q", "k", "v", "proj_out"):
            conv(f"{p}.{n}", c, c, 1)

    conv("conv_in", cfg["in_channels"], ch, 3)
    in_mult = [1] + mult
    res_now = cfg["resolution"]
    bi = ch
    for lvl in range(len(mult)):
        bi, bo = ch * in_mult[lvl], ch * mult[lvl]
        for blk in range(nrb):
            res(f"down.{lvl}.block.{blk}", bi, bo)
            bi = bo
            if res_now in cfg["attn_resolutions"]:
                attn(f"down.{lvl}.attn.{blk}", bi)
        if lvl != len(mult) - 1:
            conv(f"down.{lvl}.downsample.conv", bi, bi, 3)
            res_now //= 2
    res("mid.block_1", bi, bi)
    attn("mid.attn_1", bi)
    res("mid.block_2", bi, bi)
    norm("norm_out", bi)
    zc = cfg["z_channels"] * (2 if cfg.get("double_z", True) else 1)
    conv("conv_out", bi, zc, 3)
    if standalone:
        conv("quant_conv", zc, embed_dim * (2 if cfg.get("double_z", True) else 1), 1)
    return shapes
