"""ORACLE (test infrastructure, not product code): CPU fp32 restatement of the reference UNet.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  It restates, as plain functions over a state dict, the algorithm of
/root/reference/src/neurosis/modules/diffusion/openaimodel.py (UNetModel.forward :803-840,
ResBlock._forward :315-342, Upsample :139-143, Downsample :183-197, TimestepEmbedSequential :71-93),
modules/attention.py (SpatialTransformer.forward :642-667, BasicTransformerBlock._forward :487-511,
TorchSDPCrossAttention.forward :369-417, GEGLU :50-57, FeedForward :60-74) and
modules/diffusion/util.py (timestep_embedding :152-177).

Pinned by tests/test_oracle_vs_reference.py (live comparison against the reference modules when
/root/reference is present) and by the golden vectors under tests/golden/ that were generated from the
reference itself (tests/golden/make_golden.py).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F
from torch import Tensor


def unet_plan(cfg: dict) -> dict:
    """Static structure implied by the UNetModel constructor arguments (openaimodel.py:490-801)."""
    mc = cfg["model_channels"]
    mult = list(cfg.get("channel_mult", (1, 2, 4, 8)))
    nrb = cfg["num_res_blocks"]
    nrb = [nrb] * len(mult) if isinstance(nrb, int) else list(nrb)
    ar = cfg["attention_resolutions"]
    ar = [ar] if isinstance(ar, int) else list(ar)
    depth = cfg.get("transformer_depth", 1)
    depth = [depth] * len(mult) if isinstance(depth, int) else list(depth)
    nh, nhc = cfg.get("num_heads", -1), cfg.get("num_head_channels", -1)
    use_linear = cfg.get("use_linear_in_transformer", False)

    def heads(ch):
        return (nh, ch // nh) if nhc == -1 else (ch // nhc, nhc)

    def st(ch, d):
        h, dh = heads(ch)
        return {"kind": "st", "ch": ch, "heads": h, "dim_head": dh, "depth": d, "linear": use_linear}

    inputs = [[{"kind": "conv_in"}]]
    chans = [mc]
    ch, ds = mc, 1
    for level, m in enumerate(mult):
        for _ in range(nrb[level]):
            layers = [{"kind": "res", "cin": ch, "cout": m * mc}]
            ch = m * mc
            if ds in ar:
                layers.append(st(ch, depth[level]))
            inputs.append(layers)
            chans.append(ch)
        if level != len(mult) - 1:
            inputs.append([{"kind": "down", "ch": ch}])
            chans.append(ch)
            ds *= 2
    middle = [{"kind": "res", "cin": ch, "cout": ch}]
    if not cfg.get("disable_middle_transformer", False):
        middle.append(st(ch, depth[-1]))
    middle.append({"kind": "res", "cin": ch, "cout": ch})
    outputs = []
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nrb[level] + 1):
            ich = chans.pop()
            layers = [{"kind": "res", "cin": ch + ich, "cout": mc * m}]
            ch = mc * m
            if ds in ar:
                layers.append(st(ch, depth[level]))
            if level and i == nrb[level]:
                layers.append({"kind": "up", "ch": ch})
                ds //= 2
            outputs.append(layers)
    return {"inputs": inputs, "middle": middle, "outputs": outputs, "mc": mc,
            "has_y": cfg.get("num_classes", None) is not None}


def timestep_embedding(t: Tensor, dim: int, max_period: int = 10000) -> Tensor:
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _conv(sd, p, x, stride=1, padding=1):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=padding)


def _gn(sd, p, x, eps):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps)


def res_block(sd, p: str, x: Tensor, emb: Tensor) -> Tensor:
    """GN(1e-5) SiLU conv ; + Linear(SiLU(emb)) ; GN SiLU conv ; + skip(x)."""
    h = _conv(sd, p + ".in_layers.2", F.silu(_gn(sd, p + ".in_layers.0", x, 1e-5)))
    h = h + _lin(sd, p + ".emb_layers.1", F.silu(emb))[:, :, None, None]
    h = _conv(sd, p + ".out_layers.3", F.silu(_gn(sd, p + ".out_layers.0", h, 1e-5)))
    if p + ".skip_connection.weight" in sd:
        k = sd[p + ".skip_connection.weight"].shape[-1]
        x = _conv(sd, p + ".skip_connection", x, padding=k // 2)
    return x + h


def attention(sd, p: str, x: Tensor, context: Optional[Tensor], heads: int) -> Tensor:
    """softmax(q k^T / sqrt(d)) v with bias-free q/k/v projections and a biased output projection."""
    ctx = x if context is None else context
    q, k, v = _lin(sd, p + ".to_q", x), _lin(sd, p + ".to_k", ctx), _lin(sd, p + ".to_v", ctx)
    b, n, c = q.shape
    d = c // heads
    q, k, v = (t.view(b, -1, heads, d).transpose(1, 2) for t in (q, k, v))
    w = torch.softmax(q @ k.transpose(-1, -2) * d ** -0.5, dim=-1)
    o = (w @ v).transpose(1, 2).reshape(b, n, c)
    return _lin(sd, p + ".to_out.0", o)


def transformer_block(sd, p: str, x: Tensor, context: Optional[Tensor], heads: int) -> Tensor:
    def ln(name, t):
        return F.layer_norm(t, t.shape[-1:], sd[f"{p}.{name}.weight"], sd[f"{p}.{name}.bias"], 1e-5)

    x = x + attention(sd, p + ".attn1", ln("norm1", x), None, heads)
    x = x + attention(sd, p + ".attn2", ln("norm2", x), context, heads)
    h = _lin(sd, p + ".ff.net.0.proj", ln("norm3", x))
    a, gate = h.chunk(2, dim=-1)
    return x + _lin(sd, p + ".ff.net.2", a * F.gelu(gate))


def spatial_transformer(sd, p: str, x: Tensor, context: Optional[Tensor], spec: dict) -> Tensor:
    b, c, h, w = x.shape
    t = _gn(sd, p + ".norm", x, 1e-6)
    if not spec["linear"]:
        t = _conv(sd, p + ".proj_in", t, padding=0)
    t = t.permute(0, 2, 3, 1).reshape(b, h * w, -1)
    if spec["linear"]:
        t = _lin(sd, p + ".proj_in", t)
    for i in range(spec["depth"]):
        t = transformer_block(sd, f"{p}.transformer_blocks.{i}", t, context, spec["heads"])
    if spec["linear"]:
        t = _lin(sd, p + ".proj_out", t)
    t = t.reshape(b, h, w, -1).permute(0, 3, 1, 2)
    if not spec["linear"]:
        t = _conv(sd, p + ".proj_out", t, padding=0)
    return t + x


def _run_layers(sd, prefix: str, layers: list, h: Tensor, emb: Tensor, context) -> Tensor:
    for j, spec in enumerate(layers):
        p = f"{prefix}.{j}"
        kind = spec["kind"]
        if kind == "conv_in":
            h = _conv(sd, p, h)
        elif kind == "res":
            h = res_block(sd, p, h, emb)
        elif kind == "st":
            h = spatial_transformer(sd, p, h, context, spec)
        elif kind == "down":
            h = _conv(sd, p + ".op", h, stride=2, padding=1)
        elif kind == "up":
            h = _conv(sd, p + ".conv", F.interpolate(h, scale_factor=2, mode="nearest"))
    return h


def unet_forward(sd: dict, cfg: dict, x: Tensor, timesteps: Tensor, context: Optional[Tensor] = None,
                 y: Optional[Tensor] = None) -> Tensor:
    plan = unet_plan(cfg)
    emb = _lin(sd, "time_embed.2", F.silu(_lin(sd, "time_embed.0", timestep_embedding(timesteps, plan["mc"]))))
    if plan["has_y"]:
        emb = emb + _lin(sd, "label_emb.0.2", F.silu(_lin(sd, "label_emb.0.0", y)))
    hs = []
    h = x
    for i, layers in enumerate(plan["inputs"]):
        h = _run_layers(sd, f"input_blocks.{i}", layers, h, emb, context)
        hs.append(h)
    h = _run_layers(sd, "middle_block", plan["middle"], h, emb, context)
    for i, layers in enumerate(plan["outputs"]):
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_layers(sd, f"output_blocks.{i}", layers, h, emb, context)
    return _conv(sd, "out.2", F.silu(_gn(sd, "out.0", h, 1e-5)))


def unet_param_shapes(cfg: dict) -> dict[str, tuple]:
    """state-dict key -> shape implied by the constructor arguments (used to synthesise weights)."""
    plan = unet_plan(cfg)
    mc, ted = plan["mc"], plan["mc"] * 4
    ctx = cfg.get("context_dim")
    shapes: dict[str, tuple] = {}

    def lin(p, i, o, bias=True):
        shapes[p + ".weight"] = (o, i)
        if bias:
            shapes[p + ".bias"] = (o,)

    def conv(p, i, o, k):
        shapes[p + ".weight"] = (o, i, k, k)
        shapes[p + ".bias"] = (o,)

    def norm(p, c):
        shapes[p + ".weight"] = (c,)
        shapes[p + ".bias"] = (c,)

    lin("time_embed.0", mc, ted)
    lin("time_embed.2", ted, ted)
    if plan["has_y"]:
        lin("label_emb.0.0", cfg["adm_in_channels"], ted)
        lin("label_emb.0.2", ted, ted)

    def layers(prefix, specs):
        for j, s in enumerate(specs):
            p = f"{prefix}.{j}"
            if s["kind"] == "conv_in":
                conv(p, cfg["in_channels"], mc, 3)
            elif s["kind"] == "res":
                norm(p + ".in_layers.0", s["cin"])
                conv(p + ".in_layers.2", s["cin"], s["cout"], 3)
                lin(p + ".emb_layers.1", ted, s["cout"])
                norm(p + ".out_layers.0", s["cout"])
                conv(p + ".out_layers.3", s["cout"], s["cout"], 3)
                if s["cin"] != s["cout"]:
                    conv(p + ".skip_connection", s["cin"], s["cout"], 1)
            elif s["kind"] == "st":
                c, inner = s["ch"], s["heads"] * s["dim_head"]
                norm(p + ".norm", c)
                if s["linear"]:
                    lin(p + ".proj_in", c, inner)
                    lin(p + ".proj_out", inner, c)
                else:
                    conv(p + ".proj_in", c, inner, 1)
                    conv(p + ".proj_out", inner, c, 1)
                for i in range(s["depth"]):
                    b = f"{p}.transformer_blocks.{i}"
                    for a, cd in (("attn1", inner), ("attn2", ctx or inner)):
                        lin(f"{b}.{a}.to_q", inner, inner, False)
                        lin(f"{b}.{a}.to_k", cd, inner, False)
                        lin(f"{b}.{a}.to_v", cd, inner, False)
                        lin(f"{b}.{a}.to_out.0", inner, inner)
                    lin(f"{b}.ff.net.0.proj", inner, inner * 8)
                    lin(f"{b}.ff.net.2", inner * 4, inner)
                    for n_ in ("norm1", "norm2", "norm3"):
                        norm(f"{b}.{n_}", inner)
            elif s["kind"] == "down":
                conv(p + ".op", s["ch"], s["ch"], 3)
            elif s["kind"] == "up":
                conv(p + ".conv", s["ch"], s["ch"], 3)

    for i, specs in enumerate(plan["inputs"]):
        layers(f"input_blocks.{i}", specs)
    layers("middle_block", plan["middle"])
    for i, specs in enumerate(plan["outputs"]):
        layers(f"output_blocks.{i}", specs)
    norm("out.0", mc)
    conv("out.2", mc, cfg["out_channels"], 3)
    return shapes
