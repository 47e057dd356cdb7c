"""CPU stand-ins (plain torch, fp32 math, bf16 storage) for the raw kernel wrappers of neurosis_b200.ops that the tune probes
and the attention autograd functions call.  Every stand-in first binds its arguments against the REAL wrapper's signature, so
a call that would not fit the real function fails here too.  Test infrastructure only."""
import inspect

import torch
import torch.nn.functional as F

from neurosis_b200 import ops

BF16, F32 = torch.bfloat16, torch.float32


def _bound(real, impl):
    sig = inspect.signature(real)

    def fn(*a, **k):
        b = sig.bind(*a, **k)
        b.apply_defaults()
        return impl(**b.arguments)

    return fn


def _gelu(x):
    return 0.5 * x * (1 + torch.erf(x / 2 ** 0.5))


def _dgelu(x):
    return 0.5 * (1 + torch.erf(x / 2 ** 0.5)) + x * torch.exp(-0.5 * x * x) / (2 * torch.pi) ** 0.5


def linear_fwd(x, w, bias, residual, out_f32):
    y = x.reshape(-1, x.shape[-1]).float() @ w.float().t()
    if bias is not None:
        y = y + bias.float()
    if residual is not None:
        y = y + residual.reshape(-1, w.shape[0]).float()
    return y.to(F32 if out_f32 else BF16).view(*x.shape[:-1], w.shape[0])


def linear_dgrad(dy, w, residual):
    dx = dy.reshape(-1, w.shape[0]).float() @ w.float()
    if residual is not None:
        dx = dx + residual.reshape(-1, w.shape[1]).float()
    return dx.to(BF16).view(*dy.shape[:-1], w.shape[1])


def linear_wgrad(dy, x, out):
    dw = dy.reshape(-1, dy.shape[-1]).float().t() @ x.reshape(-1, x.shape[-1]).float()
    if out is not None:
        out += dw
        return out
    return dw


def linear_dgrad_geglu(dy, w, h):
    N, D = w.shape
    d_out = dy.reshape(-1, N).float() @ w.float()
    hv, hg = h.reshape(-1, 2 * D)[:, :D].float(), h.reshape(-1, 2 * D)[:, D:].float()
    return torch.cat([d_out * _gelu(hg), d_out * hv * _dgelu(hg)], 1).to(BF16).view(*dy.shape[:-1], 2 * D)


def colsum(x, groups, out):
    s = x.reshape(groups, -1, x.shape[-1]).float().sum(1)
    if out is not None:
        out += s
        return out
    return s


def packed_conv_weight(p, need_dgrad):
    return p.detach().to(BF16), None  # (the stand-in convolutions take the OIHW weight itself)


def conv2d_fwd(x, wp, cout, ksize, bias, bias_img, residual):
    y = F.conv2d(x.permute(0, 3, 1, 2).float(), wp.float(), padding=ksize // 2)
    if bias is not None:
        y = y + bias.float()[None, :, None, None]
    if bias_img is not None:
        y = y + bias_img.float()[:, :, None, None]
    y = y.permute(0, 2, 3, 1)
    if residual is not None:
        y = y + residual[..., :cout].float()
    out = torch.zeros(*y.shape[:3], max(cout, 64), dtype=BF16)
    out[..., :cout] = y.to(BF16)
    return out


def conv2d_stride2_fwd(x, wp, cout, ksize, bias, pad_t, pad_l, ho, wo):
    xi = F.pad(x.permute(0, 3, 1, 2).float(), (pad_l, 1, pad_t, 1))
    y = F.conv2d(xi, wp.float(), stride=2)[:, :, :ho, :wo]
    if bias is not None:
        y = y + bias.float()[None, :, None, None]
    return y.permute(0, 2, 3, 1).to(BF16).contiguous()


def layernorm_fwd(x, gamma, beta, eps):
    xf = x.float()
    mean = xf.mean(-1)
    rstd = 1.0 / torch.sqrt(xf.var(-1, unbiased=False) + eps)
    y = ((xf - mean[..., None]) * rstd[..., None] * gamma + beta).to(BF16)
    return y.contiguous(), mean.reshape(-1), rstd.reshape(-1)


def layernorm_bwd(dy, x, gamma, mean, rstd, out, dres):
    c = x.shape[-1]
    xf, df = x.reshape(-1, c).float(), dy.reshape(-1, c).float()
    xh = (xf - mean[:, None]) * rstd[:, None]
    gd = df * gamma
    dx = rstd[:, None] * (gd - gd.mean(1, keepdim=True) - xh * (gd * xh).mean(1, keepdim=True))
    dx = dx.to(BF16)
    if dres is not None:
        dx = (dx.float() + dres.reshape(-1, c).float()).to(BF16)
    dg, db = (df * xh).sum(0), df.sum(0)
    if out is not None:
        out[0].add_(dg)
        out[1].add_(db)
        dg, db = out
    return dx.view(x.shape), dg, db


def groupnorm_fwd(x, gamma, beta, groups, eps, silu):
    n, h, w_, c = x.shape
    xf = x.float().reshape(n, h * w_, groups, c // groups)
    mean = xf.mean(dim=(1, 3))
    rstd = 1.0 / torch.sqrt(xf.var(dim=(1, 3), unbiased=False) + eps)
    y = ((xf - mean[:, None, :, None]) * rstd[:, None, :, None]).reshape(n, h, w_, c) * gamma + beta
    if silu:
        y = F.silu(y)
    return y.to(BF16), mean, rstd


def groupnorm_bwd(dy, x, gamma, beta, mean, rstd, groups, silu, out):
    with torch.enable_grad():
        xf = x.float().detach().requires_grad_(True)
        g, b = gamma.detach().clone().requires_grad_(True), beta.detach().clone().requires_grad_(True)
        y = F.group_norm(xf.permute(0, 3, 1, 2), groups, g, b, 1e-5)
        if silu:
            y = F.silu(y)
        y.permute(0, 2, 3, 1).backward(dy.float())
    if out is not None:
        out[0].add_(g.grad)
        out[1].add_(b.grad)
        return xf.grad.to(BF16), out[0], out[1]
    return xf.grad.to(BF16), g.grad, b.grad


def attention_fwd(q, k, v, scale):
    s = torch.einsum("bqhd,bkhd->bhqk", q.float(), k.float()) * scale
    lse = torch.logsumexp(s, -1)
    o = torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s, -1), v.float())
    return o.to(BF16).contiguous(), lse.contiguous()


def attention_bwd(do, q, k, v, o, lse, scale, out):
    with torch.enable_grad():  # (called from inside an autograd Function's backward, where grad mode is off)
        qf, kf, vf = (t.float().detach().requires_grad_(True) for t in (q, k, v))
        s = torch.einsum("bqhd,bkhd->bhqk", qf, kf) * scale
        torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s, -1), vf).backward(do.float())
    return qf.grad.to(BF16).contiguous(), kf.grad.to(BF16).contiguous(), vf.grad.to(BF16).contiguous()  # fresh tensors, like the materialised fallback


def install(monkeypatch):
    """replace the raw wrappers (and the device-pointer helpers around them) in neurosis_b200.ops"""
    for name in ("linear_fwd", "linear_dgrad", "linear_wgrad", "linear_dgrad_geglu", "colsum", "packed_conv_weight", "conv2d_fwd",
                 "conv2d_stride2_fwd", "layernorm_fwd", "layernorm_bwd", "groupnorm_fwd", "groupnorm_bwd", "attention_fwd", "attention_bwd"):
        monkeypatch.setattr(ops, name, _bound(getattr(ops, name), globals()[name]))
    monkeypatch.setattr(ops, "cast_bf16", lambda x: x.contiguous() if x.dtype == BF16 else x.to(BF16).contiguous())
    monkeypatch.setattr(ops, "bf16_weight", lambda p: p.detach().to(BF16))
    monkeypatch.setattr(ops, "bf16_weight_group", lambda ps: torch.cat([p.detach().to(BF16) for p in ps], 0))
    monkeypatch.setattr(ops, "WGRAD_OVERLAP", False)
    monkeypatch.setattr(ops, "GRAD_SINK", None)
