"""GPU parity of the SURVEY.md §8(f) "next" rows (through the C ABI): VAE decoder + VAE training step, fused optimizer
step, device-side conditioner vector path, EMA.  Oracle = CPU fp32 restatements under oracle/, pinned by golden
vectors generated from the reference itself (tests/golden/make_golden_next.py).

Tolerances as in test_gpu_modules.py: bf16 storage with fp32 accumulation against an fp32 oracle -> relative L2
<= 3e-2 for network outputs, median per-parameter gradient error <= 4e-2; fp32 elementwise kernels <= 1e-5."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from common import ROOT, TINY_VAE
from oracle.vae import vae_decode, vae_decoder_param_shapes, vae_param_shapes, vae_train_loss
from oracle.weights import synth_state_dict, synth_tensor

pytestmark = pytest.mark.gpu
G = np.load(str(ROOT / "tests/golden/reference_golden_next.npz"))
DEV = "cuda"
BF = torch.bfloat16


def rel(a, b) -> float:
    a = torch.as_tensor(a).detach().float().cpu()
    b = torch.as_tensor(b).detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


# ---------------------------------------------------------------- row 1: VAE decoder / training step
def test_diag_gaussian_fwd_bwd():
    from neurosis_b200 import ops
    m = rnd(3, 8, 12, 10)
    m[0, 4:, :2] = 25.0   # clamp at 20
    m[1, 4:, :2] = -40.0  # clamp at -30
    m.requires_grad_(True)
    eps = rnd(3, 4, 12, 10, seed=1)
    z, kl = ops.diag_gaussian(m, eps)
    gz, gk = rnd(3, 4, 12, 10, seed=2), rnd(3, seed=3)
    (z * gz).sum().add((kl * gk).sum()).backward()
    mr = m.detach().clone().requires_grad_(True)
    mean, logvar = torch.chunk(mr, 2, dim=1)
    logvar = torch.clamp(logvar, -30.0, 20.0)
    zr = mean + torch.exp(0.5 * logvar) * eps
    klr = 0.5 * torch.sum(mean ** 2 + torch.exp(logvar) - 1.0 - logvar, dim=[1, 2, 3])
    ((zr * gz).sum() + (klr * gk).sum()).backward()
    assert rel(z, zr) < 1e-5 and rel(kl, klr) < 1e-5
    assert rel(m.grad, mr.grad) < 1e-5
    zm, _ = ops.diag_gaussian(m.detach(), None)
    assert torch.equal(zm, m.detach()[:, :4])


def test_conv1x1_thin_fwd_bwd():
    from neurosis_b200 import ops
    n, h, w, ci, co = 2, 16, 12, 8, 8
    x = rnd(n, ci, h, w)
    wt = rnd(co, ci, 1, 1, seed=1, scale=0.3).requires_grad_(True)
    b = rnd(co, seed=2).requires_grad_(True)
    xn = ops.to_nhwc(x, 64)
    y = ops.from_nhwc_f32(ops.conv1x1_thin(xn, wt, b), co)
    g = rnd(n, co, h, w, seed=3)
    y.backward(g)
    wr, br = wt.detach().to(BF).float().requires_grad_(True), b.detach().clone().requires_grad_(True)
    yr = F.conv2d(x.to(BF).float(), wr, br)
    yr.backward(g)
    assert rel(y, yr) < 6e-3
    assert rel(wt.grad, wr.grad) < 1e-2 and rel(b.grad, br.grad) < 1e-2


def test_conv3x3_thin_input_weight_gradient():
    from neurosis_b200 import ops
    n, c, h, w, co = 2, 3, 24, 20, 64
    x = rnd(n, c, h, w)
    wt = rnd(co, c, 3, 3, seed=1, scale=(9 * c) ** -0.5).requires_grad_(True)
    b = rnd(co, seed=2).requires_grad_(True)
    y = ops.conv3x3_thin_input(x, wt, b)
    g = rnd(n, h, w, co, seed=3).to(BF)
    y.backward(g)
    wr, br = wt.detach().to(BF).float().requires_grad_(True), b.detach().clone().requires_grad_(True)
    yr = F.conv2d(x.to(BF).float(), wr, br, padding=1)
    yr.backward(g.float().permute(0, 3, 1, 2))
    assert rel(y.permute(0, 3, 1, 2), yr) < 6e-3
    assert rel(wt.grad, wr.grad) < 1e-2 and rel(b.grad, br.grad) < 1e-2


def _grad_errs(module, ref_sd):
    """relative L2 error per parameter.  The key bias of an attention block has a mathematically ZERO gradient (a
    constant added to every key shifts all scores of a query equally; softmax is shift invariant), so its reference
    gradient is fp32 rounding noise: it is checked against the query-bias gradient's scale instead."""
    errs = {}
    named = dict(module.named_parameters())
    for n, p in named.items():
        if n.endswith(".k.bias"):
            qn = n[: -len(".k.bias")] + ".q.bias"
            assert float(ref_sd[n].grad.norm()) < 1e-4 * float(ref_sd[qn].grad.norm()), n
            assert float(p.grad.norm()) < 5e-2 * float(named[qn].grad.norm()), (n, float(p.grad.norm()))
            continue
        errs[n] = rel(p.grad, ref_sd[n].grad)
    return errs


def test_vae_decoder_vs_golden_and_oracle():
    from neurosis_b200.modules.vae import Decoder
    shapes = vae_decoder_param_shapes(TINY_VAE, embed_dim=4, standalone=True)
    dec = Decoder(**TINY_VAE, embed_dim=4, standalone=True)
    dec.load_state_dict(synth_state_dict(shapes, seed=5))
    dec = dec.to(DEV)
    z = synth_tensor("vaedec.z", (2, 4, 16, 16))
    xr = dec(z.to(DEV))
    assert xr.shape == (2, 3, 32, 32) and xr.dtype == torch.float32
    assert rel(xr, G["vaedec.out"]) < 3e-2, "vs the reference's own Decoder"
    g = synth_tensor("vaedec.g", (2, 3, 32, 32), scale=0.1)
    (xr * g.to(DEV)).sum().backward()
    sd = {k: v.requires_grad_(True) for k, v in synth_state_dict(shapes, seed=5).items()}
    xo = vae_decode(sd, TINY_VAE, z)
    (xo * g).sum().backward()
    errs = _grad_errs(dec, sd)
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    assert len(errs) == len(shapes) - 1  # all but mid.attn_1.k.bias (zero gradient)
    assert np.median(list(errs.values())) < 4e-2, worst
    assert worst[0][1] < 1.5e-1, worst
    keep = np.array([not n.endswith(".k.bias") for n in sorted(shapes)])
    l2 = np.array([float(dict(dec.named_parameters())[n].grad.norm()) for n in sorted(shapes)])
    assert np.allclose(l2[keep], G["vaedec.grad_l2"][keep], rtol=6e-2, atol=1e-5), "gradient norms vs the reference's"


def test_vae_training_step_vs_golden_and_oracle():
    """AutoencoderKL.training_step (encoder -> quant_conv -> sampled posterior -> post_quant_conv -> decoder -> L2)
    with the posterior noise pinned: loss within 1e-2 of the reference's fp32 value, gradients of BOTH networks."""
    from neurosis_b200.modules.vae import AutoencoderKL, DiagonalGaussianRegularizer
    eshapes = vae_param_shapes(TINY_VAE, embed_dim=4, standalone=True)
    dshapes = vae_decoder_param_shapes(TINY_VAE, embed_dim=4, standalone=True)
    esd, dsd = synth_state_dict(eshapes, seed=2), synth_state_dict(dshapes, seed=5)
    ae = AutoencoderKL(4, TINY_VAE, regularizer=DiagonalGaussianRegularizer(sample=True))
    full = {}
    for k, v in esd.items():
        full[k if k.startswith("quant_conv.") else "encoder." + k] = v
    for k, v in dsd.items():
        full[k if k.startswith("post_quant_conv.") else "decoder." + k] = v
    ae.load_state_dict(full)
    ae = ae.to(DEV)
    img = synth_tensor("vae.img", (2, 3, 32, 32), uniform=True)
    eps = synth_tensor("vaetrain.eps", (2, 4, 16, 16))
    loss = ae.training_step({"image": img.to(DEV), "posterior_eps": eps.to(DEV)})
    assert loss.ndim == 0
    np.testing.assert_allclose(float(loss), G["vaetrain.loss"], rtol=1e-2)
    np.testing.assert_allclose(float(ae.last_log["kl_loss"]), G["vaetrain.kl_loss"], rtol=5e-2)  # sum over bf16 moments
    loss.backward()
    ro = {k: v.requires_grad_(True) for k, v in esd.items()}
    rd = {k: v.requires_grad_(True) for k, v in dsd.items()}
    lo, *_ = vae_train_loss(ro, rd, TINY_VAE, img, eps)
    lo.backward()
    ref = {}
    for k, v in ro.items():
        ref[k if k.startswith("quant_conv.") else "encoder." + k] = v
    for k, v in rd.items():
        ref[k if k.startswith("post_quant_conv.") else "decoder." + k] = v
    errs = _grad_errs(ae, ref)
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    assert len(errs) == len(eshapes) + len(dshapes) - 2  # all but the two mid.attn_1.k.bias (zero gradient)
    assert np.median(list(errs.values())) < 4e-2, worst
    assert worst[0][1] < 2e-1, worst
    # without sampling noise supplied the step draws its own and still trains
    ae.zero_grad()
    l2 = ae.training_step({"image": img.to(DEV)})
    l2.backward()
    assert torch.isfinite(l2) and all(torch.isfinite(p.grad).all() for p in ae.parameters())


# ---------------------------------------------------------------- row 2: fused Adafactor step
def _opt_params(case_kw):
    from test_oracle_golden_next import OPT_SHAPES
    return {k: torch.nn.Parameter(synth_tensor(f"opt.p.{k}", s, scale=0.05).to(DEV)) for k, s in OPT_SHAPES.items()}


@pytest.mark.parametrize("case", ["yaml", "ext"])
def test_adafactor_step_vs_reference_golden(case):
    """three steps of the multi-tensor Adafactor over parameters of every layout class (large matrix, ragged matrix,
    3x3 / 1x1 conv kernels, bias, stacked matrices) against the reference optimizer's parameters, moments and RMS."""
    from neurosis_b200.optim import Adafactor
    from test_oracle_golden_next import OPT_CASES, OPT_SHAPES, check_optimizer_state
    kw = OPT_CASES[case]
    params = _opt_params(kw)
    p0 = {k: p.detach().cpu().numpy().copy() for k, p in params.items()}
    opt = Adafactor(list(params.values()), **kw)
    for step in range(3):
        for k, p in params.items():
            p.grad = synth_tensor(f"opt.g.{k}.{step}", OPT_SHAPES[k], scale=0.02 * (step + 1)).to(DEV)
        v0 = {k: p._version for k, p in params.items()}
        opt.step()
        assert all(p._version > v0[k] for k, p in params.items()), "version counters must move (weight caches key on them)"
    torch.cuda.synchronize()
    for k, p in params.items():
        st = {sk: (v.detach().cpu().numpy() if torch.is_tensor(v) else v) for sk, v in opt.state[p].items()}
        check_optimizer_state(case, k, p.detach().cpu().numpy(), st, p0[k])
        assert opt.state[p]["step"] == 3


def test_adafactor_refreshes_bf16_mirrors_and_packed_weights():
    """the apply pass rewrites the bf16 copies the GEMMs consume; packed conv weights are re-derived on next use."""
    from neurosis_b200 import ops
    from neurosis_b200.optim import Adafactor
    w = torch.nn.Parameter(rnd(128, 64, seed=1, scale=0.1))
    cw = torch.nn.Parameter(rnd(64, 64, 3, 3, seed=2, scale=0.05))
    x = rnd(32, 64, seed=3).to(BF)
    xi = rnd(1, 8, 8, 64, seed=4).to(BF)
    ops.linear(x, w)          # registers the bf16 mirror of w
    ops.conv2d(xi, cw)        # caches the packed copies of cw
    mirror = ops.registered_mirror(w)
    assert mirror is not None
    opt = Adafactor([w, cw], lr=1e-2, relative_step=False, scale_parameter=False)
    w.grad, cw.grad = rnd(128, 64, seed=5), rnd(64, 64, 3, 3, seed=6)
    opt.step()
    torch.cuda.synchronize()
    assert torch.equal(mirror, w.detach().to(BF)), "mirror rewritten by the apply pass"
    assert ops.bf16_weight(w).data_ptr() == mirror.data_ptr(), "no second cast"
    y = ops.linear(x, w)
    assert rel(y, x.float() @ w.detach().to(BF).float().t()) < 1e-2
    yc = ops.conv2d(xi, cw)
    ref = F.conv2d(xi.float().permute(0, 3, 1, 2), cw.detach().to(BF).float(), padding=1)
    assert rel(yc.permute(0, 3, 1, 2), ref) < 1e-2, "packed conv weights follow the updated parameter"


def test_adafactor_scheduler_reports_lr():
    from neurosis_b200.optim import Adafactor, AdafactorScheduler
    p = torch.nn.Parameter(rnd(70, 90, scale=0.05))
    opt = Adafactor([p], scale_parameter=True, relative_step=True, warmup_init=True)
    sched = AdafactorScheduler(opt, initial_lr=4e-7)
    assert sched.get_lr() == [4e-7]
    p.grad = rnd(70, 90, seed=1, scale=0.01)
    opt.step()
    rms = float(p.detach().norm() / p.numel() ** 0.5)
    lr = sched.get_lr()[0]
    assert abs(lr - max(1e-3, rms) * 1e-6) < 1e-9 * 0.05 + 1e-12


# ---------------------------------------------------------------- row 4: EMA
def test_lit_ema_vs_reference_golden():
    from neurosis_b200.optim import LitEma
    from test_oracle_golden_next import EMA_SHAPES

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = torch.nn.Linear(40, 30)
            self.b = torch.nn.Conv2d(8, 8, 3)

    m = M().to(DEV)
    with torch.no_grad():
        for n, p in m.named_parameters():
            p.copy_(synth_tensor(f"ema.p.{n}", tuple(p.shape)))
    ema = LitEma(m, decay=0.9999).to(DEV)
    for it in range(12):
        with torch.no_grad():
            for n, p in m.named_parameters():
                p.add_(synth_tensor(f"ema.d.{n}.{it}", tuple(p.shape), scale=0.1).to(DEV))
        ema(m)
    sh = dict(ema.named_buffers())
    for n in EMA_SHAPES:
        np.testing.assert_allclose(sh[n.replace(".", "_")].cpu().numpy(), G[f"ema.{n}"], rtol=1e-6, atol=1e-7)
    assert int(ema.num_updates) == int(G["ema.num_updates"])
    # store / copy_to / restore round trip
    before = [p.detach().clone() for p in m.parameters()]
    ema.store(m.parameters())
    ema.copy_to(m)
    assert torch.equal(m.a.weight.detach(), sh["a_weight"])
    ema.restore(m.parameters())
    assert all(torch.equal(a, b.detach()) for a, b in zip(before, m.parameters()))


# ---------------------------------------------------------------- row 3: conditioner vector path
def test_conditioner_vector_path_on_device():
    from neurosis_b200.modules.conditioner import ConcatTimestepEmbedderND, GeneralConditioner, IdentityEncoder
    emb = ConcatTimestepEmbedderND(256, input_key="original_size_as_tuple")
    sizes = torch.from_numpy(G["cond.sizes"])
    f = emb(sizes.to(DEV))
    assert f.shape == (4, 512) and f.is_cuda
    # bf16 output; arguments up to 1216 rad through the fast sin/cos path: absolute error budget 1e-2
    assert (f.float().cpu() - torch.from_numpy(G["cond.fourier"])).abs().max() < 1e-2
    cond = GeneralConditioner([IdentityEncoder(input_key="ctx"), IdentityEncoder(input_key="pooled"),
                               ConcatTimestepEmbedderND(256, input_key="original_size_as_tuple"),
                               ConcatTimestepEmbedderND(256, input_key="crop_coords_top_left", ucg_rate=0.5),
                               ConcatTimestepEmbedderND(256, input_key="target_size_as_tuple")])
    B = 4
    batch = {"image": torch.zeros(B, 3, 8, 8, device=DEV), "ctx": rnd(B, 77, 64), "pooled": rnd(B, 1280, seed=1),
             "original_size_as_tuple": [(1024, 1024), (1152, 896), (832, 1216), (1024, 1024)],   # loader lists
             "crop_coords_top_left": [(0, 0), (16, 0), (0, 32), (8, 8)],
             "target_size_as_tuple": sizes.to(DEV)}
    out = cond(batch)
    assert out["crossattn"].shape == (B, 77, 64) and out["vector"].shape == (B, 1280 + 3 * 512)
    assert torch.equal(out["vector"][:, :1280], batch["pooled"])
    ref_o = emb(torch.tensor(batch["original_size_as_tuple"], dtype=torch.float32))
    assert (out["vector"][:, 1280:1792].cpu() - ref_o).abs().max() < 1e-2
    crop = out["vector"][:, 1792:2304]
    rows_zero = (crop.abs().sum(1) == 0)
    ref_c = emb(torch.tensor(batch["crop_coords_top_left"], dtype=torch.float32))
    assert all(bool(z) or float((c.cpu() - r).abs().max()) < 1e-2 for z, c, r in zip(rows_zero, crop, ref_c))
    z = cond(batch, force_zero_embeddings=["pooled"])
    assert float(z["vector"][:, :1280].abs().sum()) == 0.0
