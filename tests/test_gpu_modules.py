"""Module-level and whole-step parity on the GPU: the drop-in CUDA modules against (a) the CPU oracle on the same
synthetic weights/inputs and (b) the golden outputs produced by the reference itself.

Tolerances: the CUDA path computes in bf16 with fp32 accumulation (as the reference does under its
`precision: bf16-mixed` autocast), the oracle/golden values are fp32 — so module outputs and gradients are
compared in relative L2 norm (<= 3e-2) and the step loss within 1e-2 relative (north-star tolerance)."""
import numpy as np
import pytest
import torch

from common import FULL_SD15, FULL_SDXL, ROOT, TINY_SD15, TINY_SDXL, TINY_VAE, fast_state_dict
from oracle import objective as O
from oracle.unet import unet_forward, unet_param_shapes
from oracle.vae import vae_param_shapes
from oracle.weights import synth_state_dict, synth_tensor

pytestmark = pytest.mark.gpu
G = np.load(str(ROOT / "tests/golden/reference_golden.npz"))
DEV = "cuda"


def rel(a, b) -> float:
    a = torch.as_tensor(a).float().cpu()
    b = torch.as_tensor(b).float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def build_unet(cfg):
    from neurosis_b200.modules import UNetModel
    m = UNetModel(**cfg)
    m.load_state_dict(synth_state_dict(unet_param_shapes(cfg), seed=1))
    return m.to(DEV)


@pytest.mark.parametrize("tag,cfg", [("sdxl", TINY_SDXL), ("sd15", TINY_SD15)])
def test_unet_forward_backward_vs_oracle_and_golden(tag, cfg):
    m = build_unet(cfg)
    x = synth_tensor(f"{tag}.x", (2, 4, 16, 16))
    ctx = synth_tensor(f"{tag}.ctx", (2, 77, cfg["context_dim"]))
    y = synth_tensor(f"{tag}.y", (2, cfg["adm_in_channels"])) if cfg.get("num_classes") else None
    ts = torch.tensor([17, 803])
    out = m(x.to(DEV), ts.to(DEV), ctx.to(DEV), y.to(DEV) if y is not None else None)
    assert out.shape == (2, 4, 16, 16)
    assert rel(out, G[f"{tag}.out"]) < 3e-2, "vs the reference's own output"
    gout = synth_tensor(f"{tag}.gout", (2, 4, 16, 16), scale=0.1)
    (out * gout.to(DEV)).sum().backward()
    # oracle on CPU, same weights / inputs
    sd = {k: v.requires_grad_(True) for k, v in synth_state_dict(unet_param_shapes(cfg), seed=1).items()}
    o_ref = unet_forward(sd, cfg, x, ts, ctx, y)
    (o_ref * gout).sum().backward()
    assert rel(out, o_ref) < 3e-2
    errs = {}
    for name, p in m.named_parameters():
        assert p.grad is not None, name
        errs[name] = rel(p.grad, sd[name].grad)
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    assert np.median(list(errs.values())) < 4e-2, worst  # bf16 rounding accumulated through ~40 layers of backward
    assert worst[0][1] < 1e-1, worst
    names = sorted(sd)
    l2 = np.array([float(dict(m.named_parameters())[n].grad.norm()) for n in names])
    assert np.allclose(l2, G[f"{tag}.grad_l2"], rtol=5e-2, atol=1e-4), "gradient norms vs the reference's"


@pytest.mark.parametrize("h,w", [(24, 16), (12, 20)])
def test_unet_aspect_bucket_shapes_vs_oracle(h, w):
    """non-square latents (aspect buckets such as 144x112 / 104x152): pixel tiles that do not
    divide the image, token counts that are not multiples of the attention tile (h*w = 384 / 240, 96 / 60 deeper)."""
    cfg = TINY_SDXL
    m = build_unet(cfg)
    x = synth_tensor("bucket.x", (2, 4, h, w))
    ctx = synth_tensor("bucket.ctx", (2, 77, cfg["context_dim"]))
    y = synth_tensor("bucket.y", (2, cfg["adm_in_channels"]))
    ts = torch.tensor([3, 977])
    out = m(x.to(DEV), ts.to(DEV), ctx.to(DEV), y.to(DEV))
    gout = synth_tensor("bucket.gout", (2, 4, h, w), scale=0.1)
    (out * gout.to(DEV)).sum().backward()
    sd = {k: v.requires_grad_(True) for k, v in synth_state_dict(unet_param_shapes(cfg), seed=1).items()}
    o_ref = unet_forward(sd, cfg, x, ts, ctx, y)
    (o_ref * gout).sum().backward()
    assert out.shape == o_ref.shape == (2, 4, h, w)
    assert rel(out, o_ref) < 3e-2
    errs = [rel(p.grad, sd[n].grad) for n, p in m.named_parameters()]
    assert np.median(errs) < 4e-2 and max(errs) < 1e-1


@pytest.mark.parametrize("tag,cfg,hw", [("sd15", FULL_SD15, 32), ("sdxl", FULL_SDXL, 32)])
def test_full_size_unet_vs_oracle(tag, cfg, hw):
    """the FULL-SIZE UNets of the example YAMLs (859.5 M / 2567.5 M parameters; SD1.5 exercises head dims 40/80/160 =
    the materialised attention path, SDXL the fused q|k|v + flash path at depth 10) at a small latent, batch 1, against
    the CPU oracle on the same weights: output, and the gradient of every parameter."""
    from neurosis_b200.modules import UNetModel
    shapes = unet_param_shapes(cfg)
    sd = fast_state_dict(shapes, seed=3)
    m = UNetModel(**cfg)
    m.load_state_dict(sd)
    m = m.to(DEV)
    x = synth_tensor(f"full.{tag}.x", (1, 4, hw, hw))
    ctx = synth_tensor(f"full.{tag}.ctx", (1, 77, cfg["context_dim"]))
    y = synth_tensor(f"full.{tag}.y", (1, cfg["adm_in_channels"])) if cfg.get("num_classes") else None
    ts = torch.tensor([481])
    gout = synth_tensor(f"full.{tag}.g", (1, 4, hw, hw), scale=0.1)
    out = m(x.to(DEV), ts.to(DEV), ctx.to(DEV), y.to(DEV) if y is not None else None)
    (out * gout.to(DEV)).sum().backward()
    ref_sd = {k: v.requires_grad_(True) for k, v in sd.items()}
    o_ref = unet_forward(ref_sd, cfg, x, ts, ctx, y)
    (o_ref * gout).sum().backward()
    errs = {n: rel(p.grad, ref_sd[n].grad) for n, p in m.named_parameters()}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    stats = dict(out=rel(out, o_ref), median=float(np.median(list(errs.values()))),
                 p90=float(np.percentile(list(errs.values()), 90)), worst=worst)
    print(tag, stats)
    assert len(errs) == len(shapes)
    # bf16 storage / fp32 accumulation against an fp32 oracle through ~70 (SDXL) transformer blocks: the error budget is
    # that of the reference's own bf16-mixed autocast, a few percent in relative L2
    assert stats["out"] < 5e-2, stats
    assert stats["median"] < 6e-2, stats
    assert worst[0][1] < 3e-1, stats


def test_full_size_sdxl_at_benchmark_shape_vs_reference_golden():
    """The SDXL UNet at the shape bench.py times — latent 128x128: 16 384-pixel 320-channel convolutions, 4 096-token /
    10-head and 1 024-token / 20-head self attention, 77-token cross attention — against the REFERENCE's own UNetModel
    run on the host in fp32 on the same weights (tests/golden/make_golden_sdxl128.py ->
    reference_golden_sdxl128.npz): output, the gradient norm of every one of the 1 680 parameters, complete gradients of
    ten parameters along the depth, and the forward at the 144x112 aspect-bucket latent (4 032 / 1 008 tokens)."""
    from neurosis_b200.modules import UNetModel
    G128 = np.load(str(ROOT / "tests/golden/reference_golden_sdxl128.npz"))
    cfg = FULL_SDXL
    shapes = unet_param_shapes(cfg)
    names = sorted(shapes)
    m = UNetModel(**cfg)
    m.load_state_dict(fast_state_dict(shapes, seed=3))
    m = m.to(DEV)
    ctx = synth_tensor("full128.ctx", (1, 77, cfg["context_dim"])).to(DEV)
    y = synth_tensor("full128.y", (1, cfg["adm_in_channels"])).to(DEV)
    ts = torch.tensor([481], device=DEV)
    x = synth_tensor("full128.x", (1, 4, 128, 128)).to(DEV)
    g = synth_tensor("full128.g", (1, 4, 128, 128), scale=0.1).to(DEV)
    out = m(x, ts, ctx, y)
    (out * g).sum().backward()
    params = dict(m.named_parameters())
    e_out = rel(out, G128["full128.out"])
    l2 = np.array([float(params[n].grad.norm()) for n in names])
    ref_l2 = G128["full128.grad_l2"]
    rel_l2 = np.abs(l2 - ref_l2) / np.maximum(ref_l2, 1e-3 * np.median(ref_l2))
    full = {}
    for key in G128.files:
        if not key.startswith("full128.grad.") or key in ("full128.grad_l2", "full128.grad_sum"):
            continue
        n = key[len("full128.grad."):]
        gr = params[n].grad
        ref = G128[key]
        if tuple(ref.shape) != tuple(gr.shape):
            gr = gr.reshape(gr.shape[0], -1)[: ref.shape[0]]
        full[n] = rel(gr, ref)
    stats = dict(out=e_out, l2_median=float(np.median(rel_l2)), l2_p99=float(np.percentile(rel_l2, 99)),
                 l2_max=float(rel_l2.max()), full=full)
    print("sdxl128", stats)
    # bf16 storage with fp32 accumulation against the fp32 reference through ~70 transformer blocks (the same budget
    # as the 32x32 full-size test above; the reference's own bf16-mixed autocast path sits at the same level)
    assert e_out < 5e-2, stats
    assert stats["l2_median"] < 3e-2 and stats["l2_p99"] < 1.5e-1, stats
    assert max(full.values()) < 3e-1 and float(np.median(list(full.values()))) < 8e-2, stats
    del out
    m.zero_grad(set_to_none=True)
    with torch.no_grad():
        xb = synth_tensor("full144x112.x", (1, 4, 144, 112)).to(DEV)
        ob = m(xb, ts, ctx, y)
    assert ob.shape == (1, 4, 144, 112)
    assert rel(ob, G128["full144x112.out"]) < 5e-2


def test_vae_encoder_vs_golden():
    from neurosis_b200.modules.vae import Encoder
    enc = Encoder(**TINY_VAE, embed_dim=4, standalone=True)
    enc.load_state_dict(synth_state_dict(vae_param_shapes(TINY_VAE, embed_dim=4, standalone=True), seed=2))
    enc = enc.to(DEV)
    img = synth_tensor("vae.img", (2, 3, 32, 32), uniform=True).to(DEV)
    with torch.no_grad():
        z = enc(img, regularize=True)
    assert z.shape == (2, 4, 16, 16)
    assert rel(z, G["vae.z"]) < 3e-2


def test_vae_mid_attention_block():
    """AttnBlock (single head, d = channels) against torch fp32 on the same weights."""
    from neurosis_b200.modules.vae import AttnBlock
    torch.manual_seed(0)
    blk = AttnBlock(128).to(DEV)
    x = torch.randn(2, 128, 16, 16, device=DEV)
    with torch.no_grad():
        y = blk(x)
        t = torch.nn.functional.group_norm(x, 32, blk.norm.weight, blk.norm.bias, 1e-6)
        q, k, v = (m(t).flatten(2).transpose(1, 2) for m in (blk.q, blk.k, blk.v))
        a = torch.softmax(q @ k.transpose(1, 2) * 128 ** -0.5, -1) @ v
        ref = x + blk.proj_out(a.transpose(1, 2).reshape(2, 128, 16, 16))
    assert rel(y, ref) < 2e-2


def test_training_step_loss_vs_golden_and_oracle():
    """loss[B] of the full objective (noise mix, sigma quantisation, c_in/c_out/c_skip, UNet, weighted MSE) for a
    fixed sigma draw and noise: bf16 step loss within 1e-2 relative of the reference's fp32 value."""
    from neurosis_b200.modules.denoiser import DiscreteDenoiser, EpsPreconditioning, EpsWeighting
    from neurosis_b200.modules.loss import OpenAIWrapper, StandardDiffusionLoss
    from neurosis_b200.modules.schedule import LegacyDDPMDiscretization
    cfg = TINY_SDXL
    m = build_unet(cfg)
    den = DiscreteDenoiser(EpsPreconditioning(), 1000, LegacyDDPMDiscretization()).to(DEV)
    sig = torch.from_numpy(G["step.sigmas"])

    class Fixed:
        def __call__(self, n, t=None):
            return sig

    loss_fn = StandardDiffusionLoss(sigma_generator=Fixed(), loss_weighting=EpsWeighting())
    lat = synth_tensor("step.latent", (2, 4, 16, 16)).to(DEV)
    noise = synth_tensor("step.noise", (2, 4, 16, 16)).to(DEV)
    cond = {"crossattn": synth_tensor("sdxl.ctx", (2, 77, cfg["context_dim"])).to(DEV),
            "vector": synth_tensor("sdxl.y", (2, cfg["adm_in_channels"])).to(DEV)}
    loss = loss_fn._forward(OpenAIWrapper(m), den, cond, lat, {}, noise=noise)
    assert loss.shape == (2,) and loss.dtype == torch.float32
    np.testing.assert_allclose(loss.detach().cpu().numpy(), G["step.loss"], rtol=1e-2)
    loss.mean().backward()
    assert rel(m.out[2].weight.grad, G["step.grad.out.2.weight"]) < 3e-2
    names = sorted(n for n, _ in m.named_parameters())
    l2 = np.array([float(dict(m.named_parameters())[n].grad.norm()) for n in names])
    assert np.allclose(l2, G["step.grad_l2"], rtol=6e-2, atol=1e-5)


def test_engine_training_step_end_to_end():
    """VAE encode (no grad) -> loss -> hook -> mean -> backward through the engine glue."""
    from neurosis_b200.engine import DiffusionEngine
    from neurosis_b200.modules.conditioner import GeneralConditioner, IdentityEncoder
    from neurosis_b200.modules.denoiser import DiscreteDenoiser, EpsPreconditioning, EpsWeighting
    from neurosis_b200.modules.loss import StandardDiffusionLoss, TagFrequencyHook
    from neurosis_b200.modules.schedule import DiscreteSigmaGenerator, LegacyDDPMDiscretization
    from neurosis_b200.modules.vae import Encoder
    cfg = TINY_SDXL
    unet = build_unet(cfg)
    enc = Encoder(**TINY_VAE, embed_dim=4, standalone=True)
    enc.load_state_dict(synth_state_dict(vae_param_shapes(TINY_VAE, embed_dim=4, standalone=True), seed=2))

    class RandIdx(DiscreteSigmaGenerator):  # harness-side generator: the SGM randint branch (t ignored)
        def __call__(self, n, t=None):
            return super().__call__(n, None).clamp_min(0.03)

    eng = DiffusionEngine(unet, DiscreteDenoiser(EpsPreconditioning(), 1000, LegacyDDPMDiscretization()), enc,
                          GeneralConditioner([IdentityEncoder(input_key="ctx"), IdentityEncoder(input_key="vec")]),
                          StandardDiffusionLoss(RandIdx(LegacyDDPMDiscretization(), 1000), EpsWeighting()),
                          scale_factor=0.13025, forward_hooks=[TagFrequencyHook(alpha=0.2)]).to(DEV)
    torch.manual_seed(42)
    batch = {"image": synth_tensor("vae.img", (2, 3, 128, 128), uniform=True).to(DEV),
             "ctx": synth_tensor("sdxl.ctx", (2, 77, cfg["context_dim"])).to(DEV),
             "vec": synth_tensor("sdxl.y", (2, cfg["adm_in_channels"])).to(DEV),
             "caption": ["1girl solo", "landscape scenery sky"]}
    loss = eng.training_step(batch)
    assert loss.ndim == 0 and torch.isfinite(loss)
    loss.backward()
    grads = [p.grad for p in unet.parameters()]
    assert all(g is not None and torch.isfinite(g).all() for g in grads)
    assert all(p.grad is None for p in enc.parameters())


def test_grad_sink_matches_autograd_accumulation():
    """weight gradients written straight into the reducer's buckets == gradients accumulated by autograd."""
    from neurosis_b200.ddp import BucketedGradReducer
    cfg = TINY_SD15
    x = synth_tensor("sd15.x", (2, 4, 16, 16)).to(DEV)
    ctx = synth_tensor("sd15.ctx", (2, 77, cfg["context_dim"])).to(DEV)
    ts = torch.tensor([17, 803], device=DEV)
    gout = synth_tensor("sd15.gout", (2, 4, 16, 16), scale=0.1).to(DEV)
    m1 = build_unet(cfg)
    (m1(x, ts, ctx) * gout).sum().backward()
    m2 = build_unet(cfg)
    red = BucketedGradReducer(m2.parameters(), bucket_mb=8.0)
    red.attach_as_grad_sink()
    try:
        for _ in range(2):  # second pass checks zero_grad + re-accumulation into the same buckets
            red.zero_grad()
            (m2(x, ts, ctx) * gout).sum().backward()
            red.finish()
    finally:
        red.detach_grad_sink()
    errs = {}
    for (n1, p1), (n2, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        assert p2.grad.data_ptr() == next(v for q, v in zip(red.buckets[red._index[p2]]["params"], red._views(red.buckets[red._index[p2]])) if q is p2).data_ptr()
        errs[n1] = rel(p2.grad, p1.grad)
    # two independent bf16 runs differ by rounding noise (fp32 atomics reorder sums -> bf16 roundings flip and the
    # flips are amplified layer by layer), so the bound is statistical: typical parameters agree to ~1e-2
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    assert np.median(list(errs.values())) < 2e-2, worst
    assert worst[0][1] < 1e-1, worst


def test_cuda_graph_step_matches_eager_gradients():
    """GraphedTrainStep (whole step in one CUDA graph) reproduces the eager step: same sigmas, same noise stream."""
    from neurosis_b200.ddp import BucketedGradReducer
    from neurosis_b200.engine import DiffusionEngine
    from neurosis_b200.graph import GraphedTrainStep
    from neurosis_b200.modules.conditioner import GeneralConditioner, IdentityEncoder
    from neurosis_b200.modules.denoiser import DiscreteDenoiser, EpsPreconditioning, EpsWeighting
    from neurosis_b200.modules.loss import StandardDiffusionLoss
    from neurosis_b200.modules.schedule import DiscreteSigmaGenerator, LegacyDDPMDiscretization
    from neurosis_b200.modules.vae import Encoder
    cfg = TINY_SDXL

    class RandIdx(DiscreteSigmaGenerator):
        def __call__(self, n, t=None):
            return super().__call__(n, None).clamp_min(0.03)

    def make():
        unet = build_unet(cfg)
        enc = Encoder(**TINY_VAE, embed_dim=4, standalone=True)
        enc.load_state_dict(synth_state_dict(vae_param_shapes(TINY_VAE, embed_dim=4, standalone=True), seed=2))
        return DiffusionEngine(unet, DiscreteDenoiser(EpsPreconditioning(), 1000, LegacyDDPMDiscretization()), enc,
                               GeneralConditioner([IdentityEncoder(input_key="ctx"), IdentityEncoder(input_key="vec")]),
                               StandardDiffusionLoss(RandIdx(LegacyDDPMDiscretization(), 1000), EpsWeighting()),
                               scale_factor=0.13025).to(DEV)

    img = synth_tensor("vae.img", (2, 3, 64, 64), uniform=True).to(DEV)
    ctx = synth_tensor("sdxl.ctx", (2, 77, cfg["context_dim"])).to(DEV)
    vec = synth_tensor("sdxl.y", (2, cfg["adm_in_channels"])).to(DEV)
    eng = make()
    red = BucketedGradReducer([p for p in eng.model.parameters() if p.requires_grad], bucket_mb=8.0)
    red.attach_as_grad_sink()
    try:
        g = GraphedTrainStep(eng, red, img, ctx, vec, warmup=2)
        assert g.launches_per_replay > 100
        l1 = float(g.step().item())
        l2 = float(g.step(img, ctx, vec).item())
        assert np.isfinite(l1) and np.isfinite(l2) and l1 != l2  # new sigma draw and noise every replay
        # ---- a replay with pinned sigmas and a fixed device RNG seed ...
        g._refresh_sigmas = lambda: None  # keep the sigmas of the last draw in the static buffer
        torch.cuda.manual_seed(1234)
        l3 = float(g.step().item())
        ps3 = g.per_sample.clone()
        grads3 = {n: p.grad.clone() for n, p in eng.model.named_parameters()}
        # ---- ... against the SAME step issued eagerly (no graph) with the same sigmas and the same RNG state: the graph
        # registers the CUDA generator, so both paths draw the same posterior sample and the same noise
        def eager():
            torch.cuda.manual_seed(1234)
            g._core()
            torch.cuda.synchronize()
            return float(g.loss.item()), g.per_sample.clone(), {n: p.grad.clone() for n, p in eng.model.named_parameters()}

        le, pse, gradse = eager()
        le2, _, gradse2 = eager()  # the yardstick: how far apart are two EAGER runs of the same step
    finally:
        red.detach_grad_sink()
    # identical kernels on identical inputs and the same noise.  What differs between a replay and an eager run is the
    # order of fp32 atomic accumulation (GroupNorm partial statistics in the forward; split-K weight gradients, the
    # attention dQ reduce-add and norm parameter gradients in the backward); an fp32 ulp in a statistic flips individual
    # bf16 roundings downstream, which the ~40-layer backward of this 64-channel miniature amplifies to ~1e-2 on
    # gradients (the loss agrees to ~1e-4).  So the criterion is: a replay is as close to an eager step as two eager
    # steps are to each other, and well inside the bf16-vs-fp32 tolerance of the parity tests.
    assert np.isfinite(l3) and abs(l3 - le) <= 1e-3 * abs(le), (l3, le)
    assert rel(ps3, pse) < 1e-3
    errs = {n: rel(grads3[n], gradse[n]) for n in grads3}
    errs_ee = {n: rel(gradse2[n], gradse[n]) for n in grads3}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    med, med_ee = float(np.median(list(errs.values()))), float(np.median(list(errs_ee.values())))
    print("graph vs eager: loss", l3, le, le2, "grad rel L2 median", med, "eager vs eager", med_ee, "worst", worst)
    assert all(float(v.abs().sum()) > 0 for v in list(gradse.values())[:8])
    assert med < max(3.0 * med_ee, 2e-3) and med < 3e-2, (med, med_ee, worst)
    assert worst[0][1] < 6e-2, worst
