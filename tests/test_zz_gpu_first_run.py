"""GPU tests of the opt-in "optimizer + EMA inside the captured step" path, the L1 / rectified-flow objectives and the
full-size VAE training step.

History: written at the end of round 1 after that round's GPU budget was spent and marked non-strict xfail; all of them
passed on the driver's round-end box (GPUTEST_r01.json: 4 xpassed), so the markers are gone and a regression now fails
the suite.  The file still sorts last: the graphed optimizer test is the heaviest capture in the suite."""
import numpy as np
import pytest
import torch

from common import ROOT, TINY_VAE
from oracle.vae import vae_param_shapes
from oracle.weights import synth_state_dict, synth_tensor
from test_gpu_next import DEV, _opt_params, rel, rnd

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120)]


# ---------------------------------------------------------------- optimizer / EMA inside the captured step (opt-in)


def test_device_side_step_scalars_match_host_side():  # seen green on a B200 (profiles/r01_pytest_gpu_first_run.log)
    """`graph_launch` (step count and hyper-parameters derived on the device) == `step` (host scalars), and the EMA
    decay warm-up from the device counter == the host expression."""
    from neurosis_b200.optim import Adafactor, LitEma
    from test_oracle_golden_next import OPT_SHAPES
    for kw in (dict(scale_parameter=True, relative_step=True, warmup_init=True),
               dict(lr=1e-3, scale_parameter=False, relative_step=False, beta1=0.9, weight_decay=0.01)):
        pa, pb = _opt_params(kw), _opt_params(kw)
        oa, ob = Adafactor(list(pa.values()), **kw), Adafactor(list(pb.values()), **kw)
        for k in pa:
            pa[k].grad = synth_tensor(f"opt.g.{k}.0", OPT_SHAPES[k], scale=0.02).to(DEV)
            pb[k].grad = pa[k].grad.clone()
        ob.graph_prepare()
        for _ in range(3):
            oa.step()
            ob.graph_launch()
        ob.sync_steps_from_device()
        for k in pa:
            assert ob.state[pb[k]]["step"] == 3
            d_a, d_b = pa[k].detach() - synth_tensor(f"opt.p.{k}", OPT_SHAPES[k], scale=0.05).to(DEV), \
                pb[k].detach() - synth_tensor(f"opt.p.{k}", OPT_SHAPES[k], scale=0.05).to(DEV)
            assert rel(d_b, d_a) < (5e-2 if kw.get("relative_step") else 1e-4), k  # ~3e-7 movements resolve to ~2 %

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = torch.nn.Linear(40, 30)

    ma, mb = M().to(DEV), M().to(DEV)
    mb.load_state_dict(ma.state_dict())
    ea, eb = LitEma(ma, decay=0.9999).to(DEV), LitEma(mb, decay=0.9999).to(DEV)
    eb.graph_prepare(mb)
    for it in range(12):
        with torch.no_grad():
            d = rnd(30, 40, seed=it, scale=0.1)
            ma.a.weight.add_(d)
            mb.a.weight.add_(d)
        ea(ma)
        eb.graph_launch()
    eb.sync_from_device()
    assert int(eb.num_updates) == 12 and eb._n_host == 12
    assert rel(eb.a_weight, ea.a_weight) < 1e-6


def test_graphed_step_with_optimizer_and_ema():
    """one replay = refresh -> encode -> loss -> backward -> reduce -> Adafactor -> EMA; counters advance per replay."""
    from test_gpu_modules import build_unet  # noqa: F401  (shared tiny-model builder)
    from common import TINY_SDXL, TINY_VAE as TV
    from neurosis_b200.ddp import BucketedGradReducer
    from neurosis_b200.engine import DiffusionEngine
    from neurosis_b200.graph import GraphedTrainStep
    from neurosis_b200.modules.conditioner import GeneralConditioner, IdentityEncoder
    from neurosis_b200.modules.denoiser import DiscreteDenoiser, EpsPreconditioning, EpsWeighting
    from neurosis_b200.modules.loss import StandardDiffusionLoss
    from neurosis_b200.modules.schedule import DiscreteSigmaGenerator, LegacyDDPMDiscretization
    from neurosis_b200.modules.vae import Encoder
    from neurosis_b200.optim import Adafactor, LitEma

    class RandIdx(DiscreteSigmaGenerator):
        def __call__(self, n, t=None):
            return super().__call__(n, None).clamp_min(0.03)

    cfg = TINY_SDXL
    unet = build_unet(cfg)
    enc = Encoder(**TV, embed_dim=4, standalone=True)
    enc.load_state_dict(synth_state_dict(vae_param_shapes(TV, embed_dim=4, standalone=True), seed=2))
    eng = DiffusionEngine(unet, DiscreteDenoiser(EpsPreconditioning(), 1000, LegacyDDPMDiscretization()), enc,
                          GeneralConditioner([IdentityEncoder(input_key="ctx"), IdentityEncoder(input_key="vec")]),
                          StandardDiffusionLoss(RandIdx(LegacyDDPMDiscretization(), 1000), EpsWeighting()),
                          scale_factor=0.13025).to(DEV)
    params = [p for p in eng.model.parameters() if p.requires_grad]
    p0 = [p.detach().clone() for p in params]
    red = BucketedGradReducer(params, bucket_mb=8.0)
    red.attach_as_grad_sink()
    opt = Adafactor(params, lr=1e-3, relative_step=False, scale_parameter=False)
    ema = LitEma(eng.model, decay=0.9999)
    img = synth_tensor("vae.img", (2, 3, 64, 64), uniform=True).to(DEV)
    ctx = synth_tensor("sdxl.ctx", (2, 77, cfg["context_dim"])).to(DEV)
    vec = synth_tensor("sdxl.y", (2, cfg["adm_in_channels"])).to(DEV)
    try:
        g = GraphedTrainStep(eng, red, img, ctx, vec, warmup=2, optimizer=opt, ema=ema)
        assert all(torch.equal(a, b.detach()) for a, b in zip(p0, params)), "warm-up and capture leave the weights alone"
        losses = [float(g.step().item()) for _ in range(3)]
        g.sync_host_state()
    finally:
        red.detach_grad_sink()
    assert all(np.isfinite(losses))
    assert all(opt.state[p]["step"] == 3 for p in params) and int(ema.num_updates) == 3
    moved = [float((a - b.detach()).abs().max()) for a, b in zip(p0, params)]
    # lr 1e-3, unit-RMS (clipped) updates, 3 steps: RMS movement ~3e-3, single elements well below 0.5; parameters whose
    # gradient is exactly zero do not move, so "almost all" instead of "all"
    assert all(np.isfinite(moved)) and np.mean(np.array(moved) > 0) > 0.95 and max(moved) < 0.5
    sh = dict(ema.named_buffers())
    w = eng.model.diffusion_model.out[2].weight
    s = sh["diffusion_model_out_2_weight"]
    assert torch.isfinite(s).all() and float((s - w.detach()).abs().max()) > 0


def test_weighted_l1_loss_fwd_bwd():
    """StandardDiffusionLoss(loss_type="l1") arithmetic: loss[b] = w[b] * mean|D - T| (BatchL1Loss) and its gradient."""
    from neurosis_b200 import ops
    D = rnd(3, 4, 16, 16).requires_grad_(True)
    T = rnd(3, 4, 16, 16, seed=1)
    with torch.no_grad():
        D[0, 0, 0, :4] = T[0, 0, 0, :4]  # exact ties: sign(0) = 0 on both sides
    w = rnd(3, seed=2).abs() + 0.1
    loss = ops.weighted_l1(D, T, w)
    g = rnd(3, seed=3)
    loss.backward(g)
    Dr = D.detach().clone().requires_grad_(True)
    lr_ = (Dr - T).abs().flatten(1).mean(1) * w
    lr_.backward(g)
    assert rel(loss, lr_) < 1e-5 and rel(D.grad, Dr.grad) < 1e-6


def test_full_size_vae_training_step_vs_reference_golden():
    """AutoencoderKL at the FULL SDXL KL-f8 configuration (108 + 140 tensors; 512-channel mid attention, d = 512) on a
    64x64 image: moments, reconstruction and loss against the reference's Encoder / Decoder (goldens), gradient norms
    of both networks."""
    from common import FULL_VAE, fast_state_dict
    from oracle.vae import vae_decoder_param_shapes
    from neurosis_b200.modules.vae import AutoencoderKL, DiagonalGaussianRegularizer
    G = np.load(str(ROOT / "tests/golden/reference_golden_next.npz"))
    es, ds = vae_param_shapes(FULL_VAE, 4, True), vae_decoder_param_shapes(FULL_VAE, 4, True)
    full = {}
    for k, v in fast_state_dict(es, seed=11).items():
        full[k if k.startswith("quant_conv.") else "encoder." + k] = v
    for k, v in fast_state_dict(ds, seed=12).items():
        full[k if k.startswith("post_quant_conv.") else "decoder." + k] = v
    ae = AutoencoderKL(4, FULL_VAE, regularizer=DiagonalGaussianRegularizer(sample=True))
    ae.load_state_dict(full)
    ae = ae.to(DEV)
    img = synth_tensor("fullvae.img", (1, 3, 64, 64), uniform=True).to(DEV)
    eps = synth_tensor("fullvae.eps", (1, 4, 8, 8)).to(DEV)
    with torch.no_grad():
        m = ae.encoder.moments(img, ae.quant_conv)
    assert rel(m, G["fullvae.moments"]) < 3e-2
    loss = ae.training_step({"image": img, "posterior_eps": eps})
    np.testing.assert_allclose(float(loss.detach()), G["fullvae.loss"], rtol=1e-2)
    loss.backward()
    named = dict(ae.named_parameters())
    enc_l2 = np.array([float(named[k if k.startswith("quant_conv.") else "encoder." + k].grad.norm()) for k in sorted(es)])
    dec_l2 = np.array([float(named[k if k.startswith("post_quant_conv.") else "decoder." + k].grad.norm())
                       for k in sorted(ds)])
    keep_e = np.array([not k.endswith(".k.bias") for k in sorted(es)])  # mathematically zero gradient
    keep_d = np.array([not k.endswith(".k.bias") for k in sorted(ds)])
    assert np.allclose(enc_l2[keep_e], G["fullvae.enc_grad_l2"][keep_e], rtol=8e-2, atol=1e-7)
    assert np.allclose(dec_l2[keep_d], G["fullvae.dec_grad_l2"][keep_d], rtol=8e-2, atol=1e-7)


def test_rectified_flow_objective_vs_reference_golden():
    """StandardDiffusionLoss(objective_type="rf") + continuous Denoiser(RectifiedFlowComfyPreconditioning) +
    RectifiedFlowComfyWeighting on the CUDA modules: bf16 step loss within 1e-2 of the reference's fp32 value, gradient
    norms of every parameter (non-integer timesteps 300.0 / 800.0 through the sinusoidal-embedding kernel)."""
    from common import TINY_SDXL
    from test_gpu_modules import build_unet
    from neurosis_b200.modules.denoiser import Denoiser, RectifiedFlowComfyPreconditioning, RectifiedFlowComfyWeighting
    from neurosis_b200.modules.loss import OpenAIWrapper, StandardDiffusionLoss
    G = np.load(str(ROOT / "tests/golden/reference_golden_next.npz"))
    cfg = TINY_SDXL
    m = build_unet(cfg)
    sig = torch.from_numpy(G["rf.sigmas"])

    class Fixed:
        def __call__(self, n, t=None):
            return sig

    loss_fn = StandardDiffusionLoss(sigma_generator=Fixed(), loss_weighting=RectifiedFlowComfyWeighting(),
                                    objective_type="rf")
    lat = synth_tensor("step.latent", (2, 4, 16, 16)).to(DEV)
    noise = synth_tensor("step.noise", (2, 4, 16, 16)).to(DEV)
    cond = {"crossattn": synth_tensor("sdxl.ctx", (2, 77, cfg["context_dim"])).to(DEV),
            "vector": synth_tensor("sdxl.y", (2, cfg["adm_in_channels"])).to(DEV)}
    loss = loss_fn._forward(OpenAIWrapper(m), Denoiser(RectifiedFlowComfyPreconditioning()), cond, lat, {}, noise=noise)
    np.testing.assert_allclose(loss.detach().cpu().numpy(), G["rf.loss"], rtol=1e-2)
    loss.mean().backward()
    assert rel(m.out[2].weight.grad, G["rf.grad.out.2.weight"]) < 3e-2
    names = sorted(n for n, _ in m.named_parameters())
    l2 = np.array([float(dict(m.named_parameters())[n].grad.norm()) for n in names])
    assert np.allclose(l2, G["rf.grad_l2"], rtol=6e-2, atol=1e-5)
