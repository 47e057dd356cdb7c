"""The oracle restatements of the SURVEY.md §8(f) rows against golden vectors produced by the reference itself
(tests/golden/make_golden_next.py).  Runs anywhere — no GPU, no /root/reference."""
import numpy as np
import torch

from common import TINY_VAE
from oracle.vae import vae_decode, vae_decoder_param_shapes, vae_param_shapes, vae_train_loss
from oracle.weights import synth_state_dict, synth_tensor

G = np.load(__file__.rsplit("/", 1)[0] + "/golden/reference_golden_next.npz")


def test_vae_decoder_forward_backward():
    shapes = vae_decoder_param_shapes(TINY_VAE, embed_dim=4, standalone=True)
    sd = {k: v.requires_grad_(True) for k, v in synth_state_dict(shapes, seed=5).items()}
    z = synth_tensor("vaedec.z", (2, 4, 16, 16))
    xr = vae_decode(sd, TINY_VAE, z)
    np.testing.assert_allclose(xr.detach().numpy(), G["vaedec.out"], rtol=1e-4, atol=2e-5)
    (xr * synth_tensor("vaedec.g", tuple(xr.shape), scale=0.1)).sum().backward()
    l2 = np.array([sd[n].grad.norm().item() for n in sorted(shapes)])
    np.testing.assert_allclose(l2, G["vaedec.grad_l2"], rtol=2e-4, atol=1e-6)
    for n in ("conv_out.weight", "post_quant_conv.weight"):
        np.testing.assert_allclose(sd[n].grad.numpy(), G[f"vaedec.grad.{n}"], rtol=1e-3, atol=1e-5)


def test_vae_training_step():
    eshapes = vae_param_shapes(TINY_VAE, embed_dim=4, standalone=True)
    dshapes = vae_decoder_param_shapes(TINY_VAE, embed_dim=4, standalone=True)
    esd = {k: v.requires_grad_(True) for k, v in synth_state_dict(eshapes, seed=2).items()}
    dsd = {k: v.requires_grad_(True) for k, v in synth_state_dict(dshapes, seed=5).items()}
    img = synth_tensor("vae.img", (2, 3, 32, 32), uniform=True)
    eps = synth_tensor("vaetrain.eps", (2, 4, 16, 16))
    loss, z, xrec, kl = vae_train_loss(esd, dsd, TINY_VAE, img, eps)
    np.testing.assert_allclose(z.detach().numpy(), G["vaetrain.z"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(xrec.detach().numpy(), G["vaetrain.xrec"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(loss.item(), G["vaetrain.loss"], rtol=1e-5)
    np.testing.assert_allclose(kl.item(), G["vaetrain.kl_loss"], rtol=1e-5)
    loss.backward()
    np.testing.assert_allclose(np.array([esd[n].grad.norm().item() for n in sorted(eshapes)]),
                               G["vaetrain.enc_grad_l2"], rtol=5e-4, atol=1e-8)
    np.testing.assert_allclose(np.array([dsd[n].grad.norm().item() for n in sorted(dshapes)]),
                               G["vaetrain.dec_grad_l2"], rtol=5e-4, atol=1e-8)
    np.testing.assert_allclose(esd["quant_conv.weight"].grad.numpy(), G["vaetrain.grad.quant_conv.weight"], rtol=1e-3,
                               atol=1e-7)
    np.testing.assert_allclose(esd["conv_in.weight"].grad.numpy(), G["vaetrain.grad.conv_in.weight"], rtol=1e-3,
                               atol=1e-7)


# ---------------------------------------------------------------- rows 2 and 4: Adafactor, EMA
OPT_SHAPES = {"lin": (96, 200), "lin_ragged": (130, 70), "conv3": (24, 16, 3, 3), "conv1": (8, 12, 1, 1),
              "bias": (300,), "stack": (2, 70, 40)}
OPT_CASES = {
    "yaml": dict(scale_parameter=True, relative_step=True, warmup_init=True),
    "ext": dict(lr=1e-3, scale_parameter=False, relative_step=False, warmup_init=False, beta1=0.9, weight_decay=0.01,
                clip_threshold=0.5),
}


def check_optimizer_state(case, name, p, st, p0):
    """shared by the CPU (oracle) and GPU (kernel) tests: parameter, moments and RMS against the reference's."""
    g = G[f"opt.{case}.{name}.p"]
    p = np.asarray(p, dtype=np.float32)
    # the relative-step case moves parameters by ~3e-7 in total: compare the values to 2 ulp (1.5e-8 at |p| ~ 0.2) and
    # the movement itself
    np.testing.assert_allclose(p, g, rtol=0, atol=3e-8 if case == "yaml" else 2e-7)
    d, dg = p - p0, g - p0
    assert np.linalg.norm(d - dg) <= (0.05 if case == "yaml" else 2e-4) * np.linalg.norm(dg), (case, name)
    for sk in ("exp_avg_sq_row", "exp_avg_sq_col", "exp_avg_sq", "exp_avg"):
        key = f"opt.{case}.{name}.{sk}"
        if key in G.files:
            # first moments mix terms of opposite sign: tolerance relative to the tensor's scale, not per element
            np.testing.assert_allclose(np.asarray(st[sk], dtype=np.float32), G[key], rtol=2e-5,
                                       atol=2e-6 * float(np.abs(G[key]).max()), err_msg=key)
    np.testing.assert_allclose(float(st["RMS"]), G[f"opt.{case}.{name}.rms"], rtol=1e-5)


def test_adafactor_oracle_vs_reference():
    from oracle.optim import adafactor_init_state, adafactor_step
    for case, kw in OPT_CASES.items():
        for name, shape in OPT_SHAPES.items():
            p = synth_tensor(f"opt.p.{name}", shape, scale=0.05)
            p0 = p.numpy().copy()
            st = adafactor_init_state(p, kw.get("beta1"))
            for step in range(3):
                adafactor_step(p, synth_tensor(f"opt.g.{name}.{step}", shape, scale=0.02 * (step + 1)), st, **kw)
            check_optimizer_state(case, name, p.numpy(), {k: (v.numpy() if torch.is_tensor(v) else v)
                                                          for k, v in st.items()}, p0)


EMA_SHAPES = {"a.weight": (30, 40), "a.bias": (30,), "b.weight": (8, 8, 3, 3), "b.bias": (8,)}


def test_ema_oracle_vs_reference():
    from oracle.optim import ema_decay, ema_update
    for name, shape in EMA_SHAPES.items():
        p = synth_tensor(f"ema.p.{name}", shape)
        shadow = p.clone()
        for it in range(12):
            p = p + synth_tensor(f"ema.d.{name}.{it}", shape, scale=0.1)
            ema_update(shadow, p, ema_decay(0.9999, it + 1))
        np.testing.assert_allclose(shadow.numpy(), G[f"ema.{name}"], rtol=1e-6, atol=1e-7)
    assert int(G["ema.num_updates"]) == 12


# ---------------------------------------------------------------- §8(a) rows 6-8: every shipped variant, bit-exact
def _family_cases():
    return [k for k in G.files if k.startswith("fam.") and not k.startswith("fam.dd.")
            and k not in ("fam.sigma", "fam.sigma01", "fam.t", "fam.errors")]


def test_discrete_denoiser_option_combinations_bit_exact():
    """DiscreteDenoiser with quantize_c_noise on/off, flip on/off, Eps / V preconditioning (denoiser.py:60-97): the
    quantised sigma and the c_noise handed to the UNet (an int64 table index, or the continuous value) are bit-exact."""
    from neurosis_b200.modules import denoiser as D
    from neurosis_b200.modules.schedule import LegacyDDPMDiscretization
    sig = torch.from_numpy(G["fam.sigma"])
    keys = [k for k in G.files if k.startswith("fam.dd.") and k.endswith(".sigma")]
    assert len(keys) == 8
    for key in keys:
        _, _, pname, q, flip, _ = key.split(".")
        den = D.DiscreteDenoiser(getattr(D, pname)(), 1000, LegacyDDPMDiscretization(), quantize_c_noise=bool(int(q)),
                                 flip=bool(int(flip)))
        s = den.possibly_quantize_sigma(sig)
        cn = den.possibly_quantize_c_noise(den.preconditioning(s)[3])
        assert np.array_equal(s.double().numpy(), G[key]), key
        assert np.array_equal(cn.double().numpy(), G[key[: -len("sigma")] + "c_noise"]), key
        assert cn.dtype == (torch.int64 if int(q) else torch.float32)


def test_schedule_and_denoiser_families_bit_exact_vs_reference():
    """every Discretization / SigmaGenerator / DenoiserPreconditioning / DenoiserWeighting class the reference ships,
    with default constructor arguments, on fixed inputs: the drop-in classes return the SAME BITS (they use the
    reference's torch op / dtype sequence), for n = 1000 and 40 steps, flipped and not, t given and drawn."""
    from neurosis_b200.modules import denoiser as D
    from neurosis_b200.modules import schedule as S
    sig, sig01, t = (torch.from_numpy(G[k]) for k in ("fam.sigma", "fam.sigma01", "fam.t"))
    cases = _family_cases()
    assert len(cases) == 68  # 26 tables + 24 preconditioning terms + 6 weightings + 12 generator draws
    seen = set()
    for key in cases:
        _, kind, name, *rest = key.split(".")
        seen.add((kind, name))
        if kind == "disc":
            val = getattr(S, name)()(int(rest[0]), flip=bool(int(rest[1])))
        elif kind == "precond":
            val = getattr(D, name)()(sig01 if "RectifiedFlow" in name else sig)[int(rest[0])]
        elif kind == "weight":
            obj = D.MinSNRGammaModifier(D.EpsWeighting()) if name == "MinSNRGammaModifier" else getattr(D, name)()
            val = obj(sig01 if "RectifiedFlow" in name else sig)
        else:
            gen = (S.DiscreteSigmaGenerator(S.LegacyDDPMDiscretization(), 1000) if name == "DiscreteSigmaGenerator"
                   else getattr(S, name)())
            if rest[0] == "t":
                val = gen(16, t.clone())
            else:
                torch.manual_seed(7)
                val = gen(16, None)
        got = torch.as_tensor(val).detach().double().numpy()
        assert got.shape == G[key].shape and np.array_equal(got, G[key]), key
    assert len({n for k, n in seen if k == "disc"}) == 7 and len({n for k, n in seen if k == "gen"}) == 6
    assert len({n for k, n in seen if k == "precond"}) == 6 and len({n for k, n in seen if k == "weight"}) == 6
    # the only reference failure on these inputs: LegacyDDPMDiscretization(n < 1000) raises (negative numpy stride);
    # the drop-in returns the table instead
    assert sorted(G["fam.errors"]) == ["fam.disc.LegacyDDPMDiscretization.40.0=ValueError",
                                       "fam.disc.LegacyDDPMDiscretization.40.1=ValueError"]
    from neurosis_b200.modules.schedule import LegacyDDPMDiscretization
    assert LegacyDDPMDiscretization()(40).shape == (41,)


# ---------------------------------------------------------------- the oracle at full size
def _full_size_oracle_vs_reference(tag, cfg):
    from common import fast_state_dict
    from oracle.unet import unet_forward, unet_param_shapes
    shapes = unet_param_shapes(cfg)
    sd = {k: v.requires_grad_(True) for k, v in fast_state_dict(shapes, seed=3).items()}
    x = synth_tensor(f"full.{tag}.x", (1, 4, 32, 32))
    ctx = synth_tensor(f"full.{tag}.ctx", (1, 77, cfg["context_dim"]))
    y = synth_tensor(f"full.{tag}.y", (1, cfg["adm_in_channels"])) if cfg.get("num_classes") else None
    o = unet_forward(sd, cfg, x, torch.tensor([481]), ctx, y)
    np.testing.assert_allclose(o.detach().numpy(), G[f"full.{tag}.out"], rtol=2e-4, atol=2e-4)
    (o * synth_tensor(f"full.{tag}.g", (1, 4, 32, 32), scale=0.1)).sum().backward()
    names = sorted(shapes)
    l2 = np.array([sd[n].grad.norm().item() for n in names])
    np.testing.assert_allclose(l2, G[f"full.{tag}.grad_l2"], rtol=1e-3, atol=1e-6)
    return len(names)


def test_oracle_full_size_sd15_vs_reference():
    """the oracle restatement against the reference's own UNetModel at the FULL SD1.5 configuration (859.5 M
    parameters, 686 tensors): output and every parameter's gradient norm."""
    from common import FULL_SD15
    assert _full_size_oracle_vs_reference("sd15", FULL_SD15) == 686


def test_oracle_full_size_sdxl_vs_reference():
    """same at the FULL SDXL configuration (2 567.5 M parameters, 1 680 tensors, transformer depth 10): this is the
    oracle the GPU full-size test (tests/test_gpu_modules.py::test_full_size_unet_vs_oracle) compares against.
    Needs ~25 GB of host memory for the fp32 weights and their gradients."""
    import psutil
    import pytest
    if psutil.virtual_memory().available < 40e9:
        pytest.skip("needs 40 GB of free host memory")
    from common import FULL_SDXL
    assert _full_size_oracle_vs_reference("sdxl", FULL_SDXL) == 1680


def test_oracle_full_size_vae_vs_reference():
    """oracle encoder + sampled posterior + decoder + L2 at the full SDXL KL-f8 configuration (ch 128, mult [1,2,4,4];
    108 + 140 tensors) against the reference's Encoder / Decoder: moments, reconstruction, loss, gradient norms."""
    from common import FULL_VAE, fast_state_dict
    from oracle.vae import vae_moments
    es, ds = vae_param_shapes(FULL_VAE, 4, True), vae_decoder_param_shapes(FULL_VAE, 4, True)
    esd = {k: v.requires_grad_(True) for k, v in fast_state_dict(es, seed=11).items()}
    dsd = {k: v.requires_grad_(True) for k, v in fast_state_dict(ds, seed=12).items()}
    img = synth_tensor("fullvae.img", (1, 3, 64, 64), uniform=True)
    eps = synth_tensor("fullvae.eps", (1, 4, 8, 8))
    with torch.no_grad():
        np.testing.assert_allclose(vae_moments(esd, FULL_VAE, img).numpy(), G["fullvae.moments"], rtol=1e-4, atol=2e-5)
    loss, _, xrec, _ = vae_train_loss(esd, dsd, FULL_VAE, img, eps)
    np.testing.assert_allclose(xrec.detach().numpy(), G["fullvae.xrec"], rtol=1e-4, atol=5e-5)
    np.testing.assert_allclose(loss.item(), G["fullvae.loss"], rtol=1e-5)
    loss.backward()
    np.testing.assert_allclose(np.array([esd[n].grad.norm().item() for n in sorted(es)]), G["fullvae.enc_grad_l2"],
                               rtol=1e-3, atol=1e-9)
    np.testing.assert_allclose(np.array([dsd[n].grad.norm().item() for n in sorted(ds)]), G["fullvae.dec_grad_l2"],
                               rtol=1e-3, atol=1e-9)


def test_rectified_flow_objective_oracle_vs_reference():
    """StandardDiffusionLoss(objective_type="rf") through a continuous Denoiser with the Comfy preconditioning /
    weighting: per-sample loss and every gradient norm of the miniature SDXL UNet against the reference."""
    from common import TINY_SDXL
    from oracle import objective as O
    from oracle.unet import unet_forward, unet_param_shapes
    cfg = TINY_SDXL
    sd = {k: v.requires_grad_(True) for k, v in synth_state_dict(unet_param_shapes(cfg), seed=1).items()}
    lat, noise = synth_tensor("step.latent", (2, 4, 16, 16)), synth_tensor("step.noise", (2, 4, 16, 16))
    cond = {"crossattn": synth_tensor("sdxl.ctx", (2, 77, cfg["context_dim"])),
            "vector": synth_tensor("sdxl.y", (2, cfg["adm_in_channels"]))}
    net = lambda x, t, c: unet_forward(sd, cfg, x, t, c["crossattn"], c["vector"])  # noqa: E731
    loss = O.rf_comfy_loss(net, lat, cond, torch.from_numpy(G["rf.sigmas"]), noise)
    np.testing.assert_allclose(loss.detach().numpy(), G["rf.loss"], rtol=1e-4)
    loss.mean().backward()
    np.testing.assert_allclose(sd["out.2.weight"].grad.numpy(), G["rf.grad.out.2.weight"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(np.array([sd[n].grad.norm().item() for n in sorted(sd)]), G["rf.grad_l2"], rtol=5e-4,
                               atol=1e-7)


def test_oracle_aspect_bucket_shapes_vs_reference():
    """non-square latents (aspect buckets): the oracle against the reference UNet on the shapes of the GPU bucket test."""
    from common import TINY_SDXL
    from oracle.unet import unet_forward, unet_param_shapes
    cfg = TINY_SDXL
    for h, w in ((24, 16), (12, 20)):
        sd = {k: v.requires_grad_(True) for k, v in synth_state_dict(unet_param_shapes(cfg), seed=1).items()}
        o = unet_forward(sd, cfg, synth_tensor("bucket.x", (2, 4, h, w)), torch.tensor([3, 977]),
                         synth_tensor("bucket.ctx", (2, 77, cfg["context_dim"])),
                         synth_tensor("bucket.y", (2, cfg["adm_in_channels"])))
        np.testing.assert_allclose(o.detach().numpy(), G[f"bucket.{h}x{w}.out"], rtol=1e-4, atol=2e-5)
        (o * synth_tensor("bucket.gout", (2, 4, h, w), scale=0.1)).sum().backward()
        np.testing.assert_allclose(np.array([sd[n].grad.norm().item() for n in sorted(sd)]),
                                   G[f"bucket.{h}x{w}.grad_l2"], rtol=2e-4, atol=1e-6)


def test_general_conditioner_assembly_and_ucg_masks_vs_reference():
    """GeneralConditioner (embedding.py:90-149) on a loader-style batch (per-sample size / crop tuples): concat order,
    Fourier features, per-sample UCG dropout masks (same torch RNG consumption under a fixed seed) and
    force_zero_embeddings — the drop-in's host path returns the reference's bits."""
    from neurosis_b200.modules.conditioner import ConcatTimestepEmbedderND, GeneralConditioner, IdentityEncoder
    gc = GeneralConditioner([IdentityEncoder(input_key="ctx"), IdentityEncoder(input_key="pooled", ucg_rate=0.3),
                             ConcatTimestepEmbedderND(256, input_key="original_size_as_tuple"),
                             ConcatTimestepEmbedderND(256, input_key="crop_coords_top_left", ucg_rate=0.5),
                             ConcatTimestepEmbedderND(256, input_key="target_size_as_tuple")])
    batch = {"image": torch.zeros(6, 3, 8, 8), "ctx": synth_tensor("gc.ctx", (6, 77, 32)),
             "pooled": synth_tensor("gc.pooled", (6, 48)),
             "original_size_as_tuple": [(1024, 1024), (1152, 896), (832, 1216), (1024, 1024), (640, 1536), (512, 512)],
             "crop_coords_top_left": [(0, 0), (16, 0), (0, 32), (8, 8), (0, 0), (64, 64)],
             "target_size_as_tuple": [(1024, 1024), (896, 1152), (1216, 832), (1024, 1024), (1536, 640), (512, 512)]}
    torch.manual_seed(1234)
    out = gc(batch)
    assert np.array_equal(out["crossattn"].numpy(), G["gc.crossattn"])
    assert out["vector"].shape == G["gc.vector"].shape == (6, 48 + 3 * 512)
    assert np.array_equal(out["vector"].numpy(), G["gc.vector"])
    assert 0 < int((G["gc.vector"][:, :48] == 0).all(1).sum()) + int((G["gc.vector"][:, 560:1072] == 0).all(1).sum())
    torch.manual_seed(1234)
    z = gc(batch, force_zero_embeddings=["pooled", "target_size_as_tuple"])
    assert np.array_equal(z["vector"].numpy(), G["gc.vector_force_zero"])
