"""Checkpoint key round trip: reference-layout and SGM-layout checkpoints load into the
drop-in engine, and the key rewrites equal the reference's converter functions (extracted from its scripts when
/root/reference exists).  Host logic only — no GPU."""
import ast
from pathlib import Path

import pytest
import torch

from common import TINY_SDXL, TINY_VAE, have_reference
from oracle.unet import unet_param_shapes
from oracle.vae import vae_decoder_param_shapes, vae_param_shapes
from oracle.weights import synth_state_dict


def _engine(first_stage):
    from neurosis_b200.engine import DiffusionEngine
    from neurosis_b200.modules import UNetModel
    from neurosis_b200.modules.conditioner import GeneralConditioner, IdentityEncoder
    from neurosis_b200.modules.denoiser import DiscreteDenoiser, EpsPreconditioning, EpsWeighting
    from neurosis_b200.modules.loss import StandardDiffusionLoss
    from neurosis_b200.modules.schedule import DiscreteSigmaGenerator, LegacyDDPMDiscretization
    return DiffusionEngine(UNetModel(**TINY_SDXL), DiscreteDenoiser(EpsPreconditioning(), 1000, LegacyDDPMDiscretization()),
                           first_stage, GeneralConditioner([IdentityEncoder(input_key="ctx")]),
                           StandardDiffusionLoss(DiscreteSigmaGenerator(LegacyDDPMDiscretization(), 1000), EpsWeighting()))


def _reference_layout_sd():
    """what a Lightning checkpoint of the reference engine holds (models/diffusion.py:73, 146-159)."""
    sd = {}
    for k, v in synth_state_dict(unet_param_shapes(TINY_SDXL), seed=1).items():
        sd["model.diffusion_model." + k] = v
    for k, v in synth_state_dict(vae_param_shapes(TINY_VAE, 4, True), seed=2).items():
        sd["vae_encoder." + k] = v
    for k, v in synth_state_dict(vae_decoder_param_shapes(TINY_VAE, 4, True), seed=5).items():
        sd["vae_decoder." + k] = v
    return sd


def test_reference_layout_loads_into_engine_with_encoder(tmp_path):
    from safetensors.torch import save_file
    from neurosis_b200.modules.vae import Encoder
    eng = _engine(Encoder(**TINY_VAE, embed_dim=4, standalone=True))
    sd = _reference_layout_sd()
    save_file(sd, str(tmp_path / "m.safetensors"))
    eng.init_from_ckpt(tmp_path / "m.safetensors")
    missing, unexpected = eng.last_ckpt_report
    assert not [m for m in missing if not m.startswith("denoiser.")], missing
    assert all(u.startswith("vae_decoder.") for u in unexpected)  # this engine holds no decoder
    assert torch.equal(eng.model.diffusion_model.out[2].weight, sd["model.diffusion_model.out.2.weight"])
    assert torch.equal(eng.first_stage_model.quant_conv.weight, sd["vae_encoder.quant_conv.weight"])
    assert torch.equal(eng.first_stage_model.down[0].block[0].conv1.weight, sd["vae_encoder.down.0.block.0.conv1.weight"])


def test_sgm_layout_loads_into_engine_with_autoencoder_and_round_trips(tmp_path):
    from neurosis_b200.checkpoint import neurosis_to_sgm, reference_state_dict, sgm_to_neurosis
    from neurosis_b200.modules.vae import AutoencoderKL
    eng = _engine(AutoencoderKL(4, TINY_VAE))
    ref_sd = _reference_layout_sd()
    sgm = neurosis_to_sgm(ref_sd)
    assert "first_stage_model.quant_conv.weight" in sgm and "first_stage_model.encoder.conv_in.weight" in sgm
    assert "first_stage_model.post_quant_conv.bias" in sgm and "first_stage_model.decoder.up.1.upsample.conv.weight" in sgm
    assert not any(k.startswith("vae_") for k in sgm)
    assert list(sgm_to_neurosis(sgm)) == list(ref_sd)  # inverse, order preserved
    torch.save({"state_dict": dict(sgm)}, str(tmp_path / "m.ckpt"))
    eng.init_from_ckpt(tmp_path / "m.ckpt")
    missing, unexpected = eng.last_ckpt_report
    assert not unexpected and not [m for m in missing if not m.startswith("denoiser.")], (missing, unexpected)
    out = reference_state_dict(eng)
    for k, v in ref_sd.items():
        assert torch.equal(out[k], v), k
    assert torch.equal(eng.first_stage_model.decoder.conv_out.weight, ref_sd["vae_decoder.conv_out.weight"])


def test_ema_shadow_keys_match_reference_naming():
    from neurosis_b200.modules.vae import Encoder
    from neurosis_b200.engine import DiffusionEngine
    eng = _engine(Encoder(**TINY_VAE, embed_dim=4, standalone=True))
    eng2 = DiffusionEngine(eng.model, eng.denoiser, None, eng.conditioner, eng.loss_fn, use_ema=True)
    keys = [k for k in eng2.state_dict() if k.startswith("model_ema.")]
    assert "model_ema.decay" in keys and "model_ema.num_updates" in keys
    assert "model_ema.diffusion_model_out_2_weight" in keys  # '.' removed, reference modules/ema.py:25-29
    assert len(keys) == 2 + len(unet_param_shapes(TINY_SDXL))


def _extract(path: Path, fn: str):
    """compile ONE function of a reference script (the scripts import typer, which is not needed for the function)."""
    tree = ast.parse(path.read_text())
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == fn)
    node.returns = None
    for a in node.args.args:
        a.annotation = None
    ns = {"OrderedDict": __import__("collections").OrderedDict, "torch": torch}
    exec(compile(ast.Module([node], []), str(path), "exec"), ns)
    return ns[fn]


@pytest.mark.skipif(not have_reference(), reason="needs /root/reference (authoring container)")
def test_key_rewrites_equal_the_reference_converters():
    from neurosis_b200.checkpoint import ldm_sd15_to_neurosis, neurosis_to_sgm
    conv = Path("/root/reference/scripts/convert")
    ref_sd = _reference_layout_sd()
    ours, theirs = neurosis_to_sgm(ref_sd), _extract(conv / "sdxl-neurosis2sgm.py", "maybe_remap_keys")(ref_sd)
    assert list(ours) == list(theirs)
    ldm = {"state_dict": {"cond_stage_model.transformer.text_model.embeddings.position_ids": torch.zeros(1),
                          "model.diffusion_model.out.2.weight": torch.zeros(1),
                          "first_stage_model.encoder.conv_in.weight": torch.zeros(1)}}
    assert list(ldm_sd15_to_neurosis(ldm)) == list(_extract(conv / "sd15-ldm2neurosis.py", "rename_keys")(ldm))


def test_configure_optimizers_param_groups():
    """reference models/diffusion.py:261-296: a "UNet" group, one group per trainable embedder, optional initial_lr."""
    from neurosis_b200.modules.vae import Encoder
    from neurosis_b200.optim import Adafactor, AdafactorScheduler
    eng = _engine(Encoder(**TINY_VAE, embed_dim=4, standalone=True))
    opt = eng.configure_optimizers()
    assert isinstance(opt, Adafactor) and [g["name"] for g in opt.param_groups] == ["UNet"]
    assert len(opt.param_groups[0]["params"]) == len(unet_param_shapes(TINY_SDXL))

    class Emb(torch.nn.Linear):
        is_trainable, base_lr, input_key, ucg_rate = True, 1e-6, "x", 0.0

    eng.conditioner.embedders.append(Emb(4, 4))
    eng.model.base_lr = 2e-6
    eng.scheduler = lambda o: AdafactorScheduler(o, initial_lr=4e-7)
    out = eng.configure_optimizers()
    groups = out["optimizer"].param_groups
    assert [g["name"] for g in groups] == ["UNet", "Emb"] and out["lr_scheduler"]["interval"] == "step"
    assert isinstance(out["lr_scheduler"]["scheduler"], AdafactorScheduler)
