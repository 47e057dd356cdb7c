"""The reference's YAML `class_path` selection mechanism picks the B200 drop-ins."""
from pathlib import Path

import pytest
import yaml

from common import have_reference, import_reference
from neurosis_b200 import config

FRAGMENT = """
model:
  class_path: neurosis.models.diffusion.DiffusionEngine
  init_args:
    model:
      class_path: neurosis.modules.diffusion.UNetModel
      init_args: {adm_in_channels: 96, num_classes: sequential, use_checkpoint: true, in_channels: 4, out_channels: 4,
                  model_channels: 64, attention_resolutions: [2], num_res_blocks: 1, channel_mult: [1, 2],
                  num_head_channels: 64, use_linear_in_transformer: true, transformer_depth: [1, 2], context_dim: 64,
                  spatial_transformer_attn_type: softmax-xformers}
    denoiser:
      class_path: neurosis.modules.diffusion.DiscreteDenoiser
      init_args:
        num_idx: 1000
        preconditioning: {class_path: neurosis.modules.diffusion.EpsPreconditioning}
        discretization: {class_path: neurosis.modules.diffusion.LegacyDDPMDiscretization}
    loss_fn:
      class_path: neurosis.modules.diffusion.StandardDiffusionLoss
      init_args:
        loss_weighting: {class_path: neurosis.modules.diffusion.EpsWeighting}
        sigma_generator:
          class_path: neurosis.modules.diffusion.sigma_sampling.DiscreteSampling
          init_args:
            num_idx: 1000
            discretization: {class_path: neurosis.modules.diffusion.LegacyDDPMDiscretization}
"""


def test_yaml_fragment_selects_dropins():
    cfg = yaml.safe_load(FRAGMENT)["model"]["init_args"]
    unet = config.instantiate(cfg["model"])
    den = config.instantiate(cfg["denoiser"])
    loss = config.instantiate(cfg["loss_fn"])
    assert type(unet).__module__ == "neurosis_b200.modules.openaimodel"
    assert type(den).__module__ == "neurosis_b200.modules.denoiser" and den.sigmas.shape == (1001,)
    assert type(loss).__module__ == "neurosis_b200.modules.loss"
    assert type(loss.sigma_generator).__name__ == "DiscreteSigmaGenerator"
    assert len(unet.state_dict()) == 292


@pytest.mark.reference
@pytest.mark.skipif(not have_reference(), reason="needs /root/reference")
@pytest.mark.parametrize("path", ["/root/reference/configs/sdxl/sdxl.example.yaml", "/root/reference/configs/sd15/sd15.example.yml"])
def test_reference_yaml_unet_denoiser_loss_nodes_resolve(path):
    cfg = config.load_yaml(path)["model"]["init_args"]
    for key in ("model", "denoiser", "loss_fn", "first_stage_model"):
        node = cfg[key]
        cls = config.resolve(node["class_path"])
        assert cls.__module__.startswith("neurosis_b200."), (key, cls)
    den = config.instantiate(cfg["denoiser"])
    assert den.sigmas.shape == (1001,)
    loss = config.instantiate(cfg["loss_fn"])
    assert loss.loss_type == "l2"


@pytest.mark.skipif(not have_reference(), reason="needs /root/reference (authoring container)")
def test_constructor_and_forward_signatures_match_the_reference():
    """argument names, order and defaults of every mirrored class (`__init__` and `forward`) equal the reference's, so a
    YAML `init_args` block or a positional call written for the reference binds the same way.  Allowed differences:
    enum defaults given as their string values, `temb=None` default of the VAE ResnetBlock."""
    import importlib
    import inspect
    import_reference()
    pairs = [
        ("neurosis.modules.diffusion.openaimodel", "neurosis_b200.modules.openaimodel",
         ["UNetModel", "ResBlock", "Upsample", "Downsample", "Timestep", "TimestepEmbedSequential"]),
        ("neurosis.modules.attention", "neurosis_b200.modules.attention",
         ["SpatialTransformer", "BasicTransformerBlock", "CrossAttention", "MemoryEfficientCrossAttention",
          "TorchSDPCrossAttention", "FeedForward", "GEGLU"]),
        ("neurosis.modules.diffusion.model", "neurosis_b200.modules.vae",
         ["Encoder", "Decoder", "ResnetBlock", "AttnBlock", "Upsample", "Downsample"]),
        ("neurosis.modules.diffusion.denoiser", "neurosis_b200.modules.denoiser", ["Denoiser", "DiscreteDenoiser"]),
        ("neurosis.modules.diffusion.loss", "neurosis_b200.modules.loss", ["StandardDiffusionLoss", "DiffusionLoss"]),
        ("neurosis.modules.regularizers", "neurosis_b200.modules.vae", ["DiagonalGaussianRegularizer"]),
        ("neurosis.optimizers.adafactor", "neurosis_b200.optim", ["Adafactor", "AdafactorScheduler"]),
        ("neurosis.modules.ema", "neurosis_b200.optim", ["LitEma"]),
        ("neurosis.modules.encoders.embedding", "neurosis_b200.modules.conditioner",
         ["GeneralConditioner", "AbstractEmbModel"]),
        ("neurosis.modules.encoders.metadata", "neurosis_b200.modules.conditioner", ["ConcatTimestepEmbedderND"]),
    ]
    allowed = {("ResnetBlock", "forward", "temb"), ("StandardDiffusionLoss", "__init__", "loss_type"),
               ("StandardDiffusionLoss", "__init__", "objective_type")}
    problems = []
    for rmod, mmod, names in pairs:
        R, M = importlib.import_module(rmod), importlib.import_module(mmod)
        for n in names:
            for meth in ("__init__", "forward"):
                if not hasattr(getattr(R, n), meth):
                    continue
                rs = inspect.signature(getattr(getattr(R, n), meth)).parameters
                ms = inspect.signature(getattr(getattr(M, n), meth)).parameters
                named = lambda ps: [(k, v.default) for k, v in ps.items()  # noqa: E731
                                    if v.kind not in (v.VAR_POSITIONAL, v.VAR_KEYWORD)]
                mine = dict(named(ms))
                var_kw = any(v.kind == v.VAR_KEYWORD for v in ms.values())
                for k, d in named(rs):
                    if k not in mine:
                        if not var_kw:
                            problems.append((n, meth, k, "missing"))
                    elif repr(mine[k]) != repr(d) and (n, meth, k) not in allowed:
                        problems.append((n, meth, k, f"default {mine[k]!r} != {d!r}"))
                shared = [k for k, _ in named(rs) if k in mine]
                if shared != [k for k in mine if k in dict(named(rs))]:
                    problems.append((n, meth, "order"))
    assert not problems, problems
