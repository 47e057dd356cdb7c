"""`class_path` / `init_args` instantiation of the reference's LightningCLI YAML configs.

The reference selects every component through jsonargparse (`trainer/cli.py:131-136`,
`subclass_mode_model=True`): a YAML node {class_path: pkg.Class, init_args: {...}} is constructed recursively.
jsonargparse / omegaconf / lightning are not part of this image, so this ~80-line loader provides the same
selection mechanism for the hot path: class paths under `neurosis.*` that belong to the training step are
redirected to their drop-in implementations in `neurosis_b200`, which is how a `configs/sdxl` or `configs/sd15`
YAML "selects the new kernels" without being edited.  Paths the reference itself renamed/removed
(e.g. `...sigma_sampling.DiscreteSampling`, configs/sdxl/sdxl.example.yaml:177) are aliased as well.
"""
from __future__ import annotations

import importlib
from typing import Any

import yaml

_M = "neurosis_b200.modules"
REDIRECT = {
    "neurosis.modules.diffusion.UNetModel": f"{_M}.openaimodel.UNetModel",
    "neurosis.modules.diffusion.openaimodel.UNetModel": f"{_M}.openaimodel.UNetModel",
    "neurosis.modules.diffusion.model.Encoder": f"{_M}.vae.Encoder",
    "neurosis.modules.diffusion.model.Decoder": f"{_M}.vae.Decoder",
    "neurosis.models.autoencoder.AutoencoderKL": f"{_M}.vae.AutoencoderKL",
    "neurosis.models.autoencoder.AutoencodingEngineLegacy": f"{_M}.vae.AutoencoderKL",
    "neurosis.modules.regularizers.DiagonalGaussianRegularizer": f"{_M}.vae.DiagonalGaussianRegularizer",
    "neurosis.modules.diffusion.wrappers.OpenAIWrapper": f"{_M}.loss.OpenAIWrapper",
    "neurosis.modules.diffusion.hooks.LossHook": f"{_M}.loss.LossHook",
    "neurosis.dataset.processing.TagFrequencyHook": f"{_M}.loss.TagFrequencyHook",
    "neurosis.dataset.processing.TagFreqScale": f"{_M}.loss.TagFreqScale",
    "neurosis.dataset.processing.TagRewards": f"{_M}.loss.TagRewards",
    "neurosis.modules.encoders.GeneralConditioner": f"{_M}.conditioner.GeneralConditioner",
    "neurosis.modules.encoders.IdentityEncoder": f"{_M}.conditioner.IdentityEncoder",
    "neurosis.modules.encoders.metadata.ConcatTimestepEmbedderND": f"{_M}.conditioner.ConcatTimestepEmbedderND",
    "neurosis.models.diffusion.DiffusionEngine": "neurosis_b200.engine.DiffusionEngine",
    # optimizer side: configs/sdxl/sdxl.example.yaml:158-169
    "neurosis.optimizers.Adafactor": "neurosis_b200.optim.Adafactor",
    "neurosis.optimizers.AdafactorScheduler": "neurosis_b200.optim.AdafactorScheduler",
    "neurosis.optimizers.adafactor.Adafactor": "neurosis_b200.optim.Adafactor",
    "neurosis.optimizers.adafactor.AdafactorScheduler": "neurosis_b200.optim.AdafactorScheduler",
    "neurosis.modules.ema.LitEma": "neurosis_b200.optim.LitEma",
    # renamed in the reference tree but still present in its example YAMLs
    "neurosis.modules.diffusion.sigma_sampling.DiscreteSampling": f"{_M}.schedule.DiscreteSigmaGenerator",
    "neurosis.modules.diffusion.sigma_sampling.EDMSampling": f"{_M}.schedule.EDMSigmaGenerator",
}
_BY_NAME = {
    f"{_M}.schedule": ["LegacyDDPMDiscretization", "EDMDiscretization", "EDMcDiscretization", "EDMcSimpleDiscretization",
                       "TanZeroSNRDiscretization", "RectifiedFlowDiscretization", "RectifiedFlowComfyDiscretization",
                       "DiscreteSigmaGenerator", "EDMSigmaGenerator", "CosineScheduleSigmaGenerator",
                       "TanScheduleSigmaGenerator", "RectifiedFlowSigmaGenerator", "RectifiedFlowComfySigmaGenerator"],
    f"{_M}.denoiser": ["Denoiser", "DiscreteDenoiser", "EpsPreconditioning", "VPreconditioning",
                       "VPreconditioningWithEDMcNoise", "EDMPreconditioning", "RectifiedFlowXLPreconditioning",
                       "RectifiedFlowComfyPreconditioning", "UnitWeighting", "EpsWeighting", "EDMWeighting",
                       "RectifiedFlowWeighting", "RectifiedFlowComfyWeighting", "MinSNRGammaModifier"],
    f"{_M}.loss": ["StandardDiffusionLoss", "DiffusionLoss", "IdentityWrapper", "OpenAIWrapper"],
}
for _mod, _names in _BY_NAME.items():
    for _n in _names:
        for _prefix in ("neurosis.modules.diffusion", "neurosis.modules.diffusion.discretization",
                        "neurosis.modules.diffusion.denoiser", "neurosis.modules.diffusion.denoiser_preconditioning",
                        "neurosis.modules.diffusion.denoiser_weighting", "neurosis.modules.diffusion.loss",
                        "neurosis.modules.diffusion.sampling", "neurosis.modules.diffusion.sampling.sigma_generators"):
            REDIRECT.setdefault(f"{_prefix}.{_n}", f"{_mod}.{_n}")


def resolve(class_path: str):
    path = REDIRECT.get(class_path, class_path)
    mod, _, name = path.rpartition(".")
    return getattr(importlib.import_module(mod), name)


def instantiate(node: Any, **overrides) -> Any:
    """recursively build {class_path, init_args} nodes; other values are returned as they are."""
    if isinstance(node, dict):
        if "class_path" in node:
            kwargs = {k: instantiate(v) for k, v in (node.get("init_args") or {}).items()}
            kwargs.update(overrides)
            return resolve(node["class_path"])(**kwargs)
        return {k: instantiate(v) for k, v in node.items()}
    if isinstance(node, list):
        return [instantiate(v) for v in node]
    return node


def load_yaml(path: str) -> dict:
    with open(path) as f:
        return yaml.safe_load(f)


def unet_from_config(path: str, **overrides):
    """build the UNet selected by `model.init_args.model` of a reference training YAML."""
    cfg = load_yaml(path)
    node = cfg["model"]["init_args"]["model"]
    return instantiate(node, **overrides)
