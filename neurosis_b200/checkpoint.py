"""Checkpoint key round trip: reference checkpoints load into the drop-in modules and back.

Mirrors, without Lightning / typer:
  * `DiffusionEngine.init_from_ckpt` (/root/reference/src/neurosis/models/diffusion.py:127-144): `.safetensors` or
    `.ckpt["state_dict"]`, `strict=False`, relocated `first_stage_model.*` keys tolerated as unexpected, `vae_*` and
    `._orig_mod.` keys tolerated as missing;
  * `scripts/convert/sdxl-neurosis2sgm.py:79-98` (`maybe_remap_keys`: `vae_encoder.quant_conv.* ->
    first_stage_model.quant_conv.*`, `vae_decoder.post_quant_conv.* -> first_stage_model.post_quant_conv.*`,
    `vae_{encoder,decoder}.* -> first_stage_model.{encoder,decoder}.*`) and its inverse;
  * `scripts/convert/sd15-ldm2neurosis.py:21-31` (`cond_stage_model.* -> conditioner.embedders.0.*`).

Reference engine layout (models/diffusion.py:73, 80-83, 146-159): `model.diffusion_model.<unet keys>`,
`vae_encoder.<encoder keys incl. quant_conv>`, `vae_decoder.<decoder keys incl. post_quant_conv>`,
`conditioner.embedders.N.*`, `model_ema.*`.  `neurosis_b200.engine.DiffusionEngine` holds the first stage as
`first_stage_model` (an `Encoder(standalone=True)` or an `AutoencoderKL`); `engine_key_map` translates.
Pure host code: dict-key rewriting only, no tensor arithmetic.
"""
from __future__ import annotations

from collections import OrderedDict
from pathlib import Path
from typing import Mapping

import torch
from torch import Tensor, nn

CHECKPOINT_EXTNS = (".ckpt", ".pt", ".pth")


def neurosis_to_sgm(state_dict: Mapping[str, Tensor]) -> "OrderedDict[str, Tensor]":
    """`maybe_remap_keys` of scripts/convert/sdxl-neurosis2sgm.py:79-98."""
    out = OrderedDict()
    for k, v in state_dict.items():
        if k.startswith("vae_decoder.post_quant_conv."):
            k = k.replace("vae_decoder.", "first_stage_model.")
        if k.startswith("vae_encoder.quant_conv."):
            k = k.replace("vae_encoder.", "first_stage_model.")
        if k.startswith("vae_"):
            k = k.replace("vae_", "first_stage_model.")
        out[k] = v
    return out


def sgm_to_neurosis(state_dict: Mapping[str, Tensor]) -> "OrderedDict[str, Tensor]":
    """inverse of `neurosis_to_sgm`: an SGM / LDM `first_stage_model.*` layout -> the engine's `vae_*` keys."""
    out = OrderedDict()
    for k, v in state_dict.items():
        if k.startswith("first_stage_model.quant_conv."):
            k = k.replace("first_stage_model.", "vae_encoder.", 1)
        elif k.startswith("first_stage_model.post_quant_conv."):
            k = k.replace("first_stage_model.", "vae_decoder.", 1)
        elif k.startswith("first_stage_model.encoder.") or k.startswith("first_stage_model.decoder."):
            k = k.replace("first_stage_model.", "vae_", 1)
        out[k] = v
    return out


def ldm_sd15_to_neurosis(state_dict: Mapping) -> dict:
    """`rename_keys` of scripts/convert/sd15-ldm2neurosis.py:21-31."""
    if "state_dict" in state_dict:
        state_dict = state_dict["state_dict"]
    return {(k.replace("cond_stage_model.", "conditioner.embedders.0.", 1) if "cond_stage_model." in k else k): v
            for k, v in state_dict.items()}


def load_state_dict_file(path) -> Mapping[str, Tensor]:
    path = Path(path)
    if path.suffix == ".safetensors":
        from safetensors.torch import load_file
        return load_file(str(path), device="cpu")
    if path.suffix in CHECKPOINT_EXTNS:
        return torch.load(str(path), map_location="cpu")["state_dict"]
    raise NotImplementedError(f"Unknown checkpoint extension {path.suffix}")


def engine_key_map(engine: nn.Module) -> dict[str, str]:
    """{reference engine key: key in `engine.state_dict()`} for every tensor the two layouts share."""
    fsm = getattr(engine, "first_stage_model", None)
    kl = fsm is not None and hasattr(fsm, "encoder")  # AutoencoderKL vs a bare Encoder(standalone=True)
    out = {}
    for k in engine.state_dict():
        r = k
        if k.startswith("first_stage_model."):
            rest = k[len("first_stage_model."):]
            if not kl:
                r = "vae_encoder." + rest
            elif rest.startswith("quant_conv."):
                r = "vae_encoder." + rest
            elif rest.startswith("post_quant_conv."):
                r = "vae_decoder." + rest
            elif rest.startswith("encoder."):
                r = "vae_encoder." + rest[len("encoder."):]
            elif rest.startswith("decoder."):
                r = "vae_decoder." + rest[len("decoder."):]
        out[r] = k
    return out


def reference_state_dict(engine: nn.Module) -> "OrderedDict[str, Tensor]":
    """the engine's tensors under the reference's key names (what a Lightning checkpoint of the reference holds)."""
    sd = engine.state_dict()
    return OrderedDict((r, sd[k]) for r, k in engine_key_map(engine).items())


def init_from_ckpt(engine: nn.Module, path) -> tuple[list, list]:
    """`DiffusionEngine.init_from_ckpt` (models/diffusion.py:127-144) for the drop-in engine; accepts the reference's
    own layout and the SGM `first_stage_model.*` layout.  Returns the filtered (missing, unexpected) key lists."""
    sd = sgm_to_neurosis(load_state_dict_file(path))
    kmap = engine_key_map(engine)
    local = OrderedDict()
    unexpected = []
    for k, v in sd.items():
        if k in kmap:
            local[kmap[k]] = v
        else:
            unexpected.append(k)
    missing, extra = engine.load_state_dict(local, strict=False)
    inv = {v: k for k, v in kmap.items()}
    missing = [inv.get(x, x) for x in missing]
    unexpected = [x for x in unexpected + list(extra) if not x.startswith("first_stage_model")]
    missing = [x for x in missing if (not x.startswith("vae_")) and "._orig_mod." not in x]
    from . import ops
    ops.invalidate_weight_cache()  # load_state_dict copies in place; the kernels' bf16 / packed copies are stale
    return missing, unexpected
