"""Shared test configurations (small enough for the CPU oracle to finish in seconds)."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

REFERENCE_SRC = Path("/root/reference/src")

# SDXL-shaped miniature: linear proj, head_dim 64, y vector, depth 2 at the lower level
TINY_SDXL = dict(in_channels=4, model_channels=64, out_channels=4, num_res_blocks=1, attention_resolutions=[2],
                 channel_mult=[1, 2], num_head_channels=64, transformer_depth=[1, 2], context_dim=64,
                 use_linear_in_transformer=True, num_classes="sequential", adm_in_channels=96,
                 spatial_transformer_attn_type="torch-sdp", use_checkpoint=False)
# SD1.5-shaped miniature: conv1x1 proj, 2 heads (head_dim 32 / 64), no y
TINY_SD15 = dict(in_channels=4, model_channels=64, out_channels=4, num_res_blocks=1, attention_resolutions=[1, 2],
                 channel_mult=[1, 2], num_heads=2, transformer_depth=1, context_dim=48,
                 use_linear_in_transformer=False, spatial_transformer_attn_type="torch-sdp", use_checkpoint=False)
TINY_VAE = dict(ch=64, out_ch=3, ch_mult=[1, 2], num_res_blocks=1, attn_resolutions=[], in_channels=3, resolution=32,
                z_channels=4, double_z=True)


# full-size UNets of the reference's example YAMLs (configs/sdxl/sdxl.example.yaml:68-84, configs/sd15/sd15.example.yml:68-81)
FULL_SDXL = dict(in_channels=4, model_channels=320, out_channels=4, num_res_blocks=2, attention_resolutions=[4, 2],
                 channel_mult=[1, 2, 4], num_head_channels=64, transformer_depth=[1, 2, 10], context_dim=2048,
                 use_linear_in_transformer=True, num_classes="sequential", adm_in_channels=2816,
                 spatial_transformer_attn_type="torch-sdp", use_checkpoint=False)
FULL_SD15 = dict(in_channels=4, model_channels=320, out_channels=4, num_res_blocks=2, attention_resolutions=[4, 2, 1],
                 channel_mult=[1, 2, 4, 4], num_heads=8, transformer_depth=1, context_dim=768,
                 use_linear_in_transformer=False, spatial_transformer_attn_type="torch-sdp", use_checkpoint=False)


# the KL-f8 VAE of configs/sdxl/sdxl.example.yaml:102-113
FULL_VAE = dict(ch=128, out_ch=3, ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], in_channels=3,
                resolution=256, z_channels=4, double_z=True)


def fast_state_dict(shapes: dict, seed: int = 0) -> dict:
    """same distributions as oracle.weights.synth_state_dict, drawn with torch's generator (seconds instead of minutes for
    the 2.57 B-parameter SDXL UNet); only for tests where both sides load the SAME dict (no golden involved)."""
    import torch
    g = torch.Generator().manual_seed(1234 + seed)
    sd = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        n = torch.randn(shape, generator=g)
        if name.endswith(".bias"):
            sd[name] = n * 0.05
        elif len(shape) == 1:
            sd[name] = 1.0 + n * 0.1
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            sd[name] = n * fan_in ** -0.5
    return sd


def have_reference() -> bool:
    return REFERENCE_SRC.exists()


def import_reference():
    """import the reference modules (works only where /root/reference exists)."""
    if str(REFERENCE_SRC) not in sys.path:
        sys.path.insert(0, str(REFERENCE_SRC))
    import neurosis.modules.diffusion as D  # noqa: F401  (must be first: circular import in the reference)
    return D
