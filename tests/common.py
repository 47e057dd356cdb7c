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


def have_reference() -> bool:
    return REFERENCE_SRC.exists()


def import_reference():
    """import the reference modules (works only where /root/reference exists)."""
    if str(REFERENCE_SRC) not in sys.path:
        sys.path.insert(0, str(REFERENCE_SRC))
    import neurosis.modules.diffusion as D  # noqa: F401  (must be first: circular import in the reference)
    return D
