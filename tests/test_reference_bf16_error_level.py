"""What error does the UNMODIFIED reference make when it runs in bf16?  (CPU, seconds.)

BASELINE.json's north-star asks for "per-module fp32 outputs and gradients within 1e-3 relative".  The CUDA path computes
its contractions in bf16 (as the reference's `precision: bf16-mixed` does), and the GPU tests judge modules that contain
a bf16 GEMM at 3e-2 (output) / 4e-2 median / 1e-1 worst (gradients) against the fp32 oracle.  This test pins the claim
that this is the level of the reference's OWN bf16 path: the reference UNetModel (baseline/_ref) under
`torch.autocast(bfloat16)` against the fp32 oracle on the same miniature, weights and inputs the GPU tests use.  Measured
here: output 1.5 - 1.7e-2, median gradient 2.6 - 2.7e-2, worst gradient 4.5e-2 — fifteen to forty times the 1e-3 figure,
and the same level as the CUDA path (`tests/test_zzz_gpu_unverified_variants.py` repeats the comparison on the GPU with
cuBLAS / cuDNN / SDPA under CUDA autocast, next to our kernels)."""
import sys

import numpy as np
import pytest
import torch

from common import ROOT, TINY_SD15, TINY_SDXL
from oracle.unet import unet_forward, unet_param_shapes
from oracle.weights import synth_state_dict, synth_tensor

sys.path.insert(0, str(ROOT / "tools"))
import ref_harness as RH  # noqa: E402


def _rel(a, b) -> float:
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


@pytest.mark.skipif(not RH.available(), reason="baseline/_ref missing (built by __graft_entry__.build() where /root/reference exists)")
@pytest.mark.parametrize("tag,cfg", [("sdxl", TINY_SDXL), ("sd15", TINY_SD15)])
def test_reference_under_bf16_autocast_sits_at_the_tolerance_the_gpu_tests_use(tag, cfg):
    RH._import()
    from neurosis.modules.diffusion import UNetModel as RefUNet
    shapes = unet_param_shapes(cfg)
    m = RefUNet(**cfg)
    m.load_state_dict(synth_state_dict(shapes, seed=1))
    x = synth_tensor(f"{tag}.x", (2, 4, 16, 16))
    ctx = synth_tensor(f"{tag}.ctx", (2, 77, cfg["context_dim"]))
    y = synth_tensor(f"{tag}.y", (2, cfg["adm_in_channels"])) if cfg.get("num_classes") else None
    gout = synth_tensor(f"{tag}.gout", (2, 4, 16, 16), scale=0.1)
    ts = torch.tensor([17, 803])
    with torch.autocast("cpu", dtype=torch.bfloat16):
        out = m(x, ts, ctx, y)
    (out.float() * gout).sum().backward()
    sd = {k: v.requires_grad_(True) for k, v in synth_state_dict(shapes, seed=1).items()}
    o = unet_forward(sd, cfg, x, ts, ctx, y)
    (o * gout).sum().backward()
    errs = np.array([_rel(p.grad, sd[n].grad) for n, p in m.named_parameters()])
    e_out, e_med, e_max = _rel(out, o), float(np.median(errs)), float(errs.max())
    print(f"[reference bf16 autocast vs fp32 oracle, {tag}] output {e_out:.3e}  grad median {e_med:.3e}  worst {e_max:.3e}")
    # far above the north-star's 1e-3 ...
    assert e_out > 5e-3 and e_med > 1e-2
    # ... and inside the bounds tests/test_gpu_modules.py applies to the CUDA path (3e-2 / 4e-2 / 1e-1)
    assert e_out < 3e-2 and e_med < 4e-2 and e_max < 1e-1
