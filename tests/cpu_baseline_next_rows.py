"""CPU baselines of the SURVEY.md §8(f) rows: the REFERENCE's own implementation (imported from /root/reference, so
this runs only in the authoring container) timed on the host cores, on a bounded sample, next to the device numbers
of tools/next_rows_bench.py.  A reported baseline, not a target.

    python tests/cpu_baseline_next_rows.py > profiles/r01_next_rows_cpu_reference.log
"""
import json
import os
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
from common import FULL_SDXL, import_reference  # noqa: E402

import_reference()
from neurosis.modules.diffusion.model import Decoder, Encoder  # noqa: E402
from neurosis.modules.ema import LitEma  # noqa: E402
from neurosis.modules.regularizers import DiagonalGaussianRegularizer  # noqa: E402
from neurosis.optimizers import Adafactor  # noqa: E402

from oracle.unet import unet_param_shapes  # noqa: E402

cores = os.cpu_count()
torch.set_num_threads(cores)
host = {"cores": cores, "torch_threads": torch.get_num_threads(), "kind": "reference",
        "where": "authoring container (the GPU box's host was not available when this was taken)"}

# ---- Adafactor / LitEma over a bounded sample of the SDXL UNet parameter set ------------------------------------
shapes = list(unet_param_shapes(FULL_SDXL).values())
sample, n = [], 0
for s in shapes[::7]:  # every 7th tensor: all layout classes (linear, conv3x3, conv1x1, bias/norm) in proportion
    sample.append(s)
    n += int(torch.tensor(s).prod())
g = torch.Generator().manual_seed(0)
params = [torch.nn.Parameter(torch.randn(s, generator=g) * 0.02) for s in sample]
for p in params:
    p.grad = torch.randn(p.shape, generator=g) * 1e-3
opt = Adafactor(params, scale_parameter=True, relative_step=True, warmup_init=True)
opt.step()
t0 = time.time()
for _ in range(2):
    opt.step()
sec = (time.time() - t0) / 2
full = 2567463684
print(json.dumps({"bench": "adafactor_step", "impl": "reference (neurosis.optimizers.Adafactor, CPU fp32)",
                  "sample": f"{len(sample)} of 1680 tensors, {n} elements", "sec_per_step_sample": sec,
                  "elements_per_s": n / sec, "sec_per_step_full_2.57G_extrapolated": sec * full / n, **host}))


class Holder(torch.nn.Module):
    def __init__(self, ps):
        super().__init__()
        self.ps = torch.nn.ParameterList(ps)


h = Holder(params)
ema = LitEma(h, decay=0.9999)
ema(h)
t0 = time.time()
for _ in range(3):
    ema(h)
sec = (time.time() - t0) / 3
print(json.dumps({"bench": "lit_ema_update", "impl": "reference (neurosis.modules.ema.LitEma, CPU fp32)",
                  "sample": f"{n} elements", "sec_per_update_sample": sec, "elements_per_s": n / sec,
                  "sec_per_update_full_2.57G_extrapolated": sec * full / n, **host}))

# ---- VAE training step (reference Encoder + Decoder, fp32) at 256^2, scaled by pixels to 1024^2 -------------------
V = dict(ch=128, out_ch=3, ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], in_channels=3, resolution=256,
         z_channels=4, double_z=True)
torch.manual_seed(0)
enc = Encoder(**V, embed_dim=4, standalone=True, attn_type="vanilla")
dec = Decoder(**V, embed_dim=4, standalone=True, attn_type="vanilla")
reg = DiagonalGaussianRegularizer(sample=True)
x = torch.rand(1, 3, 256, 256) * 2 - 1


def vae_step():
    enc.zero_grad()
    dec.zero_grad()
    z, _ = reg(enc(x))
    loss = torch.nn.functional.mse_loss(dec(z), x)
    loss.backward()


vae_step()
t0 = time.time()
vae_step()
sec = time.time() - t0
gflop = 46048.0 / 16  # 46 048 GFLOP per 1024^2 image; conv work scales with pixels
print(json.dumps({"bench": "vae_training_step", "impl": "reference (Encoder + Decoder + sampled posterior + L2, CPU fp32)",
                  "sample": "1 x 256x256 image", "sec_per_step_sample": sec, "GFLOPps": gflop / sec,
                  "sec_per_1024px_image_extrapolated_by_pixels": sec * 16, **host}))
