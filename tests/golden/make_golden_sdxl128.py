"""Golden vectors AT THE BENCHMARKED SHAPE, generated FROM THE REFERENCE ITSELF (VERDICT r01 "missing" item 1).

Run only in the authoring container (needs /root/reference):  python tests/golden/make_golden_sdxl128.py
Writes tests/golden/reference_golden_sdxl128.npz.

The reference's own `UNetModel` (modules/diffusion/openaimodel.py:803-840) at the complete SDXL configuration
(configs/sdxl/sdxl.example.yaml:68-84; 2 567.5 M parameters, 1 680 tensors) on a 128x128 latent — the shape bench.py
times: 16 384-token 320-channel convolutions, 4 096-token / 10-head and 1 024-token / 20-head attention, 77-token cross
attention — batch 1, fp32 on the host CPU.  Weights = tests/common.fast_state_dict(seed=3) (the dict the GPU test
regenerates with the same torch CPU generator), inputs = oracle.weights.synth_tensor: only the reference's OUTPUTS are
stored (output tensor, per-parameter gradient L2 norms and sums, and a few complete gradients).

Also stored: a second bucket shape of BASELINE.json configs[3] (latent 144x112 = the 896x1152 aspect bucket: 4 032 /
1 008 tokens, neither a multiple of the 128-row attention tile) — forward output only, to bound the run time.
"""
import gc
import sys
import time
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
from common import FULL_SDXL, fast_state_dict, import_reference  # noqa: E402

import_reference()
from neurosis.modules.diffusion import UNetModel as RefUNet  # noqa: E402

from oracle.unet import unet_param_shapes  # noqa: E402
from oracle.weights import synth_tensor  # noqa: E402

torch.set_num_threads(8)
out = {}
cfg = FULL_SDXL
shapes = unet_param_shapes(cfg)
names = sorted(shapes)
sd = fast_state_dict(shapes, seed=3)
ref = RefUNet(**cfg)
ref.load_state_dict(sd)
del sd
gc.collect()

ctx = synth_tensor("full128.ctx", (1, 77, cfg["context_dim"]))
y = synth_tensor("full128.y", (1, cfg["adm_in_channels"]))
ts = torch.tensor([481])

# ---- 128 x 128 latent: forward + backward -------------------------------------------------------------------
t0 = time.time()
x = synth_tensor("full128.x", (1, 4, 128, 128))
g = synth_tensor("full128.g", (1, 4, 128, 128), scale=0.1)
o = ref(x, ts, ctx, y)
print(f"forward 128x128: {time.time() - t0:.1f} s, out norm {float(o.norm()):.5f}", flush=True)
(o * g).sum().backward()
print(f"forward + backward: {time.time() - t0:.1f} s", flush=True)
out["full128.out"] = o.detach().numpy()
out["full128.grad_l2"] = np.array([ref.get_parameter(n).grad.norm().item() for n in names], dtype=np.float64)
out["full128.grad_sum"] = np.array([ref.get_parameter(n).grad.double().sum().item() for n in names], dtype=np.float64)
# complete gradients of a few parameters along the depth of the network (element-wise comparison on the GPU side)
FULL_GRADS = ("input_blocks.0.0.weight", "input_blocks.1.0.in_layers.2.weight",
              "input_blocks.4.1.transformer_blocks.0.attn1.to_q.weight",
              "input_blocks.4.1.transformer_blocks.1.attn2.to_k.weight",
              "middle_block.1.transformer_blocks.9.ff.net.0.proj.bias",
              "middle_block.1.transformer_blocks.4.norm2.weight",
              "output_blocks.8.0.skip_connection.weight", "out.2.weight", "time_embed.0.weight",
              "label_emb.0.0.weight")
for n in FULL_GRADS:
    gr = ref.get_parameter(n).grad
    if gr.numel() > 200_000:  # keep the fixture small: leading 64 rows only
        gr = gr.reshape(gr.shape[0], -1)[:64]
    out["full128.grad." + n] = gr.detach().numpy().astype(np.float32)
del o
ref.zero_grad(set_to_none=True)
gc.collect()

# ---- 144 x 112 latent (896 x 1152 bucket): forward only ------------------------------------------------------
t0 = time.time()
with torch.no_grad():
    xb = synth_tensor("full144x112.x", (1, 4, 144, 112))
    ob = ref(xb, ts, ctx, y)
print(f"forward 144x112: {time.time() - t0:.1f} s, out norm {float(ob.norm()):.5f}", flush=True)
out["full144x112.out"] = ob.numpy()

np.savez_compressed(HERE / "reference_golden_sdxl128.npz", **out)
print("wrote", HERE / "reference_golden_sdxl128.npz", len(out), "arrays")
