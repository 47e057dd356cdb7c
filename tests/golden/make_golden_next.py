"""Golden vectors for the SURVEY.md §8(f) "next" rows, generated FROM THE REFERENCE ITSELF.

Run only in the authoring container (needs /root/reference):  python tests/golden/make_golden_next.py
Writes tests/golden/reference_golden_next.npz.  As in make_golden.py, inputs and weights come from oracle.weights
(numpy MT19937 keyed by tensor name), so only the reference's OUTPUTS are stored.

  row 1  VAE Decoder + VAE training step   neurosis.modules.diffusion.model.{Encoder, Decoder},
                                           neurosis.modules.regularizers.DiagonalGaussianRegularizer
  row 2  optimizer step                    neurosis.optimizers.Adafactor (optimizers/adafactor.py:13-256)
  row 3  conditioner vector path           neurosis.modules.encoders.metadata.ConcatTimestepEmbedderND
  row 4  EMA                               neurosis.modules.ema.LitEma (modules/ema.py:11-59)
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
from common import TINY_VAE, import_reference  # noqa: E402

D = import_reference()
from neurosis.modules.diffusion.model import Decoder, Encoder  # noqa: E402
from neurosis.modules.regularizers import DiagonalGaussianRegularizer  # noqa: E402

from oracle.vae import vae_decoder_param_shapes, vae_param_shapes  # noqa: E402
from oracle.weights import synth_state_dict, synth_tensor  # noqa: E402

torch.set_num_threads(4)
out = {}

# ---- row 1: decoder forward/backward, VAE training step ------------------------------------------------------
eshapes = vae_param_shapes(TINY_VAE, embed_dim=4, standalone=True)
dshapes = vae_decoder_param_shapes(TINY_VAE, embed_dim=4, standalone=True)
esd, dsd = synth_state_dict(eshapes, seed=2), synth_state_dict(dshapes, seed=5)
enc = Encoder(**TINY_VAE, embed_dim=4, standalone=True, attn_type="vanilla")
dec = Decoder(**TINY_VAE, embed_dim=4, standalone=True, attn_type="vanilla")
assert not (set(dec.state_dict()) ^ set(dshapes)), set(dec.state_dict()) ^ set(dshapes)
enc.load_state_dict(esd)
dec.load_state_dict(dsd)

z = synth_tensor("vaedec.z", (2, 4, 16, 16))
xr = dec(z)
g = synth_tensor("vaedec.g", tuple(xr.shape), scale=0.1)
(xr * g).sum().backward()
dnames = sorted(dshapes)
out["vaedec.out"] = xr.detach().numpy()
out["vaedec.grad_l2"] = np.array([dec.get_parameter(n).grad.norm().item() for n in dnames], dtype=np.float64)
out["vaedec.grad.conv_out.weight"] = dec.get_parameter("conv_out.weight").grad.numpy()
out["vaedec.grad.post_quant_conv.weight"] = dec.get_parameter("post_quant_conv.weight").grad.numpy()
dec.zero_grad()

img = synth_tensor("vae.img", (2, 3, 32, 32), uniform=True)
eps = synth_tensor("vaetrain.eps", (2, 4, 16, 16))
reg = DiagonalGaussianRegularizer(sample=True)
_orig = torch.randn
torch.randn = lambda *a, **k: eps.clone()  # DiagonalGaussianDistribution.sample draws torch.randn(mean.shape)
try:
    zs, log = reg(enc(img))
finally:
    torch.randn = _orig
xrec = dec(zs)
loss = torch.nn.functional.mse_loss(xrec, img)
loss.backward()
enames = sorted(eshapes)
out["vaetrain.z"] = zs.detach().numpy()
out["vaetrain.kl_loss"] = np.array(log["kl_loss"].item())
out["vaetrain.xrec"] = xrec.detach().numpy()
out["vaetrain.loss"] = np.array(loss.item())
out["vaetrain.enc_grad_l2"] = np.array([enc.get_parameter(n).grad.norm().item() for n in enames], dtype=np.float64)
out["vaetrain.dec_grad_l2"] = np.array([dec.get_parameter(n).grad.norm().item() for n in dnames], dtype=np.float64)
out["vaetrain.grad.quant_conv.weight"] = enc.get_parameter("quant_conv.weight").grad.numpy()
out["vaetrain.grad.conv_in.weight"] = enc.get_parameter("conv_in.weight").grad.numpy()

for extra in (HERE / "_golden_next_extra.py",      # rows 2-4
              HERE / "_golden_next_families.py",   # all shipped schedule / preconditioning / weighting variants
              HERE / "_golden_next_fullsize.py"):  # full-size SD1.5 / SDXL UNets through the reference
    if extra.exists():  # (separate files so each generator can be read on its own)
        exec(compile(extra.read_text(), str(extra), "exec"), {"out": out, "HERE": HERE, "np": np, "torch": torch,
                                                               "synth_tensor": synth_tensor,
                                                               "synth_state_dict": synth_state_dict})

np.savez_compressed(HERE / "reference_golden_next.npz", **out)
print("wrote", HERE / "reference_golden_next.npz", len(out), "arrays")
