"""Generates the golden vectors in this directory FROM THE REFERENCE ITSELF.

Run only in the authoring container (needs /root/reference):  python tests/golden/make_golden.py
Inputs and weights come from oracle.weights (numpy MT19937, keyed by tensor name) so the tests can
regenerate them bit-identically without storing them; only the reference's OUTPUTS are stored.
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
from common import TINY_SD15, TINY_SDXL, TINY_VAE, import_reference  # noqa: E402

D = import_reference()
from neurosis.modules.diffusion import (DiscreteDenoiser, DiscreteSigmaGenerator, EpsPreconditioning,  # noqa: E402
                                        EpsWeighting, LegacyDDPMDiscretization, OpenAIWrapper,
                                        StandardDiffusionLoss, UNetModel)
from neurosis.modules.diffusion.model import Encoder  # noqa: E402

from oracle.unet import unet_param_shapes  # noqa: E402
from oracle.vae import vae_param_shapes  # noqa: E402
from oracle.weights import synth_state_dict, synth_tensor  # noqa: E402

torch.set_num_threads(4)
out = {}

# ---- sigma tables and index sampling (bit-exact targets) -------------------------------------------
disc = LegacyDDPMDiscretization()
out["table_desc"] = disc(1000, flip=False).detach().numpy()
out["table_asc"] = disc(1000, flip=True).detach().numpy()
den = DiscreteDenoiser(EpsPreconditioning(), 1000, LegacyDDPMDiscretization())
den.sigmas, den.log_sigmas = den.sigmas.detach(), den.log_sigmas.detach()
sig_in = (synth_tensor("sigma_probe", (64,), uniform=True).abs() * 15.0).float()
out["sigma_probe_idx_f32"] = den.sigma_to_idx(sig_in).numpy()
out["sigma_probe_idx_bf16"] = den.sigma_to_idx(sig_in.to(torch.bfloat16)).numpy()
out["sigma_probe_quant"] = den.possibly_quantize_sigma(sig_in).numpy()
gen = DiscreteSigmaGenerator(LegacyDDPMDiscretization(), 1000)
gen.sigmas = gen.sigmas.detach()
torch.manual_seed(42)
t = torch.rand((8,), dtype=torch.float64)
out["gen_t"] = t.numpy()
out["gen_sigma_from_t"] = gen(8, t).numpy()           # always table[0] = 0.0 (reference quirk)
torch.manual_seed(42)
out["gen_sigma_randint"] = gen(8, None).numpy()       # the SGM-style branch


# ---- UNet forward / backward ------------------------------------------------------------------------
def unet_case(tag, cfg, hw):
    shapes = unet_param_shapes(cfg)
    sd = synth_state_dict(shapes, seed=1)
    ref = UNetModel(**cfg)
    ref.load_state_dict(sd)
    B = 2
    x = synth_tensor(f"{tag}.x", (B, 4, hw, hw))
    ts = torch.tensor([17, 803])
    ctx = synth_tensor(f"{tag}.ctx", (B, 77, cfg["context_dim"]))
    y = synth_tensor(f"{tag}.y", (B, cfg["adm_in_channels"])) if cfg.get("num_classes") else None
    o = ref(x, ts, ctx, y)
    go = synth_tensor(f"{tag}.gout", tuple(o.shape), scale=0.1)
    (o * go).sum().backward()
    out[f"{tag}.out"] = o.detach().numpy()
    names = sorted(shapes)
    out[f"{tag}.grad_l2"] = np.array([ref.get_parameter(n).grad.norm().item() for n in names], dtype=np.float64)
    out[f"{tag}.grad_sum"] = np.array([ref.get_parameter(n).grad.double().sum().item() for n in names], dtype=np.float64)
    for n in ("out.2.weight", "input_blocks.0.0.bias", "time_embed.0.bias"):
        out[f"{tag}.grad.{n}"] = ref.get_parameter(n).grad.numpy()
    return ref, sd


unet_case("sdxl", TINY_SDXL, 16)
unet_case("sd15", TINY_SD15, 16)

# ---- VAE encoder ---------------------------------------------------------------------------------------
vshapes = vae_param_shapes(TINY_VAE, embed_dim=4, standalone=True)
vsd = synth_state_dict(vshapes, seed=2)
enc = Encoder(**TINY_VAE, embed_dim=4, standalone=True, attn_type="vanilla")
missing = set(enc.state_dict().keys()) ^ set(vshapes)
assert not missing, missing
enc.load_state_dict(vsd)
img = synth_tensor("vae.img", (2, 3, 32, 32), uniform=True)
with torch.no_grad():
    out["vae.z"] = enc(img, regularize=True).numpy()

# ---- full loss (fixed sigma draw and noise) -----------------------------------------------------------
cfg = TINY_SDXL
sd = synth_state_dict(unet_param_shapes(cfg), seed=1)
ref = UNetModel(**cfg)
ref.load_state_dict(sd)


class FixedSigma:
    def __init__(self, s):
        self.s = s

    def __call__(self, n, t=None):
        return self.s


class Cond(torch.nn.Module):
    def forward(self, batch):
        return {"crossattn": batch["ctx"], "vector": batch["vec"]}


lat = synth_tensor("step.latent", (2, 4, 16, 16))
noise = synth_tensor("step.noise", (2, 4, 16, 16))
sig = gen.sigmas[torch.tensor([200, 700])].clone()
loss_fn = StandardDiffusionLoss(sigma_generator=FixedSigma(sig), loss_weighting=EpsWeighting())
_orig = torch.randn_like
torch.randn_like = lambda t_, **kw: noise.to(t_)
try:
    batch = {"ctx": synth_tensor("sdxl.ctx", (2, 77, cfg["context_dim"])),
             "vec": synth_tensor("sdxl.y", (2, cfg["adm_in_channels"]))}
    loss = loss_fn(OpenAIWrapper(ref), den, Cond(), lat, batch)
finally:
    torch.randn_like = _orig
loss.mean().backward()
out["step.sigmas"] = sig.numpy()
out["step.loss"] = loss.detach().numpy()
out["step.grad.out.2.weight"] = ref.get_parameter("out.2.weight").grad.numpy()
out["step.grad_l2"] = np.array([ref.get_parameter(n).grad.norm().item() for n in sorted(sd)], dtype=np.float64)

np.savez_compressed(HERE / "reference_golden.npz", **out)
print("wrote", HERE / "reference_golden.npz", {k: v.shape for k, v in out.items()})
