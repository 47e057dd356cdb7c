# Full-size UNets of the reference's example YAMLs (configs/sd15, configs/sdxl) run through the REFERENCE's UNetModel
# at a small latent (executed by make_golden_next.py with `out`, `np`, `torch`, `synth_tensor` in scope): pins the
# oracle at the real depth / width (SDXL: 2.57 G parameters, transformer depth 10, 1 680 tensors), not only on the
# miniature configs.  Weights: tests/common.fast_state_dict(seed=3) — the same dict the GPU full-size tests load.
import gc

from common import FULL_SD15, FULL_SDXL, fast_state_dict
from neurosis.modules.diffusion import UNetModel as _RefUNet

from oracle.unet import unet_param_shapes as _shapes

for tag, cfg in (("sd15", FULL_SD15), ("sdxl", FULL_SDXL)):
    shapes = _shapes(cfg)
    sd = fast_state_dict(shapes, seed=3)
    ref = _RefUNet(**cfg)
    ref.load_state_dict(sd)
    del sd
    gc.collect()
    x = synth_tensor(f"full.{tag}.x", (1, 4, 32, 32))
    ctx = synth_tensor(f"full.{tag}.ctx", (1, 77, cfg["context_dim"]))
    y = synth_tensor(f"full.{tag}.y", (1, cfg["adm_in_channels"])) if cfg.get("num_classes") else None
    o = ref(x, torch.tensor([481]), ctx, y)
    (o * synth_tensor(f"full.{tag}.g", (1, 4, 32, 32), scale=0.1)).sum().backward()
    names = sorted(shapes)
    out[f"full.{tag}.out"] = o.detach().numpy()
    out[f"full.{tag}.grad_l2"] = np.array([ref.get_parameter(n).grad.norm().item() for n in names], dtype=np.float64)
    out[f"full.{tag}.grad_sum"] = np.array([ref.get_parameter(n).grad.double().sum().item() for n in names],
                                           dtype=np.float64)
    print("full-size", tag, len(names), "tensors, out norm", float(o.norm()))
    del ref, o
    gc.collect()

# ---- the full SDXL KL-f8 VAE (configs/sdxl/sdxl.example.yaml:102-113: ch 128, mult [1,2,4,4], 2 res blocks) ------
from common import FULL_VAE
from neurosis.modules.diffusion.model import Decoder as _RefDec
from neurosis.modules.diffusion.model import Encoder as _RefEnc

from oracle.vae import vae_decoder_param_shapes as _dshapes
from oracle.vae import vae_param_shapes as _eshapes

_es, _ds = _eshapes(FULL_VAE, 4, True), _dshapes(FULL_VAE, 4, True)
_enc = _RefEnc(**FULL_VAE, embed_dim=4, standalone=True, attn_type="vanilla")
_dec = _RefDec(**FULL_VAE, embed_dim=4, standalone=True, attn_type="vanilla")
_enc.load_state_dict(fast_state_dict(_es, seed=11))
_dec.load_state_dict(fast_state_dict(_ds, seed=12))
img = synth_tensor("fullvae.img", (1, 3, 64, 64), uniform=True)
eps = synth_tensor("fullvae.eps", (1, 4, 8, 8))
m = _enc(img)  # moments (1, 8, 8, 8)
mean, logvar = torch.chunk(m, 2, dim=1)
z = mean + torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0)) * eps
xrec = _dec(z)
loss = torch.nn.functional.mse_loss(xrec, img)
loss.backward()
out["fullvae.moments"] = m.detach().numpy()
out["fullvae.xrec"] = xrec.detach().numpy()
out["fullvae.loss"] = np.array(loss.item())
out["fullvae.enc_grad_l2"] = np.array([_enc.get_parameter(n).grad.norm().item() for n in sorted(_es)], dtype=np.float64)
out["fullvae.dec_grad_l2"] = np.array([_dec.get_parameter(n).grad.norm().item() for n in sorted(_ds)], dtype=np.float64)
print("full-size VAE", len(_es), "+", len(_ds), "tensors, loss", loss.item())
