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
