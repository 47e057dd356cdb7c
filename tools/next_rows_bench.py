"""Device timings of the SURVEY.md §8(f) rows on one B200 (CUDA events, warm-up first):
  * fused Adafactor step and LitEma update over the full SDXL UNet parameter set (2.57 G elements, 1 680 tensors)
    -> achieved GB/s of algorithmic traffic against the measured HBM peak (MEASURED_PEAKS.json)
  * VAE training step (AutoencoderKL encoder + decoder forward/backward, L2 loss) at 1024^2, batch 1
Usage: python tools/next_rows_bench.py [--skip-vae] ; prints one JSON line per measurement."""
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

SDXL_VAE = dict(ch=128, out_ch=3, ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], in_channels=3,
                resolution=256, z_channels=4, double_z=True)


def hbm_peak() -> float:
    try:
        d = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        for k in ("hbm_gbs", "hbm_gbps"):
            if k in d:
                return float(d[k])
        flat = json.dumps(d)
        import re
        m = re.search(r'"[^"]*hbm[^"]*":\s*([0-9.]+)', flat)
        if m:
            return float(m.group(1))
    except Exception:
        pass
    return 6454.0


def timed(fn, warm=2, iters=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def optimizer_bench():
    from common import FULL_SDXL
    from neurosis_b200.optim import Adafactor, LitEma
    from neurosis_b200.modules import UNetModel
    dev = "cuda"
    with torch.device("meta"):
        meta = UNetModel(**FULL_SDXL)
    shapes = [tuple(p.shape) for p in meta.parameters()]
    g = torch.Generator(device=dev).manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(s, device=dev, generator=g) * 0.02) for s in shapes]
    for p in params:
        p.grad = torch.randn(p.shape, device=dev, generator=g) * 1e-3
    n = sum(p.numel() for p in params)
    opt = Adafactor(params, scale_parameter=True, relative_step=True, warmup_init=True)
    t0 = time.time()
    opt.step()
    torch.cuda.synchronize()
    build_s = time.time() - t0
    ms = timed(opt.step)
    peak = hbm_peak()
    bytes_alg = 24.0 * n  # pass 1: p, g; pass 2: g; pass 3: g, p -> p (no bf16 mirrors registered in this bench)
    print(json.dumps({"bench": "adafactor_step", "tensors": len(params), "elements": n, "ms": ms,
                      "algorithmic_GB": bytes_alg / 1e9, "GBps": bytes_alg / ms / 1e6, "hbm_peak_GBps": peak,
                      "frac": bytes_alg / ms / 1e6 / peak, "first_step_incl_table_build_s": build_s,
                      "launches_per_step": 4, "finite": bool(torch.isfinite(params[0]).all())}))

    class Holder(torch.nn.Module):
        def __init__(self, ps):
            super().__init__()
            self.ps = torch.nn.ParameterList(ps)

    h = Holder(params)
    ema = LitEma(h, decay=0.9999)
    ms = timed(lambda: ema(h))
    print(json.dumps({"bench": "lit_ema_update", "elements": n, "ms": ms, "algorithmic_GB": 12.0 * n / 1e9,
                      "GBps": 12.0 * n / ms / 1e6, "hbm_peak_GBps": peak, "frac": 12.0 * n / ms / 1e6 / peak,
                      "launches_per_step": 1}))
    del ema, h, opt, params
    torch.cuda.empty_cache()


def vae_bench(batch=1, px=1024):
    from neurosis_b200 import ops
    from neurosis_b200.modules.vae import AutoencoderKL, DiagonalGaussianRegularizer
    dev = "cuda"
    torch.manual_seed(0)
    ae = AutoencoderKL(4, SDXL_VAE, regularizer=DiagonalGaussianRegularizer(sample=True)).to(dev)
    img = torch.rand(batch, 3, px, px, device=dev) * 2 - 1

    def step():
        for p in ae.parameters():
            p.grad = None
        loss = ae.training_step({"image": img})
        loss.backward()
        return loss

    l0 = ops.LAUNCHES
    loss = step()
    torch.cuda.synchronize()
    launches = ops.LAUNCHES - l0
    ms = timed(step, warm=1, iters=3)
    gflop = 46048.0 * batch * (px / 1024.0) ** 2  # enc+dec training step per 1024^2 image
    print(json.dumps({"bench": "vae_training_step", "batch": batch, "px": px, "ms": ms, "loss": float(loss),
                      "algorithmic_TFLOP": gflop / 1e3, "TFLOPps": gflop / ms, "kernel_launches": launches,
                      "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9,
                      "grads_finite": all(bool(torch.isfinite(p.grad).all()) for p in ae.parameters())}))


if __name__ == "__main__":
    optimizer_bench()
    if "--skip-vae" not in sys.argv:
        vae_bench()
