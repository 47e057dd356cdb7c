"""Runs the UNMODIFIED reference (neggles/neurosis, vendored by `pip install --target baseline/_ref`)
on the hot path: VAE latent encode (no-grad) -> StandardDiffusionLoss -> loss.mean().backward().

Used by
  * `bench.py --impl reference` and bench.py's `cpu_baseline` leg: the reference's own modules on the HOST cores, fp32;
  * `tools/stock_torch_bench.py`: the same modules on cuda:0 under `torch.autocast("cuda", bf16)` — the stock-PyTorch
    (cuBLAS / cuDNN / SDPA) bar on the same B200.

Nothing of neurosis_b200 is imported here.  The reference's L5 engine (`DiffusionEngine`, `AutoencoderKL`) needs
`lightning`, which this image does not have; the glue of `training_step` / `encode_first_stage`
(/root/reference/src/neurosis/models/diffusion.py:186-233: `scale_factor * vae_encoder(x, regularize=True)`,
`loss_fn(model, denoiser, conditioner, x, batch)`, `loss.mean()`) is the 6 lines of `RefStep.__call__` below; every module it
calls is the reference's.  Deviations from the example YAML, each forced by the image: attention type
`softmax-xformers` -> `torch-sdp` (xformers is absent; attention.py:369-417 is the reference's own SDPA class), VAE
`vanilla-xformers` -> `vanilla` (model.py:144-172); the sigma generator's `t in [0,1)` call always selects sigma 0 (SURVEY.md
appendix A), so — as in our own arm — the harness draws indices through the generator's own `t=None` randint branch.
"""
from __future__ import annotations

import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "baseline" / "_ref"

SDXL_UNET = dict(in_channels=4, model_channels=320, out_channels=4, num_res_blocks=2, attention_resolutions=[4, 2],
                 channel_mult=[1, 2, 4], num_head_channels=64, transformer_depth=[1, 2, 10], context_dim=2048,
                 use_linear_in_transformer=True, num_classes="sequential", adm_in_channels=2816,
                 spatial_transformer_attn_type="torch-sdp")  # configs/sdxl/sdxl.example.yaml:68-84
SD15_UNET = dict(in_channels=4, model_channels=320, out_channels=4, num_res_blocks=2, attention_resolutions=[4, 2, 1],
                 channel_mult=[1, 2, 4, 4], num_heads=8, transformer_depth=1, context_dim=768,
                 use_linear_in_transformer=False, spatial_transformer_attn_type="torch-sdp")  # configs/sd15/sd15.example.yml:68-81
KL_F8_VAE = dict(ch=128, out_ch=3, ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], in_channels=3,
                 resolution=256, z_channels=4, double_z=True)  # configs/sdxl/sdxl.example.yaml:102-113


def available() -> bool:
    return (REF / "neurosis" / "modules" / "diffusion" / "openaimodel.py").exists()


def _import():
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    import neurosis.modules.diffusion as D  # noqa: F401  (first: the reference has a circular import)
    return D


class RefStep:
    """one training step of the reference path; `__call__(image, ctx, vec)` returns the scalar loss tensor after
    `backward()` (gradients are left in `.grad`; `zero()` drops them)."""

    def __init__(self, family: str = "sdxl", device: str = "cpu", use_checkpoint: bool = True, autocast_bf16: bool = False,
                 seed: int = 42):
        import torch
        D = _import()
        from neurosis.modules.diffusion import (DiscreteDenoiser, DiscreteSigmaGenerator, EpsPreconditioning, EpsWeighting,
                                                LegacyDDPMDiscretization, OpenAIWrapper, StandardDiffusionLoss, UNetModel)
        from neurosis.modules.diffusion.model import Encoder
        self.torch = torch
        self.family = family
        cfg = dict(SDXL_UNET if family == "sdxl" else SD15_UNET, use_checkpoint=use_checkpoint)
        torch.manual_seed(seed)
        if device == "cpu":
            # construct on the meta device and draw the weights in one pass: the reference's default init of 2.57 G
            # parameters costs ~50 s of single-threaded host RNG, which is set-up, not the thing timed
            with torch.device("meta"):
                unet = UNetModel(**cfg)
                enc = Encoder(**KL_F8_VAE, embed_dim=4, standalone=True, attn_type="vanilla")
            unet, enc = unet.to_empty(device="cpu"), enc.to_empty(device="cpu")
            block = torch.randn(1 << 20) * 0.02  # tiled into the weights: the values only have to be finite and small
            with torch.no_grad():
                for m in (unet, enc):
                    for n, p in m.named_parameters():
                        if p.dim() > 1:
                            flat = p.view(-1)
                            for off in range(0, flat.numel(), block.numel()):
                                k = min(block.numel(), flat.numel() - off)
                                flat[off: off + k].copy_(block[:k])
                        elif n.endswith("weight"):
                            p.fill_(1.0)
                        else:
                            p.zero_()
        else:
            with torch.device(device):
                unet = UNetModel(**cfg)
                enc = Encoder(**KL_F8_VAE, embed_dim=4, standalone=True, attn_type="vanilla")
            with torch.no_grad():  # zero-initialised layers re-drawn, as in our arm (otherwise most gradients are zero)
                for p in unet.parameters():
                    if p.dim() > 1 and float(p.abs().sum()) == 0.0:
                        p.normal_(0.0, 0.02)
        for p in enc.parameters():
            p.requires_grad_(False)
        self.unet, self.enc = unet, enc
        self.net = OpenAIWrapper(unet)
        den = DiscreteDenoiser(EpsPreconditioning(), 1000, LegacyDDPMDiscretization())
        den.sigmas, den.log_sigmas = den.sigmas.detach(), den.log_sigmas.detach()  # (tables carry an autograd graph)
        self.den = den.to(device)

        class RandIdx(DiscreteSigmaGenerator):
            def __call__(self, n, t=None):
                return super().__call__(n, None).clamp_min(0.03)

        gen = RandIdx(LegacyDDPMDiscretization(), 1000)
        gen.sigmas = gen.sigmas.detach()
        self.loss_fn = StandardDiffusionLoss(sigma_generator=gen, loss_weighting=EpsWeighting())
        self.has_vec = family == "sdxl"

        class Cond(torch.nn.Module):
            def forward(self_c, batch):
                c = {"crossattn": batch["ctx"]}
                if "vec" in batch:
                    c["vector"] = batch["vec"]
                return c

        self.cond = Cond()
        self.scale_factor = 0.13025 if family == "sdxl" else 0.18215
        self.device = device
        self.autocast = autocast_bf16

    def ctx_dim(self) -> int:
        return 2048 if self.family == "sdxl" else 768

    def zero(self) -> None:
        for p in self.unet.parameters():
            p.grad = None

    def __call__(self, image, ctx, vec=None):
        torch = self.torch
        with torch.autocast(self.device.split(":")[0], dtype=torch.bfloat16, enabled=self.autocast):
            with torch.no_grad():
                z = self.scale_factor * self.enc(image, regularize=True)  # models/diffusion.py:186-198
            batch = {"ctx": ctx}
            if vec is not None and self.has_vec:
                batch["vec"] = vec
            loss = self.loss_fn(self.net, self.den, self.cond, z, batch)  # models/diffusion.py:200-204
        total = loss.mean()  # models/diffusion.py:233
        total.backward()
        return total


# algorithmic GFLOP per image: token-proportional terms scale with (latent/L0)^2, attention with ^4
def step_gflop(family: str, latent: int) -> float:
    if family == "sdxl":
        r2 = (latent / 128.0) ** 2
        lin = (6761.2 - 751.6 - 32.3) * r2
        attn = 751.6 * r2 * r2 + 32.3 * r2
        vae = (4879.0 - 550.0) * r2 + 550.0 * r2 * r2
        return 3 * (lin + attn) + vae
    r2 = (latent / 64.0) ** 2
    return (2409.8 + 1116.7) * r2  # SD1.5: attention is a few percent of the step; scaled with the token count


def cpu_sample(family: str, latent: int, steps: int, warmup: int, batch: int = 1) -> dict:
    """the reference on the host cores, fp32, all threads: seconds per step of a `batch` x (8*latent)^2 px sample."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rs = RefStep(family, "cpu", use_checkpoint=True, autocast_bf16=False)
    g = torch.Generator().manual_seed(42)
    px = latent * 8
    img = torch.rand(batch, 3, px, px, generator=g) * 2 - 1
    ctx = torch.randn(batch, 77, rs.ctx_dim(), generator=g)
    vec = torch.randn(batch, 2816, generator=g) if family == "sdxl" else None
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        rs(img, ctx, vec)
        rs.zero()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return {"sec_per_step": sec, "cores": cores, "threads": torch.get_num_threads(), "batch": batch, "px": px,
            "latent": latent, "gflop_per_image": step_gflop(family, latent)}
