#!/usr/bin/env python
"""The stock-PyTorch bar on the same B200 (SURVEY.md §2.1 / BASELINE.md §3): the UNMODIFIED reference
modules (baseline/_ref) under `torch.autocast("cuda", bf16)` — cuBLAS linears, cuDNN convolutions, SDPA attention — on
the benchmarked configuration, next to per-op A/B timings of the library kernels against this repository's kernels at the
top call sites of the step.

  python tools/stock_torch_bench.py [--family sdxl|sd15] [--batch 16] [--steps 5] [--ops] [--out gpurun_out/stock.json]

Everything is CUDA-event timed after warm-up; inputs are resident in HBM.  Results are written as JSON + a text table.
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import ref_harness as RH  # noqa: E402


def timed(fn, iters: int, warmup: int = 2) -> float:
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def whole_step(family: str, batch: int, steps: int, use_checkpoint: bool) -> dict:
    dev = "cuda:0"
    px = 1024 if family == "sdxl" else 512
    try:
        rs = RH.RefStep(family, dev, use_checkpoint=use_checkpoint, autocast_bf16=True)
        g = torch.Generator().manual_seed(42)
        img = (torch.rand(batch, 3, px, px, generator=g) * 2 - 1).to(dev)
        ctx = torch.randn(batch, 77, rs.ctx_dim(), generator=g).to(dev)
        vec = torch.randn(batch, 2816, generator=g).to(dev) if family == "sdxl" else None

        def one():
            rs(img, ctx, vec)
            rs.zero()

        torch.cuda.reset_peak_memory_stats()
        ms = timed(one, steps, warmup=3)
        r = {"family": family, "batch": batch, "px": px, "use_checkpoint": use_checkpoint, "ms_per_step": ms,
             "images_per_s": batch / ms * 1e3, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
        del rs
    except torch.cuda.OutOfMemoryError as e:  # report, do not die
        r = {"family": family, "batch": batch, "px": px, "use_checkpoint": use_checkpoint, "oom": str(e)[:120]}
    torch.cuda.empty_cache()
    return r


def op_ab(iters: int) -> list:
    """library kernel vs ours at the top call sites of the SDXL step (B = 16).  TFLOP/s from algorithmic FLOPs."""
    import torch.nn.functional as F
    from neurosis_b200 import ops
    dev = "cuda:0"
    bf = torch.bfloat16
    rows = []

    def rec(name, flops, lib_ms, our_ms):
        rows.append({"site": name, "gflop": flops / 1e9, "library_ms": lib_ms, "ours_ms": our_ms,
                     "library_tflops": flops / lib_ms / 1e9, "ours_tflops": flops / our_ms / 1e9})

    # ---- linears (cuBLAS via F.linear) ----
    for name, M, N, K in (("linear 16384x1280x1280 (attn out / q / proj)", 16384, 1280, 1280),
                          ("linear 16384x10240x1280 (GEGLU in)", 16384, 10240, 1280),
                          ("linear 16384x1280x5120 (FF out)", 16384, 1280, 5120),
                          ("linear 65536x640x640", 65536, 640, 640),
                          ("linear 65536x5120x640 (GEGLU in, 64x64 level)", 65536, 5120, 640),
                          ("linear 1232x1280x2048 (cross k/v)", 1232, 1280, 2048)):
        x = torch.randn(M, K, device=dev, dtype=bf)
        w = torch.randn(N, K, device=dev, dtype=bf) * K ** -0.5
        b = torch.randn(N, device=dev, dtype=bf)
        bf32 = b.float()
        lib = timed(lambda: F.linear(x, w, b), iters)
        our = timed(lambda: ops.linear_fwd(x, w, bf32), iters)
        rec(name, 2.0 * M * N * K, lib, our)
        del x, w
    # ---- convolutions (cuDNN, channels_last, bf16) ----
    for name, n, c_in, c_out, hw in (("conv3x3 16x320->320 @128^2", 16, 320, 320, 128),
                                     ("conv3x3 16x640->640 @64^2", 16, 640, 640, 64),
                                     ("conv3x3 16x1280->1280 @32^2", 16, 1280, 1280, 32),
                                     ("conv3x3 16x1920->1280 @32^2 (skip concat)", 16, 1920, 1280, 32),
                                     ("conv3x3 4x128->128 @1024^2 (VAE)", 4, 128, 128, 1024),
                                     ("conv3x3 4x256->256 @512^2 (VAE)", 4, 256, 256, 512),
                                     ("conv3x3 4x512->512 @256^2 (VAE)", 4, 512, 512, 256)):
        x = torch.randn(n, c_in, hw, hw, device=dev, dtype=bf).contiguous(memory_format=torch.channels_last)
        w = (torch.randn(c_out, c_in, 3, 3, device=dev) * (9 * c_in) ** -0.5)
        wb = w.to(bf).contiguous(memory_format=torch.channels_last)
        b = torch.randn(c_out, device=dev)
        lib = timed(lambda: F.conv2d(x, wb, b.to(bf), padding=1), iters)
        xn = x.permute(0, 2, 3, 1).contiguous()
        wp = torch.nn.Parameter(w)
        wf, _ = ops.packed_conv_weight(wp)
        our = timed(lambda: ops.conv2d_fwd(xn, wf, c_out, 3, b), iters)
        rec(name, 2.0 * n * hw * hw * c_out * 9 * c_in, lib, our)
        del x, xn, w, wb
    # ---- attention (SDPA: flash / cuDNN backends as torch picks them) ----
    for name, B, H, Nq, Nk in (("self-attn fwd 16x10x4096 d64", 16, 10, 4096, 4096),
                               ("self-attn fwd 16x20x1024 d64", 16, 20, 1024, 1024),
                               ("cross-attn fwd 16x20x1024x77 d64", 16, 20, 1024, 77)):
        q = torch.randn(B, Nq, H, 64, device=dev, dtype=bf)
        k = torch.randn(B, Nk, H, 64, device=dev, dtype=bf)
        v = torch.randn(B, Nk, H, 64, device=dev, dtype=bf)
        qt, kt, vt = (t.transpose(1, 2) for t in (q, k, v))
        lib = timed(lambda: F.scaled_dot_product_attention(qt, kt, vt), iters)
        our = timed(lambda: ops.attention_fwd(q, k, v, 0.125), iters)
        fl = 4.0 * B * H * Nq * Nk * 64
        rec(name, fl, lib, our)
        # backward
        qg, kg, vg = (t.detach().clone().requires_grad_(True) for t in (qt, kt, vt))
        o = F.scaled_dot_product_attention(qg, kg, vg)
        go = torch.randn_like(o)
        lib_b = timed(lambda: torch.autograd.grad(o, (qg, kg, vg), go, retain_graph=True), iters)
        o2, lse = ops.attention_fwd(q, k, v, 0.125)
        do = go.transpose(1, 2).contiguous()
        our_b = timed(lambda: ops.attention_bwd(do, q, k, v, o2, lse, 0.125), iters)
        rec(name.replace("fwd", "bwd"), 2.5 * fl, lib_b, our_b)
        del q, k, v, o, o2
    # ---- norms ----
    for name, shape, groups in (("GroupNorm+SiLU 16x320 @128^2", (16, 128, 128, 320), 32),
                                ("GroupNorm+SiLU 4x128 @1024^2 (VAE)", (4, 1024, 1024, 128), 32)):
        xn = torch.randn(*shape, device=dev, dtype=bf)
        x = xn.permute(0, 3, 1, 2)  # channels_last view
        gm, bt = torch.ones(shape[-1], device=dev), torch.zeros(shape[-1], device=dev)
        lib = timed(lambda: F.silu(F.group_norm(x, groups, gm.to(bf), bt.to(bf), 1e-5)), iters)
        our = timed(lambda: ops.groupnorm_fwd(xn, gm, bt, groups, 1e-5, True), iters)
        nbytes = 4.0 * xn.numel()
        rows.append({"site": name, "gbytes": nbytes / 1e9, "library_ms": lib, "ours_ms": our,
                     "library_gbs": nbytes / lib / 1e6, "ours_gbs": nbytes / our / 1e6})
        del x, xn
    x = torch.randn(16384, 1280, device=dev, dtype=bf)
    gm, bt = torch.ones(1280, device=dev), torch.zeros(1280, device=dev)
    lib = timed(lambda: F.layer_norm(x, (1280,), gm.to(bf), bt.to(bf), 1e-5), iters)
    our = timed(lambda: ops.layernorm_fwd(x, gm, bt, 1e-5), iters)
    rows.append({"site": "LayerNorm 16384x1280", "gbytes": 4.0 * x.numel() / 1e9, "library_ms": lib, "ours_ms": our,
                 "library_gbs": 4.0 * x.numel() / lib / 1e6, "ours_gbs": 4.0 * x.numel() / our / 1e6})
    return rows


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--family", default="sdxl", choices=["sdxl", "sd15"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--ops", action="store_true", help="also run the per-op library-vs-ours A/B")
    ap.add_argument("--no-step", action="store_true")
    ap.add_argument("--out", default="gpurun_out/stock_torch.json")
    args = ap.parse_args()
    assert torch.cuda.is_available(), "needs cuda:0"
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    res = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "steps": []}
    if not args.no_step:
        assert RH.available(), "baseline/_ref missing: pip install --target baseline/_ref /root/reference"
        for ck in (True, False):  # the YAML's use_checkpoint: true, and the faster no-recompute variant if it fits
            r = whole_step(args.family, args.batch, args.steps, ck)
            print(json.dumps(r), flush=True)
            res["steps"].append(r)
    if args.ops:
        res["ops"] = op_ab(10)
        for r in res["ops"]:
            print(json.dumps(r), flush=True)
    out = Path(args.out)
    out.parent.mkdir(parents=True, exist_ok=True)
    out.write_text(json.dumps(res, indent=1))
    with open(out.with_suffix(".txt"), "w") as fh:
        fh.write(f"stock PyTorch ({res['torch']}) vs neurosis_b200 on {res['gpu']}\n")
        for r in res["steps"]:
            fh.write(json.dumps(r) + "\n")
        for r in res.get("ops", []):
            if "gflop" in r:
                fh.write(f"{r['site']:52s} library {r['library_ms']:8.3f} ms {r['library_tflops']:7.1f} TF/s | ours "
                         f"{r['ours_ms']:8.3f} ms {r['ours_tflops']:7.1f} TF/s | x{r['library_ms'] / r['ours_ms']:.2f}\n")
            else:
                fh.write(f"{r['site']:52s} library {r['library_ms']:8.3f} ms {r['library_gbs']:7.0f} GB/s | ours "
                         f"{r['ours_ms']:8.3f} ms {r['ours_gbs']:7.0f} GB/s | x{r['library_ms'] / r['ours_ms']:.2f}\n")


if __name__ == "__main__":
    main()
