#!/usr/bin/env python
"""Times the fused attention kernels on the SDXL shapes (run on the GPU box).  usage: python tools/attn_bench.py [iters]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from neurosis_b200 import ops  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = "cuda"
# the VAE mid-block attention (one 512-wide head, forward only: the encoder runs without gradients)
qv, kv, vv = (torch.randn(8, 4096, 1, 512, device=dev).bfloat16() for _ in range(3))
ops.attention_fwd(qv, kv, vv, 512 ** -0.5)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(iters):
    ops.attention_fwd(qv, kv, vv, 512 ** -0.5)
ev[1].record()
torch.cuda.synchronize()
tv = ev[0].elapsed_time(ev[1]) / iters
print(f"vae 4096 d512    fwd {tv:7.3f} ms {4.0 * 8 * 4096 * 4096 * 512 / tv / 1e9:7.1f} TFLOP/s (useful flops)")
del qv, kv, vv

for name, B, H, Nq, Nk in (("self 1024", 8, 20, 1024, 1024), ("self 4096", 8, 10, 4096, 4096),
                           ("cross 1024x77", 8, 20, 1024, 77), ("cross 4096x77", 8, 10, 4096, 77)):
    q = torch.randn(B, Nq, H, 64, device=dev).bfloat16()
    k = torch.randn(B, Nk, H, 64, device=dev).bfloat16()
    v = torch.randn(B, Nk, H, 64, device=dev).bfloat16()
    do = torch.randn(B, Nq, H, 64, device=dev).bfloat16()
    o, lse = ops.attention_fwd(q, k, v, 0.125)
    ops.attention_bwd(do, q, k, v, o, lse, 0.125)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    for _ in range(iters):
        ops.attention_fwd(q, k, v, 0.125)
    e[1].record()
    for _ in range(iters):
        ops.attention_bwd(do, q, k, v, o, lse, 0.125)
    e[2].record()
    torch.cuda.synchronize()
    fl = 4.0 * B * H * Nq * Nk * 64
    tf, tb = e[0].elapsed_time(e[1]) / iters, e[1].elapsed_time(e[2]) / iters
    print(f"{name:16s} fwd {tf:7.3f} ms {fl / tf / 1e9:7.1f} TFLOP/s | bwd(+delta,+cast) {tb:7.3f} ms {2.5 * fl / tb / 1e9:7.1f} TFLOP/s")
