#!/usr/bin/env python
"""Times nk_linear_fwd / dgrad / wgrad on the step's big shapes (CUDA events, L2 flushed by size: 10 distinct buffers)."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from neurosis_b200 import ops  # noqa: E402

dev, bf = "cuda", torch.bfloat16
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("NK_GEMM"))
for M, N, K in ((16384, 10240, 1280), (16384, 1280, 1280), (16384, 1280, 5120), (65536, 640, 640), (65536, 5120, 640)):
    xs = [torch.randn(M, K, device=dev, dtype=bf) for _ in range(4)]
    w = torch.randn(N, K, device=dev, dtype=bf) * K ** -0.5
    dys = [torch.randn(M, N, device=dev, dtype=bf) for _ in range(4)]
    dw = torch.zeros(N, K, device=dev)
    for what, fn in (("fwd", lambda i: ops.linear_fwd(xs[i % 4], w)), ("dgrad", lambda i: ops.linear_dgrad(dys[i % 4], w)),
                     ("wgrad", lambda i: ops.linear_wgrad(dys[i % 4], xs[i % 4], out=dw))):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(12):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 12
        print(f"[{tag}] {what:5s} {M}x{N}x{K}: {ms:7.3f} ms {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s", flush=True)
    del xs, dys, w, dw
