#!/usr/bin/env python
"""Run under `ncu --set full --profile-from-start off`: one cuBLAS (F.linear) and one gemm_tc launch per shape inside the
profiler range, so the two kernels' DRAM / L2 / shared-memory traffic, tensor-pipe activity and clocks can be read side by
side (profiles/r02_gemm_vs_cublas.txt)."""
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from neurosis_b200 import ops  # noqa: E402

dev, bf = "cuda", torch.bfloat16
shapes = [(16384, 10240, 1280), (16384, 1280, 1280), (16384, 1280, 5120)]
data = []
for M, N, K in shapes:
    x = torch.randn(M, K, device=dev, dtype=bf)
    w = torch.randn(N, K, device=dev, dtype=bf) * K ** -0.5
    b = torch.randn(N, device=dev)
    data.append((x, w, b, b.to(bf)))
for x, w, b, bb in data:  # warm-up (cuBLAS heuristics, attribute setting)
    for _ in range(3):
        F.linear(x, w, bb)
        ops.linear_fwd(x, w, b)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for x, w, b, bb in data:
    F.linear(x, w, bb)
    ops.linear_fwd(x, w, b)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
