#!/usr/bin/env python
"""One forward + backward of the fused attention kernels on an SDXL shape (target of the ncu captures).
usage: python tools/attn_one.py [Nq] [H]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from neurosis_b200 import ops  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
H = int(sys.argv[2]) if len(sys.argv) > 2 else 10
B = 8
q, k, v, do = (torch.randn(B, N, H, 64, device="cuda").bfloat16() for _ in range(4))
for _ in range(2):
    o, lse = ops.attention_fwd(q, k, v, 0.125)
    ops.attention_bwd(do, q, k, v, o, lse, 0.125)
torch.cuda.synchronize()
