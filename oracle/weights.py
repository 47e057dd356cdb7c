"""ORACLE (test infrastructure): deterministic synthetic weights.

Random-init parity is vacuous in the reference because `zero_module` (modules/diffusion/util.py:180-186)
zeroes every ResBlock output conv, SpatialTransformer.proj_out and the final `out` conv,
so all parity work uses weights drawn here — identically for the reference modules, the oracle and the CUDA
modules — from numpy's MT19937 stream (bit-stable across platforms), keyed by the parameter name.
"""
from __future__ import annotations

import zlib

import numpy as np
import torch


def synth_state_dict(shapes: dict[str, tuple], seed: int = 0, dtype=torch.float32) -> dict[str, torch.Tensor]:
    sd = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        rs = np.random.RandomState((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
        if name.endswith(".bias"):
            w = rs.standard_normal(shape) * 0.05
        elif len(shape) == 1:  # norm scales
            w = 1.0 + rs.standard_normal(shape) * 0.1
        else:
            fan_in = int(np.prod(shape[1:]))
            w = rs.standard_normal(shape) * (1.0 / np.sqrt(fan_in))
        sd[name] = torch.from_numpy(np.ascontiguousarray(w)).to(dtype)
    return sd


def synth_tensor(name: str, shape, seed: int = 0, scale: float = 1.0, uniform: bool = False) -> torch.Tensor:
    rs = np.random.RandomState((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    a = rs.uniform(-1.0, 1.0, size=tuple(shape)) if uniform else rs.standard_normal(tuple(shape))
    return torch.from_numpy(np.ascontiguousarray(a * scale)).float()
