"""Deterministic random-init weights keyed by state-dict name.

There is no checkpoint access, so benches and parity tests use random-init weights.  Drawing
them from ``torch.manual_seed`` would tie the values to module construction order; instead
every tensor is filled from a generator seeded by ``crc32(key) ^ seed``, so the reference
model, the CPU oracle and the B200 modules all get identical weights from the key set alone.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Mapping

import numpy as np
import torch


def _canonical(key: str) -> str:
    # ResidualBlock registers norm3 twice (as ``norm3`` and as ``downsample.1``): one tensor, two keys
    return key.replace(".downsample.1.", ".norm3.")


def synth_weights_like(state: Mapping[str, torch.Tensor], seed: int = 0) -> Dict[str, torch.Tensor]:
    out: Dict[str, torch.Tensor] = {}
    for key, ref in state.items():
        ck = _canonical(key)
        rng = np.random.default_rng((zlib.crc32(ck.encode()) ^ (seed * 0x9E3779B1)) & 0xFFFFFFFF)
        shape = tuple(ref.shape)
        leaf = ck.rsplit(".", 1)[-1]
        if not ref.dtype.is_floating_point or ck.startswith("moving_dynamicness_threshold"):
            out[key] = ref.clone()
            continue
        if leaf == "running_mean":
            v = rng.uniform(-0.2, 0.2, size=shape)
        elif leaf == "running_var":
            v = rng.uniform(0.5, 1.5, size=shape)
        elif len(shape) == 1 and leaf == "weight":  # norm scale
            v = rng.uniform(0.7, 1.3, size=shape)
        elif len(shape) == 1:  # biases (conv / norm)
            v = rng.uniform(-0.1, 0.1, size=shape)
        else:  # conv / linear weight: kaiming-uniform(a=sqrt(5)) bound, as torch's default init
            fan_in = int(np.prod(shape[1:]))
            bound = 1.0 / math.sqrt(fan_in)
            v = rng.uniform(-bound, bound, size=shape)
        out[key] = torch.from_numpy(np.asarray(v, dtype=np.float32)).to(ref.dtype)
    return out
