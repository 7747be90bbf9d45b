"""Oracle (test infrastructure): deterministic synthetic weights for the reference's state-dict keys.

There are no pretrained checkpoints in the reference repository (README.md:62-64) and no network here,
so parity runs use seeded synthetic weights.  The recipe depends only on (seed, key, shape), so the
golden-vector generator (which fills the *reference's* modules), the oracle and the CUDA product all see
bit-identical fp32 parameters on any machine with the same torch build.  The 18 zero-initialised convs
of the reference UNet (openai_model_3d.py:268-270, 727) get non-zero values, otherwise every
implementation trivially outputs 0.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Mapping, Tuple

import torch


def synth_tensor(seed: int, key: str, shape: Tuple[int, ...]) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFF)
    if key.endswith("num_batches_tracked"):
        return torch.zeros((), dtype=torch.int64)
    if key.endswith("running_mean"):
        return 0.1 * torch.randn(shape, generator=g)
    if key.endswith("running_var"):
        return 1.0 + 0.2 * torch.rand(shape, generator=g)
    if "embedding" in key or "embeddings" in key:          # nn.Embedding tables (codebook, class embeddings)
        return torch.randn(shape, generator=g)
    if len(shape) >= 2:                                     # conv / linear weight: unit-gain fan-in scaling
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return torch.randn(shape, generator=g) / math.sqrt(fan_in)
    if key.endswith(".weight"):                             # norm gains
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    return 0.05 * torch.randn(shape, generator=g)           # biases


def synth_state_dict(shapes: Mapping[str, Tuple[int, ...]], seed: int) -> Dict[str, torch.Tensor]:
    return {k: synth_tensor(seed, k, tuple(s)) for k, s in shapes.items()}


@torch.no_grad()
def fill_module_(module: torch.nn.Module, seed: int) -> None:
    """Overwrite every parameter and buffer of `module` with the synthetic value for its state-dict key."""
    sd = module.state_dict()
    for k, v in sd.items():
        v.copy_(synth_tensor(seed, k, tuple(v.shape)).to(v.dtype))
