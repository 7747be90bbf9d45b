"""Golden fixtures for the concat-conditioning denoiser (config/sdfusion-txt2shape_concat.yaml),
produced by the REFERENCE's own DiffusionUNet(conditioning_key='concat') — AttentionBlock / QKVAttentionLegacy,
in_channels 4, `dims: 4`.  Build container only (imports /root/reference); same recipe as make_golden.py.

    python tests/golden/make_golden_concat.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import denoiser as D  # noqa: E402
from oracle import validate_against_reference as R  # noqa: E402

SEED = 19


@torch.no_grad()
def main():
    torch.set_num_threads(os.cpu_count() or 1)
    for tag, cfg, B in (("tiny", D.UNET_CONCAT_TINY, 2), ("full", D.UNET_CONCAT_FULL, 2)):
        m = R.ref_unet_concat(cfg, SEED)
        g = torch.Generator().manual_seed(200)
        r = cfg["image_size"]
        x = torch.randn(B, 3, r, r, r, generator=g)
        cc = torch.randn(B, cfg["in_channels"] - 3, r, r, r, generator=g)     # rel_mlp output viewed as (B, 1, 16, 16, 16)
        t = torch.randint(0, 1000, (B,), generator=g)
        eps = m(x, t, c_concat=[cc])
        np.savez_compressed(os.path.join(HERE, f"unet_concat_{tag}.npz"), x=x.numpy(), c_concat=cc.numpy(), t=t.numpy(),
                            eps=eps.numpy(), weight_seed=SEED)
        print(f"unet_concat_{tag}: eps absmax {eps.abs().max():.4f} mean|eps| {eps.abs().mean():.4f}")


if __name__ == "__main__":
    main()
