"""Golden for the whole sampling chain of the shape branch, produced by the reference's REAL
`SDFusionText2ShapeModel.rel2shape` (sdfusion_txt2shape_model.py:459-516) on the CPU: one shared x_T, DDIM S=20 eta=0 with
classifier-free guidance 3.0 in mini-batches of 7, `decode_no_quant`.  oracle/reference_diffusion_model.py builds the class
(inert stubs for its absent third-party imports); its hard-coded device='cuda' is patched to the CPU.  Build container only.

    python tests/golden/make_golden_rel2shape.py
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import denoiser as D, vqvae as V, reference_diffusion_model as RD  # noqa: E402

SEED_UNET, SEED_VQ, NOISE_SEED, STEPS = 41, 43, 777, 20
VQ_CFG = dict(V.VQ_TINY, resolution=32)          # 32^3 SDFs <-> 8^3 latents: the tiny denoiser's own grid size


@torch.no_grad()
def main():
    real = RD.build(D.UNET_TINY, VQ_CFG, seed_unet=SEED_UNET, seed_vq=SEED_VQ)
    from model.networks.diffusion_networks.samplers.ddim import DDIMSampler
    DDIMSampler.register_buffer = lambda self, name, attr: setattr(self, name, attr)
    torch.Tensor.cuda = lambda self, *a, **k: self
    randn = torch.randn
    torch.randn = lambda *a, **k: randn(*a, **{kk: vv for kk, vv in k.items() if kk != "device"})
    time.time = lambda: float(NOISE_SEED)                       # rel2shape seeds its shared noise from the clock (:489)
    g = torch.Generator().manual_seed(300)
    n = 9                                                       # two mini-batches: 7 + 2
    rel = torch.randn(n, 1, D.UNET_TINY["context_dim"], generator=g)
    uc = torch.randn(n, 1, D.UNET_TINY["context_dim"], generator=g)
    sdf = real.rel2shape({"sdf": torch.zeros(n, 1, 32, 32, 32), "rel": rel, "uc": uc}, ddim_steps=STEPS, ddim_eta=0.0, uc_scale=3.0)
    torch.manual_seed(NOISE_SEED)
    x_T = randn((1, 3, 8, 8, 8))                                # the same draw rel2shape made
    rows = [0, 6, 7, 8]                                         # keep the fixture small: objects from both mini-batches (0-6 | 7-8)
    np.savez_compressed(os.path.join(HERE, "rel2shape_tiny.npz"), rel=rel.numpy(), uc=uc.numpy(), x_T=x_T.numpy(), sdf=sdf[rows].numpy(), rows=np.asarray(rows),
                        weight_seed_unet=SEED_UNET, weight_seed_vq=SEED_VQ, steps=STEPS, resolution=32)
    print("rel2shape_tiny.npz: sdf", tuple(sdf.shape), "absmax", float(sdf.abs().max()))


if __name__ == "__main__":
    main()
