"""Golden fixture for the BENCHMARKED configuration (BASELINE cfg2: 32 objects x CFG = UNet batch 64, full 413.5 M
UNet), produced by the REFERENCE's own DiffusionUNet on the CPU (build container only: imports /root/reference).

The guided batch is built exactly as the reference's DDIMSampler.p_sample_ddim does (samplers/ddim.py:206-209):
x_in = cat([x] * 2), t_in = cat([t] * 2), c_in = cat([uc, c]); eps = model(x_in, t_in, c_in).  Inputs are seeded
(`b64_inputs`, duplicated verbatim in tests/test_unet_gpu.py) and only their checksums are stored; the fixture holds
eps (64, 3, 16, 16, 16) fp32 — ~90 s of CPU on 8 cores, 2.9 MB.

    python tests/golden/make_golden_b64.py
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

from oracle import denoiser as D  # noqa: E402
from oracle import validate_against_reference as R  # noqa: E402

SEED_W, SEED_IN, OBJECTS = 11, 640, 32


def b64_inputs(seed=SEED_IN, objects=OBJECTS):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(objects, 3, 16, 16, 16, generator=g)
    t = torch.randint(0, 1000, (objects,), generator=g)
    uc = torch.randn(objects, 1, 1280, generator=g)
    c = torch.randn(objects, 1, 1280, generator=g)
    return x, t, uc, c


@torch.no_grad()
def main():
    torch.set_num_threads(os.cpu_count() or 1)
    m = R.ref_unet(D.UNET_FULL, SEED_W)
    x, t, uc, c = b64_inputs()
    t0 = time.time()
    eps = m(torch.cat([x] * 2), torch.cat([t] * 2), c_crossattn=[torch.cat([uc, c])])
    print(f"reference DiffusionUNet, batch 64: {time.time() - t0:.1f} s on {torch.get_num_threads()} threads; "
          f"eps absmax {eps.abs().max():.4f} mean|eps| {eps.abs().mean():.4f}")
    np.savez_compressed(os.path.join(HERE, "unet_full_b64.npz"), eps=eps.numpy(), weight_seed=SEED_W, input_seed=SEED_IN,
                        objects=OBJECTS, x_sum=float(x.double().sum()), ctx_sum=float(torch.cat([uc, c]).double().sum()),
                        t=t.numpy())


if __name__ == "__main__":
    main()
