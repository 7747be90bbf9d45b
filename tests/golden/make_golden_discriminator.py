"""Golden for BoxDiscriminator (model/discriminators.py:80-163 + discriminator_regularizer :148-163) from the reference's own
class on the CPU: probabilities, gradient-penalty terms and parameter gradients of `mean(y) + mean(reg)`, train-mode
BatchNorm, with and without the `keeps` mask.  Build container only.

    python tests/golden/make_golden_discriminator.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.environ.get("CS_REFERENCE", "/root/reference"))

from oracle import weights as Wt  # noqa: E402

SEED = 37


def main():
    from model.discriminators import BoxDiscriminator
    d = BoxDiscriminator(6, 16, 36).train()
    g = torch.Generator().manual_seed(600)
    O, T = 9, 20
    objs = torch.randint(0, 36, (O,), generator=g)
    s = torch.randint(0, O, (T,), generator=g)
    o = (s + 1 + torch.randint(0, O - 1, (T,), generator=g)) % O
    triples = torch.stack([s, torch.randint(0, 16, (T,), generator=g), o], 1)
    boxes = torch.randn(O, 6, generator=g)
    keep = (torch.rand(O, 1, generator=g) > 0.3).float()
    out = dict(weight_seed=SEED, objs=objs.numpy(), triples=triples.numpy(), boxes=boxes.numpy(), keep=keep.numpy())
    modes = {"plain": dict(), "keeps": dict(keeps=keep), "real": dict(with_grad=True, is_real=True), "fake_keeps": dict(keeps=keep, with_grad=True, is_real=False)}
    for name, kw in modes.items():
        Wt.fill_module_(d, SEED)
        d.zero_grad()
        y, reg = d(objs, triples, boxes.clone(), **kw)
        loss = y.mean() + (reg.mean() if reg is not None else 0.0)
        loss.backward()
        out[f"{name}_y"] = y.detach().numpy()
        if reg is not None:
            out[f"{name}_reg"] = reg.detach().numpy()
        for k, p in d.named_parameters():       # big tensors: norm + a seeded random projection + a slice (keeps the fixture small)
            gr = p.grad.numpy().copy()
            if gr.size <= 2048:
                out[f"{name}_grad_{k}"] = gr
            else:
                proj = np.random.RandomState(SEED).standard_normal(gr.size).astype(np.float32)
                out[f"{name}_gsum_{k}"] = np.asarray([np.linalg.norm(gr), float(gr.reshape(-1) @ proj)], dtype=np.float64)
                out[f"{name}_ghead_{k}"] = gr.reshape(-1)[:512]
    np.savez_compressed(os.path.join(HERE, "box_discriminator.npz"), **out)
    print("box_discriminator.npz:", {k: v.shape for k, v in out.items() if k.endswith("_y")})


if __name__ == "__main__":
    main()
