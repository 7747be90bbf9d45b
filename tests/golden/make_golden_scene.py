"""Golden vectors for the host-side object selection of the shape branch, produced by the REAL reference
class `Sg2ScVAEModel.select_sdfs` / `balance_objects` (VAEGAN_V2FULL.py:398-463; oracle/reference_scene_model.py builds the
class with inert stubs for its absent third-party imports).  Build container only.

    python tests/golden/make_golden_scene.py
"""
from __future__ import annotations

import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import graph as G, reference_scene_model as RS  # noqa: E402


def scene_inputs(seed):
    g = torch.Generator().manual_seed(seed)
    sizes = [7, 4, 9]                                           # objects per scene (the last object of each is the _scene_ node)
    scene = torch.cat([torch.full((n,), i) for i, n in enumerate(sizes)])
    O = int(scene.numel())
    objs = torch.randint(1, 30, (O,), generator=g)
    grained = objs * 2 + torch.randint(0, 2, (O,), generator=g)  # fine-grained ids: two per coarse class
    sdfs = torch.randn(O, 1, 4, 4, 4, generator=g)
    sdfs[[6, 10, 19, 3]] = 0                                    # objects without an SDF (floor / scene nodes) are never picked
    uc = torch.randn(O, 1, 8, generator=g)
    c = torch.randn(O, 1, 8, generator=g)
    return scene, objs, grained, sdfs, uc, c


def main():
    real = RS.build(G.GCN_TINY | dict(add_dim=512, rel_hidden=960, rel_out=1280), diffusion_bs=8, seed=0)
    torch.Tensor.cuda = lambda self, *a, **k: self              # select_sdfs ends with .cuda() (:458-460); stay on the CPU
    out = {}
    scene, objs, grained, sdfs, uc, c = scene_inputs(31)
    out.update(scene=scene.numpy(), objs=objs.numpy(), grained=grained.numpy(), sdfs=sdfs.numpy(), uc=uc.numpy(), c=c.numpy())
    for bs in (8, 4, 12):
        real.diffusion_bs = bs
        random.seed(1000 + bs)
        cats, d = real.select_sdfs(scene, objs, grained, sdfs, uc, c, random=False)
        out[f"balanced_bs{bs}_cats"] = cats.numpy()
        out[f"balanced_bs{bs}_sdf"] = d["sdf"].numpy(); out[f"balanced_bs{bs}_uc"] = d["uc"].numpy(); out[f"balanced_bs{bs}_rel"] = d["rel"].numpy()
        torch.manual_seed(2000 + bs)
        cats, d = real.select_sdfs(scene, objs, grained, sdfs, uc, c, random=True)
        out[f"random_bs{bs}_cats"] = cats.numpy(); out[f"random_bs{bs}_sdf"] = d["sdf"].numpy()
    random.seed(77)
    ids = torch.tensor([5, 5, 9, 2, 9, 9, 7])
    out["balance_ids"] = ids.numpy()
    out["balance_n3"] = real.balance_objects(ids, ids, 3).numpy()
    out["balance_n6"] = real.balance_objects(ids, ids, 6).numpy()          # more than the 4 distinct ids: the rest is drawn with repeats
    # state-dict inventory of the real class in the v2_full wiring (embedding_dim 64, 5 layers, BatchNorm, residual, angles):
    # what a `model{epoch}.pth` of the reference holds besides 'epoch' / 'counter' / 'vqvae' / 'df' / 'opt'
    import json
    full = RS.build(G.GCN_FULL, diffusion_bs=8, seed=0)
    inv = [[k, list(v.shape)] for k, v in RS.module_state_dict(full).items()]      # in the class's own registration order
    with open(os.path.join(HERE, "sg2sc_v2full_state_dict_keys.json"), "w") as f:
        json.dump(inv, f, indent=0)
    print(f"sg2sc_v2full_state_dict_keys.json: {len(inv)} keys")
    np.savez_compressed(os.path.join(HERE, "select_sdfs.npz"), **out)
    print("select_sdfs.npz:", {k: v.shape for k, v in out.items() if "cats" in k or k.startswith("balance_n")})


if __name__ == "__main__":
    main()
