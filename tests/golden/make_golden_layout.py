"""Golden fixtures for the LAYOUT branch, produced by the reference's REAL Sg2ScVAEModel
(encoder :185-218, manipulate :244-258, decoder :260-289) and model/losses.py:calculate_model_losses, in the v2_full wiring
with seeded synthetic weights (oracle/weights.py) and a seeded synthetic graph.  Build container only.

    python tests/golden/make_golden_layout.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import graph as G, layout as Lo, reference_scene_model as RS, weights as Wt  # noqa: E402
from oracle.validate_against_reference import synth_graph  # noqa: E402

SEED = 23


class _W:
    def add_scalar(self, *a, **k):
        pass


@torch.no_grad()
def main():
    for tag, cfg in (("tiny", Lo.LAYOUT_TINY), ("full", Lo.LAYOUT_FULL)):
        real = RS.build(dict(cfg, rel_hidden=960, rel_out=1280), seed=0)
        from model.losses import calculate_model_losses
        shapes = Lo.layout_param_shapes(cfg)
        out = {"weight_seed": SEED}
        z, objs, triples, text, rel = synth_graph(dict(cfg), 11, 24, seed=400)
        g = torch.Generator().manual_seed(401)
        boxes, angles = torch.randn(11, 6, generator=g), torch.randint(0, 24, (11,), generator=g)
        zz = torch.randn(11, 2 * cfg["embedding_dim"], generator=g)
        out.update(z=z.numpy(), objs=objs.numpy(), triples=triples.numpy(), text=text.numpy(), rel=rel.numpy(), boxes=boxes.numpy(),
                   angles=angles.numpy(), zz=zz.numpy())
        for mode in ("eval", "train"):
            torch.nn.Module.load_state_dict(real, Wt.synth_state_dict(shapes, SEED), strict=False)    # reset running stats
            real.train(mode == "train")
            mu, logvar = real.encoder(objs, triples, boxes, None, text, rel, angles)
            torch.nn.Module.load_state_dict(real, Wt.synth_state_dict(shapes, SEED), strict=False)
            man = real.manipulate(zz, objs, triples, text, rel, None)
            torch.nn.Module.load_state_dict(real, Wt.synth_state_dict(shapes, SEED), strict=False)
            b, a = real.decoder(z, objs, triples, text, rel, None)
            tot, _ = calculate_model_losses(None, b, boxes, "box", angles=angles, angles_pred=a, mu=mu, logvar=logvar, KL_weight=0.1,
                                            writer=_W(), counter=0, withangles=True)
            out.update({f"mu_{mode}": mu.numpy(), f"logvar_{mode}": logvar.numpy(), f"man_{mode}": man.numpy(), f"boxes_{mode}": b.numpy(),
                        f"angle_logp_{mode}": a.numpy(), f"loss_{mode}": np.asarray(float(tot))})
        np.savez_compressed(os.path.join(HERE, f"layout_{tag}.npz"), **out)
        print(f"layout_{tag}.npz: loss eval {float(out['loss_eval']):.4f} train {float(out['loss_train']):.4f}")


if __name__ == "__main__":
    main()
