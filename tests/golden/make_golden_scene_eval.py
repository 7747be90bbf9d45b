"""Goldens for the evaluation entry points that tie the layout and the shape branch together (scripts/eval_3dfront.py ->
Sg2ScVAEModel.sample / decoder_with_changes / decoder_with_additions, VAEGAN_V2FULL.py:291-396, 600-616), produced by the
reference's REAL class on the CPU.  Diff.rel2shape is replaced by a recorder (the denoiser chain has its own golden,
make_golden_rel2shape.py): what is pinned here is the glue — latent sampling (numpy RNG), node insertion, change noise, the
manipulator, which objects reach the denoiser with which conditioning, the layout decode, the `keep` mask.

    python tests/golden/make_golden_scene_eval.py
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

SEED = 29


@torch.no_grad()
def main():
    cfg = dict(Lo.LAYOUT_TINY, rel_hidden=960, rel_out=1280)
    real = RS.build(cfg, seed=0)
    shapes = dict(Lo.layout_param_shapes(Lo.LAYOUT_TINY)); shapes.update(G.gcn_param_shapes(cfg))
    torch.nn.Module.load_state_dict(real, Wt.synth_state_dict(shapes, SEED), strict=True)
    real.eval()
    torch.Tensor.cuda = lambda self, *a, **k: self
    rec = {}

    def rel2shape(d, uc_scale=None):
        rec["d"] = {k: v.clone() for k, v in d.items()}
        return d["rel"].sum(dim=(1, 2))              # marker returned as gen_sdf
    real.Diff.rel2shape = rel2shape
    O, T = 9, 18
    z, objs, triples, text, rel = synth_graph(cfg, O, T, seed=500)
    g = torch.Generator().manual_seed(501)
    sdfs = torch.randn(O, 1, 4, 4, 4, generator=g)
    sdfs[[1, 8]] = 0                                  # objects without an SDF (e.g. the floor and the scene node)
    mean_est = torch.randn(64, generator=g).numpy().astype(np.float64)
    a = torch.randn(64, 64, generator=g).numpy().astype(np.float64)
    cov_est = a @ a.T / 64 + 0.1 * np.eye(64)
    out = dict(weight_seed=SEED, z=z.numpy(), objs=objs.numpy(), triples=triples.numpy(), text=text.numpy(), rel=rel.numpy(), sdfs=sdfs.numpy(),
               mean_est=mean_est, cov_est=cov_est)

    np.random.seed(7)
    (boxes, ang), gen = real.sample(None, mean_est, cov_est, objs, triples, sdfs, text, rel, None, gen_shape=True)
    out.update(sample_boxes=boxes.numpy(), sample_angles=ang.numpy(), sample_gen=gen.numpy(), sample_rel=rec["d"]["rel"].numpy(),
               sample_uc=rec["d"]["uc"].numpy(), sample_sdf=rec["d"]["sdf"].numpy())

    # manipulation: one node added at position 3 (graph inputs already have O nodes: z has O - 1), node 5 edited
    np.random.seed(8)
    (boxes, ang), gen, keep = real.decoder_with_changes(z[:-1], objs, triples, text, rel, sdfs, None, [3], [5], gen_shape=True)
    out.update(chg_boxes=boxes.numpy(), chg_angles=ang.numpy(), chg_keep=keep.numpy(), chg_rel=rec["d"]["rel"].numpy(), chg_uc=rec["d"]["uc"].numpy())
    np.random.seed(9)
    (boxes, ang), gen, keep = real.decoder_with_changes(z[:-1], objs, triples, text, rel, sdfs, None, [3], [5], distribution=(mean_est, cov_est),
                                                        gen_shape=False)
    out.update(chgd_boxes=boxes.numpy(), chgd_keep=keep.numpy())
    np.random.seed(10)
    (boxes, ang), gen, keep = real.decoder_with_additions(z[:-2], objs, triples, text, rel, sdfs, None, [2, 6], [0], gen_shape=True)
    out.update(add_boxes=boxes.numpy(), add_angles=ang.numpy(), add_keep=keep.numpy(), add_rel=rec["d"]["rel"].numpy())
    # ---- the training forward (VAEGAN_V2FULL.py:466-560): two scenes, one node added, one node manipulated ----
    import random
    real.train()
    torch.nn.Module.load_state_dict(real, Wt.synth_state_dict(shapes, SEED), strict=True)
    seen = {}
    real.Diff.set_input = lambda d: seen.update(d={k: v.clone() for k, v in d.items()})
    real.Diff.set_requires_grad = lambda *a, **k: None
    real.Diff.forward = lambda: None
    real.diffusion_bs = 6
    g2 = torch.Generator().manual_seed(502)
    enc_objs, enc_triples = objs[:-1], triples[(triples[:, 0] < O - 1) & (triples[:, 2] < O - 1)]       # the encoder graph lacks the added node
    enc_boxes, enc_angles = torch.randn(O - 1, 6, generator=g2), torch.randint(0, 24, (O - 1,), generator=g2)
    dec_boxes, dec_angles = torch.randn(O, 6, generator=g2), torch.randint(0, 24, (O,), generator=g2)
    enc_text, enc_rel = text[:-1], torch.randn(enc_triples.shape[0], 512, generator=g2)
    scene_of = torch.tensor([0, 0, 0, 0, 1, 1, 1, 1, 1])
    grained = objs * 2
    torch.manual_seed(11); np.random.seed(12); random.seed(13)
    res = real.forward(enc_objs, enc_triples, enc_boxes, enc_text, enc_rel, None, None, objs, grained, triples, dec_boxes, text, rel, None,
                       scene_of, [8], [2], sdfs, enc_angles=enc_angles, dec_angles=dec_angles)
    names = ("mu", "logvar", "orig_gt_d3", "orig_gt_angles", "orig_gt_shapes", "orig_d3", "orig_angles", "d3_pred", "angles_pred")
    out.update({f"fwd_{n}": r.numpy() for n, r in zip(names, res[:9])})
    out.update(fwd_obj_selected=res[9][0].numpy(), fwd_keep=res[10].numpy(), fwd_rel=seen["d"]["rel"].numpy(), fwd_uc=seen["d"]["uc"].numpy(),
               fwd_sdf=seen["d"]["sdf"].numpy(), fwd_enc_triples=enc_triples.numpy(), fwd_enc_boxes=enc_boxes.numpy(), fwd_enc_angles=enc_angles.numpy(),
               fwd_dec_boxes=dec_boxes.numpy(), fwd_dec_angles=dec_angles.numpy(), fwd_enc_rel=enc_rel.numpy(), fwd_scene_of=scene_of.numpy())
    out["lr_lambda"] = np.asarray([real.lr_lambda(c) for c in (0, 19999, 20000, 59999, 60000, 99999, 100000, 10 ** 7)])
    np.savez_compressed(os.path.join(HERE, "scene_eval.npz"), **out)
    print("scene_eval.npz:", {k: v.shape for k, v in out.items() if k.endswith(("boxes", "keep"))})


if __name__ == "__main__":
    main()
