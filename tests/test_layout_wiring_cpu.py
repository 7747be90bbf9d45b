"""CPU suite: HOST WIRING of the layout-branch forward in the Sg2ScVAEModel mirror.

There is no GPU here and the product has no CPU path, so for this test only the five C-ABI entry points the graph networks
call (cs_linear_small, cs_gcn_gather_triples, cs_gcn_scatter_mean, cs_batchnorm_relu, cs_add_rows) are monkeypatched with
plain-torch stand-ins.  What is being checked is everything AROUND the kernels — which embedding / feature / latent is
concatenated where, which network consumes it, the manipulator's output width, BatchNorm train/eval handling, running-stat
updates — against what the reference's REAL class computed (tests/golden/layout_*.npz).  The kernels themselves are checked
on the GPU (tests/test_gcn_gpu.py)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F
import yaml

from oracle import layout as Lo, weights as Wt
from test_checkpoint_cpu import TINY_DF, TINY_VQ

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _stand_ins(monkeypatch):
    from commonscenes_b200 import ops

    def linear_small(x, w, bias=None, act_in=0, act_out=0, out=None):
        assert act_in == 0 and act_out == 0
        return F.linear(x, w, bias)

    def gather(obj, pred, edges):
        return torch.cat([obj[edges[:, 0]], pred, obj[edges[:, 1]]], dim=1)

    def scatter_mean(tv, s_off, o_off, hd, edges, num_objs):
        pooled = torch.zeros(num_objs, hd).index_add(0, edges[:, 0], tv[:, s_off:s_off + hd]).index_add(0, edges[:, 1], tv[:, o_off:o_off + hd])
        ones = torch.ones(edges.shape[0])
        cnt = torch.zeros(num_objs).index_add(0, edges[:, 0], ones).index_add(0, edges[:, 1], ones).clamp(min=1)
        return pooled / cnt[:, None]

    def bn_relu(x, gamma, beta, rm, rv, training, momentum=0.1, eps=1e-5, relu=True):
        y = F.batch_norm(x, rm, rv, gamma, beta, training, momentum, eps)
        return F.relu(y) if relu else y
    for name, fn in (("linear_small", linear_small), ("gcn_gather_triples", gather), ("gcn_scatter_mean", scatter_mean),
                     ("batchnorm_relu", bn_relu), ("add_rows", lambda a, b: a + b)):
        monkeypatch.setattr(ops, name, fn)


@pytest.mark.parametrize("tag", ["tiny", "full"])
def test_layout_forward_wiring_matches_reference_class(tag, tmp_path, monkeypatch):
    _stand_ins(monkeypatch)
    from commonscenes_b200.model.VAEGAN_V2FULL import Sg2ScVAEModel
    from commonscenes_b200.model.sdfusion_txt2shape_model import default_opt
    cfg = Lo.LAYOUT_TINY if tag == "tiny" else Lo.LAYOUT_FULL
    g = np.load(os.path.join(GOLD, f"layout_{tag}.npz"))
    (tmp_path / "df.yaml").write_text(yaml.safe_dump(TINY_DF)); (tmp_path / "vq.yaml").write_text(yaml.safe_dump(TINY_VQ))
    vocab = {"object_idx_to_name": [f"o{i}" for i in range(cfg["num_objs"])], "pred_idx_to_name": [f"p{i}" for i in range(cfg["num_preds"])]}
    m = Sg2ScVAEModel(vocab, diff_opt=default_opt(device="cpu", df_cfg=str(tmp_path / "df.yaml"), vq_cfg=str(tmp_path / "vq.yaml")),
                      embedding_dim=cfg["embedding_dim"], mlp_normalization="batch", residual=True, gconv_num_layers=cfg["num_layers"],
                      layout_branch=True)
    shapes = Lo.layout_param_shapes(cfg)
    z, objs, triples, text, rel, boxes, angles, zz = (torch.tensor(g[k]) for k in ("z", "objs", "triples", "text", "rel", "boxes", "angles", "zz"))

    def reset():
        m.load_state_dict({k: Wt.synth_tensor(int(g["weight_seed"]), k, tuple(s)) for k, s in shapes.items()}, strict=False)
    for mode in ("eval", "train"):
        m.train(mode == "train")
        reset(); mu, logvar = m.encoder(objs, triples, boxes, None, text, rel, angles)
        assert int(m.mean_var[1].num_batches_tracked) == (1 if mode == "train" else 0)      # BatchNorm bookkeeping as in torch
        reset(); man = m.manipulate(zz, objs, triples, text, rel)
        reset(); b, a = m.decoder(z, objs, triples, text, rel)
        for got, key in ((mu, "mu"), (logvar, "logvar"), (man, "man"), (b, "boxes"), (a, "angle_logp")):
            ref = torch.tensor(g[f"{key}_{mode}"])
            assert got.shape == ref.shape and float((got - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max())), (key, mode)
