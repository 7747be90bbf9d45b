"""Parity of the CUDA scene-graph conditioning path with the reference goldens.
All fp32: tolerance 1e-4 relative to the output range (summation-order differences only)."""
import os

import numpy as np
import pytest
import torch

from oracle import graph as G, weights as Wt

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


class E2(torch.nn.Module):
    """The encoder_2 members of Sg2ScVAEModel, built from the product's graph modules with the reference's key names."""

    def __init__(self, cfg):
        super().__init__()
        from commonscenes_b200.model.graph import GraphTripleConvNet2, make_mlp
        e, add = cfg["embedding_dim"], cfg["add_dim"]
        self.obj_embeddings_dc = torch.nn.Embedding(cfg["num_objs"] + 1, e)
        self.pred_embeddings_dc = torch.nn.Embedding(cfg["num_preds"], 2 * e)
        self.gconv_net_ec_rel = GraphTripleConvNet2(input_dim_obj=2 * e + add, input_dim_pred=2 * e + add, hidden_dim=4 * e,
                                                    pooling="avg", num_layers=cfg["num_layers"], mlp_normalization="batch", residual=True)
        self.rel_mlp = make_mlp([2 * e + add, cfg["rel_hidden"], cfg["rel_out"]], batch_norm="batch", norelu=True)


@pytest.mark.parametrize("tag,cfg", [("tiny", G.GCN_TINY), ("full", G.GCN_FULL)])
def test_encoder2_matches_reference_golden(tag, cfg):
    from commonscenes_b200.model.VAEGAN_V2FULL import Sg2ScVAEModel
    g = np.load(os.path.join(GOLD, f"gcn_{tag}.npz"))
    m = E2(cfg)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == G.gcn_param_shapes(cfg)
    z, objs, triples, text, rel = (torch.tensor(g[k]).cuda() for k in ("z", "objs", "triples", "text", "rel"))
    for mode in ("eval", "train"):
        Wt.fill_module_(m, int(g["weight_seed"]))     # train mode updates running stats: reset before each pass
        m = m.cuda().train(mode == "train")
        m.clip, m.use_E2 = True, True
        uc, c = Sg2ScVAEModel.encoder_2(m, z, objs, triples, text, rel)
        for got, key in ((c, f"c_{mode}"), (uc, f"uc_{mode}")):
            ref = torch.tensor(g[key])
            err = float((got.cpu() - ref).abs().max())
            print(f"encoder_2[{tag},{mode}] {key}: max abs err {err:.3e} (ref absmax {float(ref.abs().max()):.3f})")
            assert got.shape == ref.shape and err <= 1e-4 * max(1.0, float(ref.abs().max()))


def test_graph_conv_isolated_nodes_and_running_stats():
    from commonscenes_b200.model.graph import GraphTripleConv
    torch.manual_seed(0)
    layer = GraphTripleConv(32, 32, hidden_dim=16, mlp_normalization="batch", residual=True).cuda().train()
    obj = torch.randn(6, 32, device="cuda")
    pred = torch.randn(4, 32, device="cuda")
    edges = torch.tensor([[0, 1], [1, 2], [0, 2], [2, 0]], device="cuda")          # nodes 3..5 appear in no triple
    sd = {k: v.clone().cpu() for k, v in layer.state_dict().items()}
    new_obj, new_pred = layer(obj, pred, edges)
    o_ref, p_ref = G.graph_triple_conv({f"L.{k}": v for k, v in sd.items()}, "L", obj.cpu(), pred.cpu(), edges.cpu(), 16, True)
    assert torch.allclose(new_obj.cpu(), o_ref, atol=1e-4) and torch.allclose(new_pred.cpu(), p_ref, atol=1e-4)
    bn = layer.net1[1]
    assert int(bn.num_batches_tracked) == 1 and float((bn.running_mean.cpu() - sd["net1.1.running_mean"]).abs().max()) > 0
    with pytest.raises(ValueError):
        layer(obj[:2], pred[:1], edges[:1])                                           # BatchNorm1d needs > 1 row in training
