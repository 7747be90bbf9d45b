"""Oracle (test infrastructure): CPU fp32 restatement of the scene-graph conditioning producer
(SURVEY.md §8 a17, a18): GraphTripleConv / GraphTripleConvNet2 (model/graph.py:89-288), build_mlp
(model/layers.py:21-38) and Sg2ScVAEModel.encoder_2 + rel_mlp (model/VAEGAN_V2FULL.py:152-155, 220-242)
in the v2_full wiring of model/VAE.py:57-63 (embedding_dim 64, hidden 256, CLIP 512, BatchNorm, residual, avg).
Pinned against the reference modules and against encoder_2 of the REAL Sg2ScVAEModel class
(oracle/reference_scene_model.py + validate_against_reference.py): max |diff| = 0.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

GCN_FULL = dict(embedding_dim=64, add_dim=512, num_layers=5, num_objs=36, num_preds=16, rel_hidden=960, rel_out=1280)
GCN_TINY = dict(embedding_dim=16, add_dim=32, num_layers=2, num_objs=10, num_preds=6, rel_hidden=48, rel_out=64)


def _mlp_shapes(s, name, dims, final_nonlinearity=True):
    """build_mlp with batch_norm='batch' (layers.py:21-38): Linear [+ BN + ReLU]; module indices as in nn.Sequential."""
    idx = 0
    for i in range(len(dims) - 1):
        s[f"{name}.{idx}.weight"] = (dims[i + 1], dims[i]); s[f"{name}.{idx}.bias"] = (dims[i + 1],)
        idx += 1
        if i != len(dims) - 2 or final_nonlinearity:
            for k in ("weight", "bias", "running_mean", "running_var"):
                s[f"{name}.{idx}.{k}"] = (dims[i + 1],)
            s[f"{name}.{idx}.num_batches_tracked"] = ()
            idx += 2                                          # BatchNorm1d, ReLU


def gcn_param_shapes(cfg: dict) -> Dict[str, Tuple[int, ...]]:
    e, add = cfg["embedding_dim"], cfg["add_dim"]
    dim, hid = 2 * e + add, 4 * e
    s: Dict[str, Tuple[int, ...]] = {}
    s["obj_embeddings_dc.weight"] = (cfg["num_objs"] + 1, e)
    s["pred_embeddings_dc.weight"] = (cfg["num_preds"], 2 * e)           # decoder_cat=True (VAEGAN_V2FULL.py:72-74)
    for l in range(cfg["num_layers"]):
        p = f"gconv_net_ec_rel.gconvs.{l}"
        _mlp_shapes(s, p + ".net1", [2 * dim + dim, hid, 2 * hid + dim])
        _mlp_shapes(s, p + ".net2", [hid, hid, dim])
        s[p + ".linear_projection.weight"] = (dim, dim); s[p + ".linear_projection.bias"] = (dim,)
        s[p + ".linear_projection_pred.weight"] = (dim, dim); s[p + ".linear_projection_pred.bias"] = (dim,)
    _mlp_shapes(s, "rel_mlp", [dim, cfg["rel_hidden"], cfg["rel_out"]], final_nonlinearity=False)
    return s


def _mlp(sd, name: str, x: Tensor, n_linear: int, final_nonlinearity: bool, training: bool) -> Tensor:
    idx = 0
    for i in range(n_linear):
        x = F.linear(x, sd[f"{name}.{idx}.weight"], sd[f"{name}.{idx}.bias"])
        idx += 1
        if i != n_linear - 1 or final_nonlinearity:
            # BatchNorm1d: batch statistics in train mode, running statistics in eval mode
            x = F.batch_norm(x, sd[f"{name}.{idx}.running_mean"].clone(), sd[f"{name}.{idx}.running_var"].clone(),
                             sd[f"{name}.{idx}.weight"], sd[f"{name}.{idx}.bias"], training, 0.1, 1e-5)
            x = F.relu(x)
            idx += 2
    return x


def gcn_net_shapes(s: Dict[str, Tuple[int, ...]], prefix: str, din_obj: int, din_pred: int, hidden: int, num_layers: int,
                   output_dim=None) -> None:
    """State-dict shapes of one GraphTripleConvNet (graph.py:214-243; BatchNorm + residual): every layer maps to din_obj
    except the last one when output_dim is given."""
    for l in range(num_layers):
        dout = output_dim if (output_dim is not None and l >= num_layers - 1) else din_obj
        p = f"{prefix}.gconvs.{l}"
        _mlp_shapes(s, p + ".net1", [2 * din_obj + din_pred, hidden, 2 * hidden + dout])
        _mlp_shapes(s, p + ".net2", [hidden, hidden, dout])
        s[p + ".linear_projection.weight"] = (dout, din_obj); s[p + ".linear_projection.bias"] = (dout,)
        s[p + ".linear_projection_pred.weight"] = (dout, din_pred); s[p + ".linear_projection_pred.bias"] = (dout,)


def gcn_net(sd, prefix: str, obj: Tensor, pred: Tensor, edges: Tensor, hidden: int, num_layers: int, training: bool,
            output_dim=None):
    """GraphTripleConvNet.forward (graph.py:245-249)."""
    for l in range(num_layers):
        dout = output_dim if (output_dim is not None and l >= num_layers - 1) else None
        obj, pred = graph_triple_conv(sd, f"{prefix}.gconvs.{l}", obj, pred, edges, hidden, training, dout=dout)
    return obj, pred


def graph_triple_conv(sd, name: str, obj: Tensor, pred: Tensor, edges: Tensor, hidden: int, training: bool, dout=None):
    """GraphTripleConv.forward, pooling='avg', residual=True (graph.py:124-211).  dout: output_dim (default: the object width)."""
    O = obj.shape[0]
    dout = obj.shape[1] if dout is None else dout
    s_idx, o_idx = edges[:, 0].contiguous(), edges[:, 1].contiguous()
    t = _mlp(sd, name + ".net1", torch.cat([obj[s_idx], pred, obj[o_idx]], dim=1), 2, True, training)
    new_s, new_p, new_o = t[:, :hidden], t[:, hidden:hidden + dout], t[:, hidden + dout:2 * hidden + dout]
    pooled = torch.zeros(O, hidden)
    pooled = pooled.scatter_add(0, s_idx.view(-1, 1).expand_as(new_s), new_s)
    pooled = pooled.scatter_add(0, o_idx.view(-1, 1).expand_as(new_o), new_o)
    counts = torch.zeros(O).scatter_add(0, s_idx, torch.ones(edges.shape[0])).scatter_add(0, o_idx, torch.ones(edges.shape[0]))
    pooled = pooled / counts.clamp(min=1).view(-1, 1)
    new_obj = _mlp(sd, name + ".net2", pooled, 2, True, training)
    new_obj = new_obj + F.linear(obj, sd[name + ".linear_projection.weight"], sd[name + ".linear_projection.bias"])
    new_p = new_p + F.linear(pred, sd[name + ".linear_projection_pred.weight"], sd[name + ".linear_projection_pred.bias"])
    return new_obj, new_p


def encoder_2(sd, cfg: dict, z: Tensor, objs: Tensor, triples: Tensor, text_feat: Tensor, rel_feat: Tensor,
              training: bool = False) -> Tuple[Tensor, Tensor]:
    """Sg2ScVAEModel.encoder_2 (VAEGAN_V2FULL.py:220-242), clip=True, use_E2=True -> (uc_rel, c_rel), each (O,1,rel_out)."""
    s, p, o = triples[:, 0], triples[:, 1], triples[:, 2]
    edges = torch.stack([s, o], dim=1)
    obj_vecs = torch.cat([text_feat, sd["obj_embeddings_dc.weight"][objs]], dim=1)
    pred_vecs = torch.cat([rel_feat, sd["pred_embeddings_dc.weight"][p]], dim=1)
    rel_in = torch.cat([obj_vecs, z], dim=1)
    ov, pv = rel_in, pred_vecs
    for l in range(cfg["num_layers"]):
        ov, pv = graph_triple_conv(sd, f"gconv_net_ec_rel.gconvs.{l}", ov, pv, edges, 4 * cfg["embedding_dim"], training)
    c = _mlp(sd, "rel_mlp", ov, 2, False, training).unsqueeze(1)
    uc = _mlp(sd, "rel_mlp", rel_in, 2, False, training).unsqueeze(1)
    return uc, c
