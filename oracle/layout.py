"""Oracle (test infrastructure): CPU fp32 restatement of the LAYOUT branch of Sg2ScVAEModel — the
step either side of the shape-branch hot path inside the same training iteration: the box/angle graph-VAE encoder
(model/VAEGAN_V2FULL.py:185-218), the latent manipulator (:244-258), the box/angle decoder (:260-289) and the layout
losses (model/losses.py:26-51), in the v2_full wiring of model/VAE.py:57-63 (embedding_dim 64, use_angles, decoder_cat,
CLIP 512, BatchNorm MLPs, residual GCNs).  Pinned against the reference's REAL class by validate_against_reference.py
(oracle/reference_scene_model.py builds it).  The product side of this row is not built yet: this is its first gate.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from .graph import _mlp, _mlp_shapes, gcn_net, gcn_net_shapes

Tensor = torch.Tensor

LAYOUT_FULL = dict(embedding_dim=64, add_dim=512, num_layers=5, num_objs=36, num_preds=16, num_box_params=6, n_angle=24)
LAYOUT_TINY = dict(embedding_dim=64, add_dim=512, num_layers=2, num_objs=10, num_preds=6, num_box_params=6, n_angle=24)
# (the class hard-codes the CLIP width 512; embedding_dim must be a multiple of 4 for the 3/4 : 1/4 box / angle split)


def layout_param_shapes(cfg: dict) -> Dict[str, Tuple[int, ...]]:
    e, add, L = cfg["embedding_dim"], cfg["add_dim"], cfg["num_layers"]
    box_e, ang_e, hid, dim = e * 3 // 4, e // 4, 4 * e, 2 * e + add
    s: Dict[str, Tuple[int, ...]] = {}
    s["obj_embeddings_ec.weight"] = (cfg["num_objs"] + 1, e)
    s["pred_embeddings_ec.weight"] = (cfg["num_preds"], 2 * e)
    s["obj_embeddings_dc.weight"] = (cfg["num_objs"] + 1, e)
    s["pred_embeddings_dc.weight"] = (cfg["num_preds"], 2 * e)
    s["pred_embeddings_man_dc.weight"] = (cfg["num_preds"], 3 * e)
    s["d3_embeddings.weight"] = (box_e, cfg["num_box_params"]); s["d3_embeddings.bias"] = (box_e,)
    s["angle_embeddings.weight"] = (cfg["n_angle"], ang_e)
    _mlp_shapes(s, "mean_var", [dim, hid, 2 * e])
    _mlp_shapes(s, "mean", [2 * e, box_e], final_nonlinearity=False)
    _mlp_shapes(s, "var", [2 * e, box_e], final_nonlinearity=False)
    _mlp_shapes(s, "angle_mean_var", [dim, hid, 2 * e])
    _mlp_shapes(s, "angle_mean", [2 * e, ang_e], final_nonlinearity=False)
    _mlp_shapes(s, "angle_var", [2 * e, ang_e], final_nonlinearity=False)
    gcn_net_shapes(s, "gconv_net_ec_box", dim, dim, hid, L)
    gcn_net_shapes(s, "gconv_net_dc", dim, dim, hid, L)
    gcn_net_shapes(s, "gconv_net_manipulation", 3 * e + add, 3 * e + add, hid, min(L, 5), output_dim=e)
    _mlp_shapes(s, "d3_net", [dim, hid, cfg["num_box_params"]], final_nonlinearity=False)
    _mlp_shapes(s, "angle_net", [dim, hid, cfg["n_angle"]], final_nonlinearity=False)
    return s


def _edges(triples: Tensor):
    s, p, o = triples[:, 0], triples[:, 1], triples[:, 2]
    return p, torch.stack([s, o], dim=1)


def encoder(sd, cfg: dict, objs: Tensor, triples: Tensor, boxes_gt: Tensor, text_feat: Tensor, rel_feat: Tensor, angles_gt: Tensor,
            training: bool = False) -> Tuple[Tensor, Tensor]:
    """Sg2ScVAEModel.encoder (:185-218), clip=True, use_angles=True -> (mu, logvar), each (O, embedding_dim)."""
    p, edges = _edges(triples)
    obj = torch.cat([text_feat, sd["obj_embeddings_ec.weight"][objs]], dim=1)
    pred = torch.cat([rel_feat, sd["pred_embeddings_ec.weight"][p]], dim=1)
    d3 = F.linear(boxes_gt, sd["d3_embeddings.weight"], sd["d3_embeddings.bias"])
    obj = torch.cat([obj, d3, sd["angle_embeddings.weight"][angles_gt]], dim=1)
    obj, _ = gcn_net(sd, "gconv_net_ec_box", obj, pred, edges, 4 * cfg["embedding_dim"], cfg["num_layers"], training)
    h = _mlp(sd, "mean_var", obj, 2, True, training)
    mu, logvar = _mlp(sd, "mean", h, 1, False, training), _mlp(sd, "var", h, 1, False, training)
    ha = _mlp(sd, "angle_mean_var", obj, 2, True, training)
    mu_a, logvar_a = _mlp(sd, "angle_mean", ha, 1, False, training), _mlp(sd, "angle_var", ha, 1, False, training)
    return torch.cat([mu, mu_a], dim=1), torch.cat([logvar, logvar_a], dim=1)


def manipulate(sd, cfg: dict, z: Tensor, objs: Tensor, triples: Tensor, text_feat: Tensor, rel_feat: Tensor, training: bool = False):
    """Sg2ScVAEModel.manipulate (:244-258): z is [latent | change noise] (O, 2 * embedding_dim) -> (O, embedding_dim)."""
    p, edges = _edges(triples)
    obj = torch.cat([text_feat, sd["obj_embeddings_dc.weight"][objs]], dim=1)
    pred = torch.cat([rel_feat, sd["pred_embeddings_man_dc.weight"][p]], dim=1)
    man, _ = gcn_net(sd, "gconv_net_manipulation", torch.cat([z, obj], dim=1), pred, edges, 4 * cfg["embedding_dim"],
                     min(cfg["num_layers"], 5), training, output_dim=cfg["embedding_dim"])
    return man


def decoder(sd, cfg: dict, z: Tensor, objs: Tensor, triples: Tensor, text_feat: Tensor, rel_feat: Tensor, training: bool = False):
    """Sg2ScVAEModel.decoder (:260-289), decoder_cat=True, use_angles=True -> (boxes (O, 6), log-probabilities of 24 angle bins)."""
    p, edges = _edges(triples)
    obj = torch.cat([text_feat, sd["obj_embeddings_dc.weight"][objs], z], dim=1)
    pred = torch.cat([rel_feat, sd["pred_embeddings_dc.weight"][p]], dim=1)
    obj, _ = gcn_net(sd, "gconv_net_dc", obj, pred, edges, 4 * cfg["embedding_dim"], cfg["num_layers"], training)
    return _mlp(sd, "d3_net", obj, 2, False, training), F.log_softmax(_mlp(sd, "angle_net", obj, 2, False, training), dim=1)


def layout_losses(pred: Tensor, target: Tensor, angles_pred: Tensor, angles: Tensor, mu: Tensor, logvar: Tensor, kl_weight: float = 0.1):
    """calculate_model_losses (losses.py:26-51) with withangles=True: L1 box reconstruction + NLL over the angle bins +
    KL_weight * KL(N(mu, exp(logvar)) || N(0, 1)) / O.  Returns (total, dict of the weighted terms)."""
    rec = F.l1_loss(pred, target)
    ang = F.nll_loss(angles_pred, angles)
    kld = -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp()) / mu.size(0)
    return rec + ang + kl_weight * kld, {"box": rec, "angle_pred": ang, "KLD_Gauss": kl_weight * kld}
