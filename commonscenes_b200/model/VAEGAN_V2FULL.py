"""Sg2ScVAEModel — the shape-branch slice of the v2_full scene-graph model.

The reference's model/VAEGAN_V2FULL.py:17-760 couples a layout branch (box/angle GCN-VAE, discriminators: SURVEY.md
§8f rank 2, not on this hot path) with the shape branch.  This class keeps the reference's names and semantics for
the shape branch only: the decoder-side embeddings (:69-75), the relation encoder E2 (`gconv_net_ec_rel`, :128-147),
`rel_mlp` (:152-155), `encoder_2` (:220-242), `balance_objects` / `select_sdfs` (:398-463), the denoiser call of
`forward` (:511-521) and the shape half of `sample` (:600-616).  State-dict keys of these members are the reference's,
so a v2_full checkpoint's shape-branch tensors load with strict=False.
"""
from __future__ import annotations

import random

import numpy as np
import torch
import torch.nn as nn

from .graph import GraphTripleConvNet2, make_mlp
from .layers import run_mlp
from .sdfusion_txt2shape_model import SDFusionText2ShapeModel, default_opt


class Sg2ScVAEModel(nn.Module):
    def __init__(self, vocab, diff_opt=None, diffusion_bs=8, embedding_dim=128, gconv_pooling="avg", gconv_num_layers=5,
                 mlp_normalization="none", use_E2=True, residual=False, clip=True, decoder_cat=True, **unused):
        super().__init__()
        self.embedding_dim, self.clip, self.use_E2, self.decoder_cat = embedding_dim, clip, use_E2, decoder_cat
        add_dim = 512 if clip else 0
        self.obj_classes_list = list(set(vocab["object_idx_to_name"]))
        self.edge_list = list(set(vocab["pred_idx_to_name"]))
        num_objs, num_preds = len(self.obj_classes_list), len(self.edge_list)
        gconv_dim, hidden = embedding_dim, embedding_dim * 4
        self.obj_embeddings_dc = nn.Embedding(num_objs + 1, embedding_dim)
        self.pred_embeddings_dc = nn.Embedding(num_preds, embedding_dim * 2 if decoder_cat else embedding_dim)
        self.Diff = SDFusionText2ShapeModel(default_opt() if diff_opt is None else diff_opt)
        hb = getattr(getattr(self.Diff.opt, "hyper", None), "batch_size", None)
        self.diffusion_bs = diffusion_bs if hb is None else hb
        if use_E2:
            self.gconv_net_ec_rel = GraphTripleConvNet2(input_dim_obj=gconv_dim * 2 + add_dim, input_dim_pred=gconv_dim * 2 + add_dim,
                                                        hidden_dim=hidden, pooling=gconv_pooling, num_layers=gconv_num_layers,
                                                        mlp_normalization=mlp_normalization, residual=residual)
        self.rel_mlp = make_mlp([gconv_dim * 2 + add_dim, 960, 1280], batch_norm=mlp_normalization, norelu=True)

    @torch.no_grad()
    def encoder_2(self, z, objs, triples, dec_text_feat, dec_rel_feat, attributes=None, manipulate=False):
        s, p, o = [x.squeeze(1) for x in triples.chunk(3, dim=1)]
        edges = torch.stack([s, o], dim=1)
        obj_vecs = self.obj_embeddings_dc(objs)              # embedding row gathers: index plumbing
        pred_vecs = self.pred_embeddings_dc(p)
        if self.clip:
            obj_vecs_ = torch.cat([dec_text_feat, obj_vecs], dim=1)
            pred_vecs_ = torch.cat([dec_rel_feat, pred_vecs], dim=1)
        else:
            obj_vecs_, pred_vecs_ = obj_vecs, pred_vecs
        rel_vecs_ = torch.cat([obj_vecs_, z], dim=1).float().contiguous()
        rel_vecs_2 = None
        if self.use_E2:
            rel_vecs_2, _ = self.gconv_net_ec_rel(rel_vecs_, pred_vecs_, edges)
            rel_vecs_2 = run_mlp(self.rel_mlp, rel_vecs_2).unsqueeze(1)
        return run_mlp(self.rel_mlp, rel_vecs_).unsqueeze(1), rel_vecs_2

    def balance_objects(self, id_list, object_list, n):
        """Pick n objects covering distinct fine-grained classes first (host-side, python `random`, as the reference)."""
        assert len(id_list) == len(object_list), "id_list and object_list must have the same length"
        unique_ids = torch.unique(id_list)
        if len(unique_ids) >= n:
            sampled = random.sample(unique_ids.tolist(), n)
        else:
            sampled = unique_ids.tolist() + random.choices(id_list.tolist(), k=n - len(unique_ids))
        picked = []
        for sid in sampled:
            idxs = (id_list == sid).nonzero(as_tuple=True)[0]
            picked.append(idxs[random.choice(range(len(idxs)))])
        return torch.tensor(picked)

    def select_sdfs(self, dec_objs_to_scene, dec_objs, dec_objs_grained, dec_sdfs, uc_rel_feat, c_rel_feat, random=False):
        scene = dec_objs_to_scene.detach().cpu().numpy()
        batch_size = int(np.max(scene)) + 1
        num_obj = int(np.ceil(self.diffusion_bs / batch_size))
        sdf_sel, uc_sel, c_sel, cat_sel = [], [], [], []
        for i in range(batch_size):
            rows = np.where(scene == i)[0]
            sdf_c, cat, cat_g, uc, c = dec_sdfs[rows], dec_objs[rows], dec_objs_grained[rows], uc_rel_feat[rows], c_rel_feat[rows]
            ids = torch.unique(torch.where(torch.ne(sdf_c, torch.zeros_like(sdf_c[0])))[0])     # objects that have an SDF
            if random:
                pick = ids[torch.randperm(len(ids))[:num_obj]]
                sdf_sel.append(sdf_c[pick]); uc_sel.append(uc[pick]); c_sel.append(c[pick]); cat_sel.append(cat[pick])
            else:
                sel = self.balance_objects(cat_g[ids], cat[ids], num_obj).to(ids.device)
                sdf_sel.append(sdf_c[ids][sel]); uc_sel.append(uc[ids][sel]); c_sel.append(c[ids][sel]); cat_sel.append(cat[ids][sel])
        n = self.diffusion_bs
        diff_dict = {"sdf": torch.cat(sdf_sel)[:n].cuda(), "uc": torch.cat(uc_sel)[:n].cuda(), "rel": torch.cat(c_sel)[:n].cuda()}
        return torch.cat(cat_sel)[:n], diff_dict

    def forward_shape(self, z, dec_objs, dec_objs_grained, dec_triples, dec_text_feat, dec_rel_feat, dec_sdfs, dec_objs_to_scene):
        """The shape-branch lines of Sg2ScVAEModel.forward (:511-521): conditioning -> object selection -> diffusion loss."""
        uc, c = self.encoder_2(z, dec_objs, dec_triples, dec_text_feat, dec_rel_feat)
        c = uc if c is None else c
        obj_selected, diff_dict = self.select_sdfs(dec_objs_to_scene, dec_objs, dec_objs_grained, dec_sdfs, uc, c, random=False)
        self.Diff.set_input(diff_dict)
        self.Diff.forward()
        return obj_selected, self.Diff.loss_df

    @torch.no_grad()
    def sample_shape(self, z, dec_objs, dec_triplets, dec_sdfs, dec_text_feat, dec_rel_feat, uc_scale=3., ddim_steps=100, seed=None):
        """The gen_shape=True half of Sg2ScVAEModel.sample (:604-615)."""
        uc, c = self.encoder_2(z, dec_objs, dec_triplets, dec_text_feat, dec_rel_feat)
        ids = torch.unique(torch.where(torch.ne(dec_sdfs, torch.zeros_like(dec_sdfs[0])))[0])
        c = uc if c is None else c
        diff_dict = {"sdf": dec_sdfs[ids], "rel": c[ids], "uc": uc[ids]}
        return self.Diff.rel2shape(diff_dict, ddim_steps=ddim_steps, uc_scale=uc_scale, seed=seed), dec_objs[ids]
