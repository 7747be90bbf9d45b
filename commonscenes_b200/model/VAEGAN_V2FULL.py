"""Sg2ScVAEModel — the shape-branch slice of the v2_full scene-graph model.

The reference's model/VAEGAN_V2FULL.py:17-760 couples a layout branch (box/angle GCN-VAE, discriminators: SURVEY.md
§8f rank 2, not on this hot path) with the shape branch.  This class keeps the reference's names and semantics for
the shape branch (always) and, with `layout_branch=True`, the forward of the layout branch (encoder / manipulate /
decoder: same kernels, GPU-verified against the real class's goldens in tests/test_gcn_gpu.py).  Shape branch: the decoder-side embeddings (:69-75), the relation encoder E2 (`gconv_net_ec_rel`, :128-147),
`rel_mlp` (:152-155), `encoder_2` (:220-242), `balance_objects` / `select_sdfs` (:398-463), the denoiser call of
`forward` (:511-521) and the shape half of `sample` (:600-616).  State-dict keys of these members are the reference's,
so a v2_full checkpoint's shape-branch tensors load with strict=False.
"""
from __future__ import annotations

import random

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from .graph import GraphTripleConvNet, GraphTripleConvNet2, gcn_apply, make_mlp
from .layers import mlp_apply, run_mlp, run_mlp_train
from .sdfusion_txt2shape_model import SDFusionText2ShapeModel, default_opt


class Sg2ScVAEModel(nn.Module):
    def __init__(self, vocab, diff_opt=None, diffusion_bs=8, embedding_dim=128, gconv_pooling="avg", gconv_num_layers=5,
                 mlp_normalization="none", use_E2=True, residual=False, clip=True, decoder_cat=True, layout_branch=False,
                 num_box_params=6, use_angles=True, **unused):
        super().__init__()
        self.embedding_dim, self.clip, self.use_E2, self.decoder_cat = embedding_dim, clip, use_E2, decoder_cat
        self.layout_branch, self.use_angles = layout_branch, use_angles
        add_dim = 512 if clip else 0
        self.obj_classes_list = list(set(vocab["object_idx_to_name"]))
        self.edge_list = list(set(vocab["pred_idx_to_name"]))
        num_objs, num_preds = len(self.obj_classes_list), len(self.edge_list)
        gconv_dim, hidden = embedding_dim, embedding_dim * 4
        if layout_branch:       # SURVEY.md §8f rank 2 (reference :69-88): registered in the reference's order so that
            if not (decoder_cat and use_angles and clip):     # parameters() / optimizer indices line up with its checkpoints
                raise NotImplementedError("the layout branch mirrors the v2_full wiring (decoder_cat, use_angles, clip)")
            box_e, ang_e = int(embedding_dim * 3 / 4), int(embedding_dim / 4)
            self.obj_embeddings_ec = nn.Embedding(num_objs + 1, embedding_dim)
            self.pred_embeddings_ec = nn.Embedding(num_preds, embedding_dim * 2)
        self.obj_embeddings_dc = nn.Embedding(num_objs + 1, embedding_dim)
        self.pred_embeddings_dc = nn.Embedding(num_preds, embedding_dim * 2 if decoder_cat else embedding_dim)
        if layout_branch:
            self.pred_embeddings_man_dc = nn.Embedding(num_preds, embedding_dim * 3)
            self.d3_embeddings = nn.Linear(num_box_params, box_e)
            self.angle_embeddings = nn.Embedding(24, ang_e)
            bn = mlp_normalization
            self.mean_var = make_mlp([embedding_dim * 2 + add_dim, hidden, embedding_dim * 2], batch_norm=bn)
            self.mean = make_mlp([embedding_dim * 2, box_e], batch_norm=bn, norelu=True)
            self.var = make_mlp([embedding_dim * 2, box_e], batch_norm=bn, norelu=True)
            self.angle_mean_var = make_mlp([embedding_dim * 2 + add_dim, hidden, embedding_dim * 2], batch_norm=bn)
            self.angle_mean = make_mlp([embedding_dim * 2, ang_e], batch_norm=bn, norelu=True)
            self.angle_var = make_mlp([embedding_dim * 2, ang_e], batch_norm=bn, norelu=True)
        self.Diff = SDFusionText2ShapeModel(default_opt() if diff_opt is None else diff_opt)
        hb = getattr(getattr(self.Diff.opt, "hyper", None), "batch_size", None)
        self.diffusion_bs = diffusion_bs if hb is None else hb
        if layout_branch:       # reference :100-143
            kw = dict(hidden_dim=hidden, pooling=gconv_pooling, mlp_normalization=mlp_normalization, residual=residual)
            dim = gconv_dim * 2 + add_dim
            self.gconv_net_ec_box = GraphTripleConvNet(input_dim_obj=dim, input_dim_pred=dim, num_layers=gconv_num_layers, **kw)
            self.gconv_net_dc = GraphTripleConvNet(input_dim_obj=dim, input_dim_pred=dim, num_layers=gconv_num_layers, **kw)
            self.gconv_net_manipulation = GraphTripleConvNet(input_dim_obj=embedding_dim * 3 + add_dim, input_dim_pred=embedding_dim * 3 + add_dim,
                                                             output_dim=embedding_dim, num_layers=min(gconv_num_layers, 5), **kw)
        if use_E2:
            self.gconv_net_ec_rel = GraphTripleConvNet2(input_dim_obj=gconv_dim * 2 + add_dim, input_dim_pred=gconv_dim * 2 + add_dim,
                                                        hidden_dim=hidden, pooling=gconv_pooling, num_layers=gconv_num_layers,
                                                        mlp_normalization=mlp_normalization, residual=residual)
        if layout_branch:       # reference :147-148
            self.d3_net = make_mlp([gconv_dim * 2 + add_dim, hidden, num_box_params], batch_norm=mlp_normalization, norelu=True)
        net_rel_layers = [gconv_dim * 2 + add_dim, 960, 1280]
        if self.Diff.df.conditioning_key == "concat":          # reference :152-154: 4096 = one 16^3 latent channel
            net_rel_layers = [gconv_dim * 2 + add_dim, 1280, 4096]
        self.rel_mlp = make_mlp(net_rel_layers, batch_norm=mlp_normalization, norelu=True)
        if layout_branch:       # reference :157-160
            self.angle_net = make_mlp([gconv_dim * 2 + add_dim, hidden, 24], batch_norm=mlp_normalization, norelu=True)
        # initialisation (reference :162-172): kaiming-normal Linear weights, same modules in the same order (the RNG stream
        # of a from-scratch run then matches the reference's); angle_net keeps torch's default init there too
        from .graph import _init_weights
        if layout_branch:
            for mod in (self.d3_embeddings, self.mean_var, self.mean, self.var, self.d3_net):
                mod.apply(_init_weights)
        self.rel_mlp.apply(_init_weights)
        if layout_branch and use_angles:
            for mod in (self.angle_mean_var, self.angle_mean, self.angle_var):
                mod.apply(_init_weights)

    # ---- layout branch: forward only, on the same GraphTripleConv / MLP kernels as encoder_2 ----------
    def _need_layout(self):
        if not self.layout_branch:
            raise RuntimeError("construct Sg2ScVAEModel(..., layout_branch=True) to get the box / angle branch")

    @staticmethod
    def _edges(triples):
        s, p, o = [x.squeeze(1) for x in triples.chunk(3, dim=1)]
        return p, torch.stack([s, o], dim=1)

    def encoder(self, objs, triples, boxes_gt, attributes, enc_text_feat, enc_rel_feat, angles_gt=None):
        """(mu, logvar), each (O, embedding_dim): box + angle graph-VAE encoder (reference :185-218).  Like every layout
        method it takes part in autograd: with gradients enabled `loss.backward()` runs the explicit GCN / MLP backward
        kernels (graph.gcn_apply, layers.mlp_apply); embedding lookups, concatenations and the losses are torch glue."""
        self._need_layout()
        p, edges = self._edges(triples)
        d3 = mlp_apply([self.d3_embeddings], boxes_gt.float().contiguous())
        obj = torch.cat([enc_text_feat, self.obj_embeddings_ec(objs), d3, self.angle_embeddings(angles_gt)], dim=1).float().contiguous()
        pred = torch.cat([enc_rel_feat, self.pred_embeddings_ec(p)], dim=1).float().contiguous()
        obj, _ = gcn_apply(self.gconv_net_ec_box, obj, pred, edges)
        h = mlp_apply(self.mean_var, obj)
        ha = mlp_apply(self.angle_mean_var, obj)
        mu = torch.cat([mlp_apply(self.mean, h), mlp_apply(self.angle_mean, ha)], dim=1)
        logvar = torch.cat([mlp_apply(self.var, h), mlp_apply(self.angle_var, ha)], dim=1)
        return mu, logvar

    def manipulate(self, z, objs, triples, dec_text_feat, dec_rel_feat, attributes=None):
        """[latent | change noise] (O, 2 * embedding_dim) -> manipulated latent (O, embedding_dim) (reference :244-258)."""
        self._need_layout()
        p, edges = self._edges(triples)
        obj = torch.cat([z, dec_text_feat, self.obj_embeddings_dc(objs)], dim=1).float().contiguous()
        pred = torch.cat([dec_rel_feat, self.pred_embeddings_man_dc(p)], dim=1).float().contiguous()
        man_z, _ = gcn_apply(self.gconv_net_manipulation, obj, pred, edges)
        return man_z

    def decoder(self, z, objs, triples, dec_text_feat, dec_rel_feat, attributes=None, manipulate=False):
        """(boxes (O, 6), log-probabilities over the 24 angle bins) from the latent (reference :260-289, decoder_cat)."""
        self._need_layout()
        p, edges = self._edges(triples)
        obj = torch.cat([dec_text_feat, self.obj_embeddings_dc(objs), z], dim=1).float().contiguous()
        pred = torch.cat([dec_rel_feat, self.pred_embeddings_dc(p)], dim=1).float().contiguous()
        obj, _ = gcn_apply(self.gconv_net_dc, obj, pred, edges)
        return mlp_apply(self.d3_net, obj), torch.log_softmax(mlp_apply(self.angle_net, obj), dim=1)

    # ---- evaluation entry points tying both branches together (reference :291-396, 600-616; scripts/eval_3dfront.py) --------
    def _shape_inputs(self, z, objs, triples, text_feat, rel_feat, dec_sdfs):
        """encoder_2 -> the objects that own an SDF -> the dictionary SDFusionText2ShapeModel.rel2shape consumes."""
        uc, c = self.encoder_2(z, objs, triples, text_feat, rel_feat)
        ids = torch.unique(torch.where(torch.ne(dec_sdfs, torch.zeros_like(dec_sdfs[0])))[0])
        c = uc if c is None else c
        return {"sdf": dec_sdfs[ids], "rel": c[ids], "uc": uc[ids]}

    def _insert_zero_nodes(self, z, missing_nodes, distribution=None):
        """Reference :295-309 / :338-351: a latent row per added node (zeros, or a draw from the fitted Gaussian)."""
        nodes_added = []
        for i in range(len(missing_nodes)):
            ad_id = missing_nodes[i] + i
            nodes_added.append(ad_id)
            if distribution is not None:
                mu, cov = distribution
                row = torch.from_numpy(np.random.multivariate_normal(mu, cov, 1)).float()
            else:
                row = torch.zeros(1, z.shape[1])
            z = torch.cat([z[:ad_id], row.to(z.device), z[ad_id:]], dim=0)
        return z, nodes_added

    @torch.no_grad()
    def sample(self, point_classes_idx, mean_est, cov_est, dec_objs, dec_triplets, dec_sdfs, encoded_dec_text_feat, encoded_dec_rel_feat,
               attributes=None, gen_shape=False):
        """Scene sampling (reference :600-616): z ~ N(mean_est, cov_est) per object (numpy RNG, as the reference), optional
        shape generation for the objects that own an SDF, layout decoding -> ((boxes, angle log-probs), gen_sdf)."""
        self._need_layout()
        z = torch.from_numpy(np.random.multivariate_normal(mean_est, cov_est, dec_objs.size(0))).float().to(dec_objs.device)
        gen_sdf = None
        if gen_shape:
            gen_sdf = self.Diff.rel2shape(self._shape_inputs(z, dec_objs, dec_triplets, encoded_dec_text_feat, encoded_dec_rel_feat, dec_sdfs),
                                          uc_scale=3.)
        return self.decoder(z, dec_objs, dec_triplets, encoded_dec_text_feat, encoded_dec_rel_feat, attributes), gen_sdf

    @torch.no_grad()
    def decoder_with_additions(self, z, objs, triples, encoded_dec_text_feat, encoded_dec_rel_feat, dec_sdfs, attributes, missing_nodes,
                               manipulated_nodes, distribution=None, gen_shape=False):
        """Node addition (reference :291-333): new nodes get a zero (or sampled) latent, everything is decoded as is."""
        self._need_layout()
        z, nodes_added = self._insert_zero_nodes(z, missing_nodes, distribution)
        gen_sdf = None
        if gen_shape:
            gen_sdf = self.Diff.rel2shape(self._shape_inputs(z, objs, triples, encoded_dec_text_feat, encoded_dec_rel_feat, dec_sdfs), uc_scale=3.)
        pred = self.decoder(z, objs, triples, encoded_dec_text_feat, encoded_dec_rel_feat, attributes)
        keep = [0.0 if (i in nodes_added or i in manipulated_nodes) else 1.0 for i in range(len(z))]
        return pred, gen_sdf, torch.tensor(keep, device=z.device).view(-1, 1)

    @torch.no_grad()
    def decoder_with_changes(self, z, dec_objs, dec_triples, encoded_dec_text_feat, encoded_dec_rel_feat, dec_sdfs, attributes, missing_nodes,
                             manipulated_nodes, distribution=None, gen_shape=False):
        """Scene manipulation (reference :335-396): touched nodes (added or edited) get change noise, the manipulator GCN
        predicts their new latents, untouched nodes keep theirs; then shapes (optional) and layout are decoded."""
        self._need_layout()
        z, nodes_added = self._insert_zero_nodes(z, missing_nodes, distribution)
        change = [np.random.normal(0, 1, self.embedding_dim) if (i in nodes_added or i in manipulated_nodes) else np.zeros(self.embedding_dim)
                  for i in range(len(z))]
        change_repr = torch.from_numpy(np.stack(change, axis=0)).float().to(z.device)
        z_prime = self.manipulate(torch.cat([z, change_repr], dim=1), dec_objs, dec_triples, encoded_dec_text_feat, encoded_dec_rel_feat, attributes)
        if not getattr(self, "replace_all_latent", False):
            for node in sorted(nodes_added + list(manipulated_nodes)):
                z = torch.cat([z[:node], z_prime[node:node + 1], z[node + 1:]], dim=0)
        else:
            z = z_prime
        gen_sdf = None
        if gen_shape:
            gen_sdf = self.Diff.rel2shape(self._shape_inputs(z, dec_objs, dec_triples, encoded_dec_text_feat, encoded_dec_rel_feat, dec_sdfs),
                                          uc_scale=3.)
        pred = self.decoder(z, dec_objs, dec_triples, encoded_dec_text_feat, encoded_dec_rel_feat, attributes)
        keep = [0.0 if (i in nodes_added or i in manipulated_nodes) else 1.0 for i in range(len(pred[0]))]
        return pred, gen_sdf, torch.tensor(keep, device=z.device).view(-1, 1)

    @torch.no_grad()
    def collect_train_statistics(self, train_loader, with_points=False):
        """Mean and covariance of the layout latents over a data loader (reference :700-760): the Gaussian `sample` draws
        from.  Batches are the collate dictionaries of the reference's dataset (data['decoder'][...]); a batch equal to -1
        is skipped; box rows are [6 box parameters | angle bin + 1]."""
        self._need_layout()
        dev = self.obj_embeddings_dc.weight.device
        means = []
        for data in train_loader:
            if isinstance(data, int) and data == -1:
                continue
            d = data["decoder"]
            objs, triples, tight_boxes = d["objs"].to(dev), d["tripltes"].to(dev), d["boxes"].to(dev)
            text = d["text_feats"].to(dev) if "text_feats" in d and "rel_feats" in d else None
            rel = d["rel_feats"].to(dev) if text is not None else None
            angles = tight_boxes[:, 6].long() - 1
            angles = torch.where(angles > 0, angles, torch.zeros_like(angles))
            mean, _ = self.encoder(objs, triples, tight_boxes[:, :6], None, text, rel, angles)
            means.append(mean.cpu().clone())
        mean_cat = torch.cat(means, dim=0)
        mean_est = torch.mean(mean_cat, dim=0, keepdim=True)
        cov_est = np.cov((mean_cat - mean_est).numpy().T)
        return mean_est[0], cov_est

    def optimizer_ini(self):
        """optimizerFULL over this module's parameters followed by the denoiser's (reference :634-672): torch.optim.AdamW(lr
        1e-4) + LambdaLR(lr_lambda).  With layout_branch=True the parameter order is the reference's, so the 'opt' entry of a
        reference `model{epoch}.pth` loads directly (`optimizerFULL.load_state_dict`)."""
        params = [p for p in self.parameters() if p.requires_grad]
        self.optimizerFULL = torch.optim.AdamW(params + list(self.Diff.trainable_params), lr=1e-4)
        self.scheduler = torch.optim.lr_scheduler.LambdaLR(self.optimizerFULL, lr_lambda=self.lr_lambda)
        self.optimizers = [self.optimizerFULL]

    def update_learning_rate(self):
        """reference :674-681 (called once per iteration by the trainer)."""
        self.scheduler.step()
        return self.optimizers[0].param_groups[0]["lr"]

    @staticmethod
    def lr_lambda(counter):
        """Step schedule of optimizerFULL (reference :620-633): 1e-4 -> 5e-5 (20k) -> 1e-5 (60k) -> 5e-6 (100k)."""
        if counter < 20000:
            return 1.0
        if counter < 60000:
            return 5e-5 / 1e-4
        if counter < 100000:
            return 1e-5 / 1e-4
        return 5e-6 / 1e-4

    def encoder_2(self, z, objs, triples, dec_text_feat, dec_rel_feat, attributes=None, manipulate=False):
        """(uc_rel, c_rel), each (O, 1, 1280) (reference :220-242).  With autograd enabled and trainable parameters the
        outputs carry a grad_fn whose backward runs the explicit GCN / MLP gradient kernels (encoder_2_backward) and fills
        `.grad` of rel_mlp, gconv_net_ec_rel, the two decoder embeddings and of `z`, as the reference's autograd does."""
        if torch.is_grad_enabled() and (z.requires_grad or any(p.requires_grad for p in self._enc2_params())):
            params = [p for p in self._enc2_params() if p.requires_grad]
            uc, c = _Encoder2Function.apply(self, z, objs, triples, dec_text_feat, dec_rel_feat, *params)
            return uc, (c if self.use_E2 else None)
        with torch.no_grad():
            uc, c, _ = self._encoder_2_impl(z, objs, triples, dec_text_feat, dec_rel_feat, train=False)
        return uc, c

    def _enc2_params(self):
        mods = [self.obj_embeddings_dc, self.pred_embeddings_dc, self.rel_mlp] + ([self.gconv_net_ec_rel] if self.use_E2 else [])
        return [p for m in mods for p in m.parameters()]

    @torch.no_grad()
    def _encoder_2_impl(self, z, objs, triples, dec_text_feat, dec_rel_feat, train: bool):
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
        rel_vecs_2, tape = None, None
        if not train:
            if self.use_E2:
                rel_vecs_2, _ = self.gconv_net_ec_rel(rel_vecs_, pred_vecs_, edges)
                rel_vecs_2 = run_mlp(self.rel_mlp, rel_vecs_2).unsqueeze(1)
            return run_mlp(self.rel_mlp, rel_vecs_).unsqueeze(1), rel_vecs_2, None
        gcn_tapes = mlp_c = None
        if self.use_E2:
            ov, _, gcn_tapes = self.gconv_net_ec_rel.forward_train(rel_vecs_, pred_vecs_, edges)
            rel_vecs_2, mlp_c = run_mlp_train(self.rel_mlp, ov)
            rel_vecs_2 = rel_vecs_2.unsqueeze(1)
        uc, mlp_uc = run_mlp_train(self.rel_mlp, rel_vecs_)
        tape = dict(gcn=gcn_tapes, mlp_c=mlp_c, mlp_uc=mlp_uc, objs=objs.to(torch.int64).contiguous(),
                    preds=p.to(torch.int64).contiguous(), n_text=dec_text_feat.shape[1] if self.clip else 0,
                    n_rel=dec_rel_feat.shape[1] if self.clip else 0, z_dim=z.shape[1])
        return uc.unsqueeze(1), rel_vecs_2, tape

    def encoder_2_train(self, z, objs, triples, dec_text_feat, dec_rel_feat):
        """encoder_2 keeping the activations: returns (uc_rel, c_rel, tape) for encoder_2_backward."""
        return self._encoder_2_impl(z, objs, triples, dec_text_feat, dec_rel_feat, train=True)

    @torch.no_grad()
    def encoder_2_backward(self, tape, d_c=None, d_uc=None, sink=None):
        """Gradients of every encoder_2 parameter (accumulated into `sink`, a GradSink) and d_z (O, z_dim), from the gradients
        of c_rel and / or uc_rel ((O, 1, 1280) or (O, 1280)).  The training loss only uses c_rel (sdfusion_txt2shape_model.py
        :348-361), so d_uc is normally None."""
        from .. import ops_bwd
        from .layers import mlp_backward
        from .networks.diffusion_networks.unet_train import GradSink
        sink = GradSink() if sink is None else sink
        E = self.embedding_dim
        n_obj_in = tape["n_text"] + E + tape["z_dim"]
        d_rel_in = None                                     # gradient of rel_vecs_ = [text | Emb(obj) | z]
        d_pred_in = None
        if d_c is not None and tape["gcn"] is not None:
            d_ov = mlp_backward(tape["mlp_c"], d_c.reshape(d_c.shape[0], -1).float().contiguous(), sink)
            d_rel_in, d_pred_in = self.gconv_net_ec_rel.backward(tape["gcn"], d_ov, None, sink)   # E2's predicate output is unused
        if d_uc is not None:
            d2 = mlp_backward(tape["mlp_uc"], d_uc.reshape(d_uc.shape[0], -1).float().contiguous(), sink)
            d_rel_in = d2 if d_rel_in is None else ops.add_rows(d_rel_in, d2)
        if d_rel_in is None:
            return sink, None
        if self.obj_embeddings_dc.weight.requires_grad:
            ops_bwd.embedding_bwd(d_rel_in, tape["n_text"], tape["objs"], sink.grad(self.obj_embeddings_dc.weight))
        if d_pred_in is not None and self.pred_embeddings_dc.weight.requires_grad:
            ops_bwd.embedding_bwd(d_pred_in, tape["n_rel"], tape["preds"], sink.grad(self.pred_embeddings_dc.weight))
        return sink, d_rel_in[:, n_obj_in - tape["z_dim"]:].contiguous()

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

    # ---- checkpoints: the reference's `model{epoch}.pth` layout (VAEGAN_V2FULL.py:687-699 / VAE.py:120-158, 334-340) -------
    def state_dict(self, epoch=None, counter=None, *args, optimizer_state=None, **kwargs):
        """Without arguments: the plain nn.Module state dict.  With (epoch, counter): the reference's checkpoint dictionary --
        this module's tensors plus 'epoch', 'counter', 'vqvae', 'df' and 'opt' (a torch.optim.AdamW state dict, e.g.
        DenoiserTrainStep.optimizer_state_dict(); {} if none is given)."""
        sd = super().state_dict(*args, **kwargs)
        if epoch is None and counter is None:
            return sd
        sd = dict(sd)
        sd.update({"epoch": epoch, "counter": counter, "vqvae": self.Diff.vqvae_module.state_dict(),
                   "df": self.Diff.df_module.state_dict(), "opt": optimizer_state if optimizer_state is not None else {}})
        return sd

    def save_checkpoint(self, path, epoch, counter=None, optimizer_state=None):
        torch.save(self.state_dict(epoch, counter, optimizer_state=optimizer_state), path)
        return path

    def load_checkpoint(self, ckpt, strict=False):
        """Restore from a reference-format `model{epoch}.pth` (path or loaded dict) the way VAE.load_networks does: pop
        'vqvae' / 'df' / 'opt' / 'epoch' / 'counter', load the rest into this module, then the VQ-VAE and the denoiser.
        strict=False by default: a full v2_full checkpoint also holds the layout branch (box encoder / decoder / manipulator),
        which this shape-branch class does not have; every shape-branch key must still be present and is loaded.
        Returns {'epoch', 'counter', 'opt', 'ignored_keys'}: 'opt' is the untouched optimizer state (its LAST
        len(Diff.trainable_params) entries are the denoiser's, reference :636-645 -- see denoiser_optimizer_state)."""
        ckpt = dict(torch.load(ckpt, map_location="cpu") if isinstance(ckpt, str) else ckpt)
        extra = {k: ckpt.pop(k, None) for k in ("vqvae", "df", "opt", "epoch", "counter")}
        own = set(super().state_dict().keys())
        missing = sorted(own - set(ckpt))
        if missing:
            raise KeyError(f"checkpoint lacks shape-branch tensors: {missing[:5]}{' ...' if len(missing) > 5 else ''}")
        ignored = sorted(set(ckpt) - own)
        if ignored and strict:
            raise KeyError(f"unexpected keys in checkpoint: {ignored[:5]}")
        self.load_state_dict({k: v for k, v in ckpt.items() if k in own}, strict=True)
        if extra["vqvae"] is not None:
            self.Diff.vqvae.load_state_dict(extra["vqvae"])
        if extra["df"] is not None:
            self.Diff.df.load_state_dict(extra["df"])
        return {"epoch": extra["epoch"], "counter": extra["counter"], "opt": extra["opt"], "ignored_keys": ignored}

    def denoiser_optimizer_state(self, opt_state: dict) -> dict:
        """The denoiser's slice of a reference optimizerFULL state dict, re-indexed from 0 so that
        DenoiserTrainStep.load_optimizer_state_dict (or torch.optim.AdamW over df.parameters()) accepts it.  The reference
        builds the optimizer over `params + df_params` (:636-645): the denoiser's parameters are the last entries."""
        n = len(self.Diff.trainable_params)
        group = dict(opt_state["param_groups"][0])
        ids = list(group["params"])
        if len(ids) < n:
            raise ValueError(f"optimizer state covers {len(ids)} parameters, fewer than the denoiser's {n}")
        tail = ids[-n:]
        group["params"] = list(range(n))
        return {"state": {j: opt_state["state"][i] for j, i in enumerate(tail) if i in opt_state["state"]}, "param_groups": [group]}

    def forward(self, enc_objs, enc_triples, enc, enc_text_feat, enc_rel_feat, attributes, enc_objs_to_scene, dec_objs, dec_objs_grained,
                dec_triples, dec, dec_text_feat, dec_rel_feat, dec_attributes, dec_objs_to_scene, missing_nodes, manipulated_nodes, dec_sdfs,
                enc_angles=None, dec_angles=None):
        """The training forward the reference's trainer calls (VAEGAN_V2FULL.py:466-560, via VAE.forward_mani): encode the
        layout -> reparameterise -> insert added nodes -> manipulate touched nodes -> scene-graph conditioning (encoder_2) ->
        object selection -> diffusion loss (self.Diff.loss_df, with its grad_fn) -> decode the layout.  Returns the reference's
        tuple (mu, logvar, orig_gt_d3, orig_gt_angles, orig_gt_shapes, orig_d3, orig_angles, d3_pred, angles_pred,
        [obj_selected, None], keep).  With autograd enabled every returned tensor and self.Diff.loss_df carry grad_fns: the
        trainer's `loss.backward()` (train_3dfront.py:387-390) reaches all branches through the explicit backward kernels."""
        self._need_layout()
        mu, logvar = self.encoder(enc_objs, enc_triples, enc, attributes, enc_text_feat, enc_rel_feat, enc_angles)
        if getattr(self, "use_AE", False):
            z = mu
        else:
            std = torch.exp(0.5 * logvar)
            z = torch.randn_like(std).mul(std).add_(mu)
        z, nodes_added = self._insert_zero_nodes(z, missing_nodes)
        change = [np.random.normal(0, 1, self.embedding_dim) if (i in nodes_added or i in manipulated_nodes) else np.zeros(self.embedding_dim)
                  for i in range(len(z))]
        change_repr = torch.from_numpy(np.stack(change, axis=0)).float().to(z.device)
        z_prime = self.manipulate(torch.cat([z, change_repr], dim=1), dec_objs, dec_triples, dec_text_feat, dec_rel_feat, attributes)
        if not getattr(self, "replace_all_latent", False):
            for node in sorted(nodes_added + list(manipulated_nodes)):
                z = torch.cat([z[:node], z_prime[node:node + 1], z[node + 1:]], dim=0)
        else:
            z = z_prime
        uc_rel_feat, c_rel_feat = self.encoder_2(z, dec_objs, dec_triples, dec_text_feat, dec_rel_feat, attributes)
        if c_rel_feat is None:
            c_rel_feat = uc_rel_feat
        obj_selected, diff_dict = self.select_sdfs(dec_objs_to_scene, dec_objs, dec_objs_grained, dec_sdfs, uc_rel_feat, c_rel_feat, random=False)
        self.Diff.set_input(diff_dict)
        self.Diff.set_requires_grad([self.Diff.df], requires_grad=True)
        self.Diff.forward()
        d3_pred, angles_pred = self.decoder(z, dec_objs, dec_triples, dec_text_feat, dec_rel_feat, attributes)
        kept = [i for i in range(len(d3_pred)) if i not in nodes_added and i not in manipulated_nodes]
        idx = torch.tensor(kept, dtype=torch.long, device=d3_pred.device)
        keep = torch.zeros(len(d3_pred), 1, device=d3_pred.device)
        keep[idx] = 1.0
        return (mu, logvar, dec[idx], dec_angles[idx], dec_sdfs[idx], d3_pred[idx], angles_pred[idx], d3_pred, angles_pred,
                [obj_selected, None], keep)

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


class _Encoder2Function(torch.autograd.Function):
    """Autograd bridge for encoder_2: forward = encoder_2_train, backward = encoder_2_backward (explicit kernels)."""

    @staticmethod
    def forward(ctx, model, z, objs, triples, text_feat, rel_feat, *params):
        uc, c, tape = model.encoder_2_train(z.detach(), objs, triples, text_feat.detach(), rel_feat.detach())
        ctx.model, ctx.tape, ctx.params, ctx.need_dz = model, tape, params, z.requires_grad
        ctx.set_materialize_grads(False)        # an unused output (uc_rel in training) arrives as None, not as zeros
        if c is None:
            c = uc.new_zeros(uc.shape)
        return uc, c

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_uc, d_c):
        sink, d_z = ctx.model.encoder_2_backward(ctx.tape, d_c=d_c if ctx.tape["gcn"] is not None else None, d_uc=d_uc)
        ctx.tape = None
        grads = tuple(sink.grads.get(p) for p in ctx.params)
        return (None, d_z if ctx.need_dz else None, None, None, None, None) + grads

