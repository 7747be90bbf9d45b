"""VAE — the model dispatcher the reference's scripts construct (`from model.VAE import VAE`, model/VAE.py:18-340;
scripts/train_3dfront.py, scripts/eval_3dfront.py), for the network type this framework implements: 'v2_full'
(CommonScenes: layout graph-VAE + shape diffusion).  Same constructor arguments and method names; every call forwards
to the Sg2ScVAEModel mirror (VAEGAN_V2FULL.py), so the scripts' call sites need no change beyond the import.

The other types ('v1_box', 'v1_full', 'v2_box': Graph-to-3D / box-only baselines) are outside this
framework's scope and raise NotImplementedError.
"""
from __future__ import annotations

import os
import pickle

import torch
import torch.nn as nn

from .VAEGAN_V2FULL import Sg2ScVAEModel
from .sdfusion_txt2shape_model import _Cfg, _load_yaml


class VAE(nn.Module):
    def __init__(self, root="../GT", type="v1_box", diff_opt="../config/v2_full.yaml", vocab=None, replace_latent=False, with_changes=True,
                 distribution_before=True, residual=False, gconv_pooling="avg", with_angles=False, num_box_params=6, lr_full=None,
                 deepsdf=False, clip=True, with_E2=True):
        super().__init__()
        assert type in ["v1_box", "v1_full", "v2_box", "v2_full"], "{} is not included".format(type)
        if type != "v2_full":
            raise NotImplementedError(f"network type {type!r}: only 'v2_full' (CommonScenes) is implemented here")
        self.type_, self.vocab, self.with_angles, self.epoch, self.counter = type, vocab, with_angles, 0, 0
        self.diff_opt = diff_opt
        assert distribution_before is not None and replace_latent is not None and with_changes is not None
        opt = _load_yaml(diff_opt) if isinstance(diff_opt, str) else diff_opt          # the reference passes the yaml path
        if not with_angles:
            raise NotImplementedError("the layout branch mirrors the v2_full training wiring (with_angles=True)")
        self.vae_v2 = Sg2ScVAEModel(vocab, opt, diffusion_bs=16, embedding_dim=64, decoder_cat=True, mlp_normalization="batch",
                                    gconv_num_layers=5, use_angles=with_angles, use_E2=with_E2, residual=residual, clip=clip,
                                    gconv_pooling=gconv_pooling, num_box_params=num_box_params, layout_branch=True)
        self.vae_v2.replace_all_latent = replace_latent
        self.vae_v2.optimizer_ini()

    def set_cuda(self):
        self.vae_v2.cuda()

    def forward_mani(self, enc_objs, enc_triples, enc_boxes, enc_angles, enc_shapes, encoded_enc_text_feat, encoded_enc_rel_feat, attributes,
                     enc_objs_to_scene, dec_objs, dec_objs_grained, dec_triples, dec_boxes, dec_angles, dec_sdfs, dec_shapes,
                     encoded_dec_text_feat, encoded_dec_rel_feat, dec_attributes, dec_objs_to_scene, missing_nodes, manipulated_nodes):
        """reference :88-100: the 14-tuple the training loop unpacks."""
        mu, logvar, orig_gt_boxes, orig_gt_angles, orig_gt_shapes, orig_boxes, orig_angles, boxes, angles, obj_and_shape, keep = self.vae_v2.forward(
            enc_objs, enc_triples, enc_boxes, encoded_enc_text_feat, encoded_enc_rel_feat, attributes, enc_objs_to_scene, dec_objs, dec_objs_grained,
            dec_triples, dec_boxes, encoded_dec_text_feat, encoded_dec_rel_feat, dec_attributes, dec_objs_to_scene, missing_nodes, manipulated_nodes,
            dec_sdfs, enc_angles, dec_angles)
        return mu, logvar, None, None, orig_gt_boxes, orig_gt_angles, orig_gt_shapes, orig_boxes, orig_angles, None, boxes, angles, obj_and_shape, keep

    # ---- checkpoints (reference :102-158, 334-340) -------------------------------------------------------------------
    def load_networks(self, exp, epoch, strict=True, restart_optim=False):
        info = self.vae_v2.load_checkpoint(os.path.join(exp, "checkpoint", "model{}.pth".format(epoch)), strict=False)
        if info["epoch"] is not None:
            self.epoch, self.counter = info["epoch"], info["counter"]
        self.optimizer_state = None if restart_optim else info["opt"]      # also available to DenoiserTrainStep (denoiser_optimizer_state)
        if not restart_optim and info["opt"]:
            self.vae_v2.optimizerFULL.load_state_dict(info["opt"])
            self.vae_v2.scheduler = torch.optim.lr_scheduler.LambdaLR(self.vae_v2.optimizerFULL, lr_lambda=self.vae_v2.lr_lambda,
                                                                      last_epoch=int(self.counter - 1))
        return info

    def save(self, exp, outf, epoch, counter=None, optimizer_state=None):
        if optimizer_state is None:
            optimizer_state = self.vae_v2.optimizerFULL.state_dict()
        return self.vae_v2.save_checkpoint(os.path.join(exp, outf, "model{}.pth".format(epoch)), epoch, counter, optimizer_state)

    def compute_statistics(self, exp, epoch, stats_dataloader, force=False):
        stats_f = os.path.join(exp, "checkpoint", "model_stats_{}.pkl".format(epoch))
        if os.path.exists(stats_f) and not force:
            with open(stats_f, "rb") as f:
                self.mean_est, self.cov_est = pickle.load(f)[:2]
        else:
            self.mean_est, self.cov_est = self.vae_v2.collect_train_statistics(stats_dataloader)
            with open(stats_f, "wb") as f:
                pickle.dump([self.mean_est, self.cov_est], f)

    # ---- evaluation (reference :193-300) -------------------------------------------------------------------------------
    def decoder_with_changes_boxes_and_shape(self, z_box, z_shape, objs, triples, encoded_dec_text_feat, encoded_dec_rel_feat, dec_sdfs, attributes,
                                             missing_nodes, manipulated_nodes, box_data=None, gen_shape=False):
        return self.vae_v2.decoder_with_changes(z_box, objs, triples, encoded_dec_text_feat, encoded_dec_rel_feat, dec_sdfs, attributes, missing_nodes,
                                                manipulated_nodes, gen_shape=gen_shape)

    def decoder_with_additions_boxes_and_shape(self, z_box, z_shape, objs, triples, encoded_dec_text_feat, encoded_dec_rel_feat, dec_sdfs, attributes,
                                               missing_nodes, manipulated_nodes, gen_shape=False):
        return self.vae_v2.decoder_with_additions(z_box, objs, triples, encoded_dec_text_feat, encoded_dec_rel_feat, dec_sdfs, attributes, missing_nodes,
                                                  manipulated_nodes, gen_shape=gen_shape)

    def encode_box_and_shape(self, objs, triples, encoded_enc_text_feat, encoded_enc_rel_feat, feats, boxes, angles=None, attributes=None):
        if not self.with_angles:
            angles = None
        return self.encode_box(objs, triples, encoded_enc_text_feat, encoded_enc_rel_feat, boxes, angles, attributes), (None, None)

    def encode_box(self, objs, triples, encoded_enc_text_feat, encoded_enc_rel_feat, boxes, angles=None, attributes=None):
        return self.vae_v2.encoder(objs, triples, boxes, attributes, encoded_enc_text_feat, encoded_enc_rel_feat, angles)

    def sample_box_and_shape(self, point_classes_idx, dec_objs, dec_triplets, dec_sdfs, encoded_dec_text_feat, encoded_dec_rel_feat, attributes=None,
                             gen_shape=False):
        return self.vae_v2.sample(point_classes_idx, self.mean_est, self.cov_est, dec_objs, dec_triplets, dec_sdfs, encoded_dec_text_feat,
                                  encoded_dec_rel_feat, attributes, gen_shape=gen_shape)
