"""VectorQuantizer — codebook lookup on the GPU (cs_vq_quantize).

Drop-in for the reference's model/networks/vqvae_networks/quantizer.py:10-119 on the path the shape branch
uses (is_voxel=True, no remap, eval): forward(z) -> (z_q, loss, (None, None, indices)).  The straight-through
value z + (z_q - z).detach() equals z_q in the forward pass, which is what is returned.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .... import ops


class VectorQuantizer(nn.Module):
    def __init__(self, n_e, e_dim, beta, remap=None, unknown_index="random", sane_index_shape=False, legacy=True):
        super().__init__()
        if remap is not None:
            raise NotImplementedError("index remapping is not used by the shape branch")
        self.n_e, self.e_dim, self.beta, self.legacy = n_e, e_dim, beta, legacy
        self.embedding = nn.Embedding(self.n_e, self.e_dim)
        self.embedding.weight.data.uniform_(-1.0 / self.n_e, 1.0 / self.n_e)
        self.re_embed = n_e
        self.sane_index_shape = sane_index_shape

    @torch.no_grad()
    def forward(self, z, temp=None, rescale_logits=False, return_logits=False, is_voxel=False, post_quant=None):
        assert temp is None or temp == 1.0, "Only for interface compatible with Gumbel"
        assert not rescale_logits and not return_logits, "Only for interface compatible with Gumbel"
        if not is_voxel or z.dim() != 5:
            raise NotImplementedError("only the 3-D (is_voxel=True) path of the shape branch is built")
        pw, pb = post_quant if post_quant is not None else (None, None)
        z = z.float().contiguous()
        z_q, idx = ops.vq_quantize(z, self.embedding.weight.detach().float().contiguous(), pw, pb)
        loss = None   # the embedding loss is only needed to TRAIN the VQ-VAE, which is loaded frozen (model_utils.py:28-31)
        if self.sane_index_shape:
            idx = idx.reshape(z.shape[0], z.shape[2], z.shape[3], z.shape[4])
        return z_q, loss, (None, None, idx)
