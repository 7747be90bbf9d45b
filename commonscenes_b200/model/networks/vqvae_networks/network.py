"""VQVAE — the frozen 64^3-SDF autoencoder around the denoiser, on the B200 kernels.

Drop-in for the reference's model/networks/vqvae_networks/network.py:51-140 on the paths the shape branch
calls: forward(x, forward_no_quant=True, encode_only=True) / encode_no_quant (training: SDF -> 3x16^3 latent,
sdfusion_txt2shape_model.py:357-358) and decode_no_quant (sampling: latent -> quantise -> 64^3 SDF, :511).
Same constructor, same state-dict keys (encoder.*, decoder.*, quantize.embedding.weight, quant_conv.*,
post_quant_conv.*).  quant_conv (1x1x1) is folded into the encoder's conv_out weights when they are packed;
post_quant_conv is applied inside the codebook-lookup kernel.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.nn import init

from .... import _lib, ops
from .quantizer import VectorQuantizer
from .vqvae_modules import Decoder3D, Encoder3D, _f


def init_weights(net, init_type="normal", gain=0.01):
    def init_func(m):
        classname = m.__class__.__name__
        if hasattr(m, "weight") and (classname.find("Conv") != -1 or classname.find("Linear") != -1):
            if init_type == "normal":
                init.normal_(m.weight.data, 0.0, gain)
            elif init_type == "xavier":
                init.xavier_normal_(m.weight.data, gain=gain)
            elif init_type == "kaiming":
                init.kaiming_normal_(m.weight.data, a=0, mode="fan_in")
            elif init_type == "none":
                m.reset_parameters()
            else:
                raise NotImplementedError("initialization method [%s] is not implemented" % init_type)
            if hasattr(m, "bias") and m.bias is not None:
                init.constant_(m.bias.data, 0.0)
    net.apply(init_func)


class VQVAE(nn.Module):
    def __init__(self, ddconfig, n_embed, embed_dim, remap=None, sane_index_shape=False):
        super().__init__()
        self.ddconfig, self.n_embed, self.embed_dim = ddconfig, n_embed, embed_dim
        self.encoder = Encoder3D(**ddconfig)
        self.decoder = Decoder3D(**ddconfig)
        self.quantize = VectorQuantizer(n_embed, embed_dim, beta=1.0, remap=remap, sane_index_shape=sane_index_shape, legacy=False)
        self.quant_conv = torch.nn.Conv3d(ddconfig["z_channels"], embed_dim, 1)
        self.post_quant_conv = torch.nn.Conv3d(embed_dim, ddconfig["z_channels"], 1)
        for m in (self.encoder, self.decoder, self.quant_conv, self.post_quant_conv):
            init_weights(m, "normal", 0.02)
        self._enc_out = None

    def _encoder_head(self):
        """conv_out (3x3x3) with quant_conv (1x1x1) folded in: W' = Wq Wc, b' = Wq bc + bq (exact composition)."""
        src = (self.encoder.conv_out.weight, self.encoder.conv_out.bias, self.quant_conv.weight, self.quant_conv.bias)
        key = tuple((p.data_ptr(), p._version) for p in src)
        if self._enc_out is None or self._enc_out[0] != key:
            wc, bc, wq, bq = (p.detach().double() for p in src)
            wq2 = wq.reshape(wq.shape[0], wq.shape[1])
            w = torch.einsum("oz,zcdhw->ocdhw", wq2, wc).float()
            b = (wq2 @ bc + bq).float().contiguous()
            self._enc_out = (key, ops.pack_small_cout_conv(w), b)
        return self._enc_out[1], self._enc_out[2]

    @torch.no_grad()
    def encode_no_quant(self, x):
        w, b = self._encoder_head()
        return ops.conv3d_small_cout(self.encoder.run_trunk(x), w, b, self.embed_dim)

    @torch.no_grad()
    def encode(self, x):
        return self.quantize(self.encode_no_quant(x), is_voxel=True)

    @torch.no_grad()
    def decode(self, quant):
        pw = _f(self.post_quant_conv.weight).reshape(self.post_quant_conv.weight.shape[0], -1).contiguous()
        return self.decoder(ops.channel_mix(quant.float().contiguous(), pw, _f(self.post_quant_conv.bias)))

    @torch.no_grad()
    def decode_no_quant(self, h, force_not_quantize=False):
        if force_not_quantize:
            return self.decode(h)
        pw = _f(self.post_quant_conv.weight).reshape(self.post_quant_conv.weight.shape[0], -1).contiguous()
        q, _, _ = self.quantize(h, is_voxel=True, post_quant=(pw, _f(self.post_quant_conv.bias)))
        return self.decoder(q)

    @torch.no_grad()
    def forward(self, input, verbose=False, forward_no_quant=False, encode_only=False):
        if forward_no_quant:
            z = self.encode_no_quant(input)
            if encode_only:
                return z
            return self.decode_no_quant(z), z
        quant, diff, info = self.encode(input)
        dec = self.decode(quant)
        return (dec, quant, diff, info) if verbose else (dec, diff)
