"""Oracle (test infrastructure): CPU fp32 restatement of the reference VQ-VAE.

Follows model/networks/vqvae_networks/{network.py, vqvae_modules.py, quantizer.py}; weights come in as a
state dict with the reference's keys (`encoder.*`, `decoder.*`, `quantize.embedding.weight`,
`quant_conv.*`, `post_quant_conv.*`).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# config/vqvae_snet.yaml
VQ_FULL = dict(embed_dim=3, n_embed=8192, z_channels=3, resolution=64, in_channels=1, out_ch=1, ch=64,
               ch_mult=(1, 2, 4), num_res_blocks=1)
VQ_TINY = dict(embed_dim=3, n_embed=64, z_channels=3, resolution=16, in_channels=1, out_ch=1, ch=32,
               ch_mult=(1, 2, 4), num_res_blocks=1)


def _groups(c: int) -> int:
    """Normalize() group count (vqvae_modules.py:13-21)."""
    if c <= 32:
        return c // 4
    return 32 if c % 32 == 0 else 30


def vq_param_shapes(cfg: dict) -> Dict[str, Tuple[int, ...]]:
    s: Dict[str, Tuple[int, ...]] = {}

    def conv(n, i, o, k):
        s[n + ".weight"] = (o, i, k, k, k); s[n + ".bias"] = (o,)

    def norm(n, c):
        s[n + ".weight"] = (c,); s[n + ".bias"] = (c,)

    def res(n, i, o):                                       # ResnetBlock, temb_channels = 0 (:64-101)
        norm(n + ".norm1", i); conv(n + ".conv1", i, o, 3); norm(n + ".norm2", o); conv(n + ".conv2", o, o, 3)
        if i != o:
            conv(n + ".nin_shortcut", i, o, 1)

    def attn(n, c):                                         # AttnBlock (:126-152)
        norm(n + ".norm", c)
        for k in ("q", "k", "v", "proj_out"):
            conv(f"{n}.{k}", c, c, 1)

    ch, mult, nrb, zc = cfg["ch"], cfg["ch_mult"], cfg["num_res_blocks"], cfg["z_channels"]
    nres = len(mult)
    # Encoder3D (:181-256)
    conv("encoder.conv_in", cfg["in_channels"], ch, 3)
    in_mult = (1,) + tuple(mult)
    bi = ch
    for lvl in range(nres):
        bi, bo = ch * in_mult[lvl], ch * mult[lvl]
        for b in range(nrb):
            res(f"encoder.down.{lvl}.block.{b}", bi, bo)
            bi = bo
        if lvl != nres - 1:
            conv(f"encoder.down.{lvl}.downsample.conv", bi, bi, 3)
    res("encoder.mid.block_1", bi, bi); attn("encoder.mid.attn_1", bi); res("encoder.mid.block_2", bi, bi)
    norm("encoder.norm_out", bi); conv("encoder.conv_out", bi, zc, 3)          # double_z = False
    # Decoder3D (:292-374)
    bi = ch * mult[-1]
    conv("decoder.conv_in", zc, bi, 3)
    res("decoder.mid.block_1", bi, bi); attn("decoder.mid.attn_1", bi); res("decoder.mid.block_2", bi, bi)
    for lvl in reversed(range(nres)):
        bo = ch * mult[lvl]
        for b in range(nrb):
            res(f"decoder.up.{lvl}.block.{b}", bi, bo)
            bi = bo
        if lvl != 0:
            conv(f"decoder.up.{lvl}.upsample.conv", bi, bi, 3)
    norm("decoder.norm_out", bi); conv("decoder.conv_out", bi, cfg["out_ch"], 3)
    s["quantize.embedding.weight"] = (cfg["n_embed"], cfg["embed_dim"])
    conv("quant_conv", zc, cfg["embed_dim"], 1); conv("post_quant_conv", cfg["embed_dim"], zc, 1)
    return s


def _norm(sd, n, x):
    c = x.shape[1]
    return F.group_norm(x, _groups(c), sd[n + ".weight"], sd[n + ".bias"], 1e-6)


def _conv(sd, n, x, stride=1, padding=1):
    return F.conv3d(x, sd[n + ".weight"], sd[n + ".bias"], stride=stride, padding=padding)


def _res(sd, n, x):
    """ResnetBlock.forward, temb=None, swish nonlinearity (vqvae_modules.py:103-123)."""
    h = _conv(sd, n + ".conv1", F.silu(_norm(sd, n + ".norm1", x)))
    h = _conv(sd, n + ".conv2", F.silu(_norm(sd, n + ".norm2", h)))
    if (n + ".nin_shortcut.weight") in sd:
        x = _conv(sd, n + ".nin_shortcut", x, padding=0)
    return x + h


def _attn(sd, n, x):
    """AttnBlock.forward: single head, scale c ** -0.5 (vqvae_modules.py:154-178)."""
    h = _norm(sd, n + ".norm", x)
    q, k, v = (_conv(sd, f"{n}.{t}", h, padding=0) for t in ("q", "k", "v"))
    b, c, d, hh, w = q.shape
    q = q.reshape(b, c, -1).permute(0, 2, 1)
    k = k.reshape(b, c, -1)
    w_ = torch.softmax(torch.bmm(q, k) * (int(c) ** -0.5), dim=2)
    v = v.reshape(b, c, -1)
    h = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, d, hh, w)
    return x + _conv(sd, n + ".proj_out", h, padding=0)


def encode_no_quant(sd, cfg: dict, x: Tensor) -> Tensor:
    """VQVAE.encode_no_quant = quant_conv(Encoder3D(x)) (network.py:84-88; vqvae_modules.py:258-290)."""
    nres, nrb = len(cfg["ch_mult"]), cfg["num_res_blocks"]
    h = _conv(sd, "encoder.conv_in", x)
    for lvl in range(nres):
        for b in range(nrb):
            h = _res(sd, f"encoder.down.{lvl}.block.{b}", h)
        if lvl != nres - 1:                                   # Downsample: pad (0,1)^3 then stride 2 (:54-58)
            h = _conv(sd, f"encoder.down.{lvl}.downsample.conv", F.pad(h, (0, 1, 0, 1, 0, 1)), stride=2, padding=0)
    h = _res(sd, "encoder.mid.block_1", h); h = _attn(sd, "encoder.mid.attn_1", h); h = _res(sd, "encoder.mid.block_2", h)
    h = _conv(sd, "encoder.conv_out", F.gelu(_norm(sd, "encoder.norm_out", h)))     # activ='gelu' (:199-200)
    return _conv(sd, "quant_conv", h, padding=0)


def quantize(sd, z: Tensor) -> Tuple[Tensor, Tensor]:
    """VectorQuantizer.forward, is_voxel=True (quantizer.py:68-99): argmin of z^2 + e^2 - 2 z.e; returns
    (z_q in NCDHW, indices)."""
    e = sd["quantize.embedding.weight"]
    zp = z.permute(0, 2, 3, 4, 1).contiguous()
    zf = zp.view(-1, e.shape[1])
    d = torch.sum(zf ** 2, dim=1, keepdim=True) + torch.sum(e ** 2, dim=1) - 2 * torch.einsum("bd,dn->bn", zf, e.t())
    idx = torch.argmin(d, dim=1)
    zq = e[idx].view(zp.shape)
    zq = zp + (zq - zp)                                       # straight-through estimator, forward value
    return zq.permute(0, 4, 1, 2, 3).contiguous(), idx


def decode_no_quant(sd, cfg: dict, h: Tensor, force_not_quantize: bool = False) -> Tensor:
    """VQVAE.decode_no_quant (network.py:95-103) -> Decoder3D.forward (vqvae_modules.py:376-409)."""
    nres, nrb = len(cfg["ch_mult"]), cfg["num_res_blocks"]
    q = h if force_not_quantize else quantize(sd, h)[0]
    h = _conv(sd, "decoder.conv_in", _conv(sd, "post_quant_conv", q, padding=0))
    h = _res(sd, "decoder.mid.block_1", h); h = _attn(sd, "decoder.mid.attn_1", h); h = _res(sd, "decoder.mid.block_2", h)
    for lvl in reversed(range(nres)):
        for b in range(nrb):
            h = _res(sd, f"decoder.up.{lvl}.block.{b}", h)
        if lvl != 0:                                          # Upsample: nearest x2 then conv (:36-40)
            h = _conv(sd, f"decoder.up.{lvl}.upsample.conv", F.interpolate(h, scale_factor=2.0, mode="nearest"))
    return _conv(sd, "decoder.conv_out", F.gelu(_norm(sd, "decoder.norm_out", h)))
