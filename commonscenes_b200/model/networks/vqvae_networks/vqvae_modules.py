"""Encoder3D / Decoder3D of the VQ-VAE on the B200 kernels.

Same class names, constructor keywords and state-dict keys as the reference's
model/networks/vqvae_networks/vqvae_modules.py (Normalize :13-21, Upsample :24-40, Downsample :43-61,
ResnetBlock :64-123, AttnBlock :126-178, Encoder3D :181-290, Decoder3D :292-409).  Children hold parameters;
`run()` executes on channels-last bf16 activations:

  * every conv is the tcgen05 implicit GEMM; the 1-channel / 3-channel input convs go through a bf16 patch
    matrix (cs_im2col_small); the asymmetric (0,1) pad + stride-2 downsample is expressed as front/back
    padding of the TMA box; conv_out writes fp32 NCDHW directly;
  * GroupNorm(32, eps 1e-6) + swish / GELU(erf) are one statistics pass + one apply pass (the sums of a
    conv's output come from its epilogue);
  * the single-head N=4096 mid-block attention uses one fused q|k|v GEMM and the on-chip softmax kernel
    instead of materialising the 4096 x 4096 fp32 score matrix.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .... import _lib, ops


def _f(t):
    return t.detach().float().contiguous()


def nonlinearity(x):
    return x * torch.sigmoid(x)


def Normalize(in_channels, num_groups=32):
    if in_channels <= 32:
        num_groups = in_channels // 4
    elif in_channels % num_groups != 0:
        num_groups = 30
    return torch.nn.GroupNorm(num_groups=num_groups, num_channels=in_channels, eps=1e-6, affine=True)


def _gn(norm: nn.GroupNorm, x, act, stat=None):
    return ops.groupnorm(x, _f(norm.weight), _f(norm.bias), groups=norm.num_groups, eps=norm.eps, act=act, stat_sum=stat)


def _stat(x, cout):
    S = x.shape[1] * x.shape[2] * x.shape[3]
    return ops.zero_stat_buffer(x.device, x.shape[0], cout) if S % 32 == 0 else None


class _Packed(nn.Module):
    """Caches kernel-layout weights; the VQ-VAE is frozen (model_utils.py:28-31) so they are packed once per
    load_state_dict / device move."""

    def _pk(self):
        key = tuple((p.data_ptr(), p._version) for p in self.parameters(recurse=False)) + \
            tuple((p.data_ptr(), p._version) for m in self.children() for p in m.parameters(recurse=False))
        if getattr(self, "_pk_key", None) != key:
            self._pk_val, self._pk_key = self.pack(), key
        return self._pk_val


class Upsample(_Packed):
    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if self.with_conv:
            self.conv = torch.nn.Conv3d(in_channels, in_channels, kernel_size=3, stride=1, padding=1)

    PHASE_CONV = True      # fold the nearest-upsample into the conv: eight 2x2x2 merged-tap phase convs (8 of 27 taps)

    def pack(self):
        if not self.with_conv:
            return None
        if Upsample.PHASE_CONV:
            return ops.pack_upsample_phase_weights(self.conv.weight, (2, 2, 2)), _f(self.conv.bias)
        return ops.pack_conv_weight(self.conv.weight), _f(self.conv.bias)

    def run(self, x):
        if self.with_conv and Upsample.PHASE_CONV and x.shape[1] * x.shape[2] * x.shape[3] >= 128:
            # F.interpolate(scale 2, nearest) + conv3x3x3 (vqvae_modules.py:33-40) == one 2x2x2 conv per output parity class on
            # the low-resolution tensor (ops.pack_upsample_phase_weights): the 8x larger intermediate is never written
            phases, b = self._pk()
            B, D, H, W, C = x.shape
            out = torch.empty((B, 2 * D, 2 * H, 2 * W, self.conv.out_channels), dtype=torch.bfloat16, device=x.device)
            for offs, ks, pad, pad_back, w in phases:
                ops.conv3d(x, w, ksize=ks, pad=pad, pad_back=pad_back, bias=b, out=out, phase=((2, 2, 2), offs))
            return out
        x = ops.upsample_nearest(x, (2, 2, 2))
        if self.with_conv:
            w, b = self._pk()
            if Upsample.PHASE_CONV:        # tiny grids (tests): the plain path needs the un-merged filter
                w = ops.pack_conv_weight(self.conv.weight)
            x = ops.conv3d(x, w, bias=b)
        return x


class Downsample(_Packed):
    def __init__(self, in_channels, with_conv):
        super().__init__()
        if not with_conv:
            raise NotImplementedError("avg-pool downsampling is not used (resamp_with_conv=True)")
        self.with_conv = with_conv
        self.conv = torch.nn.Conv3d(in_channels, in_channels, kernel_size=3, stride=2, padding=0)

    def pack(self):
        return ops.pack_conv_weight(self.conv.weight), _f(self.conv.bias)

    def run(self, x):
        w, b = self._pk()
        # F.pad(x, (0,1,0,1,0,1)) + stride-2 conv (vqvae_modules.py:54-58): zero padding only at the back
        return ops.conv3d(x, w, stride=(2, 2, 2), pad=(0, 0, 0), pad_back=(1, 1, 1), bias=b)


class ResnetBlock(_Packed):
    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout, temb_channels=512):
        super().__init__()
        if conv_shortcut or temb_channels > 0:
            raise NotImplementedError("conv_shortcut / timestep embedding are not used by the VQ-VAE (temb_ch = 0)")
        self.in_channels = in_channels
        out_channels = in_channels if out_channels is None else out_channels
        self.out_channels = out_channels
        self.use_conv_shortcut = conv_shortcut
        self.norm1 = Normalize(in_channels)
        self.conv1 = torch.nn.Conv3d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.norm2 = Normalize(out_channels)
        self.dropout = torch.nn.Dropout(dropout)
        self.conv2 = torch.nn.Conv3d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        if self.in_channels != self.out_channels:
            self.nin_shortcut = torch.nn.Conv3d(in_channels, out_channels, kernel_size=1, stride=1, padding=0)

    def pack(self):
        pk = {"w1": ops.pack_conv_weight(self.conv1.weight), "b1": _f(self.conv1.bias),
              "w2": ops.pack_conv_weight(self.conv2.weight), "b2": _f(self.conv2.bias)}
        if self.in_channels != self.out_channels:
            pk["ws"], pk["bs"] = ops.pack_conv_weight(self.nin_shortcut.weight), _f(self.nin_shortcut.bias)
        return pk

    def run(self, x, temb=None):
        pk = self._pk()
        stat = _stat(x, self.out_channels)
        h = ops.conv3d(_gn(self.norm1, x, ops.ACT_SILU), pk["w1"], bias=pk["b1"], stat_sum=stat)
        a = _gn(self.norm2, h, ops.ACT_SILU, stat)
        res = ops.linear_tokens(x, pk["ws"], bias=pk["bs"]) if "ws" in pk else x
        return ops.conv3d(a, pk["w2"], bias=pk["b2"], residual=res)


class AttnBlock(_Packed):
    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.norm = Normalize(in_channels)
        self.q = torch.nn.Conv3d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.k = torch.nn.Conv3d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.v = torch.nn.Conv3d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.proj_out = torch.nn.Conv3d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)

    def pack(self):
        c = self.in_channels
        if c not in (32, 64, 96, 128, 256):
            raise NotImplementedError(f"AttnBlock width {c}: cs_attention supports head dims 32/64/96/128/256")
        wqkv = torch.cat([self.q.weight, self.k.weight, self.v.weight]).detach().reshape(3 * c, c)
        bqkv = torch.cat([self.q.bias, self.k.bias, self.v.bias])
        return {"wqkv": ops.pack_linear_weight(wqkv), "bqkv": _f(bqkv),
                "wo": ops.pack_conv_weight(self.proj_out.weight), "bo": _f(self.proj_out.bias)}

    def run(self, x):
        pk = self._pk()
        B, D, H, W, c = x.shape
        qkv = ops.linear_tokens(_gn(self.norm, x, ops.ACT_NONE), pk["wqkv"], bias=pk["bqkv"]).view(B, D * H * W, 3 * c)
        q, k, v = (qkv[:, :, i * c:(i + 1) * c] for i in range(3))
        o = ops.attention(q, k, v, heads=1, head_dim=c, head_dim_padded=c, scale=int(c) ** (-0.5))
        return ops.linear_tokens(o.view(B, D, H, W, c), pk["wo"], bias=pk["bo"], residual=x)


def _small_cin_conv_pack(conv: nn.Conv3d):
    """3x3x3 conv with 1-4 input channels -> GEMM weight over the cs_im2col_small patch matrix."""
    wp, kp = ops.pack_patch_weight(conv.weight)
    return wp, _f(conv.bias), kp


class Encoder3D(_Packed):
    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, double_z=True, activ="gelu", **ignore_kwargs):
        super().__init__()
        if activ != "gelu" or len(attn_resolutions) > 0:
            raise NotImplementedError("the shape branch uses activ='gelu' and attn_resolutions=[] (config/vqvae_snet.yaml)")
        self.ch, self.temb_ch = ch, 0
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        self.resolution, self.in_channels = resolution, in_channels
        self.conv_in = torch.nn.Conv3d(in_channels, self.ch, kernel_size=3, stride=1, padding=1)
        curr_res = resolution
        in_ch_mult = (1,) + tuple(ch_mult)
        self.down = nn.ModuleList()
        for i_level in range(self.num_resolutions):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_in, block_out = ch * in_ch_mult[i_level], ch * ch_mult[i_level]
            for _ in range(self.num_res_blocks):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, temb_channels=self.temb_ch, dropout=dropout))
                block_in = block_out
            down = nn.Module()
            down.block, down.attn = block, attn
            if i_level != self.num_resolutions - 1:
                down.downsample = Downsample(block_in, resamp_with_conv)
                curr_res = curr_res // 2
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=self.temb_ch, dropout=dropout)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=self.temb_ch, dropout=dropout)
        self.norm_out = Normalize(block_in)
        self.conv_out = torch.nn.Conv3d(block_in, 2 * z_channels if double_z else z_channels, kernel_size=3, stride=1, padding=1)

    def pack(self):
        return {"in": _small_cin_conv_pack(self.conv_in)}

    def run_trunk(self, x):
        """fp32 NCDHW volume -> channels-last bf16 GELU(norm_out(h)), i.e. everything up to conv_out."""
        w, b, kp = self._pk()["in"]
        h = ops.linear_tokens(ops.im2col_small(x.float().contiguous(), kp=kp), w, bias=b)
        for i_level in range(self.num_resolutions):
            for i_block in range(self.num_res_blocks):
                h = self.down[i_level].block[i_block].run(h)
            if i_level != self.num_resolutions - 1:
                h = self.down[i_level].downsample.run(h)
        h = self.mid.block_2.run(self.mid.attn_1.run(self.mid.block_1.run(h)))
        return _gn(self.norm_out, h, ops.ACT_GELU)

    @torch.no_grad()
    def forward(self, x):
        a = self.run_trunk(x)
        return ops.conv3d_small_cout(a, ops.pack_small_cout_conv(self.conv_out.weight), _f(self.conv_out.bias),
                                     self.conv_out.weight.shape[0])


class Decoder3D(_Packed):
    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, give_pre_end=False, activ="gelu", **ignorekwargs):
        super().__init__()
        if activ != "gelu" or len(attn_resolutions) > 0 or give_pre_end:
            raise NotImplementedError("the shape branch uses activ='gelu', attn_resolutions=[] and give_pre_end=False")
        self.ch, self.temb_ch = ch, 0
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        self.resolution, self.in_channels, self.give_pre_end = resolution, in_channels, give_pre_end
        block_in = ch * ch_mult[self.num_resolutions - 1]
        curr_res = resolution // 2 ** (self.num_resolutions - 1)
        self.z_shape = (1, z_channels, curr_res, curr_res, curr_res)
        print("Working with z of shape {} = {} dimensions.".format(self.z_shape, np.prod(self.z_shape)))
        self.conv_in = torch.nn.Conv3d(z_channels, block_in, kernel_size=3, stride=1, padding=1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=self.temb_ch, dropout=dropout)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=self.temb_ch, dropout=dropout)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_out = ch * ch_mult[i_level]
            for _ in range(self.num_res_blocks):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, temb_channels=self.temb_ch, dropout=dropout))
                block_in = block_out
            up = nn.Module()
            up.block, up.attn = block, attn
            if i_level != 0:
                up.upsample = Upsample(block_in, resamp_with_conv)
                curr_res = curr_res * 2
            self.up.insert(0, up)
        self.norm_out = Normalize(block_in)
        self.conv_out = torch.nn.Conv3d(block_in, out_ch, kernel_size=3, stride=1, padding=1)

    def pack(self):
        return {"in": _small_cin_conv_pack(self.conv_in), "out": (ops.pack_small_cout_conv(self.conv_out.weight), _f(self.conv_out.bias))}

    @torch.no_grad()
    def forward(self, z):
        """z: fp32 NCDHW (B, z_channels, r, r, r) -> fp32 NCDHW (B, out_ch, R, R, R)."""
        self.last_z_shape = z.shape
        pk = self._pk()
        w, b, kp = pk["in"]
        h = ops.linear_tokens(ops.im2col_small(z.float().contiguous(), kp=kp), w, bias=b)
        h = self.mid.block_2.run(self.mid.attn_1.run(self.mid.block_1.run(h)))
        for i_level in reversed(range(self.num_resolutions)):
            for i_block in range(self.num_res_blocks):
                h = self.up[i_level].block[i_block].run(h)
            if i_level != 0:
                h = self.up[i_level].upsample.run(h)
        wo, bo = pk["out"]
        return ops.conv3d_small_cout(_gn(self.norm_out, h, ops.ACT_GELU), wo, bo, self.conv_out.weight.shape[0])
