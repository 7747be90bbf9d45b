"""UNet3DModel — the shape-branch denoiser on the B200 kernels.

Drop-in for the reference's model/networks/diffusion_networks/openai_model_3d.py: same class names,
constructor signature, child-module names (hence the same 496 state-dict keys) and the same
forward(x, timesteps, context) contract (NCDHW fp32 in / out; openai_model_3d.py:752-789).

Differences are internal only:
  * activations are channels-last bf16 between kernels (fp32 accumulation, fp32 norm statistics);
  * every 3x3x3 / 1x1x1 conv and linear is the tcgen05 implicit GEMM (cs_conv3d); `h + emb_out`, the
    skip add, the skip-concat of the decoder (two-source K loop) and the GroupNorm statistics of a conv's
    output are fused into its epilogue / operand fetch;
  * all 17 `emb_layers` share one GEMV per step and all 11 cross-attention blocks one GEMV per context;
  * no activation checkpointing, no NaN-check host syncs (attention.py:183,205), CUDA-graph capturable.
The nn.Conv3d / nn.Linear / norm children only hold parameters; packed bf16 copies are rebuilt lazily when
a parameter's version counter changes (i.e. after an optimizer step or load_state_dict).
"""
from __future__ import annotations

from abc import abstractmethod
from typing import List, Optional

import torch
import torch.nn as nn

from .... import _lib, ops
from .attention import SpatialTransformer3D, _pad_head_dim
from .ldm_diffusion_util import conv_nd, linear, normalization, timestep_embedding, zero_module


def _f(t: torch.Tensor) -> torch.Tensor:
    return t.detach().float().contiguous()


class Act:
    """A channels-last bf16 activation together with the per-(sample, channel) (sum, sum of squares) its producer's
    epilogue accumulated — everything a later GroupNorm of this tensor needs (int64 fixed point (B, C, 2), read-only for
    consumers; order-independent integer accumulation makes every forward bit-reproducible)."""
    __slots__ = ("t", "stat")

    def __init__(self, t, stat):
        self.t, self.stat = t, stat


class StatArena:
    """One zero-initialised int64 (fixed-point sums) buffer per forward pass from which the GroupNorm sum buffers are carved (a single memset
    instead of one per tensor; CUDA-graph friendly: the buffer is persistent, the memset is captured)."""

    def __init__(self, device, numel):
        self.buf = torch.zeros(numel, dtype=ops.STAT_DTYPE, device=device)
        self.off = 0

    def reset(self):
        self.buf.zero_()
        self.off = 0

    def take(self, B, C):
        n = B * C * 2
        if self.off + n > self.buf.numel():
            raise RuntimeError("StatArena exhausted (repack() under-counted the normalised tensors)")
        v = self.buf[self.off:self.off + n].view(B, C, 2)
        self.off += n
        return v


def _with_stats(arena, make, B, C, S):
    """Run `make(stat_or_None)` (a conv / linear launch) so that its output comes with GroupNorm sums: fused into the
    epilogue when whole warps of accumulator rows belong to one sample, else by a separate statistics pass."""
    stat = arena.take(B, C)
    if S % 32 == 0:
        return Act(make(stat), stat)
    t = make(None)
    ops.groupnorm_stats(t, stat)
    return Act(t, stat)


class TimestepBlock(nn.Module):
    """Any module whose run() takes the timestep-embedding vector as a second argument."""

    @abstractmethod
    def run(self, pk, x, emb_vec, skip=None):
        ...


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    """Children are applied in order; ResBlocks get the time vector, transformers the context vectors
    (reference: openai_model_3d.py:113-127)."""


class Upsample(nn.Module):
    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.dims = dims
        if use_conv:
            self.conv = conv_nd(dims, self.channels, self.out_channels, 3, padding=padding)

    PHASE_CONV = True      # inference: fold the nearest-upsample into the conv (four / eight merged-tap phase convs)

    def pack(self):
        return {"w": ops.pack_conv_weight(self.conv.weight), "b": _f(self.conv.bias), "phases": None} if self.use_conv else {}

    def run(self, pk, x, arena):
        # dims == 3 keeps D and doubles H, W (the inherited "video" convention, :150-153); dims == 4 is isotropic
        factors = (1, 2, 2) if self.dims == 3 else (2, 2, 2)
        B, D, H, W, _ = x.t.shape
        if self.use_conv and Upsample.PHASE_CONV and D * H * W >= 128:
            # F.interpolate(nearest) + conv3x3x3 (:150-158) as one merged-tap conv per output phase on the LOW-resolution
            # tensor (ops.pack_upsample_phase_weights): no up-sampled intermediate, 12 of 27 taps (8 when isotropic)
            if pk["phases"] is None:
                pk["phases"] = ops.pack_upsample_phase_weights(self.conv.weight, factors)
            out = torch.empty((B, D * factors[0], H * factors[1], W * factors[2], self.out_channels), dtype=torch.bfloat16,
                              device=x.t.device)

            def make(st):
                for offs, ks, pad, pad_back, w in pk["phases"]:
                    ops.conv3d(x.t, w, ksize=ks, pad=pad, pad_back=pad_back, bias=pk["b"], stat_sum=st, out=out, phase=(factors, offs))
                return out
            if (D * H * W) % 32 == 0:
                stat = arena.take(B, self.out_channels)
                return Act(make(stat), stat)
            stat = arena.take(B, self.out_channels)
            return Act(make(None), ops.groupnorm_stats(out, stat))
        u = ops.upsample_nearest(x.t, factors)
        B, D, H, W, _ = u.shape
        if not self.use_conv:
            stat = arena.take(B, self.channels)
            return Act(u, ops.groupnorm_stats(u, stat))
        return _with_stats(arena, lambda st: ops.conv3d(u, pk["w"], bias=pk["b"], stat_sum=st), B, self.out_channels, D * H * W)


class Downsample(nn.Module):
    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.dims = dims
        self.stride = (2, 2, 2) if dims != 3 else (1, 2, 2)
        if not use_conv:
            raise NotImplementedError("average-pool downsampling is not used by the reference configs (conv_resample=True)")
        self.op = conv_nd(dims, self.channels, self.out_channels, 3, stride=self.stride, padding=padding)

    def pack(self):
        return {"w": ops.pack_conv_weight(self.op.weight), "b": _f(self.op.bias)}

    def run(self, pk, x, arena):
        B, D, H, W, _ = x.t.shape
        So = (D // self.stride[0]) * (H // self.stride[1]) * (W // self.stride[2])
        return _with_stats(arena, lambda st: ops.conv3d(x.t, pk["w"], stride=self.stride, bias=pk["b"], stat_sum=st), B,
                           self.out_channels, So)


class ResBlock(TimestepBlock):
    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_conv=False, use_scale_shift_norm=False,
                 dims=2, use_checkpoint=False, up=False, down=False):
        super().__init__()
        if use_scale_shift_norm or up or down or use_conv:
            raise NotImplementedError("scale-shift norm / resblock up-down / conv skip are not used by the reference configs")
        self.channels = channels
        self.emb_channels = emb_channels
        self.dropout = dropout
        self.out_channels = out_channels or channels
        self.use_checkpoint = use_checkpoint
        self.in_layers = nn.Sequential(normalization(channels), nn.SiLU(), conv_nd(dims, channels, self.out_channels, 3, padding=1))
        self.emb_layers = nn.Sequential(nn.SiLU(), linear(emb_channels, self.out_channels))
        self.out_layers = nn.Sequential(normalization(self.out_channels), nn.SiLU(), nn.Dropout(p=dropout),
                                        zero_module(conv_nd(dims, self.out_channels, self.out_channels, 3, padding=1)))
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        else:
            self.skip_connection = conv_nd(dims, channels, self.out_channels, 1)

    def pack(self, skip_channels: int = 0):
        """skip_channels > 0: the block's input is cat([h, skip]) with that many skip channels (decoder blocks)."""
        split = (self.channels - skip_channels, skip_channels) if skip_channels else None
        pk = {"gn1": (_f(self.in_layers[0].weight), _f(self.in_layers[0].bias)),
              "w1": ops.pack_conv_weight(self.in_layers[2].weight), "b1": _f(self.in_layers[2].bias),
              "gn2": (_f(self.out_layers[0].weight), _f(self.out_layers[0].bias)),
              "w2": ops.pack_conv_weight(self.out_layers[3].weight), "b2": _f(self.out_layers[3].bias)}
        if not isinstance(self.skip_connection, nn.Identity):
            pk["ws"] = ops.pack_conv_weight(self.skip_connection.weight, split)
            pk["bs"] = _f(self.skip_connection.bias)
        return pk

    def run(self, pk, x, emb_vec, arena, skip=None):
        """x (and optionally the encoder skip tensor, logically concatenated after x on the channel axis): Act over
        (B, D, H, W, C) bf16; emb_vec: fp32 (B, out_channels) = emb_layers(emb).  Returns the block output as an Act."""
        B, D, H, W, _ = x.t.shape
        S, C = D * H * W, self.out_channels
        a = ops.groupnorm_fused(x.t, x.stat, *pk["gn1"], eps=self.in_layers[0].eps, act=ops.ACT_SILU,
                                x2=None if skip is None else skip.t, stat2=None if skip is None else skip.stat)
        h = _with_stats(arena, lambda st: ops.conv3d(a, pk["w1"], bias=pk["b1"], rowvec=emb_vec, stat_sum=st), B, C, S)
        a = ops.groupnorm_fused(h.t, h.stat, *pk["gn2"], eps=self.out_layers[0].eps, act=ops.ACT_SILU)
        if "ws" in pk:
            res = ops.linear_tokens(x.t, pk["ws"], bias=pk["bs"], x2=None if skip is None else skip.t)   # 1x1x1 skip on the raw concat
        else:
            res = x.t
        return _with_stats(arena, lambda st: ops.conv3d(a, pk["w2"], bias=pk["b2"], residual=res, stat_sum=st), B, C, S)


class QKVAttentionLegacy(nn.Module):
    """Parameter-free marker kept for module-tree parity (openai_model_3d.py:386-411): heads are split BEFORE q/k/v."""

    def __init__(self, n_heads):
        super().__init__()
        self.n_heads = n_heads


class QKVAttention(QKVAttentionLegacy):
    """use_new_attention_order=True (:417-447): q/k/v are split before the heads."""


class AttentionBlock(nn.Module):
    """Self-attention over all voxels, used by the concat-conditioning denoiser instead of the spatial transformer
    (reference openai_model_3d.py:317-364): x + proj_out(attention(qkv(GroupNorm32(x)))), tokens in (d h w) order,
    softmax(q k^T / sqrt(ch)) (the reference scales q and k by ch^-1/4 each).  Same children / state-dict keys
    (norm, qkv [Conv1d k=1], proj_out [Conv1d k=1]); run() = GroupNorm kernel -> one qkv GEMM (head-padded q|k|v layout,
    the legacy head-major row order is undone in the weight packing) -> flash attention kernel -> proj_out GEMM whose
    epilogue adds the residual and accumulates the next GroupNorm's sums."""

    def __init__(self, channels, num_heads=1, num_head_channels=-1, use_checkpoint=False, use_new_attention_order=False):
        super().__init__()
        self.channels = channels
        if num_head_channels == -1:
            self.num_heads = num_heads
        else:
            assert channels % num_head_channels == 0, \
                f"q,k,v channels {channels} is not divisible by num_head_channels {num_head_channels}"
            self.num_heads = channels // num_head_channels
        self.use_checkpoint = use_checkpoint
        self.norm = normalization(channels)
        self.qkv = conv_nd(1, channels, channels * 3, 1)
        self.attention = QKVAttention(self.num_heads) if use_new_attention_order else QKVAttentionLegacy(self.num_heads)
        self.proj_out = zero_module(conv_nd(1, channels, channels, 1))

    def pack(self):
        C, h = self.channels, self.num_heads
        ch = C // h
        dp = _pad_head_dim(ch)
        w = self.qkv.weight.detach().float().reshape(3 * C, C)
        b = self.qkv.bias.detach().float()
        if isinstance(self.attention, QKVAttention):                  # rows ordered [q|k|v][head][ch]
            w, b = w.reshape(3, h, ch, C), b.reshape(3, h, ch)
        else:                                                         # legacy: rows ordered [head][q|k|v][ch]
            w, b = w.reshape(h, 3, ch, C).permute(1, 0, 2, 3), b.reshape(h, 3, ch).permute(1, 0, 2)
        wp = torch.zeros(3, h, dp, C, dtype=torch.float32, device=w.device)
        bp = torch.zeros(3, h, dp, dtype=torch.float32, device=w.device)
        wp[:, :, :ch], bp[:, :, :ch] = w, b
        return {"gn": (_f(self.norm.weight), _f(self.norm.bias)), "dp": dp,
                "wqkv": ops.pack_linear_weight(wp.reshape(3 * h * dp, C)), "bqkv": bp.reshape(-1).contiguous(),
                "wo": ops.pack_linear_weight(self.proj_out.weight.detach().float().reshape(C, C)), "bo": _f(self.proj_out.bias)}

    def run(self, pk, x, arena):
        B, D, H, W, C = x.t.shape
        h, dp = self.num_heads, pk["dp"]
        ch = C // h
        a = ops.groupnorm_fused(x.t, x.stat, *pk["gn"], eps=self.norm.eps)
        qkv = ops.linear_tokens(a, pk["wqkv"], bias=pk["bqkv"]).view(B, D * H * W, 3 * h * dp)
        q, k, v = (qkv[:, :, i * h * dp:(i + 1) * h * dp] for i in range(3))
        o = ops.attention(q, k, v, heads=h, head_dim=ch, head_dim_padded=dp, scale=ch ** -0.5)
        return _with_stats(arena, lambda st: ops.linear_tokens(o.view(B, D, H, W, C), pk["wo"], bias=pk["bo"], residual=x.t,
                                                               stat_sum=st), B, C, D * H * W)


class UNet3DModel(nn.Module):
    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None,
                 use_checkpoint=False, use_fp16=False, num_heads=-1, num_head_channels=-1, num_heads_upsample=-1,
                 use_scale_shift_norm=False, resblock_updown=False, use_new_attention_order=False,
                 use_spatial_transformer=False, transformer_depth=1, context_dim=None, n_embed=None, legacy=True):
        super().__init__()
        if use_spatial_transformer:
            assert context_dim is not None, "You forgot to include the dimension of your cross-attention conditioning..."
        if context_dim is not None:
            assert use_spatial_transformer, "You forgot to use the spatial transformer for your cross-attention conditioning..."
            if not isinstance(context_dim, int):
                context_dim = list(context_dim)
        if num_classes is not None or resblock_updown or n_embed is not None:
            raise NotImplementedError("class conditioning / resblock_updown / codebook-id head are unused by the reference configs")
        if num_heads_upsample == -1:
            num_heads_upsample = num_heads
        if num_heads == -1:
            assert num_head_channels != -1, "Either num_heads or num_head_channels has to be set"
        if num_head_channels == -1:
            assert num_heads != -1, "Either num_heads or num_head_channels has to be set"

        self.image_size = image_size
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = attention_resolutions
        self.dropout = dropout
        self.channel_mult = channel_mult
        self.conv_resample = conv_resample
        self.num_classes = num_classes
        self.use_checkpoint = use_checkpoint
        self.dtype = torch.float32
        self.num_heads = num_heads
        self.num_head_channels = num_head_channels
        self.num_heads_upsample = num_heads_upsample
        self.predict_codebook_ids = False
        self.dims = dims

        time_embed_dim = model_channels * 4
        self.time_embed = nn.Sequential(linear(model_channels, time_embed_dim), nn.SiLU(), linear(time_embed_dim, time_embed_dim))

        def transformer(ch, heads):
            if num_head_channels == -1:
                dim_head = ch // heads
            else:
                heads, dim_head = ch // num_head_channels, num_head_channels
            if legacy:
                dim_head = ch // heads
            if not use_spatial_transformer:        # the concat-conditioning variant (reference :593-598, 649-654, 690-697)
                return AttentionBlock(ch, use_checkpoint=use_checkpoint, num_heads=heads, num_head_channels=dim_head,
                                      use_new_attention_order=use_new_attention_order)
            return SpatialTransformer3D(ch, heads, dim_head, depth=transformer_depth, context_dim=context_dim)

        def res(cin, cout):
            return ResBlock(cin, time_embed_dim, dropout, out_channels=cout, dims=dims, use_checkpoint=use_checkpoint)

        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(conv_nd(dims, in_channels, model_channels, 3, padding=1))])
        input_block_chans = [model_channels]
        ch, ds = model_channels, 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [res(ch, mult * model_channels)]
                ch = mult * model_channels
                if ds in attention_resolutions:
                    layers.append(transformer(ch, num_heads))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                input_block_chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch, conv_resample, dims=dims, out_channels=ch)))
                input_block_chans.append(ch)
                ds *= 2
        self.middle_block = TimestepEmbedSequential(res(ch, ch), transformer(ch, num_heads), res(ch, ch))
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                ich = input_block_chans.pop()
                layers = [res(ch + ich, model_channels * mult)]
                layers[0].skip_in_channels = ich      # this block consumes cat([h, skip]): `ich` channels come from the skip
                ch = model_channels * mult
                if ds in attention_resolutions:
                    layers.append(transformer(ch, num_heads_upsample))
                if level and i == num_res_blocks:
                    layers.append(Upsample(ch, conv_resample, dims=dims, out_channels=ch))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
        self.out = nn.Sequential(normalization(ch), nn.SiLU(), zero_module(conv_nd(dims, model_channels, out_channels, 3, padding=1)))

        self._packed = None
        self._packed_key = None

    # ------------------------------------------------------------------------------------------
    # weight packing
    # ------------------------------------------------------------------------------------------
    def _blocks(self):
        yield from self.input_blocks
        yield self.middle_block
        yield from self.output_blocks

    def _version_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def repack(self) -> None:
        """Rebuild the kernel-layout (bf16, K-major) copies of all weights."""
        dev = self.out[2].weight.device
        pk = {"blocks": [], "te": (_f(self.time_embed[0].weight), _f(self.time_embed[0].bias),
                                   _f(self.time_embed[2].weight), _f(self.time_embed[2].bias))}
        emb_w, emb_b, ca_w, ca_b = [], [], [], []
        emb_off = ca_off = 0
        stat_channels = 0            # total channels of all tensors that carry GroupNorm sums (sizes the StatArena)
        for bi, block in enumerate(self._blocks()):
            entries = []
            for layer in block:
                if isinstance(layer, ResBlock):
                    e = {"kind": "res", "pk": layer.pack(getattr(layer, "skip_in_channels", 0)), "emb": (emb_off, layer.out_channels)}
                    emb_w.append(_f(layer.emb_layers[1].weight)); emb_b.append(_f(layer.emb_layers[1].bias))
                    emb_off += layer.out_channels
                    stat_channels += 2 * layer.out_channels
                elif isinstance(layer, SpatialTransformer3D):
                    offs = []
                    for tb in layer.transformer_blocks:
                        w, b = tb.attn2.composed_single_token()
                        ca_w.append(w); ca_b.append(b)
                        offs.append((ca_off, w.shape[0])); ca_off += w.shape[0]
                    e = {"kind": "st", "pk": layer.pack(), "ca": offs}
                    stat_channels += layer.in_channels
                elif isinstance(layer, AttentionBlock):
                    e = {"kind": "attn", "pk": layer.pack()}
                    stat_channels += layer.channels
                elif isinstance(layer, (Downsample, Upsample)):
                    e = {"kind": "resample", "pk": layer.pack()}
                    stat_channels += layer.out_channels
                elif isinstance(layer, nn.Conv3d):      # the stem: few input channels -> im2col + GEMM
                    wp, kp = ops.pack_patch_weight(layer.weight)
                    e = {"kind": "stem", "pk": {"w": wp, "b": _f(layer.bias), "kp": kp}}
                    stat_channels += layer.out_channels
                else:
                    raise TypeError(f"unexpected layer {type(layer)}")
                entries.append(e)
            pk["blocks"].append(entries)
        pk["emb_w"], pk["emb_b"] = torch.cat(emb_w).contiguous(), torch.cat(emb_b).contiguous()
        if ca_w:
            pk["ca_w"], pk["ca_b"] = torch.cat(ca_w).contiguous(), torch.cat(ca_b).contiguous()
        else:                                               # no cross-attention (concat variant)
            pk["ca_w"] = pk["ca_b"] = None
        pk["stat_channels"] = stat_channels
        self._arenas = {}
        pk["out_gn"] = (_f(self.out[0].weight), _f(self.out[0].bias))
        pk["out_w"], pk["out_b"] = ops.pack_small_cout_conv(self.out[2].weight), _f(self.out[2].bias)
        self._packed, self._packed_key = pk, self._version_key()
        self._pack_generation = getattr(self, "_pack_generation", 0) + 1

    def _ensure_packed(self):
        if self._packed is None or self._packed_key != self._version_key():
            self.repack()
        return self._packed

    # ------------------------------------------------------------------------------------------
    # forward
    # ------------------------------------------------------------------------------------------
    def context_vectors(self, context: torch.Tensor) -> torch.Tensor:
        """All cross-attention outputs for a (B, 1, context_dim) conditioning: fp32 (B, sum of block widths).
        Depends only on the context, so samplers may compute it once per trajectory."""
        pk = self._ensure_packed()
        if pk["ca_w"] is None:
            raise ValueError("this UNet has no cross-attention (use_spatial_transformer=False): there is no context")
        if context.dim() != 3 or context.shape[1] != 1:
            raise ValueError("context_vectors() is the single-token fast path; pass multi-token contexts to forward()")
        return ops.linear_small(context[:, 0].float().contiguous(), pk["ca_w"], pk["ca_b"])

    def trainer(self):
        """The UNetTrainer (forward with saved activations + explicit backward kernels) bound to this module."""
        tr = getattr(self, "_trainer_obj", None)
        if tr is None:
            from .unet_train import UNetTrainer
            tr = self._trainer_obj = UNetTrainer(self)
        return tr

    def forward(self, x, timesteps=None, context=None, y=None, context_vecs=None, shared_prefix=False, **kwargs):
        """x: (B, C, D, H, W) fp32 NCDHW, timesteps: (B,) int64, context: (B, 1, context_dim) -> eps (B, C, D, H, W).

        With autograd enabled and trainable parameters (or a context that requires grad) the result carries a grad_fn:
        `loss.backward()` runs the explicit backward kernels (unet_train.py) and fills `.grad` of every parameter and of
        `context`, as the reference's autograd does (no gradient is produced for x: the reference never asks for it).

        Extensions: `context_vecs` = a cached context_vectors(context) result; x may hold B/r samples, in which
        case sample b reads x[b % (B/r)] (a guided sampler passes x once for [uncond; cond]); `shared_prefix=True`
        additionally promises timesteps[b] == timesteps[b % (B/r)], so the layers in front of the first cross-attention are
        evaluated once per distinct x (see _forward_inference).
        """
        assert (y is not None) == (self.num_classes is not None), "must specify y if and only if the model is class-conditional"
        if torch.is_grad_enabled() and context_vecs is None and context is not None and \
                (context.requires_grad or any(p.requires_grad for p in self.parameters())):
            params = [p for p in self.parameters() if p.requires_grad]
            return _UNetFunction.apply(self, x, timesteps, context, *params)
        if torch.is_grad_enabled() and context is None and context_vecs is None and (
                x.requires_grad or any(p.requires_grad for p in self.parameters())):
            # no cross-attention context (concat-conditioning variant): the conditioning gradient flows through x itself
            params = [p for p in self.parameters() if p.requires_grad]
            return _UNetInputFunction.apply(self, x, timesteps, *params)
        with torch.no_grad():
            return self._forward_inference(x, timesteps, context, context_vecs, shared_prefix)

    def _forward_inference(self, x, timesteps, context, context_vecs, shared_prefix=False):
        _lib.require_device()
        pk = self._ensure_packed()
        B = timesteps.shape[0]
        Bs = x.shape[0]
        x = x.float().contiguous()
        t_emb = timestep_embedding(timesteps, self.model_channels)
        w0, b0, w2, b2 = pk["te"]
        emb = ops.linear_small(ops.linear_small(t_emb, w0, b0, act_out=ops.ACT_SILU), w2, b2)
        emb_vecs = ops.linear_small(emb, pk["emb_w"], pk["emb_b"], act_in=ops.ACT_SILU)     # every ResBlock's emb_layers at once
        multi = context_vecs is None and context is not None and context.dim() == 3 and context.shape[1] > 1
        if pk["ca_w"] is None:          # no cross-attention layers (concat variant): the conditioning is inside x
            ca_vecs = None
        else:
            ca_vecs = None if multi else (context_vecs if context_vecs is not None else self.context_vectors(context))

        arena = self._arenas.get(B)
        if arena is None:
            arena = self._arenas[B] = StatArena(x.device, B * pk["stat_channels"] * 2)
        arena.reset()

        # Guided sampling evaluates [uncond; cond] on the SAME x and t: every layer before the first cross-attention sees
        # identical inputs in both halves (the stem, the two level-0 ResBlocks, the first Downsample and the ResBlock in front
        # of the first transformer: 11.5 % of the FLOPs).  With shared_prefix those run once on the Bs = B / r samples of x
        # and their outputs (including the encoder skips) are replicated when the conditioning first enters -- equal (up to the
        # fp32-atomic rounding noise any two launches show)
        # to evaluating them twice.  The caller vouches that timesteps[b] == timesteps[b % Bs] (the samplers build t that way).
        state = {"shared": bool(shared_prefix) and Bs < B and B % Bs == 0 and pk["ca_w"] is not None and not multi}
        hs: List[Act] = []

        def replicate(a: Act) -> Act:
            r = B // Bs
            return Act(torch.cat([a.t] * r), torch.cat([a.stat] * r))     # contiguous block copies (vectorised), not a strided repeat

        def run_block(block, entries, h, skip=None):
            for layer, e in zip(block, entries):
                nb = Bs if state["shared"] else B
                if e["kind"] == "res":
                    off, n = e["emb"]
                    h = layer.run(e["pk"], h, emb_vecs[:nb, off:off + n], arena, skip=skip)
                    skip = None
                elif e["kind"] == "st":
                    if state["shared"]:           # the conditioning enters here: leave the shared prefix
                        h = replicate(h)
                        hs[:] = [replicate(a) for a in hs]
                        state["shared"] = False
                    if multi:    # generic cross-attention over all context tokens (attention.py:172-219)
                        h = layer.run(e["pk"], h, [None] * len(e["ca"]), arena, context=context)
                    else:
                        h = layer.run(e["pk"], h, [ca_vecs[:, o:o + n] for o, n in e["ca"]], arena)
                elif e["kind"] in ("resample", "attn"):
                    h = layer.run(e["pk"], h, arena)
                else:
                    col = ops.im2col_small(h, batch=nb, kp=e["pk"]["kp"])
                    S = col.shape[1] * col.shape[2] * col.shape[3]
                    h = _with_stats(arena, lambda st: ops.linear_tokens(col, e["pk"]["w"], bias=e["pk"]["b"], stat_sum=st), nb,
                                    layer.out_channels, S)
            return h

        entries = pk["blocks"]
        n_in = len(self.input_blocks)
        h = x
        for i, block in enumerate(self.input_blocks):
            h = run_block(block, entries[i], h)
            hs.append(h)
        h = run_block(self.middle_block, entries[n_in], h)
        for i, block in enumerate(self.output_blocks):
            h = run_block(block, entries[n_in + 1 + i], h, skip=hs.pop())    # th.cat([h, hs.pop()], dim=1), never materialised raw
        a = ops.groupnorm_fused(h.t, h.stat, *pk["out_gn"], eps=self.out[0].eps, act=ops.ACT_SILU)
        return ops.conv3d_small_cout(a, pk["out_w"], pk["out_b"], self.out_channels)


class _UNetFunction(torch.autograd.Function):
    """Autograd bridge: forward = UNetTrainer.forward_train, backward = UNetTrainer.backward (explicit kernels)."""

    @staticmethod
    def forward(ctx, unet, x, timesteps, context, *params):
        eps, tape = unet.trainer().forward_train(x.detach(), timesteps, context.detach())
        ctx.unet, ctx.tape, ctx.params = unet, tape, params
        ctx.need_dcontext = context.requires_grad
        return eps

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_eps):
        sink, dctx = ctx.unet.trainer().backward(ctx.tape, d_eps, need_dcontext=ctx.need_dcontext)
        ctx.tape = None
        grads = tuple(sink.grads.get(p) for p in ctx.params)
        return (None, None, None, dctx) + grads


class _UNetInputFunction(torch.autograd.Function):
    """Autograd bridge for a UNet without cross-attention context (concat variant): gradients for the parameters and for x
    (whose extra channels carry the conditioning, network.py:25-27)."""

    @staticmethod
    def forward(ctx, unet, x, timesteps, *params):
        eps, tape = unet.trainer().forward_train(x.detach(), timesteps, None)
        ctx.unet, ctx.tape, ctx.params, ctx.need_dx = unet, tape, params, x.requires_grad
        return eps

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_eps):
        sink, _ = ctx.unet.trainer().backward(ctx.tape, d_eps, need_dcontext=False, need_dx=ctx.need_dx)
        dx = ctx.tape.get("dx") if ctx.need_dx else None
        ctx.tape = None
        grads = tuple(sink.grads.get(p) for p in ctx.params)
        return (None, dx, None) + grads

