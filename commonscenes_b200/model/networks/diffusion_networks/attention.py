"""Transformer blocks of the denoiser on the B200 kernels.

Same class names, constructor signatures and state-dict keys as the reference's
model/networks/diffusion_networks/attention.py (GEGLU :39-46, FeedForward :49-66, CrossAttention :154-219,
BasicTransformerBlock :222-245, SpatialTransformer3D :298-351).  The nn.Linear / nn.Conv3d / norm children
are parameter holders; `run()` methods execute on channels-last bf16 activations through the C ABI:

    GroupNorm(eps 1e-6) -> 1x1x1 proj_in -> [LN -> fused qkv GEMM -> fused softmax(QK^T)V -> to_out GEMM
    (+bias +cross-attention vector +residual)] -> [LN -> GEGLU GEMM -> x*gelu(g) -> GEMM (+residual)]
    -> 1x1x1 proj_out (+residual)

Cross-attention: the scene-graph conditioning is ONE token per object (VAEGAN_V2FULL.py:237-240), so
softmax over that single key is exactly 1 and attn2(x, ctx) == to_out(to_v(ctx)) for every query
(SURVEY.md §0.3).  That per-sample vector is computed once for all 11 blocks by one small GEMV and added
in the epilogue of attn1's to_out GEMM.  Longer contexts take the generic path (q/k/v GEMMs + the same
attention kernel).
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch
import torch.nn as nn

from .... import ops


def exists(val):
    return val is not None


def default(val, d):
    return val if exists(val) else (d() if callable(d) else d)


def zero_module(module):
    for p in module.parameters():
        p.detach().zero_()
    return module


def Normalize(in_channels):
    return torch.nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)


def _pad_head_dim(d: int) -> int:
    for dp in (32, 64, 96, 128, 256):
        if d <= dp:
            return dp
    raise ValueError(f"head dim {d} > 256 is not supported by cs_attention")


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(nn.Module):
    FUSE_GEGLU = os.environ.get("CS_FUSE_GEGLU", "1") != "0"

    def __init__(self, dim, dim_out=None, mult=4, glu=False, dropout=0.):
        super().__init__()
        if not glu:
            raise NotImplementedError("the reference's transformer blocks always use the gated feed-forward")
        inner_dim = int(dim * mult)
        dim_out = default(dim_out, dim)
        self.net = nn.Sequential(GEGLU(dim, inner_dim), nn.Dropout(dropout), nn.Linear(inner_dim, dim_out))

    def pack(self):
        # "w1" / "b1": the plain projection (the training path keeps the pre-activation for the backward: unet_train.py).
        # "w1f" / "b1f": rows interleaved 16 value / 16 gate columns for the ACT_GEGLU epilogue, packed on first inference use.
        return {"w1": ops.pack_linear_weight(self.net[0].proj.weight), "b1": self.net[0].proj.bias.detach().float().contiguous(),
                "w1f": None, "b1f": None, "fused": False,
                "w2": ops.pack_linear_weight(self.net[2].weight), "b2": self.net[2].bias.detach().float().contiguous()}

    def run(self, pk, x_ln, residual):
        # FUSE_GEGLU: x * gelu(gate) inside the projection's epilogue (ACT_GEGLU): the (tokens, 8 C) pre-activation -- 470 MB
        # per launch at batch 64 -- is never written or re-read.  Round 1 measured this SLOWER (454 us vs 264 + 110 us) because
        # the erf evaluation (exp + reciprocal) made the epilogue issue-bound; with gelu_tanh_fit (cs_common.cuh: one
        # MUFU.TANH + 9 FP instructions, 2.9e-5 from the erf form) the epilogue fits under the K = 448 / 672 main loop.
        if FeedForward.FUSE_GEGLU:
            if pk["w1f"] is None:
                pk["w1f"], pk["b1f"] = ops.pack_geglu_weight(self.net[0].proj.weight, self.net[0].proj.bias)
            h = ops.linear_tokens(x_ln, pk["w1f"], bias=pk["b1f"], act=ops.ACT_GEGLU)
        else:
            h = ops.geglu(ops.linear_tokens(x_ln, pk["w1"], bias=pk["b1"]))
        return ops.linear_tokens(h, pk["w2"], bias=pk["b2"], residual=residual)


class CrossAttention(nn.Module):
    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, dropout=0.):
        super().__init__()
        inner_dim = dim_head * heads
        self.is_self = context_dim is None
        context_dim = default(context_dim, query_dim)
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.dim_head = dim_head
        self.to_q = nn.Linear(query_dim, inner_dim, bias=False)
        self.to_k = nn.Linear(context_dim, inner_dim, bias=False)
        self.to_v = nn.Linear(context_dim, inner_dim, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, query_dim), nn.Dropout(dropout))

    # --- self-attention: one GEMM produces q|k|v with every head zero-padded to an MMA-friendly width ---
    def pack_self(self):
        h, d = self.heads, self.dim_head
        dp = _pad_head_dim(d)
        cin = self.to_q.weight.shape[1]
        w = torch.zeros(3, h, dp, cin, dtype=torch.float32, device=self.to_q.weight.device)
        for i, lin in enumerate((self.to_q, self.to_k, self.to_v)):
            w[i, :, :d] = lin.weight.detach().float().reshape(h, d, cin)
        return {"wqkv": ops.pack_linear_weight(w.reshape(3 * h * dp, cin)), "dp": dp,
                "wo": ops.pack_linear_weight(self.to_out[0].weight), "bo": self.to_out[0].bias.detach().float().contiguous()}

    def run_self(self, pk, x_ln, residual, rowvec=None):
        """to_out(softmax(q k^T * scale) v) + bias (+ rowvec[b]) + residual on (B, D, H, W, C) tokens."""
        B, D, H, W, _ = x_ln.shape
        h, d, dp = self.heads, self.dim_head, pk["dp"]
        qkv = ops.linear_tokens(x_ln, pk["wqkv"]).view(B, D * H * W, 3 * h * dp)
        q, k, v = (qkv[:, :, i * h * dp:(i + 1) * h * dp] for i in range(3))
        o = ops.attention(q, k, v, heads=h, head_dim=d, head_dim_padded=dp, scale=self.scale)
        return ops.linear_tokens(o.view(B, D, H, W, h * d), pk["wo"], bias=pk["bo"], rowvec=rowvec, residual=residual)

    # --- generic cross-attention over a multi-token context (API completeness; v2_full always has ONE token) ---
    def pack_cross(self):
        h, d = self.heads, self.dim_head
        dp = _pad_head_dim(d)

        def padded(lin):
            w = torch.zeros(h, dp, lin.weight.shape[1], dtype=torch.float32, device=lin.weight.device)
            w[:, :d] = lin.weight.detach().float().reshape(h, d, -1)
            return w.reshape(h * dp, -1)
        return {"wq": ops.pack_linear_weight(padded(self.to_q)), "wk": padded(self.to_k).contiguous(),
                "wv": padded(self.to_v).contiguous(), "dp": dp,
                "wo": ops.pack_linear_weight(self.to_out[0].weight), "bo": self.to_out[0].bias.detach().float().contiguous()}

    def run_cross(self, pk, x_ln, context, residual):
        """to_out(softmax(q k^T * scale) v) + bias + residual with k, v projected from `context` (B, M, context_dim) fp32."""
        B, D, H, W, _ = x_ln.shape
        h, d, dp = self.heads, self.dim_head, pk["dp"]
        M = context.shape[1]
        q = ops.linear_tokens(x_ln, pk["wq"]).view(B, D * H * W, h * dp)
        ctx = context.float().reshape(B * M, -1).contiguous()
        k = ops.cast_bf16(ops.linear_small(ctx, pk["wk"])).view(B, M, h * dp)
        v = ops.cast_bf16(ops.linear_small(ctx, pk["wv"])).view(B, M, h * dp)
        o = ops.attention(q, k, v, heads=h, head_dim=d, head_dim_padded=dp, scale=self.scale)
        return ops.linear_tokens(o.view(B, D, H, W, h * d), pk["wo"], bias=pk["bo"], residual=residual)

    # --- cross-attention with a single context token: attn2(x, ctx) == to_out(to_v(ctx)) ---
    def composed_single_token(self):
        """(W_out @ W_v, b_out) in fp32: maps the context token straight to the block's additive vector."""
        # fp32 on the library's own small-GEMM kernel (K = inner dim <= 672, no split-K: deterministic).  This fold runs
        # after every optimizer step of the native training loop; torch's fp64 matmul put a cuBLAS/cutlass kernel there.
        from .... import ops_bwd
        wo = self.to_out[0].weight.detach().float().contiguous()
        wv = self.to_v.weight.detach().float().contiguous()
        return ops_bwd.sgemm(wo, wv), self.to_out[0].bias.detach().float().contiguous()


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, n_heads, d_head, dropout=0., context_dim=None, gated_ff=True, checkpoint=True):
        super().__init__()
        self.attn1 = CrossAttention(query_dim=dim, heads=n_heads, dim_head=d_head, dropout=dropout)
        self.ff = FeedForward(dim, dropout=dropout, glu=gated_ff)
        self.attn2 = CrossAttention(query_dim=dim, context_dim=context_dim, heads=n_heads, dim_head=d_head, dropout=dropout)
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.norm3 = nn.LayerNorm(dim)
        self.checkpoint = checkpoint

    def pack(self):
        f = lambda t: t.detach().float().contiguous()
        return {"attn1": self.attn1.pack_self(), "ff": self.ff.pack(), "attn2": None,
                "ln1": (f(self.norm1.weight), f(self.norm1.bias)), "ln2": (f(self.norm2.weight), f(self.norm2.bias)),
                "ln3": (f(self.norm3.weight), f(self.norm3.bias))}

    def run(self, pk, x, ctx_vec, context=None):
        """ctx_vec: the block's single-token cross-attention vector (fast path), or None with `context` (B, M > 1, dim)."""
        if ctx_vec is None:
            if pk["attn2"] is None:
                pk["attn2"] = self.attn2.pack_cross()          # packed on first use: the v2_full path never needs it
            x = self.attn1.run_self(pk["attn1"], ops.layernorm(x, *pk["ln1"], eps=self.norm1.eps), residual=x)
            x = self.attn2.run_cross(pk["attn2"], ops.layernorm(x, *pk["ln2"], eps=self.norm2.eps), context, residual=x)
            return self.ff.run(pk["ff"], ops.layernorm(x, *pk["ln3"], eps=self.norm3.eps), residual=x)
        # x = attn1(norm1(x)) + x ; x = attn2(norm2(x), ctx) + x   [attn2 == ctx_vec, independent of x]
        x = self.attn1.run_self(pk["attn1"], ops.layernorm(x, *pk["ln1"], eps=self.norm1.eps), residual=x, rowvec=ctx_vec)
        # x = ff(norm3(x)) + x
        return self.ff.run(pk["ff"], ops.layernorm(x, *pk["ln3"], eps=self.norm3.eps), residual=x)


def init_weights(m):
    if isinstance(m, nn.Conv3d):
        nn.init.xavier_normal_(m.weight)


class SpatialTransformer3D(nn.Module):
    def __init__(self, in_channels, n_heads, d_head, depth=1, dropout=0., context_dim=None):
        super().__init__()
        self.in_channels = in_channels
        inner_dim = n_heads * d_head
        self.norm = Normalize(in_channels)
        self.proj_in = nn.Conv3d(in_channels, inner_dim, kernel_size=1, stride=1, padding=0)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner_dim, n_heads, d_head, dropout=dropout, context_dim=context_dim) for _ in range(depth)])
        self.proj_out = zero_module(nn.Conv3d(inner_dim, in_channels, kernel_size=1, stride=1, padding=0))
        self.apply(init_weights)   # as in the reference: xavier on the 1x1x1 convs (this un-zeroes proj_out.weight)

    def pack(self):
        f = lambda t: t.detach().float().contiguous()
        return {"gn": (f(self.norm.weight), f(self.norm.bias)),
                "w_in": ops.pack_conv_weight(self.proj_in.weight), "b_in": f(self.proj_in.bias),
                "w_out": ops.pack_conv_weight(self.proj_out.weight), "b_out": f(self.proj_out.bias),
                "blocks": [b.pack() for b in self.transformer_blocks]}

    def run(self, pk, x, ctx_vecs, arena, context=None):
        """x: Act over (B, D, H, W, C) bf16 (with its GroupNorm sums); ctx_vecs: one fp32 (B, inner) vector per block
        (single-token context), or None entries with the raw multi-token `context`."""
        from .openai_model_3d import _with_stats
        B, D, H, W, C = x.t.shape
        t = ops.linear_tokens(ops.groupnorm_fused(x.t, x.stat, *pk["gn"], eps=self.norm.eps), pk["w_in"], bias=pk["b_in"])
        for blk, bpk, vec in zip(self.transformer_blocks, pk["blocks"], ctx_vecs):
            t = blk.run(bpk, t, vec, context)
        return _with_stats(arena, lambda st: ops.linear_tokens(t, pk["w_out"], bias=pk["b_out"], residual=x.t, stat_sum=st),
                           B, C, D * H * W)
