"""Training path of UNet3DModel on the B200 kernels: forward with saved activations + explicit backward.

Reference: the denoiser is trained by autograd through UNet3DModel.forward (openai_model_3d.py:752-789) from
SDFusionText2ShapeModel.p_losses / backward (sdfusion_txt2shape_model.py:311-345, 568-575), with every block under
activation checkpointing (ldm_diffusion_util.py:142-171).  Here the backward is written out by hand on the same
channels-last bf16 activations as the forward kernels:

  * data gradients of convs / linears  = the tcgen05 implicit GEMM with flipped / transposed weights (cs_conv3d),
  * weight gradients                   = the tcgen05 weight-gradient kernel (cs_conv3d_wgrad), fp32,
  * GroupNorm+SiLU, LayerNorm, GEGLU, attention, upsample gradients = their own kernels (cs_*_bwd),
  * per-sample vectors (time embedding, emb_layers, single-token cross-attention) = small fp32 GEMMs (cs_sgemm_small).

No activation checkpointing (180 GB of HBM holds every activation of a batch-32 step several times over), so the step
costs 3x a forward instead of the reference's 4x.  Gradients land in fp32 tensors with the parameters' own shapes
(`GradSink`), so either torch.optim.AdamW or cs_adamw can apply them.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .... import _lib, ops, ops_bwd
from .attention import SpatialTransformer3D
from .ldm_diffusion_util import timestep_embedding
from .openai_model_3d import Act, AttentionBlock, Downsample, QKVAttention, ResBlock, StatArena, Upsample, _f, _with_stats


class GradSink:
    """fp32 gradient tensors keyed by parameter (created zeroed on first use, or views into a caller-provided flat buffer)."""

    def __init__(self, views: Optional[Dict[nn.Parameter, torch.Tensor]] = None,
                 packed: Optional[Dict[nn.Parameter, torch.Tensor]] = None):
        """packed: parameters whose gradient is kept in the weight-gradient kernel's PACKED layout (Cout, taps, padded Cin)
        -- slots of a flat buffer that the fused optimizer (cs_adamw_repack) reads directly; they have no parameter-layout
        gradient tensor at all."""
        self.views = views
        self.packed = packed or {}
        self.grads: Dict[nn.Parameter, torch.Tensor] = {}

    def packed_slot(self, p: Optional[nn.Parameter]) -> Optional[torch.Tensor]:
        return None if p is None else self.packed.get(p)

    def grad(self, p: nn.Parameter) -> torch.Tensor:
        g = self.grads.get(p)
        if g is None:
            if p in self.packed:
                raise KeyError("this parameter's gradient lives in a packed slot (GradSink.packed_slot); use ops_bwd.unpack_wgrad")
            g = self.views[p] if self.views is not None else torch.zeros(p.shape, dtype=torch.float32, device=p.device)
            self.grads[p] = g
        return g


class _TrainGrad(Act):
    """An Act that also carries the gradient flowing back into it."""
    __slots__ = ("grad",)

    def __init__(self, t, stat):
        super().__init__(t, stat)
        self.grad = None


def _acc(a: _TrainGrad, g: torch.Tensor) -> None:
    if a.grad is None:
        a.grad = g
    else:
        ops_bwd.add_(a.grad, g)


def _tg(act: Act) -> _TrainGrad:
    return _TrainGrad(act.t, act.stat)


class UNetTrainer:
    """forward_train() / backward() for one UNet3DModel.  Weight packs (forward + data-gradient layouts) are rebuilt when a
    parameter's version counter changes, like the inference packs."""

    def __init__(self, unet):
        self.unet = unet
        self._key = None
        self._dpk = None
        self._scratch = None
        self._stat_ws = {}
        self._zero_arena = None

    # ------------------------------------------------------------------------------------------
    # packs
    # ------------------------------------------------------------------------------------------
    def _ensure(self):
        u = self.unet
        pk = u._ensure_packed()
        if self._dpk is not None and self._key == u._pack_generation:
            return pk, self._dpk
        d: Dict[nn.Parameter, torch.Tensor] = {}

        def dg(p):
            d[p] = ops_bwd.pack_dgrad_weight(p if p.dim() == 5 else p.reshape(p.shape[0], p.shape[1], 1, 1, 1), owner=p)

        biggest = 0
        for block in u._blocks():
            for layer in block:
                if isinstance(layer, ResBlock):
                    convs = [layer.in_layers[2].weight, layer.out_layers[3].weight]
                    if not isinstance(layer.skip_connection, nn.Identity):
                        convs.append(layer.skip_connection.weight)
                elif isinstance(layer, SpatialTransformer3D):
                    convs = [layer.proj_in.weight, layer.proj_out.weight]
                    for tb in layer.transformer_blocks:
                        convs += [tb.attn1.to_out[0].weight, tb.ff.net[0].proj.weight, tb.ff.net[2].weight]
                        a = tb.attn1
                        h, dd, cin = a.heads, a.dim_head, a.to_q.weight.shape[1]
                        dp = _pad_head(dd)
                        w = torch.zeros(3, h, dp, cin, dtype=torch.float32, device=a.to_q.weight.device)
                        for i, lin in enumerate((a.to_q, a.to_k, a.to_v)):
                            w[i, :, :dd] = lin.weight.detach().float().reshape(h, dd, cin)
                        d[a.to_q.weight] = ops_bwd.pack_dgrad_weight(w.reshape(3 * h * dp, cin, 1, 1, 1))
                        biggest = max(biggest, 3 * h * dp * ops._pad64(cin))
                elif isinstance(layer, AttentionBlock):     # concat-conditioning variant: qkv in the head-padded q|k|v layout
                    C, hh = layer.channels, layer.num_heads
                    ch = C // hh
                    dp = _pad_head(ch)
                    w = torch.zeros(3, hh, dp, C, dtype=torch.float32, device=layer.qkv.weight.device)
                    w[:, :, :ch] = _qkv_rows_to_packed(layer, layer.qkv.weight.detach().float().reshape(3 * C, C))
                    d[layer.qkv.weight] = ops_bwd.pack_dgrad_weight(w.reshape(3 * hh * dp, C, 1, 1, 1))
                    pw = layer.proj_out.weight
                    d[pw] = ops_bwd.pack_dgrad_weight(pw.reshape(C, C, 1, 1, 1), owner=pw)
                    biggest = max(biggest, 3 * hh * dp * ops._pad64(C), C * (ops._pad64(C) + 64))
                    convs = []
                elif isinstance(layer, Downsample):
                    convs = [layer.op.weight]
                elif isinstance(layer, Upsample):
                    convs = [layer.conv.weight]
                else:
                    convs = []
                for w in convs:
                    dg(w)
                    taps = w[0, 0].numel() if w.dim() == 5 else 1
                    biggest = max(biggest, w.shape[0] * taps * (ops._pad64(w.shape[1]) + 64))
        # out head: d a[v][ci] = sum_{tap', co} col(d_eps)[v][tap' * Co + co] * W[co][ci][26 - tap']
        w = u.out[2].weight.detach().float()
        co, ci = w.shape[0], w.shape[1]
        kp = (27 * co + 15) // 16 * 16
        wd = torch.zeros(ci, 1, kp, dtype=torch.float32, device=w.device)
        wd[:, 0, :27 * co] = w.reshape(co, ci, 27).flip(2).permute(1, 2, 0).reshape(ci, 27 * co)
        d[u.out[2].weight] = ops._pad_k(wd)
        biggest = max(biggest, 128 * ops._pad64(ci), u.model_channels * 128)
        # stem data gradient (needed when the conditioning is concatenated to x): a 3x3x3 conv model_channels -> in_channels
        # with the transposed, flipped filter, through the few-output-channel path (one GEMM over the taps + cs_tap_gather)
        stem = u.input_blocks[0][0]
        if stem.weight.shape[1] <= 4:
            d["stem_dx"] = ops.pack_small_cout_conv(stem.weight.detach().float().transpose(0, 1).flip(2, 3, 4).contiguous())
        if self._scratch is None or self._scratch.numel() < biggest:
            self._scratch = torch.empty(biggest, dtype=torch.float32, device=w.device)
        self._dpk, self._key = d, u._pack_generation
        return pk, d

    # ------------------------------------------------------------------------------------------
    # gradient helpers
    # ------------------------------------------------------------------------------------------
    def _wgrad(self, sink, param, x, dy, *, ksize=(3, 3, 3), stride=(1, 1, 1), pad=(1, 1, 1), x2=None):
        c1 = x.shape[-1]
        c2 = 0 if x2 is None else x2.shape[-1]
        co, taps = dy.shape[-1], ksize[0] * ksize[1] * ksize[2]
        cp = ops._pad64(c1) + ops._pad64(c2)
        slot = sink.packed_slot(param)
        if slot is not None:
            # fused-optimizer mode: accumulate straight into the parameter's packed slot (zeroed by the previous update)
            if tuple(slot.shape) != (co, taps, cp):
                raise RuntimeError(f"packed gradient slot {tuple(slot.shape)} does not match the layer ({co}, {taps}, {cp})")
            ops_bwd.conv3d_wgrad(x, dy, slot, ksize=ksize, stride=stride, pad=pad, x2=x2)
            return None
        if param is not None and taps == 1 and x2 is None and cp == c1:
            # a linear layer without channel padding: the packed gradient layout IS the parameter's (out, in) layout
            ops_bwd.conv3d_wgrad(x, dy, sink.grad(param).view(co, 1, c1), ksize=ksize, stride=stride, pad=pad)
            return None
        dw = self._scratch[:co * taps * cp].view(co, taps, cp)
        dw.zero_()
        ops_bwd.conv3d_wgrad(x, dy, dw, ksize=ksize, stride=stride, pad=pad, x2=x2)
        if param is None:
            return dw
        ops_bwd.unpack_wgrad_into(dw, sink.grad(param), (c1, c2) if x2 is not None else None)
        return None

    def _lin_wgrad(self, sink, param, x, dy, x2=None):
        return self._wgrad(sink, param, x, dy, ksize=(1, 1, 1), pad=(0, 0, 0), x2=x2)

    def _colsums(self, dy):
        """fp32 (B, C, 2) per-sample channel sums of a bf16 channels-last tensor (component 0 = sum)."""
        B, C = dy.shape[0], dy.shape[-1]
        return ops_bwd.channel_sums(dy, ops_bwd.zero_f32((B, C, 2), dy.device))

    def _bias_grad(self, sink, param, dy, sums=None):
        if sums is None and dy.shape[-1] > 2048:       # cs_channel_sums handles <= 2048 channels per call
            g = sink.grad(param)
            for c0 in range(0, dy.shape[-1], 2048):
                c1 = min(dy.shape[-1], c0 + 2048)
                ops_bwd.batch_reduce(self._colsums(dy[..., c0:c1]), 0, g[c0:c1])
            return None
        sums = self._colsums(dy) if sums is None else sums
        ops_bwd.batch_reduce(sums, 0, sink.grad(param))
        return sums

    # ------------------------------------------------------------------------------------------
    # forward with saved activations
    # ------------------------------------------------------------------------------------------
    def forward_train(self, x: torch.Tensor, timesteps: torch.Tensor, context: torch.Tensor):
        """x (B, C, D, H, W) fp32, timesteps (B,) int64, context (B, 1, context_dim) fp32 -> (eps fp32 NCDHW, tape)."""
        _lib.require_device()
        u = self.unet
        pk, _ = self._ensure()
        has_ctx = pk["ca_w"] is not None          # False for the concat-conditioning variant (no cross-attention layers)
        if has_ctx and (context is None or context.dim() != 3 or context.shape[1] != 1):
            raise NotImplementedError("the training path covers the single-token conditioning of v2_full (VAEGAN_V2FULL.py:237-240)")
        B = timesteps.shape[0]
        x = x.float().contiguous()
        tape: dict = {"B": B}
        # --- per-sample vectors -------------------------------------------------------------
        t_emb = timestep_embedding(timesteps, u.model_channels)
        w0, b0, w2, b2 = pk["te"]
        h1 = ops.linear_small(t_emb, w0, b0)
        emb = ops.linear_small(h1, w2, b2, act_in=ops.ACT_SILU)
        emb_vecs = ops.linear_small(emb, pk["emb_w"], pk["emb_b"], act_in=ops.ACT_SILU)
        ctx = context[:, 0].float().contiguous() if has_ctx else None
        tape.update(t_emb=t_emb, h1=h1, emb=emb, ctx=ctx)
        ca = []      # per transformer block: (attn2 module, v2, vec)
        arena = StatArena(x.device, B * pk["stat_channels"] * 2)
        tape["arena"] = arena
        records: List[dict] = []

        def run_block(block, entries, h, skip=None):
            for layer, e in zip(block, entries):
                if e["kind"] == "res":
                    off, n = e["emb"]
                    h = self._res_fwd(layer, e["pk"], h, emb_vecs[:, off:off + n], arena, skip, records, (off, n))
                    skip = None
                elif e["kind"] == "st":
                    h = self._st_fwd(layer, e["pk"], h, ctx, arena, records, ca)
                elif e["kind"] == "attn":
                    h = self._attn_fwd(layer, e["pk"], h, arena, records)
                elif e["kind"] == "resample":
                    h = self._resample_fwd(layer, e["pk"], h, arena, records)
                else:
                    col = ops.im2col_small(h, batch=B, kp=e["pk"]["kp"])
                    S = col.shape[1] * col.shape[2] * col.shape[3]
                    out = _tg(_with_stats(arena, lambda st: ops.linear_tokens(col, e["pk"]["w"], bias=e["pk"]["b"], stat_sum=st), B,
                                          layer.out_channels, S))
                    records.append({"kind": "stem", "layer": layer, "col": col, "out": out})
                    h = out
            return h

        entries = pk["blocks"]
        n_in = len(u.input_blocks)
        hs: List[_TrainGrad] = []
        h = x
        for i, block in enumerate(u.input_blocks):
            h = run_block(block, entries[i], h)
            hs.append(h)
        h = run_block(u.middle_block, entries[n_in], h)
        for i, block in enumerate(u.output_blocks):
            h = run_block(block, entries[n_in + 1 + i], h, skip=hs.pop())
        a = ops.groupnorm_fused(h.t, h.stat, *pk["out_gn"], eps=u.out[0].eps, act=ops.ACT_SILU)
        eps = ops.conv3d_small_cout(a, pk["out_w"], pk["out_b"], u.out_channels)
        tape.update(records=records, head_in=h, head_a=a, ca=ca, emb_vecs_shape=emb_vecs.shape)
        return eps, tape

    def _res_fwd(self, layer, pk, x, emb_vec, arena, skip, records, emb_slot):
        B, D, H, W, _ = x.t.shape
        S, C = D * H * W, layer.out_channels
        a1 = ops.groupnorm_fused(x.t, x.stat, *pk["gn1"], eps=layer.in_layers[0].eps, act=ops.ACT_SILU,
                                 x2=None if skip is None else skip.t, stat2=None if skip is None else skip.stat)
        h = _with_stats(arena, lambda st: ops.conv3d(a1, pk["w1"], bias=pk["b1"], rowvec=emb_vec, stat_sum=st), B, C, S)
        a2 = ops.groupnorm_fused(h.t, h.stat, *pk["gn2"], eps=layer.out_layers[0].eps, act=ops.ACT_SILU)
        if "ws" in pk:
            res = ops.linear_tokens(x.t, pk["ws"], bias=pk["bs"], x2=None if skip is None else skip.t)
        else:
            res = x.t
        out = _tg(_with_stats(arena, lambda st: ops.conv3d(a2, pk["w2"], bias=pk["b2"], residual=res, stat_sum=st), B, C, S))
        records.append({"kind": "res", "layer": layer, "pk": pk, "x": x, "skip": skip, "a1": a1, "h": h, "a2": a2, "out": out,
                        "emb": emb_slot})
        return out

    def _resample_fwd(self, layer, pk, x, arena, records):
        B = x.t.shape[0]
        if isinstance(layer, Upsample):
            f = (1, 2, 2) if layer.dims == 3 else (2, 2, 2)
            up = ops.upsample_nearest(x.t, f)
            _, D, H, W, _ = up.shape
            out = _tg(_with_stats(arena, lambda st: ops.conv3d(up, pk["w"], bias=pk["b"], stat_sum=st), B, layer.out_channels,
                                  D * H * W))
            records.append({"kind": "up", "layer": layer, "x": x, "u": up, "f": f, "out": out})
        else:
            _, D, H, W, _ = x.t.shape
            s = layer.stride
            So = (D // s[0]) * (H // s[1]) * (W // s[2])
            out = _tg(_with_stats(arena, lambda st: ops.conv3d(x.t, pk["w"], stride=s, bias=pk["b"], stat_sum=st), B,
                                  layer.out_channels, So))
            records.append({"kind": "down", "layer": layer, "x": x, "out": out})
        return out

    def _st_fwd(self, layer, pk, x, ctx, arena, records, ca):
        B, D, H, W, C = x.t.shape
        g = ops.groupnorm_fused(x.t, x.stat, *pk["gn"], eps=layer.norm.eps)
        t = ops.linear_tokens(g, pk["w_in"], bias=pk["b_in"])
        blocks = []
        for blk, bpk in zip(layer.transformer_blocks, pk["blocks"]):
            a2 = blk.attn2
            v2 = ops.linear_small(ctx, _f(a2.to_v.weight))
            vec = ops.linear_small(v2, _f(a2.to_out[0].weight), _f(a2.to_out[0].bias))
            apk = bpk["attn1"]
            hh, dd, dp = blk.attn1.heads, blk.attn1.dim_head, apk["dp"]
            l1 = ops.layernorm(t, *bpk["ln1"], eps=blk.norm1.eps)
            qkv = ops.linear_tokens(l1, apk["wqkv"]).view(B, D * H * W, 3 * hh * dp)
            q, k, v = (qkv[:, :, i * hh * dp:(i + 1) * hh * dp] for i in range(3))
            o, lse = ops_bwd.attention_lse(q, k, v, heads=hh, head_dim=dd, head_dim_padded=dp, scale=blk.attn1.scale)
            o5 = o.view(B, D, H, W, hh * dd)
            t1 = ops.linear_tokens(o5, apk["wo"], bias=apk["bo"], rowvec=vec, residual=t)
            l3 = ops.layernorm(t1, *bpk["ln3"], eps=blk.norm3.eps)
            fpk = bpk["ff"]
            if fpk["fused"]:
                raise NotImplementedError("FeedForward.FUSE_GEGLU is an inference-only option")
            uu = ops.linear_tokens(l3, fpk["w1"], bias=fpk["b1"])
            f = ops.geglu(uu)
            t2 = ops.linear_tokens(f, fpk["w2"], bias=fpk["b2"], residual=t1)
            blocks.append({"blk": blk, "pk": bpk, "t0": t, "l1": l1, "qkv": qkv, "o": o5, "lse": lse, "t1": t1, "l3": l3, "u": uu,
                           "f": f, "v2": v2})
            t = t2
        out = _tg(_with_stats(arena, lambda st: ops.linear_tokens(t, pk["w_out"], bias=pk["b_out"], residual=x.t, stat_sum=st),
                              B, C, D * H * W))
        records.append({"kind": "st", "layer": layer, "pk": pk, "x": x, "g": g, "blocks": blocks, "t_last": t, "out": out})
        return out

    def _attn_fwd(self, layer, pk, x, arena, records):
        """AttentionBlock (concat variant): x + proj_out(attention(qkv(GroupNorm(x)))), keeping what the backward needs."""
        B, D, H, W, C = x.t.shape
        hh, dp = layer.num_heads, pk["dp"]
        ch = C // hh
        g = ops.groupnorm_fused(x.t, x.stat, *pk["gn"], eps=layer.norm.eps)
        qkv = ops.linear_tokens(g, pk["wqkv"], bias=pk["bqkv"]).view(B, D * H * W, 3 * hh * dp)
        q, k, v = (qkv[:, :, i * hh * dp:(i + 1) * hh * dp] for i in range(3))
        o, lse = ops_bwd.attention_lse(q, k, v, heads=hh, head_dim=ch, head_dim_padded=dp, scale=ch ** -0.5)
        o5 = o.view(B, D, H, W, C)
        out = _tg(_with_stats(arena, lambda st: ops.linear_tokens(o5, pk["wo"], bias=pk["bo"], residual=x.t, stat_sum=st),
                              B, C, D * H * W))
        records.append({"kind": "attn", "layer": layer, "pk": pk, "x": x, "g": g, "qkv": qkv, "o": o5, "lse": lse, "out": out})
        return out

    # ------------------------------------------------------------------------------------------
    # backward
    # ------------------------------------------------------------------------------------------
    def backward(self, tape: dict, d_eps: torch.Tensor, sink: Optional[GradSink] = None, need_dcontext: bool = True,
                 on_block_done=None, need_dx: bool = False):
        """See _backward_impl.  Runs it inside the trainer's ZeroArena: the small zero-initialised reduction buffers of the
        whole backward come from ONE memset.  need_dx: also compute the gradient wrt the network input (fp32 NCDHW, left in
        tape["dx"]) -- the path of the conditioning gradient when it is concatenated to x (concat variant)."""
        if self._zero_arena is None or self._zero_arena.buf.device != d_eps.device:
            self._zero_arena = ops_bwd.ZeroArena(d_eps.device)
        # the tiny per-parameter reductions (d beta / d gamma, bias gradients) are queued and issued 24 per launch at the end of
        # every block (before on_block_done, whose caller may send those gradients off) and at the end of the backward
        with self._zero_arena, ops_bwd.ReduceQueue():
            return self._backward_impl(tape, d_eps, sink, need_dcontext, on_block_done, need_dx)

    def _backward_impl(self, tape: dict, d_eps: torch.Tensor, sink: Optional[GradSink] = None, need_dcontext: bool = True,
                       on_block_done=None, need_dx: bool = False):
        """d_eps: gradient wrt the fp32 NCDHW output of forward_train.  Accumulates every parameter gradient into `sink`
        (created if None) and returns (sink, d_context (B, 1, context_dim) fp32 or None).

        Layers are visited in exactly the reverse of `unet.parameters()` order; `on_block_done(p)` is called after each
        layer with its first parameter p: the gradients of p and of every parameter after it are final at that point
        (what a data-parallel caller needs to start all-reducing them while the rest of the backward runs)."""
        u = self.unet
        pk, dpk = self._ensure()
        sink = sink or GradSink()
        B = tape["B"]
        d_eps = d_eps.float().contiguous()
        dev = d_eps.device
        d_emb_vecs = torch.zeros(tape["emb_vecs_shape"], dtype=torch.float32, device=dev)
        d_ctx = torch.zeros_like(tape["ctx"]) if tape["ctx"] is not None else None
        se = F.silu(tape["emb"])

        # --- out head: conv3x3x3 224 -> 3 (im2col of d_eps on both sides), GroupNorm + SiLU -----
        h, a = tape["head_in"], tape["head_a"]
        co = u.out_channels
        wh = u.out[2].weight
        sink.grad(u.out[2].bias).add_(d_eps.sum(dim=(0, 2, 3, 4)))
        col = ops.im2col_small(d_eps, kp=(27 * co + 15) // 16 * 16)
        da = ops.linear_tokens(col, dpk[wh])
        dwh = self._lin_wgrad(sink, None, a, col)                 # (kp, 1, pad64(Cin)): row tap'*Co + co
        ci = wh.shape[1]
        gh = dwh[:27 * co, 0, :ci].reshape(27, co, ci).flip(0).permute(1, 2, 0)
        sink.grad(wh).add_(gh.reshape(wh.shape))
        dh, _ = ops_bwd.groupnorm_bwd(h.t, h.stat, *pk["out_gn"], da, eps=u.out[0].eps, act=ops.ACT_SILU,
                                      dgamma=sink.grad(u.out[0].weight), dbeta=sink.grad(u.out[0].bias))
        _acc(h, dh)
        ops_bwd.flush_reductions()
        if on_block_done is not None:
            on_block_done(u.out[0].weight)

        for rec in reversed(tape["records"]):
            kind = rec["kind"]
            if kind == "res":
                self._res_bwd(rec, sink, dpk, d_emb_vecs, se)
            elif kind == "st":
                self._st_bwd(rec, sink, dpk, tape["ctx"], d_ctx)
            elif kind == "attn":
                self._attn_bwd(rec, sink, dpk)
            elif kind == "up":
                layer, out = rec["layer"], rec["out"]
                dy = out.grad
                self._wgrad(sink, layer.conv.weight, rec["u"], dy)
                self._bias_grad(sink, layer.conv.bias, dy)
                du = ops_bwd.conv3d_dgrad(dy, dpk[layer.conv.weight])
                _acc(rec["x"], ops_bwd.upsample_nearest_bwd(du, rec["f"]))
            elif kind == "down":
                layer, out = rec["layer"], rec["out"]
                dy = out.grad
                self._wgrad(sink, layer.op.weight, rec["x"].t, dy, stride=layer.stride)
                self._bias_grad(sink, layer.op.bias, dy)
                _acc(rec["x"], ops_bwd.conv3d_dgrad_strided(dy, dpk[layer.op.weight], layer.stride))
            else:   # stem: weight / bias gradients only (x_t needs none)
                conv, dy = rec["layer"], rec["out"].grad
                dws = self._lin_wgrad(sink, None, rec["col"], dy)
                cin = conv.weight.shape[1]
                g = dws[:, 0, :27 * cin].reshape(conv.weight.shape[0], 27, cin).permute(0, 2, 1)
                sink.grad(conv.weight).add_(g.reshape(conv.weight.shape))
                self._bias_grad(sink, conv.bias, dy)
                if need_dx:
                    if "stem_dx" not in dpk:
                        raise NotImplementedError("input gradient of a stem with more than 4 input channels")
                    tape["dx"] = ops.conv3d_small_cout(dy, dpk["stem_dx"], None, cin)
            rec["out"].grad = None
            ops_bwd.flush_reductions()
            if on_block_done is not None:
                on_block_done(next(rec["layer"].parameters()))

        # --- time_embed (the emb_layers weights were handled inside their ResBlocks) --------------
        emb, h1, t_emb = tape["emb"], tape["h1"], tape["t_emb"]
        s1 = F.silu(h1)
        d_emb = ops_bwd.sgemm(d_emb_vecs, pk["emb_w"], silu_pre=emb)                      # (B, 896), through SiLU(emb)
        te0, te2 = u.time_embed[0], u.time_embed[2]
        ops_bwd.sgemm(d_emb, s1, trans_a=True, out=sink.grad(te2.weight), accumulate=True)
        ops_bwd.batch_reduce(d_emb.view(B, -1, 1), 0, sink.grad(te2.bias))
        d_h1 = ops_bwd.sgemm(d_emb, _f(te2.weight), silu_pre=h1)
        ops_bwd.sgemm(d_h1, t_emb, trans_a=True, out=sink.grad(te0.weight), accumulate=True)
        ops_bwd.batch_reduce(d_h1.view(B, -1, 1), 0, sink.grad(te0.bias))
        return sink, (d_ctx[:, None, :] if need_dcontext and d_ctx is not None else None)

    def _res_bwd(self, rec, sink, dpk, d_emb_vecs, se):
        layer, pk, x, skip, out = rec["layer"], rec["pk"], rec["x"], rec["skip"], rec["out"]
        dy = out.grad
        B = dy.shape[0]
        conv1, conv2 = layer.in_layers[2], layer.out_layers[3]
        # out = conv2(a2) + b2 + res
        self._wgrad(sink, conv2.weight, rec["a2"], dy)
        sums = self._bias_grad(sink, conv2.bias, dy)
        da2 = ops_bwd.conv3d_dgrad(dy, dpk[conv2.weight])
        h = rec["h"]
        dh, _ = ops_bwd.groupnorm_bwd(h.t, h.stat, *pk["gn2"], da2, eps=layer.out_layers[0].eps, act=ops.ACT_SILU,
                                      dgamma=sink.grad(layer.out_layers[0].weight), dbeta=sink.grad(layer.out_layers[0].bias))
        # h = conv1(a1) + b1 + emb_vec[b]
        hs = self._bias_grad(sink, conv1.bias, dh)
        off, n = rec["emb"]
        dvec = hs[:, :, 0].contiguous()
        d_emb_vecs[:, off:off + n] = dvec
        lin = layer.emb_layers[1]                  # emb_vec = Linear(SiLU(emb))
        ops_bwd.sgemm(dvec, se, trans_a=True, out=sink.grad(lin.weight), accumulate=True)
        ops_bwd.batch_reduce(hs, 0, sink.grad(lin.bias))
        self._wgrad(sink, conv1.weight, rec["a1"], dh)
        da1 = ops_bwd.conv3d_dgrad(dh, dpk[conv1.weight])
        c1 = x.t.shape[-1]
        if "ws" in pk:
            sc = layer.skip_connection
            self._lin_wgrad(sink, sc.weight, x.t, dy, x2=None if skip is None else skip.t)
            ops_bwd.batch_reduce(sums, 0, sink.grad(sc.bias))
            dcat = ops_bwd.conv3d_dgrad(dy, dpk[sc.weight], ksize=(1, 1, 1), pad=(0, 0, 0))
            e1 = dcat[..., :c1]
            e2 = dcat[..., c1:] if skip is not None else None
        else:
            e1, e2 = dy, None
        dx, dskip = ops_bwd.groupnorm_bwd(x.t, x.stat, *pk["gn1"], da1, eps=layer.in_layers[0].eps, act=ops.ACT_SILU,
                                          x2=None if skip is None else skip.t, stat2=None if skip is None else skip.stat,
                                          extra=e1, extra2=e2, dgamma=sink.grad(layer.in_layers[0].weight),
                                          dbeta=sink.grad(layer.in_layers[0].bias))
        _acc(x, dx)
        if skip is not None:
            _acc(skip, dskip)

    def _chan_sums(self, dy):
        """fp32 (C,) sum over every token of a bf16 channels-last tensor (bias gradients), 2048 channels per launch."""
        Ct = dy.shape[-1]
        parts = [self._colsums(dy[..., c0:min(Ct, c0 + 2048)])[:, :, 0].sum(0) for c0 in range(0, Ct, 2048)]
        return parts[0] if len(parts) == 1 else torch.cat(parts)

    def _attn_bwd(self, rec, sink, dpk):
        layer, pk, x, out = rec["layer"], rec["pk"], rec["x"], rec["out"]
        dy = out.grad
        B, D, H, W, C = dy.shape
        hh, dp = layer.num_heads, pk["dp"]
        ch, N = C // hh, D * H * W
        # out = o Wo + bo + x
        self._lin_wgrad(sink, layer.proj_out.weight, rec["o"], dy)
        self._bias_grad(sink, layer.proj_out.bias, dy)
        do = ops_bwd.conv3d_dgrad(dy, dpk[layer.proj_out.weight], ksize=(1, 1, 1), pad=(0, 0, 0))
        dqkv = ops_bwd.attention_bwd(rec["qkv"], rec["o"].view(B, N, C), do.view(B, N, C), rec["lse"], heads=hh, head_dim=ch,
                                     head_dim_padded=dp, scale=ch ** -0.5)
        dqkv5 = dqkv.view(B, D, H, W, 3 * hh * dp)
        # qkv = g Wqkv + bqkv in the head-padded q|k|v layout: map the packed gradients back to the parameter's row order
        dwq = self._lin_wgrad(sink, None, rec["g"], dqkv5)
        gw = dwq[:, 0, :C].view(3, hh, dp, C)[:, :, :ch]
        sink.grad(layer.qkv.weight).add_(_qkv_packed_to_rows(layer, gw).reshape(layer.qkv.weight.shape))
        gb = self._chan_sums(dqkv5).view(3, hh, dp)[:, :, :ch]
        sink.grad(layer.qkv.bias).add_(_qkv_packed_to_rows(layer, gb))
        dg = ops_bwd.conv3d_dgrad(dqkv5, dpk[layer.qkv.weight], ksize=(1, 1, 1), pad=(0, 0, 0))
        dx, _ = ops_bwd.groupnorm_bwd(x.t, x.stat, *pk["gn"], dg, eps=layer.norm.eps, act=ops.ACT_NONE, extra=dy,
                                      dgamma=sink.grad(layer.norm.weight), dbeta=sink.grad(layer.norm.bias))
        _acc(x, dx)

    def _st_bwd(self, rec, sink, dpk, ctx, d_ctx):
        layer, pk, x, out = rec["layer"], rec["pk"], rec["x"], rec["out"]
        dy = out.grad
        B, D, H, W, C = dy.shape
        self._lin_wgrad(sink, layer.proj_out.weight, rec["t_last"], dy)
        self._bias_grad(sink, layer.proj_out.bias, dy)
        dt = ops_bwd.conv3d_dgrad(dy, dpk[layer.proj_out.weight], ksize=(1, 1, 1), pad=(0, 0, 0))
        for b in reversed(rec["blocks"]):
            blk, bpk = b["blk"], b["pk"]
            ff, a1, a2 = blk.ff, blk.attn1, blk.attn2
            # t2 = f W2 + b2 + t1
            self._lin_wgrad(sink, ff.net[2].weight, b["f"], dt)
            self._bias_grad(sink, ff.net[2].bias, dt)
            df = ops_bwd.conv3d_dgrad(dt, dpk[ff.net[2].weight], ksize=(1, 1, 1), pad=(0, 0, 0))
            du = ops_bwd.geglu_bwd(b["u"], df)
            self._lin_wgrad(sink, ff.net[0].proj.weight, b["l3"], du)
            self._bias_grad(sink, ff.net[0].proj.bias, du)
            dl3 = ops_bwd.conv3d_dgrad(du, dpk[ff.net[0].proj.weight], ksize=(1, 1, 1), pad=(0, 0, 0))
            dt1 = ops_bwd.layernorm_bwd(b["t1"], bpk["ln3"][0], dl3, sink.grad(blk.norm3.weight), sink.grad(blk.norm3.bias),
                                        eps=blk.norm3.eps, extra=dt)
            # t1 = o Wo + bo + vec[b] + t0,  vec = to_out2(to_v2(ctx)) + b_out2
            sums = self._bias_grad(sink, a1.to_out[0].bias, dt1)
            dvec = sums[:, :, 0].contiguous()
            ops_bwd.batch_reduce(sums, 0, sink.grad(a2.to_out[0].bias))
            ops_bwd.sgemm(dvec, b["v2"], trans_a=True, out=sink.grad(a2.to_out[0].weight), accumulate=True)
            dv2 = ops_bwd.sgemm(dvec, _f(a2.to_out[0].weight))
            ops_bwd.sgemm(dv2, ctx, trans_a=True, out=sink.grad(a2.to_v.weight), accumulate=True)
            ops_bwd.sgemm(dv2, _f(a2.to_v.weight), out=d_ctx, accumulate=True)
            self._lin_wgrad(sink, a1.to_out[0].weight, b["o"], dt1)
            do = ops_bwd.conv3d_dgrad(dt1, dpk[a1.to_out[0].weight], ksize=(1, 1, 1), pad=(0, 0, 0))
            hh, dd, dp = a1.heads, a1.dim_head, bpk["attn1"]["dp"]
            N = D * H * W
            dqkv = ops_bwd.attention_bwd(b["qkv"], b["o"].view(B, N, hh * dd), do.view(B, N, hh * dd), b["lse"], heads=hh,
                                         head_dim=dd, head_dim_padded=dp, scale=a1.scale)
            dqkv5 = dqkv.view(B, D, H, W, 3 * hh * dp)
            dwq = self._lin_wgrad(sink, None, b["l1"], dqkv5)
            cin = a1.to_q.weight.shape[1]
            g3 = dwq[:, 0, :cin].view(3, hh, dp, cin)[:, :, :dd].reshape(3, hh * dd, cin)
            for i, lin in enumerate((a1.to_q, a1.to_k, a1.to_v)):
                sink.grad(lin.weight).add_(g3[i])
            dl1 = ops_bwd.conv3d_dgrad(dqkv5, dpk[a1.to_q.weight], ksize=(1, 1, 1), pad=(0, 0, 0))
            dt = ops_bwd.layernorm_bwd(b["t0"], bpk["ln1"][0], dl1, sink.grad(blk.norm1.weight), sink.grad(blk.norm1.bias),
                                       eps=blk.norm1.eps, extra=dt1)
        # t0 = proj_in(GN(x)); out = proj_out(t_last) + x
        self._lin_wgrad(sink, layer.proj_in.weight, rec["g"], dt)
        self._bias_grad(sink, layer.proj_in.bias, dt)
        dg = ops_bwd.conv3d_dgrad(dt, dpk[layer.proj_in.weight], ksize=(1, 1, 1), pad=(0, 0, 0))
        dx, _ = ops_bwd.groupnorm_bwd(x.t, x.stat, *pk["gn"], dg, eps=layer.norm.eps, act=ops.ACT_NONE, extra=dy,
                                      dgamma=sink.grad(layer.norm.weight), dbeta=sink.grad(layer.norm.bias))
        _acc(x, dx)


def _qkv_rows_to_packed(layer, rows: torch.Tensor) -> torch.Tensor:
    """AttentionBlock.qkv rows (3C, ...) -> (3, heads, ch, ...): legacy order is [head][q|k|v][ch], new order [q|k|v][head][ch]."""
    C, hh = layer.channels, layer.num_heads
    ch = C // hh
    tail = rows.shape[1:]
    if isinstance(layer.attention, QKVAttention):
        return rows.reshape(3, hh, ch, *tail)
    return rows.reshape(hh, 3, ch, *tail).transpose(0, 1)


def _qkv_packed_to_rows(layer, packed: torch.Tensor) -> torch.Tensor:
    """Inverse of _qkv_rows_to_packed: (3, heads, ch, ...) -> (3C, ...) in the parameter's own row order."""
    C = layer.channels
    tail = packed.shape[3:]
    if isinstance(layer.attention, QKVAttention):
        return packed.reshape(3 * C, *tail)
    return packed.transpose(0, 1).reshape(3 * C, *tail)


def _pad_head(d: int) -> int:
    from .attention import _pad_head_dim
    return _pad_head_dim(d)
