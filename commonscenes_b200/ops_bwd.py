"""Backward (training-path) operators on libcsb200.so.

The reference trains the denoiser with autograd (`loss.backward()`, sdfusion_txt2shape_model.py:568-575;
train_3dfront.py:387-401).  These wrappers are the explicit gradients of the forward kernels in ops.py, on the same
channels-last bf16 activations (fp32 accumulation, fp32 parameter gradients).  No CPU / PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import WgradArgs, check
from .ops import _check_act, _pad64, _pad_k, _ptr, _stream, conv3d


# ----------------------------------------------------------------------------------------------
# zero-initialised fp32 scratch (reduction targets of the backward kernels)
# ----------------------------------------------------------------------------------------------
class ZeroArena:
    """One fp32 buffer zeroed by ONE memset per training step, from which the many small accumulation targets of the backward
    (GroupNorm reduction pairs, per-sample channel sums for bias gradients) are carved -- instead of ~330 tiny fill launches."""
    current: Optional["ZeroArena"] = None

    def __init__(self, device, numel: int = 24 << 20):
        self.buf = torch.zeros(numel, dtype=torch.float32, device=device)
        self.off = 0

    def reset(self):
        self.buf.zero_()
        self.off = 0

    def take(self, shape):
        n = 1
        for d in shape:
            n *= int(d)
        n4 = (n + 3) // 4 * 4
        if self.off + n4 > self.buf.numel():
            return None
        v = self.buf[self.off:self.off + n].view(shape)
        self.off += n4
        return v

    def __enter__(self):
        self.reset()
        ZeroArena.current = self
        return self

    def __exit__(self, *exc):
        ZeroArena.current = None


def zero_f32(shape, device) -> torch.Tensor:
    """A zeroed fp32 tensor: a slice of the active ZeroArena when there is one (and it has room), else torch.zeros."""
    a = ZeroArena.current
    if a is not None and a.buf.device == torch.device(device):
        v = a.take(shape)
        if v is not None:
            return v
    return torch.zeros(shape, dtype=torch.float32, device=device)


# ----------------------------------------------------------------------------------------------
# GEMM-class gradients
# ----------------------------------------------------------------------------------------------
def pack_dgrad_weight(w: torch.Tensor, owner=None) -> torch.Tensor:
    """(Cout, Cin, kd, kh, kw) fp32 -> packed bf16 (Cin, taps, pad64(Cout)) with the filter flipped: the weight with which
    cs_conv3d maps dY (Cout channels) to dX (Cin channels) for a stride-1 convolution."""
    co, ci = w.shape[0], w.shape[1]
    wd = w.detach()
    if wd.is_cuda and wd.dtype == torch.float32 and wd.is_contiguous():
        taps = wd.numel() // (co * ci)
        from .ops import _pack_buffer, _pack_key, _pack_is_maintained
        own, shape = owner if owner is not None else w, (ci, taps, _pad64(co))
        out = _pack_buffer("dgrad", own, shape, w.device)
        if _pack_is_maintained(_pack_key("dgrad", own, shape), own):
            return out
        check(_lib.load().cs_pack_weight(wd.data_ptr(), co, ci, taps, ci, None, out.data_ptr(), _stream()), "cs_pack_weight")
        return out
    wt = wd.reshape(co, ci, -1).flip(2).permute(1, 2, 0)     # (Cin, taps flipped, Cout)
    return _pad_k(wt)


def conv3d_dgrad(dy: torch.Tensor, w_dgrad: torch.Tensor, *, ksize: Sequence[int] = (3, 3, 3),
                 pad: Sequence[int] = (1, 1, 1), **kw) -> torch.Tensor:
    """dX of a stride-1 conv3d: conv3d(dY, flipped weight) with padding k - 1 - pad.  Keyword arguments (residual=,
    out=, stat_sum= ...) are those of ops.conv3d: the epilogue can add another gradient stream for free."""
    bpad = tuple(k - 1 - p for k, p in zip(ksize, pad))
    return conv3d(dy, w_dgrad, ksize=ksize, pad=bpad, **kw)


def conv3d_wgrad(x: torch.Tensor, dy: torch.Tensor, dw: torch.Tensor, *, ksize: Sequence[int] = (3, 3, 3),
                 stride: Sequence[int] = (1, 1, 1), pad: Sequence[int] = (1, 1, 1),
                 pad_back: Optional[Sequence[int]] = None, x2: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dw (fp32, packed (Cout, taps, pad64(C1) + pad64(C2))) += sum_v dy[v] (x) cat(x, x2)[shift_tap(v)]."""
    lib = _lib.load()
    B, D, H, W, C1, p1 = _check_act(x, "conv3d_wgrad.x")
    C2, p2 = 0, 0
    if x2 is not None:
        B2, D2, H2, W2, C2, p2 = _check_act(x2, "conv3d_wgrad.x2")
        if (B2, D2, H2, W2) != (B, D, H, W):
            raise _lib.CsError("conv3d_wgrad: x and x2 must share batch and spatial dims")
    Bo, Do, Ho, Wo, Cout, pdy = _check_act(dy, "conv3d_wgrad.dy")
    kd, kh, kw_ = ksize
    pb = tuple(pad) if pad_back is None else tuple(pad_back)
    want = ((D + pad[0] + pb[0] - kd) // stride[0] + 1, (H + pad[1] + pb[1] - kh) // stride[1] + 1,
            (W + pad[2] + pb[2] - kw_) // stride[2] + 1)
    if Bo != B or (Do, Ho, Wo) != want:
        raise _lib.CsError(f"conv3d_wgrad: dy grid {(Bo, Do, Ho, Wo)} does not match the conv output {(B,) + want}")
    if dw.dtype != torch.float32 or not dw.is_contiguous() or tuple(dw.shape) != (Cout, kd * kh * kw_, _pad64(C1) + _pad64(C2)):
        raise _lib.CsError(f"conv3d_wgrad: dw must be contiguous fp32 ({Cout}, {kd * kh * kw_}, {_pad64(C1) + _pad64(C2)})")
    a = WgradArgs()
    a.x1, a.C1, a.x1_pitch = x.data_ptr(), C1, p1
    a.x2, a.C2, a.x2_pitch = _ptr(x2), C2, p2
    a.B, a.D, a.H, a.W = B, D, H, W
    a.dy, a.Cout, a.dy_pitch = dy.data_ptr(), Cout, pdy
    a.kd, a.kh, a.kw = kd, kh, kw_
    a.sd, a.sh, a.sw = stride
    a.pd, a.ph, a.pw = pad
    a.pd_back, a.ph_back, a.pw_back = pb
    a.dw = dw.data_ptr()
    check(lib.cs_conv3d_wgrad(C.byref(a), _stream()), "cs_conv3d_wgrad")
    return dw


def unpack_wgrad(dw: torch.Tensor, shape: Sequence[int], split=None) -> torch.Tensor:
    """packed fp32 (Cout, taps, pad64(C1) [+ pad64(C2)]) -> the parameter's own layout (Cout, Cin, kd, kh, kw) / (out, in)."""
    co = shape[0]
    ci = shape[1]
    parts = [ci] if split is None else [int(c) for c in split]
    cols, off = [], 0
    for c in parts:
        cols.append(dw[:, :, off:off + c])
        off += _pad64(c)
    g = cols[0] if len(cols) == 1 else torch.cat(cols, dim=2)
    return g.permute(0, 2, 1).reshape(shape)


def unpack_wgrad_into(dw: torch.Tensor, grad: torch.Tensor, split=None) -> torch.Tensor:
    """grad (the parameter's layout: (Cout, Cin, kd, kh, kw) or (out, in), fp32 contiguous) += packed dw."""
    co, taps, _ = dw.shape
    ci = grad.shape[1]
    c1, c2 = (ci, 0) if split is None else (int(split[0]), int(split[1]))
    if grad.dtype != torch.float32 or not grad.is_contiguous() or grad.numel() != co * ci * taps or c1 + c2 != ci:
        raise _lib.CsError("unpack_wgrad_into: grad must be contiguous fp32 of the parameter's shape")
    check(_lib.load().cs_unpack_wgrad(dw.data_ptr(), co, taps, c1, c2, grad.data_ptr(), _stream()), "cs_unpack_wgrad")
    return grad


# ----------------------------------------------------------------------------------------------
# normalisation / pointwise gradients
# ----------------------------------------------------------------------------------------------
def groupnorm_bwd(x: torch.Tensor, stat: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, dy: torch.Tensor, *,
                  groups: int = 32, eps: float = 1e-5, act: int = 0, x2: Optional[torch.Tensor] = None,
                  stat2: Optional[torch.Tensor] = None, extra: Optional[torch.Tensor] = None,
                  extra2: Optional[torch.Tensor] = None, dgamma: Optional[torch.Tensor] = None,
                  dbeta: Optional[torch.Tensor] = None, need_dx: bool = True):
    """Backward of ops.groupnorm_fused: dy is the gradient of act(GN(cat(x, x2))) over all C1 + C2 channels.
    Returns (dx, dx2) (bf16, + extra / extra2 when given) and accumulates dgamma / dbeta (fp32 (C1 + C2,), +=)."""
    lib = _lib.load()
    B, D, H, W, C1, p1 = _check_act(x, "groupnorm_bwd.x")
    S = D * H * W
    C2, p2 = 0, 0
    if x2 is not None:
        _, _, _, _, C2, p2 = _check_act(x2, "groupnorm_bwd.x2")
    Ct = C1 + C2
    dB, dD, dH, dW, dC, pdy = _check_act(dy, "groupnorm_bwd.dy")
    if (dB, dD * dH * dW, dC) != (B, S, Ct):
        raise _lib.CsError("groupnorm_bwd: dy must cover all channels of the normalised tensor")
    st = _stream()
    red = zero_f32((B, Ct, 2), x.device)
    g, bta = _ptr(gamma), _ptr(beta)
    srcs = [(x, C1, p1, 0, extra)] + ([(x2, C2, p2, C1, extra2)] if x2 is not None else [])
    for t, c, p, off, _ in srcs:
        check(lib.cs_groupnorm_bwd(t.data_ptr(), B, S, c, p, off, dy.data_ptr(), pdy, off, stat.data_ptr(), C1, _ptr(stat2), C2,
                                   g, bta, groups, eps, act, red.data_ptr(), None, 0, None, 0, 0, st), "cs_groupnorm_bwd")
    outs = []
    if need_dx:
        for t, c, p, off, ex in srcs:
            dx = torch.empty((B, D, H, W, c), dtype=torch.bfloat16, device=x.device)
            check(lib.cs_groupnorm_bwd(t.data_ptr(), B, S, c, p, off, dy.data_ptr(), pdy, off, stat.data_ptr(), C1, _ptr(stat2),
                                       C2, g, bta, groups, eps, act, red.data_ptr(), _ptr(ex), 0 if ex is None else ex.stride(3),
                                       dx.data_ptr(), c, 1, st), "cs_groupnorm_bwd")
            outs.append(dx)
    if dbeta is not None:
        batch_reduce(red, 0, dbeta)
    if dgamma is not None:
        batch_reduce(red, 1, dgamma)
    if not need_dx:
        return None, None
    return outs[0], (outs[1] if len(outs) > 1 else None)


def channel_sums(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out (B, C, 2) fp32 += per-sample (sum, sum of squares) over the voxels of a bf16 channels-last tensor (bias gradients)."""
    B, D, H, W, Cc, p = _check_act(x, "channel_sums.x")
    if out.dtype != torch.float32 or not out.is_contiguous() or tuple(out.shape) != (B, Cc, 2):
        raise _lib.CsError("channel_sums: out must be contiguous fp32 (B, C, 2)")
    check(_lib.load().cs_channel_sums(x.data_ptr(), B, D * H * W, Cc, p, out.data_ptr(), Cc, _stream()), "cs_channel_sums")
    return out


class _ReduceItem(C.Structure):
    """cs_reduce_item of include/cs_b200.h."""
    _fields_ = [("inp", C.c_void_p), ("out", C.c_void_p), ("B", C.c_int32), ("C", C.c_int32), ("comp", C.c_int32), ("ncomp", C.c_int32)]


class ReduceQueue:
    """Collects batch_reduce() calls and issues them together (cs_batch_reduce_many: 24 per launch) at flush().  The backward
    of the denoiser ends every block with a handful of these tiny reductions; queued, ~244 launches per step become ~40.
    Inputs and outputs are kept alive until the flush; two items with the same output never share a launch, so the order of
    the additions into a parameter gradient is the program order (bit-reproducible, replicas stay identical)."""
    current: Optional["ReduceQueue"] = None

    def __init__(self):
        self.items = []

    def add(self, stat, comp, out):
        self.items.append((stat, comp, out))

    def flush(self):
        if not self.items:
            return
        lib, st = _lib.load(), _stream()
        batch, seen = [], set()

        def go():
            if batch:
                arr = (_ReduceItem * len(batch))(*batch)
                check(lib.cs_batch_reduce_many(arr, len(batch), st), "cs_batch_reduce_many")
                batch.clear(); seen.clear()
        for stat, comp, out in self.items:
            key = out.data_ptr()
            if key in seen:
                go()
            B, Cc, n = stat.shape
            batch.append(_ReduceItem(stat.data_ptr(), key, B, Cc, comp, n))
            seen.add(key)
        go()
        self.items = []

    def __enter__(self):
        self._prev = ReduceQueue.current
        ReduceQueue.current = self
        return self

    def __exit__(self, *exc):
        ReduceQueue.current = self._prev
        if exc[0] is None:
            self.flush()
        else:
            self.items = []


def flush_reductions() -> None:
    """Issue the queued batch reductions now (a point where their outputs must be final: a gradient bucket goes out)."""
    if ReduceQueue.current is not None:
        ReduceQueue.current.flush()


def batch_reduce(stat: torch.Tensor, comp: int, out: torch.Tensor) -> torch.Tensor:
    """out (C,) += sum_b stat[b, :, comp]   (stat fp32 (B, C, ncomp) contiguous).  Inside a ReduceQueue the launch is deferred
    to its next flush()."""
    if stat.dtype != torch.float32 or out.dtype != torch.float32 or not stat.is_contiguous() or not out.is_contiguous():
        raise _lib.CsError("batch_reduce: contiguous fp32 tensors required")
    if ReduceQueue.current is not None:
        ReduceQueue.current.add(stat, comp, out)
        return out
    B, Cc, n = stat.shape
    check(_lib.load().cs_batch_reduce(stat.data_ptr(), B, Cc, comp, n, out.data_ptr(), _stream()), "cs_batch_reduce")
    return out


def layernorm_bwd(x: torch.Tensor, gamma: torch.Tensor, dy: torch.Tensor, dgamma: torch.Tensor, dbeta: torch.Tensor,
                  eps: float = 1e-5, extra: Optional[torch.Tensor] = None) -> torch.Tensor:
    B, D, H, W, Cc, p = _check_act(x, "layernorm_bwd.x")
    _, _, _, _, _, pdy = _check_act(dy, "layernorm_bwd.dy")
    dx = torch.empty((B, D, H, W, Cc), dtype=torch.bfloat16, device=x.device)
    check(_lib.load().cs_layernorm_bwd(x.data_ptr(), B * D * H * W, Cc, p, dy.data_ptr(), pdy, gamma.data_ptr(), eps, _ptr(extra),
                                       0 if extra is None else extra.stride(3), dx.data_ptr(), Cc, dgamma.data_ptr(),
                                       dbeta.data_ptr(), _stream()), "cs_layernorm_bwd")
    return dx


def geglu_bwd(u: torch.Tensor, df: torch.Tensor) -> torch.Tensor:
    B, D, H, W, C2, p = _check_act(u, "geglu_bwd.u")
    _, _, _, _, I, pdf = _check_act(df, "geglu_bwd.df")
    if I * 2 != C2:
        raise _lib.CsError("geglu_bwd: df must have half the channels of u")
    du = torch.empty((B, D, H, W, C2), dtype=torch.bfloat16, device=u.device)
    check(_lib.load().cs_geglu_bwd(u.data_ptr(), B * D * H * W, I, p, df.data_ptr(), pdf, du.data_ptr(), C2, _stream()),
          "cs_geglu_bwd")
    return du


def upsample_nearest_bwd(dy: torch.Tensor, factors: Sequence[int]) -> torch.Tensor:
    B, Do, Ho, Wo, Cc, p = _check_act(dy, "upsample_bwd.dy")
    fd, fh, fw = factors
    D, H, W = Do // fd, Ho // fh, Wo // fw
    dx = torch.empty((B, D, H, W, Cc), dtype=torch.bfloat16, device=dy.device)
    check(_lib.load().cs_upsample_nearest_bwd(dy.data_ptr(), B, D, H, W, Cc, fd, fh, fw, p, dx.data_ptr(), Cc, _stream()),
          "cs_upsample_nearest_bwd")
    return dx


def zero_insert(x: torch.Tensor, stride: Sequence[int]) -> torch.Tensor:
    B, D, H, W, Cc, p = _check_act(x, "zero_insert.x")
    sd, sh, sw = stride
    out = torch.empty((B, D * sd, H * sh, W * sw, Cc), dtype=torch.bfloat16, device=x.device)
    check(_lib.load().cs_zero_insert(x.data_ptr(), B, D, H, W, Cc, sd, sh, sw, p, out.data_ptr(), Cc, _stream()), "cs_zero_insert")
    return out


def conv3d_dgrad_strided(dy: torch.Tensor, w_dgrad: torch.Tensor, stride: Sequence[int], **kw) -> torch.Tensor:
    """dX of a 3x3x3 / pad 1 conv with stride (sd, sh, sw) over an input whose extent is stride * output extent
    (Downsample, openai_model_3d.py:186-190): zero insertion followed by the stride-1 data-gradient conv."""
    return conv3d_dgrad(zero_insert(dy, stride), w_dgrad, **kw)


def add_(y: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """y += x (bf16 channels-last, same logical shape)."""
    B, D, H, W, Cc, py = _check_act(y, "add_.y")
    _, _, _, _, Cx, px = _check_act(x, "add_.x")
    if Cx != Cc or x.numel() != y.numel():
        raise _lib.CsError("add_: shape mismatch")
    check(_lib.load().cs_add_bf16(y.data_ptr(), py, x.data_ptr(), px, B * D * H * W, Cc, _stream()), "cs_add_bf16")
    return y


def sgemm(a: torch.Tensor, b: torch.Tensor, *, trans_a: bool = False, trans_b: bool = False, out: Optional[torch.Tensor] = None,
          accumulate: bool = False, silu_pre: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 out (M, N) [+]= op(a) @ op(b), optionally times SiLU'(silu_pre) elementwise; row-pitched 2-D operands."""
    for t in (a, b):
        if t.dtype != torch.float32 or t.dim() != 2 or t.stride(1) != 1 or not t.is_cuda:
            raise _lib.CsError("sgemm: operands must be fp32 2-D CUDA tensors with a contiguous last dim")
    M, K = (a.shape[1], a.shape[0]) if trans_a else (a.shape[0], a.shape[1])
    K2, N = (b.shape[1], b.shape[0]) if trans_b else (b.shape[0], b.shape[1])
    if K != K2:
        raise _lib.CsError(f"sgemm: inner dimensions differ ({K} vs {K2})")
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
        accumulate = False
    if out.dtype != torch.float32 or tuple(out.shape) != (M, N) or out.stride(1) != 1:
        raise _lib.CsError("sgemm: bad output")
    check(_lib.load().cs_sgemm_small(a.data_ptr(), a.stride(0), int(trans_a), b.data_ptr(), b.stride(0), int(trans_b),
                                     out.data_ptr(), out.stride(0), M, N, K, int(accumulate), _ptr(silu_pre),
                                     0 if silu_pre is None else silu_pre.stride(0), _stream()), "cs_sgemm_small")
    return out


def mse_loss_grad(pred: torch.Tensor, target: torch.Tensor, loss: torch.Tensor, loss_scale: float = 1.0,
                  need_grad: bool = True):
    """loss (fp32 scalar tensor) += mean((pred - target)^2); returns d(loss_scale * mean)/d pred (fp32) or None."""
    if pred.dtype != torch.float32 or target.dtype != torch.float32 or not pred.is_contiguous() or not target.is_contiguous() \
            or pred.shape != target.shape:
        raise _lib.CsError("mse_loss_grad: pred / target must be contiguous fp32 of the same shape")
    grad = torch.empty_like(pred) if need_grad else None
    check(_lib.load().cs_mse_loss_grad(pred.data_ptr(), target.data_ptr(), pred.numel(), loss_scale, _ptr(grad), loss.data_ptr(),
                                       _stream()), "cs_mse_loss_grad")
    return grad


_SUMSQ_WS = {}


def sumsq(g: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out (fp32 scalar) += sum(g^2), summed in a fixed order (bit-identical on every data-parallel replica)."""
    key = g.device
    ws = _SUMSQ_WS.get(key)
    if ws is None:      # block partials + ticket; one per device (calls are stream-ordered), zeroed once, left zeroed by the kernel
        ws = _SUMSQ_WS[key] = torch.zeros(2048, dtype=torch.float32, device=g.device)
    check(_lib.load().cs_sumsq(g.data_ptr(), g.numel(), out.data_ptr(), ws.data_ptr(), ws.numel() * 4, _stream()), "cs_sumsq")
    return out


def adamw_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, *, lr: float, betas=(0.9, 0.999),
               eps: float = 1e-8, weight_decay: float = 0.01, step: int = 0, sumsq_buf: Optional[torch.Tensor] = None,
               max_norm: float = 0.0, grad_scale: float = 1.0, step_dev: Optional[torch.Tensor] = None) -> None:
    """In-place torch.optim.AdamW step over flat fp32 buffers (same update rule; clip factor from sumsq_buf if given).
    `step_dev`: int32 device scalar holding the step number (overrides `step`; lets a CUDA graph be replayed)."""
    if step_dev is not None and (step_dev.dtype != torch.int32 or not step_dev.is_cuda):
        raise _lib.CsError("adamw_step: step_dev must be an int32 CUDA scalar")
    for t in (p, g, m, v):
        if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != p.numel():
            raise _lib.CsError("adamw_step: p, g, m, v must be contiguous fp32 of equal size")
    check(_lib.load().cs_adamw(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, betas[0], betas[1], eps,
                               weight_decay, step, _ptr(sumsq_buf), max_norm, grad_scale, _ptr(step_dev), _stream()), "cs_adamw")


# ----------------------------------------------------------------------------------------------
# attention
# ----------------------------------------------------------------------------------------------
def attention_lse(q, k, v, *, heads: int, head_dim: int, head_dim_padded: int, scale: float):
    """ops.attention that also returns the (B, heads, N) base-2 log-sum-exp rows the backward needs."""
    B, Nq = q.shape[0], q.shape[1]
    Nk = k.shape[1]
    out = torch.empty((B, Nq, heads * head_dim), dtype=torch.bfloat16, device=q.device)
    lse = torch.empty((B, heads, Nq), dtype=torch.float32, device=q.device)
    check(_lib.load().cs_attention_lse(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, heads, Nq, Nk,
                                       head_dim_padded, q.stride(1), k.stride(1), out.stride(1), head_dim, scale, lse.data_ptr(),
                                       _stream()), "cs_attention_lse")
    return out, lse


def attention_bwd(qkv: torch.Tensor, o: torch.Tensor, do: torch.Tensor, lse: torch.Tensor, *, heads: int, head_dim: int,
                  head_dim_padded: int, scale: float) -> torch.Tensor:
    """qkv: (B, N, 3 * heads * Dp) bf16 (the fused projection), o / do: (B, N, heads * head_dim) bf16 -> d_qkv like qkv."""
    B, N, W3 = qkv.shape
    hd = heads * head_dim_padded
    if W3 != 3 * hd or qkv.dtype != torch.bfloat16 or not qkv.is_contiguous():
        raise _lib.CsError("attention_bwd: qkv must be contiguous bf16 (B, N, 3 * heads * Dp)")
    if do.dtype != torch.bfloat16 or do.stride(-1) != 1 or o.stride(-1) != 1:
        raise _lib.CsError("attention_bwd: o / do must be bf16 with a contiguous last dim")
    dqkv = torch.empty_like(qkv)
    dsum = torch.empty((B, heads, N), dtype=torch.float32, device=qkv.device)
    es = qkv.element_size()
    check(_lib.load().cs_attention_bwd(qkv.data_ptr(), qkv.data_ptr() + hd * es, qkv.data_ptr() + 2 * hd * es, o.data_ptr(),
                                       do.data_ptr(), lse.data_ptr(), dsum.data_ptr(), dqkv.data_ptr(), dqkv.data_ptr() + hd * es,
                                       dqkv.data_ptr() + 2 * hd * es, B, heads, N, head_dim_padded, W3, o.stride(-2),
                                       do.stride(-2), W3, head_dim, scale, _stream()), "cs_attention_bwd")
    return dqkv


# ----------------------------------------------------------------------------------------------
# scene-graph conditioning gradients (fp32 row matrices; graphs have tens of rows)
# ----------------------------------------------------------------------------------------------
def _mat(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float32 or t.dim() != 2 or not t.is_cuda or (t.shape[1] > 1 and t.stride(1) != 1):
        raise _lib.CsError(f"{name}: expected an fp32 (M, C) CUDA matrix with contiguous rows")
    return t


def batchnorm_relu_bwd(x: torch.Tensor, y: Optional[torch.Tensor], dy: torch.Tensor, gamma, running_mean, running_var,
                       training: bool, eps: float = 1e-5, relu: bool = True, dgamma: Optional[torch.Tensor] = None,
                       dbeta: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dx of y = [relu](BatchNorm1d(x)); dgamma / dbeta (fp32 (C,)) are accumulated in place when given."""
    _mat(x, "batchnorm_bwd.x"), _mat(dy, "batchnorm_bwd.dy")
    M, Cc = x.shape
    if tuple(dy.shape) != (M, Cc) or (y is not None and tuple(_mat(y, "batchnorm_bwd.y").shape) != (M, Cc)):
        raise _lib.CsError("batchnorm_bwd: shape mismatch")
    dx = torch.empty((M, Cc), dtype=torch.float32, device=x.device)
    check(_lib.load().cs_batchnorm_relu_bwd(x.data_ptr(), M, Cc, x.stride(0), _ptr(gamma), _ptr(running_mean), _ptr(running_var),
                                            int(training), eps, int(relu), _ptr(y), 0 if y is None else y.stride(0),
                                            dy.data_ptr(), dy.stride(0), dx.data_ptr(), dx.stride(0), _ptr(dgamma), _ptr(dbeta),
                                            _stream()), "cs_batchnorm_relu_bwd")
    return dx


def gcn_scatter_mean_bwd(d_pooled: torch.Tensor, edges: torch.Tensor, hidden: int, mid: Optional[torch.Tensor],
                         mid_w: Optional[int] = None) -> torch.Tensor:
    """Gradient of net1's output (T, 2*hidden + mid_w) = [s | mid | o] from d_pooled (O, hidden) and the gradient of the middle
    (predicate) slice (`mid` None = that slice received no gradient: zeros of width mid_w)."""
    _mat(d_pooled, "scatter_mean_bwd.d_pooled")
    if not d_pooled.is_contiguous() or d_pooled.shape[1] != hidden:
        raise _lib.CsError("scatter_mean_bwd: d_pooled must be contiguous (O, hidden)")
    T, O = edges.shape[0], d_pooled.shape[0]
    mid_w = (mid_w or 0) if mid is None else _mat(mid, "scatter_mean_bwd.mid").shape[1]
    out = torch.empty((T, 2 * hidden + mid_w), dtype=torch.float32, device=d_pooled.device)
    check(_lib.load().cs_gcn_scatter_mean_bwd(d_pooled.data_ptr(), hidden, edges.data_ptr(), T, O, _ptr(mid),
                                              0 if mid is None else mid.stride(0), mid_w, out.data_ptr(), out.stride(0), 0, hidden,
                                              hidden + mid_w, _stream()), "cs_gcn_scatter_mean_bwd")
    return out


def gcn_gather_triples_bwd(d_in: torch.Tensor, edges: torch.Tensor, Do: int, Dp: int, d_obj: torch.Tensor, d_pred: torch.Tensor,
                           accumulate: bool = True) -> None:
    """d_obj (O, Do) / d_pred (T, Dp) (+)= the gradient of cat([obj[s], pred, obj[o]]) = d_in (T, 2*Do + Dp)."""
    _mat(d_in, "gather_bwd.d_in")
    if not (d_in.is_contiguous() and d_obj.is_contiguous() and d_pred.is_contiguous()) or d_in.shape[1] != 2 * Do + Dp:
        raise _lib.CsError("gather_triples_bwd: contiguous matrices of widths 2*Do+Dp / Do / Dp expected")
    check(_lib.load().cs_gcn_gather_triples_bwd(d_in.data_ptr(), d_obj.shape[0], Do, d_in.shape[0], Dp, edges.data_ptr(),
                                                int(accumulate), d_obj.data_ptr(), d_pred.data_ptr(), _stream()),
          "cs_gcn_gather_triples_bwd")


def embedding_bwd(d_rows: torch.Tensor, col_off: int, idx: torch.Tensor, d_weight: torch.Tensor) -> None:
    """d_weight (V, D) += scatter of d_rows[:, col_off:col_off+D] by idx (int64 (R,))."""
    _mat(d_rows, "embedding_bwd.d_rows")
    if idx.dtype != torch.int64 or not idx.is_contiguous() or idx.shape[0] != d_rows.shape[0] or not d_weight.is_contiguous():
        raise _lib.CsError("embedding_bwd: idx must be contiguous int64 (R,), d_weight contiguous")
    V, Dm = d_weight.shape
    check(_lib.load().cs_embedding_bwd(d_rows.data_ptr(), d_rows.stride(0), col_off, Dm, idx.data_ptr(), idx.shape[0], V,
                                       d_weight.data_ptr(), _stream()), "cs_embedding_bwd")
