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
# GEMM-class gradients
# ----------------------------------------------------------------------------------------------
def pack_dgrad_weight(w: torch.Tensor) -> torch.Tensor:
    """(Cout, Cin, kd, kh, kw) fp32 -> packed bf16 (Cin, taps, pad64(Cout)) with the filter flipped: the weight with which
    cs_conv3d maps dY (Cout channels) to dX (Cin channels) for a stride-1 convolution."""
    co, ci = w.shape[0], w.shape[1]
    wt = w.detach().reshape(co, ci, -1).flip(2).permute(1, 2, 0)     # (Cin, taps flipped, Cout)
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
