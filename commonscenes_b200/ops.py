"""Torch-tensor front end of the C ABI (include/cs_b200.h).

Everything here is plumbing: it validates shapes/dtypes (mirroring the CHECK_INPUT style of the
reference's own native code, scripts/pytorch_structural_losses/src/structural_loss.cpp:10-12),
allocates outputs with torch, and passes raw device pointers plus the current CUDA stream to
libcsb200.so.  No arithmetic of the hot path is done by torch.

Activation convention: channels-last bf16 tensors of shape (B, D, H, W, C) whose last dim is
contiguous; the row pitch (stride of W) may be larger than C, so a tensor may be a channel slice
of a wider buffer.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import (ACT_GEGLU, ACT_GELU, ACT_NONE, ACT_SILU, OUT_BF16_NDHWC, OUT_F32_NCDHW, OUT_F32_NDHWC,
                   Conv3dArgs, check)

__all__ = [
    "ACT_NONE", "ACT_SILU", "ACT_GELU", "ACT_GEGLU", "pack_geglu_weight", "pack_patch_weight", "pack_upsample_phase_weights", "conv3d", "linear_tokens", "groupnorm", "layernorm", "attention",
    "geglu", "upsample_nearest", "im2col_small", "timestep_embedding", "linear_small", "ddim_step",
    "q_sample", "to_channels_last", "to_ncdhw", "pack_conv_weight", "pack_linear_weight", "launch_count",
    "reset_launch_count", "zero_stat_buffer", "groupnorm_stats", "groupnorm_fused", "ConvProfiler", "vq_quantize", "channel_mix", "pack_small_cout_conv", "conv3d_small_cout", "gcn_gather_triples", "gcn_scatter_mean",
    "batchnorm_relu", "add_rows", "cast_bf16",
]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _check_act(x: torch.Tensor, name: str) -> Tuple[int, int, int, int, int, int]:
    """Validate a channels-last bf16 activation; returns (B, D, H, W, C, pitch)."""
    if not x.is_cuda:
        raise _lib.CsError(f"{name}: expected a CUDA tensor (commonscenes_b200 has no CPU path)")
    if x.dtype != torch.bfloat16 or x.dim() != 5:
        raise _lib.CsError(f"{name}: expected a 5-D bf16 channels-last tensor, got {x.dtype} {tuple(x.shape)}")
    B, D, H, W, Cc = x.shape
    sb, sd, sh, sw, sc = x.stride()
    if sc != 1 and Cc != 1:
        raise _lib.CsError(f"{name}: channel dim must be contiguous")
    pitch = sw if W > 1 else (sh if H > 1 else (sd if D > 1 else (sb if B > 1 else Cc)))
    dense = (W == 1 or sw == pitch) and (H == 1 or sh == W * pitch) and (D == 1 or sd == H * W * pitch) \
        and (B == 1 or sb == D * H * W * pitch)
    if not dense or pitch < Cc:
        raise _lib.CsError(f"{name}: spatial dims must be densely packed (strides {x.stride()})")
    return B, D, H, W, Cc, pitch


def _f32(t: Optional[torch.Tensor], name: str) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
        raise _lib.CsError(f"{name}: expected a contiguous fp32 CUDA tensor")
    return t


def launch_count() -> int:
    return int(_lib.load().cs_launch_count())


def reset_launch_count() -> None:
    _lib.load().cs_reset_launch_count()


CONV_VARIANTS = ("one tile per CTA", "pair / hybrid work list", "CTA-pair kernel (cta_group::2)", "CTA pairs x two accumulators")


def conv3d_variant_counts(reset: bool = False) -> dict:
    """cs_conv3d launches per kernel variant since the last reset (which code path a shape / batch size selects)."""
    buf = (C.c_uint64 * 4)()
    _lib.load().cs_conv3d_variant_counts(buf, int(reset))
    return dict(zip(CONV_VARIANTS, (int(v) for v in buf)))


# ----------------------------------------------------------------------------------------------
# weight packing (one-off, at load / after an optimizer step) — layout plumbing, done with torch
# ----------------------------------------------------------------------------------------------
def _pad64(c: int) -> int:
    return (c + 63) // 64 * 64


def _pad_k(wt: torch.Tensor, split=None) -> torch.Tensor:
    """(Cout, taps, Cin) -> bf16 (Cout, taps, pad64(C1) [+ pad64(C2)]): each source's channel run is zero-padded to a
    multiple of 64 so that every 64-channel slab the kernel fetches starts on a 128-byte boundary."""
    co, taps, ci = wt.shape
    parts = [ci] if split is None else [int(c) for c in split]
    if sum(parts) != ci:
        raise _lib.CsError(f"pack: channel split {parts} does not sum to {ci}")
    out = torch.zeros((co, taps, sum(_pad64(c) for c in parts)), dtype=torch.bfloat16, device=wt.device)
    src = dst = 0
    for c in parts:
        out[:, :, dst:dst + c] = wt[:, :, src:src + c]
        src, dst = src + c, dst + _pad64(c)
    return out


_PACK_BUFFERS = {}


# Keys of cached pack buffers that somebody else keeps current (the fused optimizer kernel cs_adamw_repack rewrites the bf16
# packs of the weights it updates): pack_conv_weight / pack_dgrad_weight return such a buffer without launching.
# key -> (weak reference to the owning parameter, its autograd version when the registration was made).  An id() can be reused
# by a new tensor once the old one is collected, and a parameter written through torch (load_state_dict, an initialiser, an
# in-place op) bumps its version while the fused kernel does not: either way the registration is void and the pack is rebuilt.
_PACK_MAINTAINED = {}


def _pack_maintain(key, owner) -> None:
    _PACK_MAINTAINED[key] = (weakref.ref(owner), owner._version)


def _pack_is_maintained(key, owner) -> bool:
    hit = _PACK_MAINTAINED.get(key)
    if hit is None:
        return False
    if hit[0]() is owner and owner._version == hit[1]:
        return True
    del _PACK_MAINTAINED[key]
    return False


def _pack_key(kind: str, owner, shape):
    return (kind, id(owner), tuple(shape))


def _pack_buffer(kind: str, owner, shape, device) -> torch.Tensor:
    """Destination of a device-side weight pack.  For an nn.Parameter `owner` the buffer is cached per (layout, parameter):
    the pack kernels never write the pad columns, so it is zero-filled ONCE and re-packed in place after every optimizer
    step -- no per-step memsets (1.7 GB per training step of the denoiser) and stable addresses for CUDA-graph replays.
    Anything else (temporaries) gets a fresh zeroed tensor."""
    if not isinstance(owner, torch.nn.Parameter):
        return torch.zeros(shape, dtype=torch.bfloat16, device=device)
    key = (kind, id(owner), tuple(shape))
    hit = _PACK_BUFFERS.get(key)
    if hit is not None and hit[0]() is owner and hit[1].device == device:
        return hit[1]
    buf = torch.zeros(shape, dtype=torch.bfloat16, device=device)
    _PACK_BUFFERS[key] = (weakref.ref(owner), buf)
    if len(_PACK_BUFFERS) > 4096:          # drop the packs of parameters that no longer exist
        for k in [k for k, (r, _) in _PACK_BUFFERS.items() if r() is None]:
            del _PACK_BUFFERS[k]
    return buf


def pack_conv_weight(w: torch.Tensor, split=None, owner=None) -> torch.Tensor:
    """(Cout, Cin, kd, kh, kw) fp32 -> the K-major bf16 layout cs_conv3d reads (see _pad_k).  `split` = (C1, C2) when the
    conv consumes the channel concatenation of two tensors."""
    co, ci = w.shape[0], w.shape[1]
    wd = w.detach()
    if wd.is_cuda and wd.dtype == torch.float32 and wd.is_contiguous():      # device kernel (training re-packs every step)
        taps = wd.numel() // (co * ci)
        parts = [ci] if split is None else [int(c) for c in split]
        if sum(parts) != ci:
            raise _lib.CsError(f"pack: channel split {parts} does not sum to {ci}")
        own, kind, shape = owner if owner is not None else w, "fwd" + str(tuple(parts)), (co, taps, sum(_pad64(c) for c in parts))
        out = _pack_buffer(kind, own, shape, w.device)
        if _pack_is_maintained(_pack_key(kind, own, shape), own):
            return out
        check(_lib.load().cs_pack_weight(wd.data_ptr(), co, ci, taps, parts[0], out.data_ptr(), None, _stream()), "cs_pack_weight")
        return out
    return _pad_k(wd.reshape(co, ci, -1).permute(0, 2, 1), split)


_PHASE_TAPS = {0: ((1., 0., 0.), (0., 1., 1.)), 1: ((1., 1., 0.), (0., 0., 1.))}     # offset -> (2 merged taps) x (3 taps)


def pack_upsample_phase_weights(w: torch.Tensor, factors: Sequence[int]):
    """nearest-upsample by `factors` (each 1 or 2) followed by a 3x3x3 / pad 1 conv == one conv per output phase over the
    LOW-resolution tensor: along an axis with factor 2, output index 2i reads up-sampled rows 2i-1, 2i, 2i+1 = low-res rows
    i-1, i, i, and 2i+1 reads 2i, 2i+1, 2i+2 = rows i, i, i+1, so the three taps merge into two:
        offset 0: (w0, w1 + w2) with one zero row in front;   offset 1: (w0 + w1, w2) with one zero row behind.
    (Upsample, openai_model_3d.py:150-158; vqvae_modules.py:42-47.)  Taps are summed in fp32 and rounded to bf16 once.
    Returns [(offsets, ksize, pad, pad_back, packed weight)], one entry per phase."""
    if w.dim() != 5 or tuple(w.shape[2:]) != (3, 3, 3) or any(f not in (1, 2) for f in factors):
        raise _lib.CsError("pack_upsample_phase_weights: a (Cout, Cin, 3, 3, 3) filter and factors in {1, 2}")
    wf = w.detach().float()
    out = []
    for od in range(factors[0]):
        for oh in range(factors[1]):
            for ow in range(factors[2]):
                offs = (od, oh, ow)
                m = wf
                ksize, pad, pad_back = [], [], []
                for axis, (f, o) in enumerate(zip(factors, offs)):
                    if f == 1:
                        ksize.append(3); pad.append(1); pad_back.append(1)
                        continue
                    t = torch.tensor(_PHASE_TAPS[o], dtype=torch.float32, device=wf.device)          # (2, 3)
                    m = torch.tensordot(m, t, dims=([2 + axis], [1]))                                 # merged axis goes last
                    m = m.movedim(-1, 2 + axis)
                    ksize.append(2); pad.append(1 - o); pad_back.append(o)
                out.append((offs, tuple(ksize), tuple(pad), tuple(pad_back), pack_conv_weight(m.contiguous())))
    return out


def pack_geglu_weight(w: torch.Tensor, b: torch.Tensor):
    """GEGLU projection (2*inner, in): reorder rows so every 32-column group of the GEMM output holds 16 value
    columns followed by their 16 gate columns (what the ACT_GEGLU epilogue expects).  Returns (packed w, fp32 bias)."""
    inner = w.shape[0] // 2
    if inner % 16:
        raise _lib.CsError("pack_geglu_weight: inner dim must be a multiple of 16")
    idx = torch.arange(inner, device=w.device).view(-1, 16)
    perm = torch.cat([idx, idx + inner], dim=1).reshape(-1)
    return pack_linear_weight(w.detach()[perm]), b.detach().float()[perm].contiguous()


def pack_patch_weight(w: torch.Tensor):
    """3x3x3 conv with a few input channels -> GEMM weight over the cs_im2col_small patch matrix (column = tap*C + c).
    Returns (packed weight, Kp) with Kp = patch-matrix width (27*C rounded up to 16)."""
    w = w.detach().float()
    k = 27 * w.shape[1]
    kp = (k + 15) // 16 * 16
    wp = torch.zeros(w.shape[0], 1, kp, device=w.device)
    wp[:, 0, :k] = w.permute(0, 2, 3, 4, 1).reshape(w.shape[0], k)
    return _pad_k(wp), kp


def pack_linear_weight(w: torch.Tensor, owner=None) -> torch.Tensor:
    """(out, in) fp32 -> (out, 1, pad64(in)) bf16."""
    if w.is_cuda and w.dtype == torch.float32 and w.is_contiguous():
        return pack_conv_weight(w, owner=owner)
    return _pad_k(w.detach().reshape(w.shape[0], 1, w.shape[1]))


# ----------------------------------------------------------------------------------------------
# GEMM-class
# ----------------------------------------------------------------------------------------------
class ConvProfiler:
    """Context manager: brackets every cs_conv3d launch with CUDA events on the launching stream and records its
    algorithmic FLOPs (2*MAC over real channels/taps).  Used by bench.py for the live roofline numbers."""
    active: Optional["ConvProfiler"] = None

    def __init__(self):
        self.records = []   # (start_event, end_event, flops, tag)
        self.bytes = 0      # algorithmic bytes of the recorded launches: activations in + weights + residual + output

    def __enter__(self):
        ConvProfiler.active = self
        return self

    def __exit__(self, *exc):
        ConvProfiler.active = None

    def summary(self):
        """(total ms, total TFLOP, launches) — call after a synchronize."""
        ms = sum(a.elapsed_time(b) for a, b, _, _ in self.records)
        return ms, sum(f for _, _, f, _ in self.records) / 1e12, len(self.records)

    def table(self):
        return [(tag, a.elapsed_time(b), f) for a, b, f, tag in self.records]


def conv3d(x: torch.Tensor, weight: torch.Tensor, *, ksize: Sequence[int] = (3, 3, 3),
           stride: Sequence[int] = (1, 1, 1), pad: Sequence[int] = (1, 1, 1),
           pad_back: Optional[Sequence[int]] = None, bias: Optional[torch.Tensor] = None,
           rowvec: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
           x2: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, out_mode: int = OUT_BF16_NDHWC,
           act: int = ACT_NONE, stat_sum: Optional[torch.Tensor] = None, bn_hint: int = 0,
           phase: Optional[Tuple[Sequence[int], Sequence[int]]] = None) -> torch.Tensor:
    """act(conv3d(cat(x, x2)) + bias + rowvec[b] + residual) on the tcgen05 implicit-GEMM kernel.

    `weight` is the packed (Cout, taps, Cin) bf16 tensor of `pack_conv_weight`.
    `phase` = ((f_d, f_h, f_w), (o_d, o_h, o_w)): phase launch of a nearest-upsample + conv (pack_upsample_phase_weights):
    the rows of this (low-resolution) conv are scattered to the voxels of `out` (required, full resolution) congruent to
    the offsets modulo the factors.
    """
    lib = _lib.load()
    B, D, H, W, C1, p1 = _check_act(x, "conv3d.x")
    C2, p2 = 0, 0
    if x2 is not None:
        B2, D2, H2, W2, C2, p2 = _check_act(x2, "conv3d.x2")
        if (B2, D2, H2, W2) != (B, D, H, W):
            raise _lib.CsError("conv3d: x and x2 must share batch and spatial dims")
    kd, kh, kw = ksize
    if weight.dtype != torch.bfloat16 or weight.dim() != 3 or not weight.is_contiguous() or \
            weight.shape[1] != kd * kh * kw or weight.shape[2] != _pad64(C1) + _pad64(C2):
        raise _lib.CsError(f"conv3d: packed weight must be bf16 (Cout, {kd * kh * kw}, {_pad64(C1) + _pad64(C2)}), "
                           f"got {tuple(weight.shape)} (use pack_conv_weight / pack_linear_weight)")
    Cout = weight.shape[0]
    Cres = Cout // 2 if act == ACT_GEGLU else Cout      # channels of the tensor written
    pb = tuple(pad) if pad_back is None else tuple(pad_back)
    Do = (D + pad[0] + pb[0] - kd) // stride[0] + 1
    Ho = (H + pad[1] + pb[1] - kh) // stride[1] + 1
    Wo = (W + pad[2] + pb[2] - kw) // stride[2] + 1
    fd, fh, fw = (1, 1, 1) if phase is None else tuple(int(f) for f in phase[0])
    if phase is not None and (out is None or out_mode != OUT_BF16_NDHWC or residual is not None):
        raise _lib.CsError("conv3d: a phase launch writes into a caller-provided bf16 channels-last `out` and takes no residual")
    if out is None:
        if out_mode == OUT_BF16_NDHWC:
            out = torch.empty((B, Do, Ho, Wo, Cres), dtype=torch.bfloat16, device=x.device)
        elif out_mode == OUT_F32_NCDHW:
            out = torch.empty((B, Cout, Do, Ho, Wo), dtype=torch.float32, device=x.device)
        else:
            out = torch.empty((B, Do, Ho, Wo, Cout), dtype=torch.float32, device=x.device)
    if out_mode == OUT_F32_NCDHW:
        if out.dtype != torch.float32 or not out.is_contiguous() or tuple(out.shape) != (B, Cout, Do, Ho, Wo):
            raise _lib.CsError("conv3d: NCDHW output must be contiguous fp32 (B, Cout, Do, Ho, Wo)")
        out_pitch = 0
    else:
        want = torch.bfloat16 if out_mode == OUT_BF16_NDHWC else torch.float32
        if out.dtype != want or tuple(out.shape) != (B, Do * fd, Ho * fh, Wo * fw, Cres) or out.stride(-1) != 1:
            raise _lib.CsError(f"conv3d: output must be {want} (B, Do, Ho, Wo, Cout)")
        if phase is not None:
            _check_act(out, "conv3d.out")
        out_pitch = out.stride(3)
    a = Conv3dArgs()
    a.in1, a.C1, a.in1_pitch = x.data_ptr(), C1, p1
    a.in2, a.C2, a.in2_pitch = _ptr(x2), C2, p2
    a.B, a.D, a.H, a.W = B, D, H, W
    a.weight, a.Cout = weight.data_ptr(), Cout
    a.kd, a.kh, a.kw = kd, kh, kw
    a.sd, a.sh, a.sw = stride
    a.pd, a.ph, a.pw = pad
    a.pd_back, a.ph_back, a.pw_back = pb
    a.bias = _ptr(_f32(bias, "conv3d.bias"))
    if rowvec is not None:
        if rowvec.dtype != torch.float32 or rowvec.dim() != 2 or rowvec.shape[0] != B or rowvec.shape[1] != Cout \
                or rowvec.stride(1) != 1:
            raise _lib.CsError("conv3d: rowvec must be fp32 (B, Cout)")
        a.rowvec, a.rowvec_pitch = rowvec.data_ptr(), rowvec.stride(0)
    if residual is not None:
        rB, rD, rH, rW, rC, rp = _check_act(residual, "conv3d.residual")
        if (rB, rD, rH, rW, rC) != (B, Do, Ho, Wo, Cout):
            raise _lib.CsError("conv3d: residual shape must equal the output shape")
        a.residual, a.res_pitch = residual.data_ptr(), rp
    a.out, a.out_pitch, a.out_mode, a.act = out.data_ptr(), out_pitch, out_mode, act
    if stat_sum is not None:
        a.stat_sum, a.stat_pitch = _check_stat(stat_sum, "conv3d.stat_sum").data_ptr(), stat_sum.shape[1]
    a.bn_hint = bn_hint
    if phase is not None:
        a.up_f[:] = [fd, fh, fw]
        a.up_o[:] = [int(o) for o in phase[1]]
    prof = ConvProfiler.active
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(lib.cs_conv3d(C.byref(a), _stream()), "cs_conv3d")
    if prof is not None:
        e1.record()
        flops = 2.0 * B * Do * Ho * Wo * Cout * (C1 + C2) * kd * kh * kw
        prof.bytes += 2 * B * D * H * W * (C1 + C2) + 2 * weight.numel() + out.numel() * out.element_size() \
            + (0 if residual is None else 2 * residual.numel())
        prof.records.append((e0, e1, flops, f"{C1 + C2}->{Cout} k{kd}{kh}{kw} s{stride[0]}{stride[1]}{stride[2]} @{Do}x{Ho}x{Wo} B{B}"))
    return out


def linear_tokens(x: torch.Tensor, weight: torch.Tensor, **kw) -> torch.Tensor:
    """nn.Linear / 1x1x1 conv over a (B, D, H, W, C) token grid (same kernel, one tap)."""
    return conv3d(x, weight, ksize=(1, 1, 1), pad=(0, 0, 0), **kw)


# ----------------------------------------------------------------------------------------------
# normalisation
# ----------------------------------------------------------------------------------------------
_ws: dict = {}


def _workspace(device: torch.device, key: str, numel: int, zero: bool = False, dtype=torch.float32) -> torch.Tensor:
    k = (device, key)
    t = _ws.get(k)
    if t is None or t.numel() < numel:
        t = torch.zeros(numel, dtype=dtype, device=device)
        _ws[k] = t
    return t


STAT_DTYPE = torch.int64     # GroupNorm sum buffers: 64-bit fixed point, see include/cs_b200.h (cs_groupnorm_stats)
STAT_SUM_SCALE, STAT_SQ_SCALE = float(1 << 20), float(1 << 12)


def stat_to_float(stat: torch.Tensor) -> torch.Tensor:
    """(B, C, 2) fixed-point GroupNorm sums -> fp64 (sum, sum of squares) (tests / diagnostics)."""
    return torch.stack([stat[..., 0].double() / STAT_SUM_SCALE, stat[..., 1].double() / STAT_SQ_SCALE], dim=-1)


def _check_stat(stat: Optional[torch.Tensor], name: str) -> Optional[torch.Tensor]:
    if stat is not None and (stat.dtype != STAT_DTYPE or not stat.is_cuda or not stat.is_contiguous()):
        raise _lib.CsError(f"{name}: GroupNorm sum buffers are contiguous int64 (B, C, 2) CUDA tensors (fixed point)")
    return stat


def zero_stat_buffer(device: torch.device, B: int, C: int) -> torch.Tensor:
    """(B, C, 2) int64 (fixed-point) view of the shared GroupNorm-sum workspace.  It is all-zero on return (the finalize
    kernel clears what it consumes); pass it as `stat_sum` to conv3d and then to the groupnorm that follows,
    with no other groupnorm in between."""
    return _workspace(device, "gn_stat", B * C * 2, dtype=STAT_DTYPE)[:B * C * 2].view(B, C, 2)


def groupnorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, *, groups: int = 32, eps: float = 1e-5,
              act: int = ACT_NONE, x2: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
              stat_sum: Optional[torch.Tensor] = None) -> torch.Tensor:
    """act(GroupNorm(cat(x, x2))) -> bf16 channels-last.  If `stat_sum` (B, C, 2) already holds the
    per-channel sums (produced by a conv epilogue) the statistics pass is skipped."""
    lib = _lib.load()
    B, D, H, W, C1, p1 = _check_act(x, "groupnorm.x")
    S = D * H * W
    C2 = 0
    if x2 is not None:
        _, _, _, _, C2, p2 = _check_act(x2, "groupnorm.x2")
    Ct = C1 + C2
    if out is None:
        out = torch.empty((B, D, H, W, Ct), dtype=torch.bfloat16, device=x.device)
    oB, oD, oH, oW, oC, op = _check_act(out, "groupnorm.out")
    if (oB, oD, oH, oW, oC) != (B, D, H, W, Ct):
        raise _lib.CsError("groupnorm: bad output shape")
    st = _stream()
    if stat_sum is None:
        stat = _workspace(x.device, "gn_stat", B * Ct * 2, dtype=STAT_DTYPE)  # kept all-zero between calls by finalize
        check(lib.cs_groupnorm_stats(x.data_ptr(), B, S, C1, p1, stat.data_ptr(), Ct, st), "cs_groupnorm_stats")
        if x2 is not None:
            check(lib.cs_groupnorm_stats(x2.data_ptr(), B, S, C2, p2, stat.data_ptr() + C1 * 16, Ct, st),
                  "cs_groupnorm_stats")
    else:
        stat = _check_stat(stat_sum, "groupnorm.stat_sum")
    ss = _workspace(x.device, "gn_ss", B * Ct * 2)
    check(lib.cs_groupnorm_finalize(stat.data_ptr(), _ptr(_f32(gamma, "gamma")), _ptr(_f32(beta, "beta")), B, Ct,
                                    groups, S, eps, ss.data_ptr(), st), "cs_groupnorm_finalize")
    check(lib.cs_groupnorm_apply(x.data_ptr(), B, S, C1, p1, ss.data_ptr(), Ct, out.data_ptr(), op, act, st),
          "cs_groupnorm_apply")
    if x2 is not None:
        check(lib.cs_groupnorm_apply(x2.data_ptr(), B, S, C2, p2, ss.data_ptr() + C1 * 8, Ct,
                                     out.data_ptr() + C1 * 2, op, act, st), "cs_groupnorm_apply")
    return out


def groupnorm_stats(x: torch.Tensor, stat: torch.Tensor) -> torch.Tensor:
    """Accumulate the per-(sample, channel) sum / sum of squares of x into the ZEROED fixed-point buffer stat (B, C, 2)."""
    B, D, H, W, Cc, p = _check_act(x, "groupnorm_stats.x")
    _check_stat(stat, "groupnorm_stats.stat")
    check(_lib.load().cs_groupnorm_stats(x.data_ptr(), B, D * H * W, Cc, p, stat.data_ptr(), Cc, _stream()), "cs_groupnorm_stats")
    return stat


def groupnorm_fused(x: torch.Tensor, stat: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, *, groups: int = 32,
                    eps: float = 1e-5, act: int = ACT_NONE, x2: Optional[torch.Tensor] = None,
                    stat2: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """act(GroupNorm(cat(x, x2))) from per-channel sums that already exist (`stat`, `stat2`: int64 fixed-point (B, C, 2) written by the
    producing conv's epilogue or by groupnorm_stats).  One kernel per source; nothing is cleared."""
    lib = _lib.load()
    _check_stat(stat, "groupnorm_fused.stat"), _check_stat(stat2, "groupnorm_fused.stat2")
    B, D, H, W, C1, p1 = _check_act(x, "groupnorm_fused.x")
    S = D * H * W
    C2 = 0
    if x2 is not None:
        _, _, _, _, C2, p2 = _check_act(x2, "groupnorm_fused.x2")
    Ct = C1 + C2
    if out is None:
        out = torch.empty((B, D, H, W, Ct), dtype=torch.bfloat16, device=x.device)
    op = out.stride(3)
    st = _stream()
    g, bta = _ptr(_f32(gamma, "gamma")), _ptr(_f32(beta, "beta"))
    check(lib.cs_groupnorm_apply_fused(x.data_ptr(), B, S, C1, p1, 0, stat.data_ptr(), C1, _ptr(stat2), C2, g, bta, groups, eps,
                                       out.data_ptr(), op, act, st), "cs_groupnorm_apply_fused")
    if x2 is not None:
        check(lib.cs_groupnorm_apply_fused(x2.data_ptr(), B, S, C2, p2, C1, stat.data_ptr(), C1, stat2.data_ptr(), C2, g, bta,
                                           groups, eps, out.data_ptr() + C1 * 2, op, act, st), "cs_groupnorm_apply_fused")
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    B, D, H, W, Cc, p = _check_act(x, "layernorm.x")
    if out is None:
        out = torch.empty((B, D, H, W, Cc), dtype=torch.bfloat16, device=x.device)
    check(_lib.load().cs_layernorm(x.data_ptr(), B * D * H * W, Cc, p, _f32(gamma, "gamma").data_ptr(),
                                   _f32(beta, "beta").data_ptr(), eps, out.data_ptr(), out.stride(3), _stream()),
          "cs_layernorm")
    return out


# ----------------------------------------------------------------------------------------------
# attention / pointwise
# ----------------------------------------------------------------------------------------------
def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, heads: int, head_dim: int, head_dim_padded: int,
              scale: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q: (B, Nq, >=heads*Dp) / k,v: (B, Nk, >=heads*Dp) bf16 row-pitched views -> (B, Nq, heads*head_dim)."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        if t.dtype != torch.bfloat16 or t.dim() != 3 or t.stride(2) != 1 or not t.is_cuda:
            raise _lib.CsError(f"attention.{n}: expected bf16 (B, N, heads*Dp) with contiguous last dim")
    B, Nq = q.shape[0], q.shape[1]
    Nk = k.shape[1]
    if q.stride(0) != Nq * q.stride(1) or k.stride(0) != Nk * k.stride(1) or v.stride() != k.stride():
        raise _lib.CsError("attention: batch stride must equal N * row pitch; k and v must share strides")
    if out is None:
        out = torch.empty((B, Nq, heads * head_dim), dtype=torch.bfloat16, device=q.device)
    check(_lib.load().cs_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, heads, Nq, Nk,
                                   head_dim_padded, q.stride(1), k.stride(1), out.stride(1), head_dim, scale,
                                   _stream()), "cs_attention")
    return out


def geglu(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    B, D, H, W, C2, p = _check_act(x, "geglu.x")
    Ch = C2 // 2
    if out is None:
        out = torch.empty((B, D, H, W, Ch), dtype=torch.bfloat16, device=x.device)
    check(_lib.load().cs_geglu(x.data_ptr(), B * D * H * W, Ch, p, out.data_ptr(), out.stride(3), _stream()),
          "cs_geglu")
    return out


def upsample_nearest(x: torch.Tensor, factors: Sequence[int], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    B, D, H, W, Cc, p = _check_act(x, "upsample.x")
    fd, fh, fw = factors
    if out is None:
        out = torch.empty((B, D * fd, H * fh, W * fw, Cc), dtype=torch.bfloat16, device=x.device)
    check(_lib.load().cs_upsample_nearest(x.data_ptr(), B, D, H, W, Cc, p, fd, fh, fw, out.data_ptr(),
                                          out.stride(3), _stream()), "cs_upsample_nearest")
    return out


def im2col_small(x: torch.Tensor, batch: Optional[int] = None, kp: Optional[int] = None) -> torch.Tensor:
    """fp32 NCDHW (few channels) -> bf16 (B, D, H, W, Kp) patch matrix for a 3x3x3 / pad 1 conv."""
    _f32(x, "im2col_small.x")
    Bs, Cc, D, H, W = x.shape
    B = Bs if batch is None else batch
    if kp is None:
        kp = (27 * Cc + 15) // 16 * 16
    col = torch.empty((B, D, H, W, kp), dtype=torch.bfloat16, device=x.device)
    check(_lib.load().cs_im2col_small(x.data_ptr(), Bs, B, Cc, D, H, W, kp, col.data_ptr(), _stream()),
          "cs_im2col_small")
    return col


def timestep_embedding(t: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    if t.dtype != torch.int64 or not t.is_cuda or not t.is_contiguous():
        raise _lib.CsError("timestep_embedding: t must be a contiguous int64 CUDA tensor")
    out = torch.empty((t.shape[0], dim), dtype=torch.float32, device=t.device)
    check(_lib.load().cs_timestep_embedding(t.data_ptr(), t.shape[0], dim, max_period, out.data_ptr(), _stream()),
          "cs_timestep_embedding")
    return out


def linear_small(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, act_in: int = ACT_NONE,
                 act_out: int = ACT_NONE, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 (M, K) @ (N, K)^T for a handful of rows (time embedding / context projections)."""
    if x.dtype != torch.float32 or x.dim() != 2 or x.stride(1) != 1 or not x.is_cuda:
        raise _lib.CsError("linear_small: x must be fp32 (M, K)")
    _f32(w, "linear_small.w")
    M, K = x.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise _lib.CsError("linear_small: weight must be (N, K)")
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=x.device)
    check(_lib.load().cs_linear_small(x.data_ptr(), M, K, x.stride(0), w.data_ptr(), _ptr(_f32(bias, "bias")), N,
                                      act_in, act_out, out.data_ptr(), out.stride(0), _stream()), "cs_linear_small")
    return out


def ddim_step(x: torch.Tensor, eps: torch.Tensor, *, guided: bool, scale: float, a_t: float, a_prev: float,
              sigma: float, sqrt_one_minus_at: float, noise: Optional[torch.Tensor] = None,
              want_pred_x0: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    _f32(x, "ddim_step.x"), _f32(eps, "ddim_step.eps")
    n = x.numel()
    if eps.numel() != (2 * n if guided else n):
        raise _lib.CsError("ddim_step: eps must hold [uncond; cond] (2x) when guided, else 1x")
    x_prev = torch.empty_like(x)
    pred = torch.empty_like(x) if want_pred_x0 else None
    check(_lib.load().cs_ddim_step(x.data_ptr(), eps.data_ptr(), n, int(guided), scale, a_t, a_prev, sigma,
                                   sqrt_one_minus_at, _ptr(_f32(noise, "noise")), x_prev.data_ptr(), _ptr(pred),
                                   _stream()), "cs_ddim_step")
    return x_prev, pred


def q_sample(x0: torch.Tensor, noise: torch.Tensor, t: torch.Tensor, sqrt_ac: torch.Tensor,
             sqrt_1mac: torch.Tensor) -> torch.Tensor:
    _f32(x0, "x0"), _f32(noise, "noise"), _f32(sqrt_ac, "sqrt_ac"), _f32(sqrt_1mac, "sqrt_1mac")
    out = torch.empty_like(x0)
    B = x0.shape[0]
    check(_lib.load().cs_q_sample(x0.data_ptr(), noise.data_ptr(), t.data_ptr(), sqrt_ac.data_ptr(),
                                  sqrt_1mac.data_ptr(), x0.numel() // B, B, out.data_ptr(), _stream()), "cs_q_sample")
    return out


def to_channels_last(x: torch.Tensor, c_pad: Optional[int] = None) -> torch.Tensor:
    """NCDHW fp32 -> (B, D, H, W, Cp) bf16, channels zero-padded to Cp."""
    _f32(x, "to_channels_last.x")
    B, Cc, D, H, W = x.shape
    cp = Cc if c_pad is None else c_pad
    y = torch.empty((B, D, H, W, cp), dtype=torch.bfloat16, device=x.device)
    check(_lib.load().cs_ncdhw_to_ndhwc(x.data_ptr(), B, Cc, D * H * W, cp, y.data_ptr(), _stream()),
          "cs_ncdhw_to_ndhwc")
    return y


def to_ncdhw(x: torch.Tensor, channels: Optional[int] = None) -> torch.Tensor:
    B, D, H, W, Cc, p = _check_act(x, "to_ncdhw.x")
    cc = Cc if channels is None else channels
    y = torch.empty((B, cc, D, H, W), dtype=torch.float32, device=x.device)
    check(_lib.load().cs_ndhwc_to_ncdhw(x.data_ptr(), B, cc, D * H * W, p, y.data_ptr(), _stream()),
          "cs_ndhwc_to_ncdhw")
    return y


def vq_quantize(z: torch.Tensor, codebook: torch.Tensor, post_w: Optional[torch.Tensor] = None,
                post_b: Optional[torch.Tensor] = None, want_indices: bool = True):
    """z: fp32 NCDHW (B, E, D, H, W) -> (z_q [optionally through post_quant_conv], int64 indices (B*D*H*W,))."""
    _f32(z, "vq_quantize.z"), _f32(codebook, "vq_quantize.codebook")
    B, E = z.shape[0], z.shape[1]
    S = z.numel() // (B * E)
    zc = E if post_w is None else post_w.shape[0]
    out = torch.empty((B, zc) + tuple(z.shape[2:]), dtype=torch.float32, device=z.device)
    idx = torch.empty(B * S, dtype=torch.int64, device=z.device) if want_indices else None
    check(_lib.load().cs_vq_quantize(z.data_ptr(), B, E, S, codebook.data_ptr(), codebook.shape[0],
                                     _ptr(_f32(post_w, "post_w")), _ptr(_f32(post_b, "post_b")), zc, out.data_ptr(),
                                     _ptr(idx), _stream()), "cs_vq_quantize")
    return out, idx


def channel_mix(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """1x1x1 conv on a few-channel fp32 NCDHW tensor: (B, Ci, ...) x (Co, Ci) -> (B, Co, ...)."""
    _f32(x, "channel_mix.x"), _f32(w, "channel_mix.w")
    B, Ci = x.shape[0], x.shape[1]
    S = x.numel() // (B * Ci)
    y = torch.empty((B, w.shape[0]) + tuple(x.shape[2:]), dtype=torch.float32, device=x.device)
    check(_lib.load().cs_channel_mix(x.data_ptr(), B, Ci, w.shape[0], S, w.data_ptr(), _ptr(_f32(bias, "bias")),
                                     y.data_ptr(), _stream()), "cs_channel_mix")
    return y


def pack_small_cout_conv(w: torch.Tensor):
    """(Co <= 4, Cin, 3, 3, 3) -> per-tap GEMM weight (27*Co padded to 16, 1, Cin) bf16: row tap*Co + co = w[co, :, tap]."""
    co, ci = w.shape[0], w.shape[1]
    rows = (27 * co + 15) // 16 * 16
    wt = torch.zeros(rows, 1, ci, dtype=torch.float32, device=w.device)
    wt[:27 * co, 0] = w.detach().float().reshape(co, ci, 27).permute(2, 0, 1).reshape(27 * co, ci)
    return _pad_k(wt)


def conv3d_small_cout(x: torch.Tensor, w_taps: torch.Tensor, bias: Optional[torch.Tensor], cout: int) -> torch.Tensor:
    """3x3x3 / pad 1 conv to <= 4 channels: one tensor-core GEMM over the taps + a gather of the 27 shifted planes.
    x: channels-last bf16 (B, D, H, W, Cin) -> fp32 NCDHW (B, cout, D, H, W)."""
    B, D, H, W, _, _ = _check_act(x, "conv3d_small_cout.x")
    y = conv3d(x, w_taps, ksize=(1, 1, 1), pad=(0, 0, 0), out_mode=OUT_F32_NCDHW)
    out = torch.empty((B, cout, D, H, W), dtype=torch.float32, device=x.device)
    check(_lib.load().cs_tap_gather(y.data_ptr(), B, y.shape[1], cout, D, H, W, _ptr(_f32(bias, "bias")), out.data_ptr(),
                                    _stream()), "cs_tap_gather")
    return out


# ----------------------------------------------------------------------------------------------
# scene-graph conditioning (fp32, tiny)
# ----------------------------------------------------------------------------------------------
def _mat(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float32 or t.dim() != 2 or not t.is_cuda or (t.shape[1] > 1 and t.stride(1) != 1):
        raise _lib.CsError(f"{name}: expected an fp32 (M, C) CUDA matrix with contiguous rows")
    return t


def gcn_gather_triples(obj: torch.Tensor, pred: torch.Tensor, edges: torch.Tensor) -> torch.Tensor:
    _f32(obj, "obj"), _f32(pred, "pred")
    if edges.dtype != torch.int64 or not edges.is_contiguous() or edges.dim() != 2 or edges.shape[1] != 2:
        raise _lib.CsError("gcn_gather_triples: edges must be contiguous int64 (T, 2)")
    O, Do = obj.shape
    T, Dp = pred.shape
    out = torch.empty((T, 2 * Do + Dp), dtype=torch.float32, device=obj.device)
    check(_lib.load().cs_gcn_gather_triples(obj.data_ptr(), O, Do, pred.data_ptr(), T, Dp, edges.data_ptr(), out.data_ptr(),
                                            _stream()), "cs_gcn_gather_triples")
    return out


def gcn_scatter_mean(tv: torch.Tensor, s_off: int, o_off: int, hidden: int, edges: torch.Tensor, num_objs: int) -> torch.Tensor:
    _mat(tv, "tv")
    pooled = torch.empty((num_objs, hidden), dtype=torch.float32, device=tv.device)
    check(_lib.load().cs_gcn_scatter_mean(tv.data_ptr(), tv.stride(0), s_off, o_off, hidden, edges.data_ptr(), tv.shape[0],
                                          num_objs, pooled.data_ptr(), _stream()), "cs_gcn_scatter_mean")
    return pooled


def batchnorm_relu(x: torch.Tensor, gamma, beta, running_mean, running_var, training: bool, momentum: float = 0.1,
                   eps: float = 1e-5, relu: bool = True) -> torch.Tensor:
    _mat(x, "batchnorm.x")
    M, Cc = x.shape
    if training and M < 2:
        raise ValueError(f"Expected more than 1 value per channel when training, got input size {tuple(x.shape)}")
    y = torch.empty((M, Cc), dtype=torch.float32, device=x.device)
    check(_lib.load().cs_batchnorm_relu(x.data_ptr(), M, Cc, x.stride(0), _ptr(gamma), _ptr(beta), _ptr(running_mean),
                                        _ptr(running_var), int(training), momentum, eps, int(relu), y.data_ptr(), y.stride(0),
                                        _stream()), "cs_batchnorm_relu")
    return y


def add_rows(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    _mat(a, "add_rows.a"), _mat(b, "add_rows.b")
    M, Cc = a.shape
    y = torch.empty((M, Cc), dtype=torch.float32, device=a.device)
    check(_lib.load().cs_add_rows(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), M, Cc, y.data_ptr(), y.stride(0),
                                  _stream()), "cs_add_rows")
    return y


def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    """contiguous fp32 CUDA tensor -> bf16 copy of the same shape."""
    _f32(x, "cast_bf16.x")
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    check(_lib.load().cs_cast_f32_to_bf16(x.data_ptr(), x.numel(), y.data_ptr(), _stream()), "cs_cast_f32_to_bf16")
    return y
