"""ctypes binding of libcsb200.so (the C ABI declared in include/cs_b200.h).

The library is the product; there is no CPU or PyTorch fallback.  Importing this module works
without a GPU (so the CPU test-suite can check that the library loads and exports every
declared symbol), but any compute call on a machine without an sm_100 device raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libcsb200.so"

CS_OK, CS_ERR_INVALID, CS_ERR_CUDA, CS_ERR_UNSUPPORTED, CS_ERR_NO_DEVICE = range(5)
OUT_BF16_NDHWC, OUT_F32_NCDHW, OUT_F32_NDHWC = 0, 1, 2
ACT_NONE, ACT_SILU, ACT_GELU, ACT_GEGLU = 0, 1, 2, 3

_vp, _i32, _i64, _f32, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double


class Conv3dArgs(C.Structure):
    """Mirror of `cs_conv3d_args` (include/cs_b200.h)."""
    _fields_ = [
        ("in1", _vp), ("C1", _i32), ("in1_pitch", _i32),
        ("in2", _vp), ("C2", _i32), ("in2_pitch", _i32),
        ("B", _i32), ("D", _i32), ("H", _i32), ("W", _i32),
        ("weight", _vp), ("Cout", _i32),
        ("kd", _i32), ("kh", _i32), ("kw", _i32), ("sd", _i32), ("sh", _i32), ("sw", _i32),
        ("pd", _i32), ("ph", _i32), ("pw", _i32), ("pd_back", _i32), ("ph_back", _i32), ("pw_back", _i32),
        ("bias", _vp),
        ("rowvec", _vp), ("rowvec_pitch", _i32),
        ("residual", _vp), ("res_pitch", _i32),
        ("out", _vp), ("out_pitch", _i32), ("out_mode", _i32), ("act", _i32),
        ("stat_sum", _vp), ("stat_pitch", _i32),
        ("bn_hint", _i32),
        ("up_f", _i32 * 3), ("up_o", _i32 * 3),
    ]


class WgradArgs(C.Structure):
    """Mirror of `cs_conv3d_wgrad_args` (include/cs_b200.h)."""
    _fields_ = [
        ("x1", _vp), ("C1", _i32), ("x1_pitch", _i32),
        ("x2", _vp), ("C2", _i32), ("x2_pitch", _i32),
        ("B", _i32), ("D", _i32), ("H", _i32), ("W", _i32),
        ("dy", _vp), ("Cout", _i32), ("dy_pitch", _i32),
        ("kd", _i32), ("kh", _i32), ("kw", _i32), ("sd", _i32), ("sh", _i32), ("sw", _i32),
        ("pd", _i32), ("ph", _i32), ("pw", _i32), ("pd_back", _i32), ("ph_back", _i32), ("pw_back", _i32),
        ("dw", _vp),
    ]


# name -> (restype, argtypes); must list EVERY symbol include/cs_b200.h declares
SIGNATURES = {
    "cs_abi_version": (_i32, []),
    "cs_last_error": (C.c_char_p, []),
    "cs_device_check": (_i32, []),
    "cs_launch_count": (C.c_uint64, []),
    "cs_reset_launch_count": (None, []),
    "cs_conv3d": (_i32, [C.POINTER(Conv3dArgs), _vp]),
    "cs_conv3d_wgrad": (_i32, [C.POINTER(WgradArgs), _vp]),
    "cs_unpack_wgrad": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "cs_groupnorm_bwd": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _i32, _f32,
                                _i32, _vp, _vp, _i32, _vp, _i32, _i32, _vp]),
    "cs_batch_reduce": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "cs_layernorm_bwd": (_i32, [_vp, _i64, _i32, _i32, _vp, _i32, _vp, _f32, _vp, _i32, _vp, _i32, _vp, _vp, _vp]),
    "cs_geglu_bwd": (_i32, [_vp, _i64, _i32, _i32, _vp, _i32, _vp, _i32, _vp]),
    "cs_upsample_nearest_bwd": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp]),
    "cs_zero_insert": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp]),
    "cs_add_bf16": (_i32, [_vp, _i32, _vp, _i32, _i64, _i32, _vp]),
    "cs_cast_rows": (_i32, [_vp, _i32, _i64, _i32, _vp, _i32, _vp]),
    "cs_sgemm_small": (_i32, [_vp, _i32, _i32, _vp, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp]),
    "cs_mse_loss_grad": (_i32, [_vp, _vp, _i64, _f32, _vp, _vp, _vp]),
    "cs_sumsq": (_i32, [_vp, _i64, _vp, _vp, _i64, _vp]),
    "cs_adamw": (_i32, [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _f32, _i32, _vp, _f32, _f32, _vp, _vp]),
    "cs_pack_weight": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "cs_attention_lse": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _vp, _vp]),
    "cs_attention_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32,
                                _i32, _i32, _f32, _vp]),
    "cs_groupnorm_stats": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp, _i32, _vp]),
    "cs_channel_sums": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp, _i32, _vp]),
    "cs_groupnorm_finalize": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _vp, _vp]),
    "cs_groupnorm_apply": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _i32, _i32, _vp]),
    "cs_groupnorm_apply_fused": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _i32, _f32, _vp, _i32,
                                        _i32, _vp]),
    "cs_layernorm": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp, _f32, _vp, _i32, _vp]),
    "cs_attention": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _vp]),
    "cs_geglu": (_i32, [_vp, _i64, _i32, _i32, _vp, _i32, _vp]),
    "cs_upsample_nearest": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp]),
    "cs_im2col_small": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "cs_timestep_embedding": (_i32, [_vp, _i32, _i32, _f32, _vp, _vp]),
    "cs_linear_small": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp, _i32, _i32, _i32, _vp, _i32, _vp]),
    "cs_ddim_step": (_i32, [_vp, _vp, _i64, _i32, _f32, _f32, _f32, _f32, _f32, _vp, _vp, _vp, _vp]),
    "cs_q_sample": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "cs_ncdhw_to_ndhwc": (_i32, [_vp, _i32, _i32, _i64, _i32, _vp, _vp]),
    "cs_ndhwc_to_ncdhw": (_i32, [_vp, _i32, _i32, _i64, _i32, _vp, _vp]),
    "cs_channel_mix": (_i32, [_vp, _i32, _i32, _i32, _i64, _vp, _vp, _vp, _vp]),
    "cs_gcn_gather_triples": (_i32, [_vp, _i32, _i32, _vp, _i32, _i32, _vp, _vp, _vp]),
    "cs_gcn_scatter_mean": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _vp, _vp]),
    "cs_batchnorm_relu": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _f32, _f32, _i32, _vp, _i32, _vp]),
    "cs_add_rows": (_i32, [_vp, _i32, _vp, _i32, _i32, _i32, _vp, _i32, _vp]),
    "cs_batchnorm_relu_bwd": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _i32, _f32, _i32, _vp, _i32, _vp, _i32, _vp, _i32,
                                     _vp, _vp, _vp]),
    "cs_gcn_scatter_mean_bwd": (_i32, [_vp, _i32, _vp, _i32, _i32, _vp, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _vp]),
    "cs_gcn_gather_triples_bwd": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _vp, _vp]),
    "cs_embedding_bwd": (_i32, [_vp, _i32, _i32, _i32, _vp, _i32, _i32, _vp, _vp]),
    "cs_tap_gather": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "cs_cast_f32_to_bf16": (_i32, [_vp, _i64, _vp, _vp]),
    "cs_nn_distance": (_i32, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "cs_nn_distance_grad": (_i32, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cs_approx_match": (_i32, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "cs_match_cost": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "cs_match_cost_grad": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "cs_batch_reduce_many": (_i32, [_vp, _i32, _vp]),
    "cs_adamw_repack": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _f32, _f32, _f32, _f32, _f32, _i32, _vp, _f32, _f32, _vp, _vp]),
    "cs_surface_count": (_i32, [_vp, _i32, _i32, _i32, _i32, _f64, _vp, _vp, _vp, _vp, _vp]),
    "cs_surface_emit": (_i32, [_vp, _i32, _i32, _i32, _i32, _f64, _f64, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cs_debug_set": (None, [_i32]),
    "cs_conv3d_variant_counts": (None, [_vp, _i32]),
    "cs_vq_quantize": (_i32, [_vp, _i32, _i32, _i64, _vp, _i32, _vp, _vp, _i32, _vp, _vp, _vp]),
}

_lib = None


class CsError(RuntimeError):
    """A C-ABI call returned a non-zero status."""


def load() -> C.CDLL:
    """Load libcsb200.so (once).  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise CsError(
            f"{LIB_PATH} is missing: run `python -m commonscenes_b200.build` (nvcc, sm_100a). "
            "commonscenes_b200 has no CPU/PyTorch fallback.")
    try:
        import torch  # noqa: F401  (makes sure the process-wide libcudart.so.12 is the one torch uses)
    except Exception:  # pragma: no cover
        pass
    lib = C.CDLL(os.fspath(LIB_PATH), mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    flags = os.environ.get("CS_DEBUG_FLAGS")
    if flags:   # tuning experiments only (see cs_debug_set in include/cs_b200.h)
        lib.cs_debug_set(int(flags))
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != CS_OK:
        msg = load().cs_last_error()
        raise CsError(f"{what or 'libcsb200'} failed (status {status}): {msg.decode() if msg else ''}")


def require_device() -> None:
    """Fail loudly unless the current CUDA device is an sm_100 part."""
    check(load().cs_device_check(), "cs_device_check")
