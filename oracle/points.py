"""TEST INFRASTRUCTURE -- ctypes front end of the two compiled checkers of the point-cloud path:

* `oracle/points.c` (CPU restatement, numpy in / numpy out): nn_distance, nn_distance_grad, approx_match, match_cost,
  match_cost_grad;
* `oracle/_ref/libref_points.so` (the reference's OWN CUDA kernels, built by oracle/build_ref.py from the sources under
  /root/reference; torch CUDA tensors in / out): `ref_*`.

Only tests/, smoke() and the cpu_baseline / comparator legs of the benches may import this.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import build_ref

_f = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_lib = None
_ref = None


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(str(build_ref.build_oracle_c()))
        lib.oracle_nn_distance.argtypes = [C.c_int] * 3 + [_f, _f, _f, _i, _f, _i]
        lib.oracle_nn_distance_grad.argtypes = [C.c_int] * 3 + [_f, _f, _f, _i, _f, _i, _f, _f]
        lib.oracle_approx_match.argtypes = [C.c_int] * 3 + [_f, _f, _f]
        lib.oracle_match_cost.argtypes = [C.c_int] * 3 + [_f, _f, _f, _f]
        lib.oracle_match_cost_grad.argtypes = [C.c_int] * 3 + [_f, _f, _f, _f, _f]
        for fn in ("oracle_nn_distance", "oracle_nn_distance_grad", "oracle_approx_match", "oracle_match_cost",
                   "oracle_match_cost_grad"):
            getattr(lib, fn).restype = None
        _lib = lib
    return _lib


def _c(a, dt=np.float32):
    return np.ascontiguousarray(a, dtype=dt)


def nn_distance(xyz1, xyz2):
    xyz1, xyz2 = _c(xyz1), _c(xyz2)
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    d1, i1 = np.zeros((b, n), np.float32), np.zeros((b, n), np.int32)
    d2, i2 = np.zeros((b, m), np.float32), np.zeros((b, m), np.int32)
    _load().oracle_nn_distance(b, n, m, xyz1, xyz2, d1, i1, d2, i2)
    return d1, i1, d2, i2


def nn_distance_grad(xyz1, xyz2, idx1, idx2, g1, g2):
    xyz1, xyz2 = _c(xyz1), _c(xyz2)
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    o1, o2 = np.zeros((b, n, 3), np.float32), np.zeros((b, m, 3), np.float32)
    _load().oracle_nn_distance_grad(b, n, m, xyz1, xyz2, _c(g1), _c(idx1, np.int32), _c(g2), _c(idx2, np.int32), o1, o2)
    return o1, o2


def approx_match(xyz1, xyz2):
    xyz1, xyz2 = _c(xyz1), _c(xyz2)
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    match = np.zeros((b, m, n), np.float32)
    _load().oracle_approx_match(b, n, m, xyz1, xyz2, match)
    return match


def match_cost(xyz1, xyz2, match):
    xyz1, xyz2 = _c(xyz1), _c(xyz2)
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    out = np.zeros((b,), np.float32)
    _load().oracle_match_cost(b, n, m, xyz1, xyz2, _c(match), out)
    return out


def match_cost_grad(xyz1, xyz2, match):
    xyz1, xyz2 = _c(xyz1), _c(xyz2)
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    g1, g2 = np.zeros((b, n, 3), np.float32), np.zeros((b, m, 3), np.float32)
    _load().oracle_match_cost_grad(b, n, m, xyz1, xyz2, _c(match), g1, g2)
    return g1, g2


# ---------------------------------------------------------------------------------------------------------------------
# the reference's own kernels (GPU box only)
# ---------------------------------------------------------------------------------------------------------------------
def reference_available() -> bool:
    return build_ref.REF_LIB.exists()


def _load_ref():
    global _ref
    if _ref is None:
        path = build_ref.build_reference_points()
        if path is None:
            raise RuntimeError("oracle/_ref/libref_points.so is missing: run `python oracle/build_ref.py` where /root/reference exists")
        import torch  # noqa: F401  (libcudart)
        _ref = C.CDLL(str(path))
    return _ref


def _chk(rc, what):
    if rc != 0:
        raise RuntimeError(f"reference {what} failed (status {rc})")


def ref_nn_distance(xyz1, xyz2):
    import torch
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    d1 = torch.zeros(b, n, device=xyz1.device); i1 = torch.zeros(b, n, dtype=torch.int32, device=xyz1.device)
    d2 = torch.zeros(b, m, device=xyz1.device); i2 = torch.zeros(b, m, dtype=torch.int32, device=xyz1.device)
    st = torch.cuda.current_stream().cuda_stream
    _chk(_load_ref().ref_nndistance(b, n, C.c_void_p(xyz1.data_ptr()), m, C.c_void_p(xyz2.data_ptr()), C.c_void_p(d1.data_ptr()),
                                    C.c_void_p(i1.data_ptr()), C.c_void_p(d2.data_ptr()), C.c_void_p(i2.data_ptr()), C.c_void_p(st)),
         "nndistance")
    return d1, i1, d2, i2


def ref_nn_distance_grad(xyz1, xyz2, idx1, idx2, g1, g2):
    import torch
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    o1 = torch.zeros(b, n, 3, device=xyz1.device); o2 = torch.zeros(b, m, 3, device=xyz1.device)
    torch.cuda.synchronize()            # the reference zero-fills with a synchronous cudaMemset on the legacy stream
    st = torch.cuda.current_stream().cuda_stream
    _chk(_load_ref().ref_nndistancegrad(b, n, C.c_void_p(xyz1.data_ptr()), m, C.c_void_p(xyz2.data_ptr()), C.c_void_p(g1.data_ptr()),
                                        C.c_void_p(idx1.data_ptr()), C.c_void_p(g2.data_ptr()), C.c_void_p(idx2.data_ptr()),
                                        C.c_void_p(o1.data_ptr()), C.c_void_p(o2.data_ptr()), C.c_void_p(st)), "nndistancegrad")
    return o1, o2


def ref_approx_match(xyz1, xyz2):
    import torch
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    match = torch.empty(b, m, n, device=xyz1.device)
    temp = torch.empty(max(b, 32), (n + m) * 2, device=xyz1.device)
    st = torch.cuda.current_stream().cuda_stream
    _chk(_load_ref().ref_approxmatch(b, n, m, C.c_void_p(xyz1.data_ptr()), C.c_void_p(xyz2.data_ptr()), C.c_void_p(match.data_ptr()),
                                     C.c_void_p(temp.data_ptr()), C.c_void_p(st)), "approxmatch")
    return match


def ref_match_cost(xyz1, xyz2, match):
    import torch
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    out = torch.empty(b, device=xyz1.device)
    st = torch.cuda.current_stream().cuda_stream
    _chk(_load_ref().ref_matchcost(b, n, m, C.c_void_p(xyz1.data_ptr()), C.c_void_p(xyz2.data_ptr()), C.c_void_p(match.data_ptr()),
                                   C.c_void_p(out.data_ptr()), C.c_void_p(st)), "matchcost")
    return out


def ref_match_cost_grad(xyz1, xyz2, match):
    import torch
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    g1 = torch.empty(b, n, 3, device=xyz1.device); g2 = torch.empty(b, m, 3, device=xyz1.device)
    st = torch.cuda.current_stream().cuda_stream
    _chk(_load_ref().ref_matchcostgrad(b, n, m, C.c_void_p(xyz1.data_ptr()), C.c_void_p(xyz2.data_ptr()), C.c_void_p(match.data_ptr()),
                                       C.c_void_p(g1.data_ptr()), C.c_void_p(g2.data_ptr()), C.c_void_p(st)), "matchcostgrad")
    return g1, g2
