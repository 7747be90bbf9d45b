"""Torch-tensor front end of the point-cloud entry points of the C ABI (include/cs_b200.h: cs_nn_distance,
cs_nn_distance_grad, cs_approx_match, cs_match_cost, cs_match_cost_grad).

Plumbing only: the checks are the CHECK_INPUT of the reference's own binding
(scripts/pytorch_structural_losses/src/structural_loss.cpp:10-12 -- CUDA + contiguous), outputs are allocated with torch
exactly as that binding does (:21-37, :39-52, :54-70, :81-101, :103-125) and raw pointers go to libcsb200.so on the
current stream.  There is no CPU path.
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import _lib
from ._lib import check

__all__ = ["nn_distance", "nn_distance_grad", "approx_match", "match_cost", "match_cost_grad"]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _points(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.CsError(f"{name} must be a CUDA tensor (commonscenes_b200 has no CPU path)")
    if t.dtype != torch.float32 or t.dim() != 3 or t.shape[2] != 3:
        raise _lib.CsError(f"{name}: expected a (batch, points, 3) fp32 tensor, got {t.dtype} {tuple(t.shape)}")
    if not t.is_contiguous():
        raise _lib.CsError(f"{name} must be contiguous")
    return t


def _pair(a: torch.Tensor, b: torch.Tensor, who: str) -> Tuple[int, int, int]:
    _points(a, f"{who}: set_d")
    _points(b, f"{who}: set_q")
    if a.shape[0] != b.shape[0] or a.device != b.device:
        raise _lib.CsError(f"{who}: both point sets need the same batch size and device")
    return a.shape[0], a.shape[1], b.shape[1]


def nn_distance(xyz1: torch.Tensor, xyz2: torch.Tensor):
    """-> dist1 (b, n) fp32, idx1 (b, n) int32, dist2 (b, m), idx2 (b, m): squared distance to / index of the nearest point of
    the other set."""
    b, n, m = _pair(xyz1, xyz2, "nn_distance")
    dev = xyz1.device
    dist1 = torch.empty(b, n, dtype=torch.float32, device=dev)
    idx1 = torch.empty(b, n, dtype=torch.int32, device=dev)
    dist2 = torch.empty(b, m, dtype=torch.float32, device=dev)
    idx2 = torch.empty(b, m, dtype=torch.int32, device=dev)
    check(_lib.load().cs_nn_distance(xyz1.data_ptr(), xyz2.data_ptr(), b, n, m, dist1.data_ptr(), idx1.data_ptr(),
                                     dist2.data_ptr(), idx2.data_ptr(), _stream()), "cs_nn_distance")
    return dist1, idx1, dist2, idx2


def nn_distance_grad(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2):
    b, n, m = _pair(xyz1, xyz2, "nn_distance_grad")
    for t, shape, dt, name in ((idx1, (b, n), torch.int32, "idx1"), (idx2, (b, m), torch.int32, "idx2"),
                               (grad_dist1, (b, n), torch.float32, "grad_dist1"), (grad_dist2, (b, m), torch.float32, "grad_dist2")):
        if not t.is_cuda or t.dtype != dt or tuple(t.shape) != shape or not t.is_contiguous():
            raise _lib.CsError(f"nn_distance_grad: {name} must be a contiguous CUDA {dt} tensor of shape {shape}")
    g1 = torch.empty(b, n, 3, dtype=torch.float32, device=xyz1.device)
    g2 = torch.empty(b, m, 3, dtype=torch.float32, device=xyz1.device)
    check(_lib.load().cs_nn_distance_grad(xyz1.data_ptr(), xyz2.data_ptr(), b, n, m, grad_dist1.data_ptr(), idx1.data_ptr(),
                                          grad_dist2.data_ptr(), idx2.data_ptr(), g1.data_ptr(), g2.data_ptr(), _stream()),
          "cs_nn_distance_grad")
    return g1, g2


def approx_match(xyz1: torch.Tensor, xyz2: torch.Tensor):
    """-> match (b, m, n) fp32 and the (b, 2 (n + m)) scratch tensor the reference binding returns next to it."""
    b, n, m = _pair(xyz1, xyz2, "approx_match")
    match = torch.empty(b, m, n, dtype=torch.float32, device=xyz1.device)
    temp = torch.empty(b, (n + m) * 2, dtype=torch.float32, device=xyz1.device)
    check(_lib.load().cs_approx_match(xyz1.data_ptr(), xyz2.data_ptr(), b, n, m, match.data_ptr(), temp.data_ptr(), _stream()),
          "cs_approx_match")
    return match, temp


def _check_match(match, b, n, m, who):
    if not match.is_cuda or match.dtype != torch.float32 or tuple(match.shape) != (b, m, n) or not match.is_contiguous():
        raise _lib.CsError(f"{who}: match must be a contiguous fp32 CUDA tensor of shape {(b, m, n)}")


def match_cost(xyz1, xyz2, match) -> torch.Tensor:
    b, n, m = _pair(xyz1, xyz2, "match_cost")
    _check_match(match, b, n, m, "match_cost")
    out = torch.empty(b, dtype=torch.float32, device=xyz1.device)
    check(_lib.load().cs_match_cost(xyz1.data_ptr(), xyz2.data_ptr(), match.data_ptr(), b, n, m, out.data_ptr(), _stream()),
          "cs_match_cost")
    return out


def match_cost_grad(xyz1, xyz2, match):
    b, n, m = _pair(xyz1, xyz2, "match_cost_grad")
    _check_match(match, b, n, m, "match_cost_grad")
    g1 = torch.empty(b, n, 3, dtype=torch.float32, device=xyz1.device)
    g2 = torch.empty(b, m, 3, dtype=torch.float32, device=xyz1.device)
    check(_lib.load().cs_match_cost_grad(xyz1.data_ptr(), xyz2.data_ptr(), match.data_ptr(), b, n, m, g1.data_ptr(),
                                         g2.data_ptr(), _stream()), "cs_match_cost_grad")
    return g1, g2
