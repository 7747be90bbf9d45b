"""Mirror of the reference's `extension/dist_chamfer.py` (chamferFunction :12-46, chamferDist :48-53) over
`cs_nn_distance` / `cs_nn_distance_grad` of libcsb200.so instead of the `chamfer` torch extension
(extension/chamfer_cuda.cpp:17-32, extension/chamfer.cu).

    import commonscenes_b200.extension.dist_chamfer as ext      # scripts/eval_3dfront.py:24-25
    chamfer = ext.chamferDist()
    dist1, dist2 = chamfer(points_a, points_b)                  # (B, n, 3), (B, m, 3) CUDA fp32 -> (B, n), (B, m)

GPU tensors only, as the reference.  Values (distances, indices) are bit-equal to the reference kernels'.
"""
import torch
from torch import nn
from torch.autograd import Function

from .. import ops_points


class chamferFunction(Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        xyz1 = xyz1.contiguous()
        xyz2 = xyz2.contiguous()
        dist1, idx1, dist2, idx2 = ops_points.nn_distance(xyz1, xyz2)
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        ctx.mark_non_differentiable(idx1, idx2)
        return dist1, dist2

    @staticmethod
    def backward(ctx, graddist1, graddist2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        gradxyz1, gradxyz2 = ops_points.nn_distance_grad(xyz1, xyz2, idx1, idx2, graddist1.contiguous(), graddist2.contiguous())
        return gradxyz1, gradxyz2


class chamferDist(nn.Module):
    def __init__(self):
        super(chamferDist, self).__init__()

    def forward(self, input1, input2):
        return chamferFunction.apply(input1, input2)
