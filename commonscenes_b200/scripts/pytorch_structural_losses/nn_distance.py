"""Mirror of `scripts/pytorch_structural_losses/nn_distance.py:6-42` (NNDistanceFunction, nn_distance) as used by
`scripts/compute_mmd_cov_1nn.py:25-28` (distChamferCUDA)."""
from torch.autograd import Function

from ... import ops_points


class NNDistanceFunction(Function):
    @staticmethod
    def forward(ctx, seta, setb):
        """seta (B, n, 3), setb (B, m, 3) -> dist1 (B, n), dist2 (B, m): squared nearest-neighbour distances."""
        seta = seta.contiguous()
        setb = setb.contiguous()
        ctx.save_for_backward(seta, setb)
        dist1, idx1, dist2, idx2 = ops_points.nn_distance(seta, setb)
        ctx.idx1 = idx1
        ctx.idx2 = idx2
        return dist1, dist2

    @staticmethod
    def backward(ctx, grad_dist1, grad_dist2):
        seta, setb = ctx.saved_tensors
        grada, gradb = ops_points.nn_distance_grad(seta, setb, ctx.idx1, ctx.idx2, grad_dist1.contiguous(),
                                                   grad_dist2.contiguous())
        return grada, gradb


nn_distance = NNDistanceFunction.apply
