"""Mirror of `scripts/pytorch_structural_losses/match_cost.py:6-45` (MatchCostFunction, match_cost) as used by
`scripts/compute_mmd_cov_1nn.py:55-62` (emd_approx_cuda: `match_cost(sample, ref) / N`)."""
from torch.autograd import Function

from ... import ops_points


class MatchCostFunction(Function):
    @staticmethod
    def forward(ctx, seta, setb):
        """seta (B, n, 3), setb (B, m, 3) -> cost (B,): approximate earth-mover matching cost."""
        seta = seta.contiguous()
        setb = setb.contiguous()
        ctx.save_for_backward(seta, setb)
        match, _temp = ops_points.approx_match(seta, setb)
        ctx.match = match
        return ops_points.match_cost(seta, setb, match)

    @staticmethod
    def backward(ctx, grad_output):
        seta, setb = ctx.saved_tensors
        grada, gradb = ops_points.match_cost_grad(seta, setb, ctx.match)
        grad_output_expand = grad_output.unsqueeze(1).unsqueeze(2)
        return grada * grad_output_expand, gradb * grad_output_expand


match_cost = MatchCostFunction.apply
