"""build_mlp — Linear [+ BatchNorm1d + ReLU] stacks of the scene-graph networks, on the GPU kernels.

Same signature and nn.Sequential layout (hence state-dict keys) as the reference's model/layers.py:21-38;
`run_mlp` executes such a stack through cs_linear_small / cs_batchnorm_relu (fp32: the graphs have tens of rows)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops


def build_mlp(dim_list, activation="relu", batch_norm="none", dropout=0, final_nonlinearity=True):
    layers = []
    for i in range(len(dim_list) - 1):
        dim_in, dim_out = dim_list[i], dim_list[i + 1]
        layers.append(nn.Linear(dim_in, dim_out))
        final_layer = (i == len(dim_list) - 2)
        if not final_layer or final_nonlinearity:
            if batch_norm == "batch":
                layers.append(nn.BatchNorm1d(dim_out))
            if activation == "relu":
                layers.append(nn.ReLU())
            elif activation == "leakyrelu":
                layers.append(nn.LeakyReLU())
        if dropout > 0:
            layers.append(nn.Dropout(p=dropout))
    return nn.Sequential(*layers)


def _bn_momentum(m: nn.BatchNorm1d) -> float:
    """torch semantics: momentum=None is a cumulative moving average, factor 1 / num_batches_tracked (already bumped)."""
    if m.momentum is not None:
        return float(m.momentum)
    n = int(m.num_batches_tracked) if m.num_batches_tracked is not None else 0
    return 1.0 / max(n, 1)


@torch.no_grad()
def run_mlp(mlp, x: torch.Tensor) -> torch.Tensor:
    """Execute a build_mlp stack (nn.Sequential or list of its layers) on an fp32 (M, C) CUDA matrix."""
    mods = list(mlp)
    i = 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, nn.Linear):
            x = ops.linear_small(x, m.weight.detach().float().contiguous(),
                                 None if m.bias is None else m.bias.detach().float().contiguous())
            i += 1
        elif isinstance(m, nn.BatchNorm1d):
            relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
            training = m.training or not m.track_running_stats
            if training and m.track_running_stats:
                m.num_batches_tracked += 1
            x = ops.batchnorm_relu(x, m.weight, m.bias, m.running_mean, m.running_var, training,
                                   momentum=_bn_momentum(m), eps=m.eps, relu=relu)
            i += 2 if relu else 1
        elif isinstance(m, nn.ReLU):
            x = ops.batchnorm_relu(x, None, None, torch.zeros(x.shape[1], device=x.device), torch.ones(x.shape[1], device=x.device),
                                   False, eps=0.0, relu=True)   # plain ReLU (batch_norm='none' stacks)
            i += 1
        elif isinstance(m, nn.Dropout):
            if m.training and m.p > 0:
                raise NotImplementedError("dropout > 0 is not used by the v2_full configuration")
            i += 1
        else:
            raise NotImplementedError(f"build_mlp layer {type(m).__name__}")
    return x


# ----------------------------------------------------------------------------------------------------------------------
# training path: the same stacks with the activations kept, and their explicit backward (what the reference's autograd
# does for `loss.backward()` through rel_mlp / the GraphTripleConv nets, train_3dfront.py:387-391)
# ----------------------------------------------------------------------------------------------------------------------
_ONES = {}


def _ones_row(m: int, device) -> torch.Tensor:
    key = (m, str(device))
    t = _ONES.get(key)
    if t is None:
        t = _ONES[key] = torch.ones((1, m), dtype=torch.float32, device=device)
    return t


def linear_backward(lin: nn.Linear, x: torch.Tensor, dy: torch.Tensor, sink, need_dx: bool = True):
    """Gradients of y = x W^T + b: dW += dy^T x, db += column sums of dy (both into `sink`), returns dx = dy W (or None)."""
    from .. import ops_bwd
    if lin.weight.requires_grad:
        ops_bwd.sgemm(dy, x, trans_a=True, out=sink.grad(lin.weight), accumulate=True)
    if lin.bias is not None and lin.bias.requires_grad:
        ops_bwd.sgemm(_ones_row(dy.shape[0], dy.device), dy, out=sink.grad(lin.bias).view(1, -1), accumulate=True)
    return ops_bwd.sgemm(dy, lin.weight.detach().float().contiguous()) if need_dx else None


@torch.no_grad()
def run_mlp_train(mlp, x: torch.Tensor):
    """run_mlp that also returns the tape [(module, input, output, training)] its backward needs (mlp: an nn.Sequential of
    build_mlp, or a plain list of such layers)."""
    mods = list(mlp)
    tape = []
    i = 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, nn.Linear):
            y = ops.linear_small(x, m.weight.detach().float().contiguous(),
                                 None if m.bias is None else m.bias.detach().float().contiguous())
            tape.append(("linear", m, x, None, False))
            i += 1
        elif isinstance(m, nn.BatchNorm1d):
            relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
            training = m.training or not m.track_running_stats
            if training and m.track_running_stats:
                m.num_batches_tracked += 1
            y = ops.batchnorm_relu(x, m.weight, m.bias, m.running_mean, m.running_var, training,
                                   momentum=_bn_momentum(m), eps=m.eps, relu=relu)
            tape.append(("bn_relu" if relu else "bn", m, x, y, training))
            i += 2 if relu else 1
        elif isinstance(m, nn.ReLU):
            y = ops.batchnorm_relu(x, None, None, torch.zeros(x.shape[1], device=x.device), torch.ones(x.shape[1], device=x.device),
                                   False, eps=0.0, relu=True)
            tape.append(("relu", m, x, y, False))
            i += 1
        elif isinstance(m, nn.Dropout):
            if m.training and m.p > 0:
                raise NotImplementedError("dropout > 0 is not used by the v2_full configuration")
            y = x
            i += 1
        else:
            raise NotImplementedError(f"build_mlp layer {type(m).__name__}")
        x = y
    return x, tape


@torch.no_grad()
def mlp_backward(tape, dy: torch.Tensor, sink, need_dx: bool = True):
    """Backward through a run_mlp_train tape; parameter gradients accumulate into `sink`; returns dx (or None)."""
    from .. import ops_bwd
    for k, (kind, m, x, y, training) in enumerate(reversed(tape)):
        first = k == len(tape) - 1
        if kind == "linear":
            dy = linear_backward(m, x, dy, sink, need_dx=need_dx or not first)
        elif kind in ("bn", "bn_relu"):
            dg = sink.grad(m.weight) if m.affine and m.weight.requires_grad else None
            db = sink.grad(m.bias) if m.affine and m.bias.requires_grad else None
            dy = ops_bwd.batchnorm_relu_bwd(x, y, dy, m.weight.detach() if m.affine else None, m.running_mean, m.running_var,
                                            training, eps=m.eps, relu=(kind == "bn_relu"), dgamma=dg, dbeta=db)
        else:   # plain ReLU
            dy = ops_bwd.batchnorm_relu_bwd(x, y, dy, None, torch.zeros(x.shape[1], device=x.device),
                                            torch.ones(x.shape[1], device=x.device), False, eps=0.0, relu=True)
    return dy


class _MlpFunction(torch.autograd.Function):
    """Autograd bridge for a build_mlp stack: forward = run_mlp_train, backward = mlp_backward (explicit kernels)."""

    @staticmethod
    def forward(ctx, mlp, x, *params):
        y, tape = run_mlp_train(mlp, x.detach().float().contiguous())
        ctx.tape, ctx.params, ctx.need_dx = tape, params, x.requires_grad
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        from .networks.diffusion_networks.unet_train import GradSink
        sink = GradSink()
        dx = mlp_backward(ctx.tape, dy.float().contiguous(), sink, need_dx=ctx.need_dx)     # (tape kept: retain_graph re-runs work)
        return (None, dx if ctx.need_dx else None) + tuple(sink.grads.get(p) for p in ctx.params)


def mlp_apply(mlp, x: torch.Tensor) -> torch.Tensor:
    """run_mlp that takes part in autograd: with gradients enabled and something to differentiate, the result carries a
    grad_fn whose backward runs the explicit MLP gradient kernels (what `loss.backward()` does in the reference)."""
    mods = list(mlp)
    params = [p for m in mods for p in m.parameters() if p.requires_grad]
    if torch.is_grad_enabled() and (x.requires_grad or params):
        return _MlpFunction.apply(mods, x, *params)
    return run_mlp(mods, x)

