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


@torch.no_grad()
def run_mlp(mlp: nn.Sequential, x: torch.Tensor) -> torch.Tensor:
    """Execute a build_mlp stack on an fp32 (M, C) CUDA matrix."""
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
                                   momentum=0.1 if m.momentum is None else m.momentum, eps=m.eps, relu=relu)
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
