"""GraphTripleConv / GraphTripleConvNet(2) — scene-graph convolutions producing the denoiser's conditioning.

Drop-in for the reference's model/graph.py (make_mlp :27-28, _init_weights :31-34, GraphTripleConv :89-211,
GraphTripleConvNet :214-249, GraphTripleConvNet2 :252-288): same constructors, child names and
forward(obj_vecs, pred_vecs, edges) -> (new_obj_vecs, new_pred_vecs).  pooling is 'avg' or 'sum'
('wAvg' is not used by v2_full, model/VAE.py:57-63).  Each layer is: gather kernel -> net1 (GEMV + BatchNorm/ReLU
kernels) -> deterministic scatter-mean kernel -> net2 -> two residual projections.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .layers import build_mlp, linear_backward, mlp_backward, run_mlp, run_mlp_train


def make_mlp(dim_list, activation="relu", batch_norm="none", dropout=0, norelu=False):
    return build_mlp(dim_list, activation, batch_norm, dropout, final_nonlinearity=(not norelu))


def _init_weights(module):
    if hasattr(module, "weight") and isinstance(module, nn.Linear):
        nn.init.kaiming_normal_(module.weight)


class GraphTripleConv(nn.Module):
    def __init__(self, input_dim_obj, input_dim_pred, output_dim=None, hidden_dim=512, pooling="avg",
                 mlp_normalization="none", residual=True):
        super().__init__()
        if output_dim is None:
            output_dim = input_dim_obj
        self.input_dim_obj, self.input_dim_pred = input_dim_obj, input_dim_pred
        self.output_dim, self.hidden_dim, self.residual = output_dim, hidden_dim, residual
        assert pooling in ["sum", "avg", "wAvg"], 'Invalid pooling "%s"' % pooling
        if pooling != "avg":
            raise NotImplementedError("only pooling='avg' (the v2_full setting) is built")
        self.pooling = pooling
        self.net1 = build_mlp([2 * input_dim_obj + input_dim_pred, hidden_dim, 2 * hidden_dim + output_dim], batch_norm=mlp_normalization)
        self.net1.apply(_init_weights)
        self.net2 = build_mlp([hidden_dim, hidden_dim, output_dim], batch_norm=mlp_normalization)
        self.net2.apply(_init_weights)
        if self.residual:
            self.linear_projection = nn.Linear(input_dim_obj, output_dim)
            self.linear_projection_pred = nn.Linear(input_dim_pred, output_dim)

    @torch.no_grad()
    def forward(self, obj_vecs, pred_vecs, edges):
        H, Dout = self.hidden_dim, self.output_dim
        obj_vecs, pred_vecs = obj_vecs.float().contiguous(), pred_vecs.float().contiguous()
        edges = edges.to(torch.int64).contiguous()
        num_objs = obj_vecs.size(0)
        new_t = run_mlp(self.net1, ops.gcn_gather_triples(obj_vecs, pred_vecs, edges))      # (T, 2H + Dout) = [s | p | o]
        pooled = ops.gcn_scatter_mean(new_t, 0, H + Dout, H, edges, num_objs)
        new_obj = run_mlp(self.net2, pooled)
        new_p = new_t[:, H:H + Dout]
        if self.residual:
            f = lambda t: t.detach().float().contiguous()
            new_obj = ops.add_rows(new_obj, ops.linear_small(obj_vecs, f(self.linear_projection.weight), f(self.linear_projection.bias)))
            new_p = ops.add_rows(new_p, ops.linear_small(pred_vecs, f(self.linear_projection_pred.weight), f(self.linear_projection_pred.bias)))
        else:
            new_p = new_p.contiguous()
        return new_obj, new_p

    # ---- training path: forward keeping the activations + explicit backward (the reference relies on autograd) ----
    @torch.no_grad()
    def forward_train(self, obj_vecs, pred_vecs, edges):
        """forward() that also returns the tape for backward()."""
        H, Dout = self.hidden_dim, self.output_dim
        obj_vecs, pred_vecs = obj_vecs.float().contiguous(), pred_vecs.float().contiguous()
        edges = edges.to(torch.int64).contiguous()
        num_objs = obj_vecs.size(0)
        new_t, tape1 = run_mlp_train(self.net1, ops.gcn_gather_triples(obj_vecs, pred_vecs, edges))
        pooled = ops.gcn_scatter_mean(new_t, 0, H + Dout, H, edges, num_objs)
        new_obj, tape2 = run_mlp_train(self.net2, pooled)
        new_p = new_t[:, H:H + Dout]
        if self.residual:
            f = lambda t: t.detach().float().contiguous()
            new_obj = ops.add_rows(new_obj, ops.linear_small(obj_vecs, f(self.linear_projection.weight), f(self.linear_projection.bias)))
            new_p = ops.add_rows(new_p, ops.linear_small(pred_vecs, f(self.linear_projection_pred.weight), f(self.linear_projection_pred.bias)))
        else:
            new_p = new_p.contiguous()
        return new_obj, new_p, (obj_vecs, pred_vecs, edges, tape1, tape2)

    @torch.no_grad()
    def backward(self, tape, d_new_obj, d_new_p, sink):
        """(d_obj_vecs, d_pred_vecs) from the gradients of the two outputs; parameter gradients accumulate into `sink`."""
        from .. import ops_bwd
        obj_vecs, pred_vecs, edges, tape1, tape2 = tape
        H, Do, Dp = self.hidden_dim, self.input_dim_obj, self.input_dim_pred
        d_new_obj = d_new_obj.float().contiguous()
        d_new_p = None if d_new_p is None else d_new_p.float().contiguous()      # None: the predicate output is unused
        d_obj = linear_backward(self.linear_projection, obj_vecs, d_new_obj, sink) if self.residual else torch.zeros_like(obj_vecs)
        if self.residual and d_new_p is not None:
            d_pred = linear_backward(self.linear_projection_pred, pred_vecs, d_new_p, sink)
        else:
            d_pred = torch.zeros_like(pred_vecs)
        d_pooled = mlp_backward(tape2, d_new_obj, sink)
        d_new_t = ops_bwd.gcn_scatter_mean_bwd(d_pooled, edges, H, d_new_p, mid_w=self.output_dim)
        d_in = mlp_backward(tape1, d_new_t, sink)
        ops_bwd.gcn_gather_triples_bwd(d_in, edges, Do, Dp, d_obj, d_pred, accumulate=True)
        return d_obj, d_pred


class GraphTripleConvNet(nn.Module):
    """A sequence of scene graph convolution layers."""

    def __init__(self, input_dim_obj, input_dim_pred, num_layers=2, hidden_dim=512, residual=False, pooling="avg",
                 mlp_normalization="none", output_dim=None):
        super().__init__()
        self.num_layers = num_layers
        self.gconvs = nn.ModuleList()
        kw = dict(input_dim_obj=input_dim_obj, input_dim_pred=input_dim_pred, hidden_dim=hidden_dim, pooling=pooling,
                  residual=residual, mlp_normalization=mlp_normalization)
        for i in range(self.num_layers):
            last = output_dim is not None and i >= self.num_layers - 1
            self.gconvs.append(GraphTripleConv(output_dim=output_dim, **kw) if last else GraphTripleConv(**kw))

    def forward(self, obj_vecs, pred_vecs, edges):
        for gconv in self.gconvs:
            obj_vecs, pred_vecs = gconv(obj_vecs, pred_vecs, edges)
        return obj_vecs, pred_vecs

    def forward_train(self, obj_vecs, pred_vecs, edges):
        tapes = []
        for gconv in self.gconvs:
            obj_vecs, pred_vecs, tape = gconv.forward_train(obj_vecs, pred_vecs, edges)
            tapes.append(tape)
        return obj_vecs, pred_vecs, tapes

    def backward(self, tapes, d_obj, d_pred, sink):
        for gconv, tape in zip(reversed(self.gconvs), reversed(tapes)):
            d_obj, d_pred = gconv.backward(tape, d_obj, d_pred, sink)
        return d_obj, d_pred


class _GcnNetFunction(torch.autograd.Function):
    """Autograd bridge for a whole GraphTripleConvNet: forward = forward_train, backward = the explicit backward kernels."""

    @staticmethod
    def forward(ctx, net, obj_vecs, pred_vecs, edges, *params):
        o, p, tapes = net.forward_train(obj_vecs.detach(), pred_vecs.detach(), edges)
        ctx.net, ctx.tapes, ctx.params = net, tapes, params
        ctx.need = (obj_vecs.requires_grad, pred_vecs.requires_grad)
        ctx.set_materialize_grads(False)
        return o, p

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_o, d_p):
        from .networks.diffusion_networks.unet_train import GradSink
        sink = GradSink()
        if d_o is None:       # only the predicate output was used downstream
            last = ctx.net.gconvs[-1]
            d_o = torch.zeros((ctx.tapes[-1][0].shape[0], last.output_dim), dtype=torch.float32, device=d_p.device)
        d_obj, d_pred = ctx.net.backward(ctx.tapes, d_o, d_p, sink)
        return (None, d_obj if ctx.need[0] else None, d_pred if ctx.need[1] else None, None) + tuple(sink.grads.get(p) for p in ctx.params)


def gcn_apply(net: GraphTripleConvNet, obj_vecs, pred_vecs, edges):
    """net(obj_vecs, pred_vecs, edges) that takes part in autograd (see layers.mlp_apply)."""
    params = [p for p in net.parameters() if p.requires_grad]
    if torch.is_grad_enabled() and (obj_vecs.requires_grad or pred_vecs.requires_grad or params):
        return _GcnNetFunction.apply(net, obj_vecs.float().contiguous(), pred_vecs.float().contiguous(), edges, *params)
    return net(obj_vecs, pred_vecs, edges)


class GraphTripleConvNet2(GraphTripleConvNet):
    """Identical stack under the name the relation encoder E2 uses (model/graph.py:252-288)."""
