"""BoxDiscriminator — the relationship discriminator of the layout branch's adversarial loss (used by the trainer when
--weight_D_box > 0, scripts/train_3dfront.py:230-236, 359-375).

Drop-in for the reference's model/discriminators.py:80-163: same constructor, the same `D` nn.Sequential (hence state-dict
keys D.0 / D.1 / D.3 / D.4 / D.6), same forward(objs, triples, boxes, keeps, with_grad, is_real) -> (probabilities, reg).
The Linear + BatchNorm1d pairs run on the CUDA MLP kernels through layers.mlp_apply (explicit forward and backward behind an
autograd bridge); LeakyReLU / Sigmoid / one-hot / the gradient-penalty arithmetic are elementwise torch glue on T x 512
matrices.  `discriminator_regularizer` keeps the reference's semantics: the input gradient of the logits is obtained with a
backward pass (which, as in the reference, also deposits gradients on D's parameters) and enters the penalty as a constant.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .layers import mlp_apply


def _init_weights(module):
    if hasattr(module, "weight") and isinstance(module, nn.Linear):
        nn.init.kaiming_normal_(module.weight)


def to_one_hot_vector(num_class, label):
    return F.one_hot(label, num_class).float()


def discriminator_regularizer(logits, arg, is_real):
    """Gradient penalty (reference :148-163): (1 - D)^2 |dD/dx|^2 for real pairs, D^2 |dD/dx|^2 for generated ones."""
    logits.backward(torch.ones_like(logits), retain_graph=True)
    grad_norm = torch.norm(arg.grad, dim=1).unsqueeze(1)
    assert grad_norm.shape == logits.shape
    return ((1.0 - logits) if is_real else logits) ** 2 * grad_norm ** 2


class BoxDiscriminator(nn.Module):
    def __init__(self, box_dim, rel_dim, obj_dim, with_obj_labels=True):
        super().__init__()
        self.rel_dim, self.obj_dim, self.with_obj_labels = rel_dim, obj_dim, with_obj_labels
        in_size = box_dim * 2 + rel_dim + (obj_dim * 2 if with_obj_labels else 0)
        self.D = nn.Sequential(nn.Linear(in_size, 512), nn.BatchNorm1d(512), nn.LeakyReLU(),
                               nn.Linear(512, 512), nn.BatchNorm1d(512), nn.LeakyReLU(),
                               nn.Linear(512, 1), nn.Sigmoid())
        self.D.apply(_init_weights)

    def _run_D(self, x):
        h = F.leaky_relu(mlp_apply([self.D[0], self.D[1]], x), self.D[2].negative_slope)
        h = F.leaky_relu(mlp_apply([self.D[3], self.D[4]], h), self.D[5].negative_slope)
        return torch.sigmoid(mlp_apply([self.D[6]], h))

    def forward(self, objs, triples, boxes, keeps=None, with_grad=False, is_real=False):
        s_idx, predicates, o_idx = [t.squeeze(1) for t in triples.chunk(3, dim=1)]
        parts = [to_one_hot_vector(self.rel_dim, predicates), boxes[s_idx], boxes[o_idx]]
        if self.with_obj_labels:
            parts = [to_one_hot_vector(self.obj_dim, objs[s_idx]), to_one_hot_vector(self.obj_dim, objs[o_idx])] + parts
        x = torch.cat(parts, 1).float().contiguous()
        keep_t = None
        if keeps is not None:
            keep_t = ((1 - keeps[s_idx]) + (1 - keeps[o_idx])) > 0
        reg = None
        if with_grad:
            x.requires_grad = True
            y = self._run_D(x)
            reg = discriminator_regularizer(y, x, is_real)
            x.requires_grad = False
        else:
            y = self._run_D(x)
        if keep_t is not None:
            return y[keep_t], (reg[keep_t] if reg is not None else None)
        return y, reg
