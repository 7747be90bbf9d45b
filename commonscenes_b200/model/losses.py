"""Layout losses of the training loop — same call signatures as the reference's model/losses.py (bce_loss :5-24,
calculate_model_losses :26-51, add_loss :54-60), so `from model.losses import calculate_model_losses, bce_loss` in
scripts/train_3dfront.py can point here.  Tiny reductions over (objects x 6) / (objects x 24) tensors: plain torch glue on
the device; the networks that produce their inputs run on the CUDA kernels (VAEGAN_V2FULL.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def bce_loss(input, target, reduce=True):
    """Binary cross-entropy on logits, numerically stable: max(x, 0) - x t + log(1 + exp(-|x|))."""
    loss = F.binary_cross_entropy_with_logits(input, target, reduction="none")
    return loss.mean() if reduce else loss


def add_loss(total_loss, curr_loss, loss_dict, loss_name, weight=1):
    weighted = curr_loss * weight
    loss_dict[loss_name] = weighted.item()
    return weighted if total_loss is None else total_loss + weighted


def calculate_model_losses(args, pred, target, name, angles=None, angles_pred=None, mu=None, logvar=None, KL_weight=None, writer=None,
                           counter=None, withangles=False):
    """total = L1(pred, target) [+ NLL(angles_pred, angles)] + KL_weight * KL(N(mu, e^logvar) || N(0, I)) / objects.
    Returns (total, {name, 'angle_pred', 'KLD_Gauss'} -> weighted python floats); scalars go to `writer` when given."""
    losses = {}
    rec = F.l1_loss(pred, target)
    total = add_loss(0.0, rec, losses, name, 1)
    ang = None
    if withangles:
        ang = F.nll_loss(angles_pred, angles)
        total = add_loss(total, ang, losses, "angle_pred", 1)
    kld = -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp()) / mu.size(0)
    total = add_loss(total, kld, losses, "KLD_Gauss", KL_weight)
    if writer is not None:
        writer.add_scalar("Train_Loss_KL_{}".format(name), kld, counter)
        writer.add_scalar("Train_Loss_Rec_{}".format(name), rec, counter)
        if withangles:
            writer.add_scalar("Train_Loss_Angle_{}".format(name), ang, counter)
    return total, losses
