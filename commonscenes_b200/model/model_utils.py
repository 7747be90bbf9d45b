"""load_vqvae — build the frozen VQ-VAE from its config and optional checkpoint (API mirror of the
reference's model/model_utils.py:7-32: accepts a raw state dict or a {'vqvae': ...} checkpoint, freezes and
puts the model in eval mode)."""
from __future__ import annotations

import torch

from .networks.vqvae_networks.network import VQVAE


def _get(cfg, *path):
    for p in path:
        cfg = cfg[p] if isinstance(cfg, dict) else getattr(cfg, p)
    return cfg


def load_vqvae(vq_conf, vq_ckpt=None, opt=None, device=None):
    mparam = _get(vq_conf, "model", "params")
    n_embed, embed_dim, ddconfig = _get(mparam, "n_embed"), _get(mparam, "embed_dim"), _get(mparam, "ddconfig")
    ddconfig = {k: (list(v) if isinstance(v, (list, tuple)) or type(v).__name__ == "ListConfig" else v) for k, v in dict(ddconfig).items()}
    vqvae = VQVAE(ddconfig, n_embed, embed_dim)
    if vq_ckpt is not None:
        state_dict = torch.load(vq_ckpt, map_location=lambda storage, loc: storage)
        vqvae.load_state_dict(state_dict["vqvae"] if "vqvae" in state_dict else state_dict)
        print("[*] VQVAE: weight successfully load from: %s" % vq_ckpt)
    if device is None and opt is not None:
        device = _get(opt, "hyper", "device")
    if device is not None:
        vqvae = vqvae.to(device)
    vqvae.eval()
    for param in vqvae.parameters():
        param.requires_grad = False
    return vqvae
