"""DiffusionUNet — conditioning-key dispatch around UNet3DModel.

Drop-in for the reference's model/networks/diffusion_networks/network.py:11-42 (same constructor,
same forward(x, t, c_concat, c_crossattn), same `diffusion_net.` state-dict prefix)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .openai_model_3d import UNet3DModel


class DiffusionUNet(nn.Module):
    def __init__(self, unet_params, vq_conf=None, conditioning_key=None):
        super().__init__()
        self.diffusion_net = UNet3DModel(**unet_params)
        self.conditioning_key = conditioning_key

    def forward(self, x, t, c_concat: list = None, c_crossattn: list = None, **kw):
        if self.conditioning_key is None:
            return self.diffusion_net(x, t, **kw)
        if self.conditioning_key == "crossattn":
            cc = c_crossattn[0] if len(c_crossattn) == 1 else torch.cat(c_crossattn, 1)
            return self.diffusion_net(x, t, context=cc, **kw)
        if self.conditioning_key == "concat":           # reference network.py:25-27
            xc = torch.cat([x.float()] + [c.float() for c in c_concat], dim=1)
            return self.diffusion_net(xc, t, **kw)
        if self.conditioning_key in ("hybrid", "adm"):
            raise NotImplementedError(f"conditioning_key={self.conditioning_key!r} is not used by any reference config")
        raise NotImplementedError()
