"""Tagged per-launch timing (CUDA events) of every cs_conv3d launch of one eager UNet evaluation at batch 64."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from commonscenes_b200 import ops
from commonscenes_b200.model.networks.diffusion_networks.network import DiffusionUNet
from commonscenes_b200.model.sdfusion_txt2shape_model import UNET_PARAMS
objs = 32
with torch.device("cuda"):
    m = DiffusionUNet(dict(UNET_PARAMS), conditioning_key="crossattn")
    for p in m.parameters():
        if p.dim() > 1 and float(p.abs().max()) == 0:
            torch.nn.init.normal_(p, std=0.02)
unet = m.eval().diffusion_net
x = torch.randn(objs, 3, 16, 16, 16, device="cuda")
t = torch.full((2 * objs,), 500, dtype=torch.int64, device="cuda")
ca = unet.context_vectors(torch.randn(2 * objs, 1, 1280, device="cuda"))
for _ in range(2):
    unet(x, t, context_vecs=ca)
torch.cuda.synchronize()
prof = ops.ConvProfiler()
with prof:
    unet(x, t, context_vecs=ca)
torch.cuda.synchronize()
tot = 0
for i, (tag, ms, fl) in enumerate(prof.table()):
    tot += ms
    print(f"{i:3d} {ms * 1e3:8.1f} us {fl / ms / 1e9:8.1f} TF/s  {tag}")
print("total", tot)
