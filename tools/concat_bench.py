"""Guided DDIM step of the concat-conditioning denoiser (config/sdfusion-txt2shape_concat.yaml) at
cfg2's batch: 32 objects x CFG = UNet batch 64, through the sampler's CUDA-graph replay + the fused CFG / x_prev kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from commonscenes_b200 import ops
from commonscenes_b200.model.networks.diffusion_networks.network import DiffusionUNet
from commonscenes_b200.model.networks.diffusion_networks.samplers.ddim import DDIMSampler
from commonscenes_b200.model.sdfusion_txt2shape_model import UNET_PARAMS_CONCAT, diffusion_schedule

objs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)
with torch.device("cuda"):
    df = DiffusionUNet(dict(UNET_PARAMS_CONCAT), conditioning_key="concat")
    for p in df.parameters():
        if p.dim() > 1 and float(p.abs().max()) == 0:
            torch.nn.init.normal_(p, std=0.02)
df.eval()
sched = diffusion_schedule()


class Host:
    num_timesteps = 1000
    betas = sched["betas"].cuda()
    alphas_cumprod = sched["alphas_cumprod"].cuda()
Host.df = df
s = DDIMSampler(Host())
s.make_schedule(100, ddim_eta=0.0, verbose=False)
x = torch.randn(objs, 3, 16, 16, 16, device="cuda")
c, uc = torch.randn(objs, 1, 16, 16, 16, device="cuda"), torch.randn(objs, 1, 16, 16, 16, device="cuda")
with torch.no_grad():
    _, concat = s._conditioning(c, uc, True)
    t = torch.full((2 * objs,), 991, dtype=torch.int64, device="cuda")
    for _ in range(3):
        s._step(x, t, None, 99, True, 3.0, concat=concat)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        out, _ = s._step(x, t, None, 99 - i, True, 3.0, concat=concat)
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"concat denoiser, guided DDIM step, {objs} objects (UNet batch {2 * objs}): {ms:.2f} ms -> {1000 / ms:.2f} steps/s "
      f"({s.kernels_per_eval + 1} kernels per step); finite: {bool(torch.isfinite(out).all())}")
