"""Optimizer step of the denoiser alone (clip-norm reduction + AdamW over the plain parameters + cs_adamw_repack over the
packed-gradient weights), CUDA events; `--unfused` times the separate-kernel sequence it replaces (un-pack is not included
there: it happens inside the backward).  Under ncu: `ncu --set full -k regex:adamw_repack -c 1 python tools/optim_bench.py`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from commonscenes_b200.model.sdfusion_txt2shape_model import SDFusionText2ShapeModel, UNET_PARAMS, diffusion_schedule
from commonscenes_b200.model.networks.diffusion_networks.network import DiffusionUNet
from commonscenes_b200.train import DenoiserTrainStep

torch.manual_seed(111)


class Stub:
    q_sample = SDFusionText2ShapeModel.q_sample

    def __init__(self, df):
        self.df, self.num_timesteps, self.device = df, 1000, "cuda"


df = DiffusionUNet(dict(UNET_PARAMS, use_spatial_transformer=True), conditioning_key="crossattn").cuda()
step = DenoiserTrainStep(Stub(df), fused_update="--unfused" not in sys.argv)
print(f"fused: {step.fused}; {len(step.packed_views)} packed-gradient weights, plain region {step.n_plain / 1e6:.1f} M of "
      f"{step.flat_p.numel() / 1e6:.1f} M parameters, gradient buffer {step.flat_g.numel() * 4 / 2**30:.2f} GiB")


def fill():
    step.flat_g.normal_(0, 1e-3)
    if step.fused:                        # pad columns of the packed slots stay zero in real use
        for pv in step.packed_views.values():
            pass


ts = []
for i in range(8):
    fill()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step._clip_and_update(None)
    if not step.fused:
        step.unet._packed = None
        step.trainer._ensure()
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts = sorted(ts[2:])
n = step.flat_p.numel()
print(f"optimizer step{'' if step.fused else ' + stand-alone re-pack'}: {ts[len(ts) // 2]:.2f} ms (median of 6) = "
      f"{36 * n / ts[len(ts) // 2] / 1e9:.2f} TB/s of the fused pass's 36 B per weight")
