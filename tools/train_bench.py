"""Training-step timing of the denoiser at cfg3 (32 objects per GPU): forward_train, backward, optimizer,
weight re-packing.  `--b N` changes the batch.  Prints ms per phase and train-steps/s."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from commonscenes_b200 import ops, ops_bwd
from commonscenes_b200.model.sdfusion_txt2shape_model import SDFusionText2ShapeModel, UNET_PARAMS, diffusion_schedule
from commonscenes_b200.model.networks.diffusion_networks.network import DiffusionUNet
from commonscenes_b200.train import DenoiserTrainStep

B = int(sys.argv[sys.argv.index("--b") + 1]) if "--b" in sys.argv else 32
torch.manual_seed(111)


class Stub:
    q_sample = SDFusionText2ShapeModel.q_sample

    def __init__(self, df):
        self.df, self.num_timesteps, self.device = df, 1000, "cuda"
        for k, v in diffusion_schedule(1000, 0.00085, 0.012).items():
            setattr(self, k, v.cuda())


params = dict(UNET_PARAMS, use_spatial_transformer=True)
df = DiffusionUNet(params, conditioning_key="crossattn").cuda()
with torch.no_grad():
    for n, p in df.named_parameters():      # the reference zero-initialises 18 convs; use non-zero weights like the goldens
        if p.dim() > 1 and float(p.abs().max()) == 0.0:
            p.normal_(0, 0.02)
m = Stub(df)
step = DenoiserTrainStep(m)
z = torch.randn(B, 3, 16, 16, 16, device="cuda")
ctx = torch.randn(B, 1, 1280, device="cuda")


def timed(fn, reps=3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r


for _ in range(2):
    loss, _ = step.step(z, ctx)
print("loss after warm-up:", loss.item())
if "--ncu" in sys.argv:      # one eager training step between cudaProfilerStart/Stop (ncu --profile-from-start off)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step.step(z, ctx)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)
ms_eager, _ = timed(lambda: step.step(z, ctx), reps=3)
print(f"eager train step B={B}: {ms_eager:.1f} ms")
step.capture(B, 1280)
for _ in range(2):
    loss, _ = step.step_graphed(z, ctx)
ms_step, _ = timed(lambda: step.step_graphed(z, ctx), reps=5)
print("loss (graphed):", loss.item())
print(f"graphed train step B={B}: {ms_step:.1f} ms -> {1000 / ms_step:.2f} steps/s, {B * 1000 / ms_step:.1f} objects/s "
      f"({3 * B * 557.6 / ms_step:.0f} TFLOP/s algorithmic, fwd+bwd = 3 x 557.6 GF/sample)")
tr = step.trainer
t = torch.randint(0, 1000, (B,), device="cuda")
noise = torch.randn_like(z)
x_t = m.q_sample(z, t, noise)
ms_pack, _ = timed(lambda: (setattr(step.unet, "_packed", None), tr._ensure()))
ms_fwd, (eps, tape) = timed(lambda: tr.forward_train(x_t, t, ctx))
lbuf = torch.zeros((), device="cuda")
d_eps = ops_bwd.mse_loss_grad(eps, noise, lbuf, loss_scale=100.0)
ops.reset_launch_count()
ms_bwd, _ = timed(lambda: tr.backward(tr.forward_train(x_t, t, ctx)[1], d_eps, need_dcontext=False))
launches = ops.launch_count() / 3
ms_opt, _ = timed(lambda: step._clip_and_update(None))      # sumsq + AdamW (+ the fused re-pack of the packed-gradient weights)
print(f"re-pack (fwd + dgrad layouts): {ms_pack:.1f} ms | forward_train {ms_fwd:.1f} ms | fwd+backward {ms_bwd:.1f} ms "
      f"(backward ~{ms_bwd - ms_fwd:.1f} ms, {launches:.0f} launches fwd+bwd) | clip + optimizer (+ fused re-pack) {ms_opt:.2f} ms")
print(f"memory: allocated {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB peak")
if "--profile" in sys.argv:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step.step_graphed(z, ctx)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=60))
