"""Scratch timing of one guided denoising step (UNet at batch 2*objects + DDIM update) on the current GPU."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from commonscenes_b200 import ops
from commonscenes_b200.model.networks.diffusion_networks.network import DiffusionUNet
from commonscenes_b200.model.sdfusion_txt2shape_model import UNET_PARAMS

objs = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 32
if "--debug" in sys.argv:      # cs_debug_set switches (tuning experiments)
    from commonscenes_b200 import _lib as _l
    _l.load().cs_debug_set(int(sys.argv[sys.argv.index("--debug") + 1]))
cfg = UNET_PARAMS
torch.manual_seed(0)
with torch.device("cuda"):
    m = DiffusionUNet(dict(cfg, use_spatial_transformer=True, legacy=False), conditioning_key="crossattn")
    for p in m.parameters():
        if p.dim() > 1 and float(p.abs().max()) == 0:
            torch.nn.init.normal_(p, std=0.02)
m.eval()
unet = m.diffusion_net
B = 2 * objs
x = torch.randn(objs, 3, 16, 16, 16, device="cuda")
t = torch.full((B,), 500, dtype=torch.int64, device="cuda")
ctx = torch.randn(B, 1, 1280, device="cuda")
ca = unet.context_vectors(ctx)
for _ in range(2):
    eps = unet(x, t, context_vecs=ca)
torch.cuda.synchronize()
n0 = ops.launch_count()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(3):
    eps = unet(x, t, context_vecs=ca)
ev[1].record(); torch.cuda.synchronize()
print(f"eager: {ev[0].elapsed_time(ev[1]) / 3:.2f} ms per UNet(B={B}); launches/eval {(ops.launch_count() - n0) // 3}")
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    eps = unet(x, t, context_vecs=ca)
g.replay(); torch.cuda.synchronize()
ev[0].record()
for _ in range(5):
    g.replay()
ev[1].record(); torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 5
print(f"graph: {ms:.2f} ms per UNet(B={B}) -> {1000 / ms:.2f} denoising-steps/s; {B * 557.6e9 / ms / 1e9:.1f} TFLOP/s")
print("finite:", bool(torch.isfinite(eps).all()), "absmean", float(eps.abs().mean()))

# box speed index: sustained cuBLAS bf16 GEMM on the same box right after the run (boxes differ by +-20 %)
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16); b2 = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
for _ in range(10): a @ b2
torch.cuda.synchronize(); ev[0].record()
for _ in range(100): a @ b2
ev[1].record(); torch.cuda.synchronize()
cub = 100 * 2 * 8192 ** 3 / ev[0].elapsed_time(ev[1]) / 1e9
print(f"box index: cuBLAS bf16 8192^3 sustained {cub:.0f} TFLOP/s -> whole-step fraction {B * 557.6e9 / ms / 1e9 / cub:.3f}")
