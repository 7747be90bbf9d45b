"""One eager guided UNet evaluation ([uncond; cond] = batch 64 for 32 objects, shared conditioning-free prefix: exactly what
bench.py's timed steps replay) between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from commonscenes_b200.model.networks.diffusion_networks.network import DiffusionUNet
from commonscenes_b200.model.sdfusion_txt2shape_model import UNET_PARAMS
objs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
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
    unet(x, t, context_vecs=ca, shared_prefix=True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
unet(x, t, context_vecs=ca, shared_prefix=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
if "--algo-bytes" in sys.argv:      # algorithmic traffic of the implicit-GEMM launches of this step (for tools/dram_summary.py)
    import json
    from commonscenes_b200 import ops
    prof = ops.ConvProfiler()
    with prof:
        unet(x, t, context_vecs=ca, shared_prefix=True)
    torch.cuda.synchronize()
    ms, tflop, n = prof.summary()
    with open(sys.argv[sys.argv.index("--algo-bytes") + 1], "w") as f:
        json.dump({"launches": n, "algorithmic_bytes_per_step": prof.bytes, "tflop_per_step": tflop}, f)
