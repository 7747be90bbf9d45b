"""VQ-VAE encode / decode timing at full size (64^3 SDFs, config/vqvae_snet.yaml) for a batch of objects."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from commonscenes_b200 import ops
from commonscenes_b200.model.sdfusion_txt2shape_model import VQ_CONF
from commonscenes_b200.model.model_utils import load_vqvae
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)
vq = load_vqvae(VQ_CONF, device="cuda")
with torch.no_grad():
    vq.quantize.embedding.weight.normal_()
x = (torch.randn(n, 1, 64, 64, 64, device="cuda") * 0.1).clamp(-0.2, 0.2)
z = torch.randn(n, 3, 16, 16, 16, device="cuda")
for name, fn, gf in (("encode_no_quant", lambda: vq.encode_no_quant(x), 271.3), ("decode_no_quant", lambda: vq.decode_no_quant(z), 722.6)):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    n0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{name}: {ms:.2f} ms for {n} objects ({ms / n:.3f} ms/object), {n * gf / ms:.1f} TFLOP/s, {(ops.launch_count() - n0) // 3} launches, out {tuple(out.shape)}")
if "--table" in sys.argv:
    prof = ops.ConvProfiler()
    with prof:
        vq.decode_no_quant(z)
    torch.cuda.synchronize()
    for tag, ms, fl in prof.table():
        print(f"  {ms * 1e3:9.1f} us {fl / ms / 1e9:8.1f} TF/s  {tag}")
