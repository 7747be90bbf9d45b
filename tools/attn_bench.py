"""Self-attention core timing at the denoiser's level-1 shape (B=64, 8 heads, N=1024, d=56 padded to 64)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from commonscenes_b200 import _lib, ops
B, H, N, d, dp = 64, 8, 1024, 56, 64
qkv = torch.randn(B, N, 3 * H * dp, device="cuda").to(torch.bfloat16)
q, k, v = (qkv[:, :, i * H * dp:(i + 1) * H * dp] for i in range(3))
fl = 4.0 * B * H * N * N * d
for name, flag in (("tcgen05 gen 3 (staggered pipelines, P in TMEM)", 0), ("gen 3, no stagger", 8 << 16),
                   ("tcgen05 two-sweep (gen 2)", 32768), ("tcgen05 gen 1", 16384), ("mma.sync", 128)):
    _lib.load().cs_debug_set(flag)
    for _ in range(3):
        o = ops.attention(q, k, v, heads=H, head_dim=d, head_dim_padded=dp, scale=d ** -0.5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        o = ops.attention(q, k, v, heads=H, head_dim=d, head_dim_padded=dp, scale=d ** -0.5)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name:46s}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s (unpadded FLOPs)")
    if flag == 0:
        o_tc = o.float()
    else:
        print("max |tc - mma.sync| =", float((o_tc - o.float()).abs().max()))
_lib.load().cs_debug_set(0)

# ---- backward (training path): the two transformer shapes of the denoiser at the cfg3 batch (32 objects) ----
from commonscenes_b200 import ops_bwd
for (B, H, N, d, dp) in ((32, 8, 1024, 56, 64), (32, 8, 256, 84, 96)):
    qkv = (torch.randn(B, N, 3 * H * dp, device="cuda") * 0.5).to(torch.bfloat16)
    q, k, v = (qkv[:, :, i * H * dp:(i + 1) * H * dp] for i in range(3))
    o, lse = ops_bwd.attention_lse(q, k, v, heads=H, head_dim=d, head_dim_padded=dp, scale=d ** -0.5)
    do = torch.randn_like(o)
    fl = 10.0 * B * H * N * N * d            # 5 GEMMs of 2 N^2 d (algorithmic; the kernels execute 7)
    for name, flag in (("8 warps/CTA", 0), ("4 warps/CTA", 512)):
        _lib.load().cs_debug_set(flag)
        for _ in range(3):
            ops_bwd.attention_bwd(qkv, o, do, lse, heads=H, head_dim=d, head_dim_padded=dp, scale=d ** -0.5)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops_bwd.attention_bwd(qkv, o, do, lse, heads=H, head_dim=d, head_dim_padded=dp, scale=d ** -0.5)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"attention backward N={N} d={d} B={B} ({name}): {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s algorithmic")
_lib.load().cs_debug_set(0)
