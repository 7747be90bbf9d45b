"""One launch of the N = 1024 self-attention core at the cfg2 batch (B = 64, 8 heads, d = 56 padded to 64) for `ncu --set full`:
    ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 3 -c 1 -o gpurun_out/x python tools/attn_one.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from commonscenes_b200 import ops
B, H, N, d, dp = 64, 8, 1024, 56, 64
qkv = torch.randn(B, N, 3 * H * dp, device="cuda").to(torch.bfloat16)
q, k, v = (qkv[:, :, i * H * dp:(i + 1) * H * dp] for i in range(3))
for _ in range(6):
    ops.attention(q, k, v, heads=H, head_dim=d, head_dim_padded=dp, scale=d ** -0.5)
torch.cuda.synchronize()
