"""Where the two-sweep attention kernel spends its time: cs_debug_set bits 16-18 switch parts of it off (results are then
wrong): +1 no exponentials, +2 no P stores, +4 no sweep 1."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from commonscenes_b200 import _lib, ops
B, H, N, d, dp = 64, 8, 1024, 56, 64
qkv = torch.randn(B, N, 3 * H * dp, device="cuda").to(torch.bfloat16)
q, k, v = (qkv[:, :, i * H * dp:(i + 1) * H * dp] for i in range(3))
for name, flag in (("full kernel", 0), ("no exponentials", 1), ("no P stores", 2), ("no sweep 1", 4), ("no exp, no P stores", 3),
                   ("no exp, no stores, no sweep 1", 7)):
    _lib.load().cs_debug_set(flag << 16)
    for _ in range(3):
        ops.attention(q, k, v, heads=H, head_dim=d, head_dim_padded=dp, scale=d ** -0.5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.attention(q, k, v, heads=H, head_dim=d, head_dim_padded=dp, scale=d ** -0.5)
    e1.record(); torch.cuda.synchronize()
    print(f"{name:32s}: {e0.elapsed_time(e1) / 10:.3f} ms")
_lib.load().cs_debug_set(0)
