"""GroupNorm-apply (+SiLU) pass at the three spatial levels of the denoiser, batch 64: tanh-form SiLU (default) vs ex2 + rcp
(cs_debug_set bit 21), one-wave launch geometry (default) vs 8 CTAs per SM (bit 22), and a same-size torch copy_ as the practical ceiling.  CUDA events, L2-cold (a 512 MB buffer is written between launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from commonscenes_b200 import _lib, ops
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for (hw, C) in ((16, 224), (8, 448), (4, 672), (16, 448)):
    B, S = 64, 16 * hw * hw
    x = torch.randn(B, 16, hw, hw, C, device="cuda").to(torch.bfloat16)
    stat = ops.zero_stat_buffer(x.device, B, C)
    ops.groupnorm_stats(x, stat)
    ga, be = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
    outs = {}
    for name, flag in (("tanh-form SiLU", 0), ("ex2 + rcp SiLU", 1 << 21), ("8-CTA/SM geometry", 1 << 22)):
        _lib.load().cs_debug_set(flag)
        ts = []
        for _ in range(6):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            y = ops.groupnorm_fused(x, stat, ga, be, groups=32, eps=1e-5, act=ops.ACT_SILU)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        us = 1e3 * sorted(ts)[len(ts) // 2]
        outs[name] = y.float()
        print(f"S={S:5d} C={C:4d} {name:16s}: {us:7.1f} us  {2 * x.numel() * 2 / us / 1e6:6.2f} TB/s (read + write)")
    _lib.load().cs_debug_set(0)
    ts, y2 = [], torch.empty_like(x)
    for _ in range(6):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); y2.copy_(x); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    us = 1e3 * sorted(ts)[len(ts) // 2]
    print(f"S={S:5d} C={C:4d} {'torch copy_ (same bytes)':16s}: {us:7.1f} us  {2 * x.numel() * 2 / us / 1e6:6.2f} TB/s")
    a, b = outs["tanh-form SiLU"], outs["ex2 + rcp SiLU"]
    print(f"        tanh-form vs ex2+rcp: max |diff| {float((a - b).abs().max()):.3e}, rel-L2 {float((a - b).norm() / b.norm()):.3e}")

# GroupNorm(+SiLU) backward (both passes + the (B, C, 2) reduction buffer) at the training batch, L2-cold
from commonscenes_b200 import ops_bwd
for (hw, C) in ((16, 224), (8, 448), (4, 672)):
    B, S = 32, 16 * hw * hw
    x = torch.randn(B, 16, hw, hw, C, device="cuda").to(torch.bfloat16)
    dy = torch.randn_like(x)
    stat = ops.zero_stat_buffer(x.device, B, C)
    ops.groupnorm_stats(x, stat)
    ga, be = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    ts = []
    for _ in range(6):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops_bwd.groupnorm_bwd(x, stat, ga, be, dy, groups=32, eps=1e-5, act=ops.ACT_SILU, dgamma=dg, dbeta=db)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    us = 1e3 * sorted(ts)[len(ts) // 2]
    print(f"backward B={B} S={S:5d} C={C:4d}: {us:7.1f} us  {5 * x.numel() * 2 / us / 1e6:6.2f} TB/s (x, dy read twice + dx written)")
