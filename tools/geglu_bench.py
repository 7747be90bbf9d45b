"""GEGLU feed-forward projection: GEMM + separate x * gelu(gate) pass vs the fused ACT_GEGLU epilogue, at the two
transformer widths of the denoiser (batch 64)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from commonscenes_b200 import ops


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for C, grid in ((448, (16, 8, 8)), (672, (16, 4, 4))):
    B = 64
    x = torch.randn(B, *grid, C, device="cuda").to(torch.bfloat16)
    w = torch.randn(8 * C, C, device="cuda") / C ** 0.5
    b = torch.randn(8 * C, device="cuda")
    wp, bp = ops.pack_linear_weight(w), b
    wf, bf = ops.pack_geglu_weight(w, b)
    t_gemm = timed(lambda: ops.linear_tokens(x, wp, bias=bp))
    u = ops.linear_tokens(x, wp, bias=bp)
    t_act = timed(lambda: ops.geglu(u))
    t_fused = timed(lambda: ops.linear_tokens(x, wf, bias=bf, act=ops.ACT_GEGLU))
    a, f = ops.geglu(u).float(), ops.linear_tokens(x, wf, bias=bf, act=ops.ACT_GEGLU).float()
    fl = 2.0 * B * grid[0] * grid[1] * grid[2] * C * 8 * C
    print(f"GEGLU projection {C} -> {8 * C}, batch {B}: GEMM {t_gemm * 1e3:.0f} us + geglu pass {t_act * 1e3:.0f} us = {(t_gemm + t_act) * 1e3:.0f} us "
          f"| fused epilogue {t_fused * 1e3:.0f} us ({fl / t_fused / 1e9:.0f} TFLOP/s); fused vs two-pass rel-L2 {float((a - f).norm() / a.norm()):.2e}")
