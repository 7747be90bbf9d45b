"""Per-shape timing of the implicit-GEMM kernel on the conv / linear shapes of one UNet evaluation at batch 64
(SURVEY.md Appendix A.1/A.2).  Prints TFLOP/s per shape; `--only i` runs a single shape (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from commonscenes_b200 import ops

B = 64
SHAPES = [  # (Cin, Cout, (D,H,W), k, stride, count per UNet eval)
    (224, 224, (16, 16, 16), 3, (1, 1, 1), 7), (448, 448, (16, 8, 8), 3, (1, 1, 1), 6), (672, 672, (16, 4, 4), 3, (1, 1, 1), 10),
    (448, 448, (16, 16, 16), 3, (1, 1, 1), 1), (448, 224, (16, 16, 16), 3, (1, 1, 1), 2), (672, 224, (16, 16, 16), 3, (1, 1, 1), 1),
    (1120, 448, (16, 8, 8), 3, (1, 1, 1), 1), (1344, 672, (16, 4, 4), 3, (1, 1, 1), 2), (672, 672, (16, 8, 8), 3, (1, 1, 1), 1),
    (896, 448, (16, 8, 8), 3, (1, 1, 1), 1), (672, 448, (16, 8, 8), 3, (1, 1, 1), 1), (1120, 672, (16, 4, 4), 3, (1, 1, 1), 1),
    (224, 448, (16, 8, 8), 3, (1, 1, 1), 1), (448, 672, (16, 4, 4), 3, (1, 1, 1), 1),
    (224, 224, (16, 16, 16), 3, (1, 2, 2), 1), (448, 448, (16, 8, 8), 3, (1, 2, 2), 1),
    (448, 3584, (16, 8, 8), 1, (1, 1, 1), 5), (448, 1536, (16, 8, 8), 1, (1, 1, 1), 5), (448, 448, (16, 8, 8), 1, (1, 1, 1), 15),
    (1792, 448, (16, 8, 8), 1, (1, 1, 1), 5), (672, 5376, (16, 4, 4), 1, (1, 1, 1), 6), (672, 2304, (16, 4, 4), 1, (1, 1, 1), 6),
    (672, 672, (16, 4, 4), 1, (1, 1, 1), 18), (2688, 672, (16, 4, 4), 1, (1, 1, 1), 6),
]
only = int(sys.argv[sys.argv.index("--only") + 1]) if "--only" in sys.argv else None
if "--debug" in sys.argv:
    from commonscenes_b200 import _lib
    _lib.load().cs_debug_set(int(sys.argv[sys.argv.index("--debug") + 1]))
first = int(sys.argv[sys.argv.index("--from") + 1]) if "--from" in sys.argv else 0
reps = 1 if only is not None else 5
mode = "wgrad" if "--wgrad" in sys.argv else ("dgrad" if "--dgrad" in sys.argv else "fwd")
if "--b" in sys.argv:
    B = int(sys.argv[sys.argv.index("--b") + 1])
if mode != "fwd":
    from commonscenes_b200 import ops_bwd
tot_ms = tot_fl = 0.0
for i, (ci, co, (D, H, W), k, st, cnt) in enumerate(SHAPES):
    if (only is not None and i != only) or i < first:
        continue
    x = torch.randn(B, D, H, W, ci, device="cuda").to(torch.bfloat16)
    w = ops.pack_conv_weight(torch.randn(co, ci, k, k, k, device="cuda") / (ci * k ** 3) ** 0.5)
    b = torch.randn(co, device="cuda")
    pad = (k // 2,) * 3
    y = ops.conv3d(x, w, ksize=(k, k, k), stride=st, pad=pad, bias=b)
    if mode == "wgrad":
        dw = torch.zeros(co, k ** 3, ops._pad64(ci), device="cuda")
        run = lambda: ops_bwd.conv3d_wgrad(x, y, dw, ksize=(k, k, k), stride=st, pad=pad)
    elif mode == "dgrad":
        if st != (1, 1, 1):
            continue
        wd = ops_bwd.pack_dgrad_weight(torch.randn(co, ci, k, k, k, device="cuda") / (ci * k ** 3) ** 0.5)
        run = lambda: ops_bwd.conv3d_dgrad(y, wd, ksize=(k, k, k), pad=pad)
    else:
        run = lambda: ops.conv3d(x, w, ksize=(k, k, k), stride=st, pad=pad, bias=b)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * y.shape[0] * y.shape[1] * y.shape[2] * y.shape[3] * co * ci * k ** 3
    tot_ms += ms * cnt; tot_fl += fl * cnt
    print(f"[{i:2d}] {ci:5d}->{co:5d} k{k} s{st} @{D}x{H}x{W}: {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s  x{cnt}  ({ms * cnt:7.2f} ms/eval)")
if only is None:
    print(f"sum over one UNet eval: {tot_ms:.2f} ms, {tot_fl / tot_ms / 1e9:.1f} TFLOP/s")

a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16); b2 = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
for _ in range(10): a @ b2
torch.cuda.synchronize(); e0.record()
for _ in range(50): a @ b2
e1.record(); torch.cuda.synchronize()
print(f"box index: cuBLAS bf16 8192^3 {50 * 2 * 8192 ** 3 / e0.elapsed_time(e1) / 1e9:.0f} TFLOP/s")
