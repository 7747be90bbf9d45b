"""Gradient kernels vs PyTorch autograd (fp32) on the same bf16-rounded operands (floating-point kernels: torch fp32 is
the reference, SURVEY.md §8c suggests cosine >= 0.999 for gradients; we assert rel-L2 <= 1e-2)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


def _cl(t):   # NCDHW fp32 -> channels-last bf16
    return t.permute(0, 2, 3, 4, 1).contiguous().to(torch.bfloat16)


CASES = [
    # B, C1, C2, Cout, (D, H, W), ksize, stride
    (2, 64, 0, 64, (8, 8, 8), 3, (1, 1, 1)),
    (2, 224, 0, 448, (16, 8, 8), 3, (1, 1, 1)),
    (2, 448, 224, 224, (4, 8, 8), 3, (1, 1, 1)),
    (2, 224, 0, 224, (8, 16, 16), 3, (1, 2, 2)),
    (2, 448, 0, 3584, (4, 8, 8), 1, (1, 1, 1)),
    (3, 96, 32, 40, (2, 4, 4), 3, (1, 1, 1)),
    (1, 672, 672, 672, (16, 4, 4), 3, (1, 1, 1)),
]


@pytest.mark.parametrize("B,C1,C2,Cout,grid,k,stride", CASES)
def test_conv3d_wgrad_and_dgrad(B, C1, C2, Cout, grid, k, stride):
    from commonscenes_b200 import ops, ops_bwd
    torch.manual_seed(0)
    dev = "cuda"
    D, H, W = grid
    pad = (k // 2,) * 3
    x = torch.randn(B, C1 + C2, D, H, W, device=dev).to(torch.bfloat16).float()
    w = (torch.randn(Cout, C1 + C2, k, k, k, device=dev) * 0.05).to(torch.bfloat16).float().requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    y = F.conv3d(xr, w, stride=stride, padding=pad)
    dy = torch.randn_like(y).to(torch.bfloat16).float()
    gx, gw = torch.autograd.grad(y, (xr, w), dy)

    x_cl = _cl(x)
    x1, x2 = (x_cl[..., :C1].contiguous(), x_cl[..., C1:].contiguous()) if C2 else (x_cl, None)
    dy_cl = _cl(dy)
    dw = torch.zeros(Cout, k ** 3, ops._pad64(C1) + ops._pad64(C2), device=dev)
    ops_bwd.conv3d_wgrad(x1, dy_cl, dw, ksize=(k,) * 3, stride=stride, pad=pad, x2=x2)
    got = ops_bwd.unpack_wgrad(dw, w.shape, (C1, C2) if C2 else None)
    assert _rel(got, gw) < 2e-3
    # accumulation semantics: a second call doubles the gradient
    ops_bwd.conv3d_wgrad(x1, dy_cl, dw, ksize=(k,) * 3, stride=stride, pad=pad, x2=x2)
    assert _rel(ops_bwd.unpack_wgrad(dw, w.shape, (C1, C2) if C2 else None), 2 * gw) < 2e-3

    if stride == (1, 1, 1):
        wd = ops_bwd.pack_dgrad_weight(w)
        dx = ops_bwd.conv3d_dgrad(dy_cl, wd, ksize=(k,) * 3, pad=pad)
        assert _rel(dx.permute(0, 4, 1, 2, 3), gx) < 6e-3     # bf16 output rounding
