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
    # the training batch of BASELINE cfg3 (32 objects): the dgrad takes the pair / quad igemm variants, the wgrad its full K split
    (32, 224, 0, 224, (16, 16, 16), 3, (1, 1, 1)),
    (32, 448, 0, 448, (16, 8, 8), 3, (1, 1, 1)),
    (32, 448, 224, 224, (16, 16, 16), 3, (1, 1, 1)),
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


def _ncdhw(t):   # channels-last -> NCDHW fp32
    return t.permute(0, 4, 1, 2, 3).float()


@pytest.mark.parametrize("C1,C2,act", [(64, 0, 1), (224, 0, 1), (448, 224, 1), (96, 32, 0), (672, 0, 0)])
def test_groupnorm_bwd(C1, C2, act):
    from commonscenes_b200 import ops, ops_bwd
    torch.manual_seed(1)
    dev = "cuda"
    B, D, H, W = 3, 4, 4, 8
    Ct = C1 + C2
    x = (torch.randn(B, Ct, D, H, W, device=dev) * 1.5 + 0.3).to(torch.bfloat16).float()
    gamma = (1 + 0.2 * torch.randn(Ct, device=dev))
    beta = 0.1 * torch.randn(Ct, device=dev)
    xr, gr, br = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = F.group_norm(xr, 32, gr, br, eps=1e-5)
    if act:
        y = F.silu(y)
    dy = torch.randn_like(y).to(torch.bfloat16).float()
    ex = torch.randn_like(x).to(torch.bfloat16).float()
    gx, gg, gb = torch.autograd.grad(y, (xr, gr, br), dy)

    x_cl = _cl(x)
    x1 = x_cl[..., :C1].contiguous()
    x2 = x_cl[..., C1:].contiguous() if C2 else None
    s1 = ops.groupnorm_stats(x1, torch.zeros(B, C1, 2, dtype=ops.STAT_DTYPE, device=dev))
    s2 = ops.groupnorm_stats(x2, torch.zeros(B, C2, 2, dtype=ops.STAT_DTYPE, device=dev)) if C2 else None
    ex_cl = _cl(ex)
    e1 = ex_cl[..., :C1].contiguous()
    e2 = ex_cl[..., C1:].contiguous() if C2 else None
    dg, db = torch.zeros(Ct, device=dev), torch.zeros(Ct, device=dev)
    dx1, dx2 = ops_bwd.groupnorm_bwd(x1, s1, gamma, beta, _cl(dy), act=act, x2=x2, stat2=s2, extra=e1, extra2=e2,
                                     dgamma=dg, dbeta=db)
    got = torch.cat([_ncdhw(dx1)] + ([_ncdhw(dx2)] if C2 else []), dim=1)
    assert _rel(got, gx + ex) < 6e-3
    assert _rel(dg, gg) < 2e-3 and _rel(db, gb) < 2e-3


@pytest.mark.parametrize("C", [64, 448, 672])
def test_layernorm_bwd(C):
    from commonscenes_b200 import ops_bwd
    torch.manual_seed(2)
    dev = "cuda"
    x = (torch.randn(2, 4, 4, 8, C, device=dev) * 2 + 0.5).to(torch.bfloat16)
    gamma, beta = 1 + 0.2 * torch.randn(C, device=dev), 0.1 * torch.randn(C, device=dev)
    xr, gr, br = x.float().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = F.layer_norm(xr, (C,), gr, br)
    dy = torch.randn_like(y).to(torch.bfloat16)
    ex = torch.randn_like(y).to(torch.bfloat16)
    gx, gg, gb = torch.autograd.grad(y, (xr, gr, br), dy.float())
    dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    dx = ops_bwd.layernorm_bwd(x, gamma, dy, dg, db, extra=ex)
    assert _rel(dx, gx + ex.float()) < 6e-3
    assert _rel(dg, gg) < 2e-3 and _rel(db, gb) < 2e-3


def test_geglu_upsample_zero_insert_add():
    from commonscenes_b200 import ops, ops_bwd
    torch.manual_seed(3)
    dev = "cuda"
    u = torch.randn(2, 2, 4, 4, 256, device=dev).to(torch.bfloat16)
    ur = u.float().requires_grad_(True)
    a, g = ur.chunk(2, dim=-1)
    f = a * F.gelu(g)
    df = torch.randn_like(f).to(torch.bfloat16)
    (gu,) = torch.autograd.grad(f, ur, df.float())
    assert _rel(ops_bwd.geglu_bwd(u, df), gu) < 6e-3

    x = torch.randn(2, 4, 4, 4, 64, device=dev).to(torch.bfloat16)
    xr = x.float().requires_grad_(True)
    up = xr.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    dy = torch.randn_like(up).to(torch.bfloat16)
    (gx,) = torch.autograd.grad(up, xr, dy.float())
    assert _rel(ops_bwd.upsample_nearest_bwd(dy, (1, 2, 2)), gx) < 6e-3

    z = ops_bwd.zero_insert(x, (1, 2, 2))
    ref = torch.zeros(2, 4, 8, 8, 64, device=dev, dtype=torch.bfloat16)
    ref[:, :, ::2, ::2] = x
    assert torch.equal(z, ref)
    y = x.clone()
    ops_bwd.add_(y, x)
    assert torch.equal(y, (x.float() * 2).to(torch.bfloat16))


def test_strided_conv_dgrad():
    from commonscenes_b200 import ops_bwd
    torch.manual_seed(4)
    dev = "cuda"
    x = torch.randn(2, 64, 4, 8, 8, device=dev).to(torch.bfloat16).float().requires_grad_(True)
    w = (torch.randn(96, 64, 3, 3, 3, device=dev) * 0.05).to(torch.bfloat16).float()
    y = F.conv3d(x, w, stride=(1, 2, 2), padding=1)
    dy = torch.randn_like(y).to(torch.bfloat16).float()
    (gx,) = torch.autograd.grad(y, x, dy)
    dx = ops_bwd.conv3d_dgrad_strided(_cl(dy), ops_bwd.pack_dgrad_weight(w), (1, 2, 2))
    assert _rel(_ncdhw(dx), gx) < 6e-3


@pytest.mark.parametrize("N,heads,d", [(256, 4, 84), (1024, 2, 56), (64, 2, 16), (200, 2, 56)])
def test_attention_bwd(N, heads, d):
    from commonscenes_b200 import ops_bwd
    from commonscenes_b200.model.networks.diffusion_networks.attention import _pad_head_dim
    torch.manual_seed(5)
    dev = "cuda"
    B = 2
    dp = _pad_head_dim(d)
    scale = d ** -0.5
    qkv_real = torch.randn(B, N, 3, heads, d, device=dev).to(torch.bfloat16)
    qkv = torch.zeros(B, N, 3, heads, dp, device=dev, dtype=torch.bfloat16)
    qkv[..., :d] = qkv_real
    qkv = qkv.view(B, N, 3 * heads * dp)
    hd = heads * dp
    q, k, v = (qkv[:, :, i * hd:(i + 1) * hd] for i in range(3))
    o, lse = ops_bwd.attention_lse(q, k, v, heads=heads, head_dim=d, head_dim_padded=dp, scale=scale)

    qr = qkv_real.float().requires_grad_(True)
    qq, kk, vv = (qr[:, :, i].permute(0, 2, 1, 3) for i in range(3))      # (B, h, N, d)
    p = torch.softmax(qq @ kk.transpose(-1, -2) * scale, dim=-1)
    oref = (p @ vv).permute(0, 2, 1, 3).reshape(B, N, heads * d)
    assert _rel(o, oref) < 1e-2
    lse_ref = torch.logsumexp(qq @ kk.transpose(-1, -2) * scale, dim=-1) * 1.4426950408889634
    assert (lse - lse_ref).abs().max().item() < 2e-2
    do = torch.randn_like(oref).to(torch.bfloat16)
    (g,) = torch.autograd.grad(oref, qr, do.float())
    dqkv = ops_bwd.attention_bwd(qkv, o, do, lse, heads=heads, head_dim=d, head_dim_padded=dp, scale=scale)
    got = dqkv.view(B, N, 3, heads, dp)
    assert _rel(got[..., :d], g) < 2e-2
    assert got[..., d:].float().abs().max().item() == 0.0


def test_sgemm_mse_adamw_unpack():
    from commonscenes_b200 import ops_bwd
    torch.manual_seed(6)
    dev = "cuda"
    a, b = torch.randn(37, 70, device=dev), torch.randn(70, 45, device=dev)
    assert _rel(ops_bwd.sgemm(a, b), a @ b) < 1e-5
    assert _rel(ops_bwd.sgemm(a.t().contiguous(), b, trans_a=True), a @ b) < 1e-5
    assert _rel(ops_bwd.sgemm(a, b.t().contiguous(), trans_b=True), a @ b) < 1e-5
    c = torch.randn(37, 45, device=dev)
    pre = torch.randn(37, 45, device=dev)
    pr = pre.clone().requires_grad_(True)
    (sg,) = torch.autograd.grad(F.silu(pr), pr, torch.ones_like(pr))
    want = c + (a @ b) * sg
    assert _rel(ops_bwd.sgemm(a, b, out=c.clone(), accumulate=True, silu_pre=pre), want) < 1e-5

    pred, tgt = torch.randn(4, 3, 8, 8, 8, device=dev), torch.randn(4, 3, 8, 8, 8, device=dev)
    loss = torch.zeros((), device=dev)
    g = ops_bwd.mse_loss_grad(pred, tgt, loss, loss_scale=100.0)
    pr = pred.clone().requires_grad_(True)
    ref = F.mse_loss(pr, tgt)
    (gr,) = torch.autograd.grad(100.0 * ref, pr)
    assert abs(loss.item() - ref.item()) < 1e-5 and _rel(g, gr) < 1e-5

    n = 100003
    p0, g0 = torch.randn(n, device=dev), torch.randn(n, device=dev) * 3
    pt = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([pt], lr=1e-3, weight_decay=0.01)
    p, m, v = p0.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    for step in (1, 2, 3):
        gs = g0 * step
        pt.grad = gs.clone()
        torch.nn.utils.clip_grad_norm_([pt], 5.0)
        opt.step()
        ss = ops_bwd.sumsq(gs, torch.zeros((), device=dev))
        assert abs(ss.item() - (gs.double() ** 2).sum().item()) / ss.item() < 1e-4
        ops_bwd.adamw_step(p, gs, m, v, lr=1e-3, weight_decay=0.01, step=step, sumsq_buf=ss, max_norm=5.0)
    assert _rel(p, pt.detach()) < 1e-5

    dw = torch.randn(10, 27, 128 + 64, device=dev)
    grad = torch.zeros(10, 96 + 40, 3, 3, 3, device=dev)
    ops_bwd.unpack_wgrad_into(dw, grad, (96, 40))
    assert torch.equal(grad, ops_bwd.unpack_wgrad(dw, grad.shape, (96, 40)).contiguous())
