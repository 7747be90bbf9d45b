"""Per-kernel numerics on a real B200: every C-ABI entry point against a plain PyTorch fp32 reference
of the same op (inputs pre-rounded to bf16 where the kernel consumes bf16, TF32 off).

Tolerances (stated per test): GEMM-class kernels accumulate in fp32 and round the result to bf16 once,
so |err| <= 2^-8 * |ref| + small absolute slack; fp32 kernels are compared at 1e-5.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from commonscenes_b200 import _lib, ops as o
    _lib.require_device()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return o


def _bf(x):
    return x.to(torch.bfloat16)


def _cl(x):  # NCDHW fp32 -> channels-last bf16
    return _bf(x.permute(0, 2, 3, 4, 1).contiguous())


def _ncdhw(x):  # channels-last bf16 -> NCDHW fp32
    return x.float().permute(0, 4, 1, 2, 3).contiguous()


def _report(name, got, ref, rtol, atol):
    err = (got - ref).abs()
    bound = atol + rtol * ref.abs()
    bad = (err > bound)
    rel_l2 = (got - ref).norm() / ref.norm().clamp_min(1e-12)
    msg = (f"{name}: max_abs_err={err.max().item():.4e} rel_l2={rel_l2.item():.4e} "
           f"violations={int(bad.sum())}/{bad.numel()} ref_absmax={ref.abs().max().item():.3e}")
    print(msg)
    assert not bad.any(), msg


CONV_CASES = [
    # name, B, Cin, Cout, (D,H,W), ksize, stride, pad
    ("res224_16^3", 2, 224, 224, (16, 16, 16), 3, (1, 1, 1), 1),
    ("res448_16x8x8", 2, 448, 448, (16, 8, 8), 3, (1, 1, 1), 1),
    ("res672_16x4x4", 4, 672, 672, (16, 4, 4), 3, (1, 1, 1), 1),
    ("in4_224to448", 2, 224, 448, (16, 8, 8), 3, (1, 1, 1), 1),
    ("lin448to3584", 2, 448, 3584, (16, 8, 8), 1, (1, 1, 1), 0),
    ("lin1792to448", 2, 1792, 448, (16, 8, 8), 1, (1, 1, 1), 0),
    ("vq64_32^3", 1, 64, 128, (32, 32, 32), 3, (1, 1, 1), 1),
    ("small_c32", 2, 32, 32, (16, 4, 4), 3, (1, 1, 1), 1),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv3d_matches_torch(ops, case):
    name, B, Cin, Cout, (D, H, W), k, stride, pad = case
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.randn(B, Cin, D, H, W, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, k, k, k, device="cuda", generator=g) / math.sqrt(Cin * k ** 3)
    b = torch.randn(Cout, device="cuda", generator=g)
    xb, wb = _bf(x).float(), _bf(w).float()
    ref = F.conv3d(xb, wb, b, stride=stride, padding=pad)
    got = ops.conv3d(_cl(x), ops.pack_conv_weight(w), ksize=(k, k, k), stride=stride, pad=(pad,) * 3, bias=b)
    torch.cuda.synchronize()
    _report(name, _ncdhw(got), ref, rtol=2 ** -7, atol=2e-3)


# The benchmarked batch (32 objects x CFG = 64): the 16^3-level convs take the CTA-pair / two-accumulator ("quad") kernel, the
# deeper levels the pair / hybrid work lists -- variants the B <= 4 cases above never select (cs_igemm.cu: igemm_launch).
CONV_CASES_B64 = [
    # name, Cin (C1, C2), Cout, (D, H, W), expected variant
    ("b64_224to224_16^3", (224, 0), 224, (16, 16, 16), "CTA pairs x two accumulators"),
    ("b64_448to448_16^3", (448, 0), 448, (16, 16, 16), "CTA pairs x two accumulators"),
    ("b64_448+224to224_16^3", (448, 224), 224, (16, 16, 16), "CTA pairs x two accumulators"),
    ("b64_672to224_16^3", (672, 0), 224, (16, 16, 16), "CTA pairs x two accumulators"),
    ("b64_448to448_16x8x8", (448, 0), 448, (16, 8, 8), "pair / hybrid work list"),
    ("b64_672to672_16x4x4", (672, 0), 672, (16, 4, 4), "pair / hybrid work list"),
]


@pytest.mark.parametrize("case", CONV_CASES_B64, ids=[c[0] for c in CONV_CASES_B64])
def test_conv3d_benchmark_batch_matches_torch(ops, case):
    name, (C1, C2), Cout, (D, H, W), want = case
    B = 64
    g = torch.Generator(device="cuda").manual_seed(4321)
    x = torch.randn(B, C1 + C2, D, H, W, device="cuda", generator=g)
    w = torch.randn(Cout, C1 + C2, 3, 3, 3, device="cuda", generator=g) / math.sqrt((C1 + C2) * 27)
    b = torch.randn(Cout, device="cuda", generator=g)
    ref = F.conv3d(_bf(x).float(), _bf(w).float(), b, padding=1)
    stat = torch.zeros(B, Cout, 2, dtype=ops.STAT_DTYPE, device="cuda")
    ops.conv3d_variant_counts(reset=True)
    if C2:
        got = ops.conv3d(_cl(x[:, :C1]), ops.pack_conv_weight(w, split=(C1, C2)), x2=_cl(x[:, C1:]), bias=b, stat_sum=stat)
    else:
        got = ops.conv3d(_cl(x), ops.pack_conv_weight(w), bias=b, stat_sum=stat)
    torch.cuda.synchronize()
    ran = [k for k, v in ops.conv3d_variant_counts(reset=True).items() if v]
    print(f"{name}: kernel variant {ran}")
    _report(name, _ncdhw(got), ref, rtol=2 ** -7, atol=2e-3)
    gotf = got.float()
    s_ref = torch.stack([gotf.sum(dim=(1, 2, 3)), (gotf * gotf).sum(dim=(1, 2, 3))], dim=-1)     # sums of the fp32 values
    _report(name + "_stats", ops.stat_to_float(stat).float(), s_ref, rtol=2e-3, atol=1.0)   # that were rounded to bf16
    if want is not None:
        assert ran == [want], f"{name}: expected the '{want}' variant at batch 64, got {ran}"


@pytest.mark.parametrize("case", [("unet_448_1x2x2", 4, 448, 448, (16, 8, 8), (1, 2, 2)), ("unet_672_1x2x2", 4, 672, 672, (16, 4, 4), (1, 2, 2)),
                                  ("vq_256_2x2x2", 1, 256, 256, (16, 16, 16), (2, 2, 2)), ("b64_448_1x2x2", 64, 448, 448, (16, 8, 8), (1, 2, 2))],
                         ids=lambda c: c[0])
def test_upsample_conv_as_phase_convs_matches_torch(ops, case):
    """nearest-upsample + 3x3x3 conv (Upsample: openai_model_3d.py:150-158, vqvae_modules.py:33-40) evaluated as merged-tap
    phase convs on the low-resolution tensor vs F.interpolate + F.conv3d in fp32; fused GroupNorm sums included."""
    name, B, Cin, Cout, (D, H, W), f = case
    g = torch.Generator(device="cuda").manual_seed(99)
    x = torch.randn(B, Cin, D, H, W, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, 3, 3, 3, device="cuda", generator=g) / math.sqrt(Cin * 27)
    b = torch.randn(Cout, device="cuda", generator=g)
    ref = F.conv3d(F.interpolate(_bf(x).float(), scale_factor=tuple(float(v) for v in f), mode="nearest"), _bf(w).float(), b, padding=1)
    out = torch.full((B, D * f[0], H * f[1], W * f[2], Cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    stat = torch.zeros(B, Cout, 2, dtype=ops.STAT_DTYPE, device="cuda")
    phases = ops.pack_upsample_phase_weights(w, f)
    assert len(phases) == f[0] * f[1] * f[2]
    xc = _cl(x)
    for offs, ks, pad, pad_back, wp in phases:
        ops.conv3d(xc, wp, ksize=ks, pad=pad, pad_back=pad_back, bias=b, stat_sum=stat, out=out, phase=(f, offs))
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all(), "every output voxel must be written by exactly one phase"
    # the merged taps are rounded to bf16 AFTER summation while the fp32 comparison rounds every tap separately: the two filter
    # roundings differ by up to 2^-8 per weight, so the element-wise bound carries an absolute term that scales with
    # sqrt(K) * |w| * |x| (measured max 2.2e-2 at |y| ~ 7); the relative L2 error is the sharp criterion (measured 2.2e-3)
    _report(name, _ncdhw(out), ref, rtol=2 ** -6, atol=3e-2)
    rel = float((_ncdhw(out) - ref).norm() / ref.norm())
    print(f"{name}: rel-L2 {rel:.3e}")
    assert rel < 4e-3, f"{name}: rel-L2 {rel}"
    of = out.float()
    s_ref = torch.stack([of.sum(dim=(1, 2, 3)), (of * of).sum(dim=(1, 2, 3))], dim=-1)
    _report(name + "_stats", ops.stat_to_float(stat).float(), s_ref, rtol=2e-3, atol=1.0)


def test_conv3d_epilogue_rowvec_residual_stats(ops):
    B, C, D, H, W = 2, 224, 16, 8, 8
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(B, C, D, H, W, device="cuda", generator=g)
    w = torch.randn(C, C, 3, 3, 3, device="cuda", generator=g) / math.sqrt(C * 27)
    b = torch.randn(C, device="cuda", generator=g)
    rv = torch.randn(B, C, device="cuda", generator=g)
    res = torch.randn(B, C, D, H, W, device="cuda", generator=g)
    ref = F.conv3d(_bf(x).float(), _bf(w).float(), b, padding=1) + rv[:, :, None, None, None] + _bf(res).float()
    stat = torch.zeros(B, C, 2, dtype=ops.STAT_DTYPE, device="cuda")
    got = ops.conv3d(_cl(x), ops.pack_conv_weight(w), bias=b, rowvec=rv, residual=_cl(res), stat_sum=stat)
    stat_again = torch.zeros_like(stat)
    ops.conv3d(_cl(x), ops.pack_conv_weight(w), bias=b, rowvec=rv, residual=_cl(res), stat_sum=stat_again)
    torch.cuda.synchronize()
    _report("epilogue", _ncdhw(got), ref, rtol=2 ** -7, atol=4e-3)
    s_ref = torch.stack([ref.sum(dim=(2, 3, 4)), (ref * ref).sum(dim=(2, 3, 4))], dim=-1)
    _report("fused_stats", ops.stat_to_float(stat).float(), s_ref, rtol=1e-3, atol=0.5)
    assert torch.equal(stat, stat_again), "fixed-point GroupNorm sums must be bit-identical from launch to launch"


def test_conv3d_two_sources_equals_concat(ops):
    B, C1, C2, Cout, D, H, W = 2, 448, 224, 224, 16, 16, 16
    g = torch.Generator(device="cuda").manual_seed(11)
    x1 = torch.randn(B, C1, D, H, W, device="cuda", generator=g)
    x2 = torch.randn(B, C2, D, H, W, device="cuda", generator=g)
    w = torch.randn(Cout, C1 + C2, 3, 3, 3, device="cuda", generator=g) / math.sqrt((C1 + C2) * 27)
    ref = F.conv3d(_bf(torch.cat([x1, x2], 1)).float(), _bf(w).float(), None, padding=1)
    got = ops.conv3d(_cl(x1), ops.pack_conv_weight(w, split=(C1, C2)), x2=_cl(x2))
    torch.cuda.synchronize()
    _report("two_source", _ncdhw(got), ref, rtol=2 ** -7, atol=2e-3)


@pytest.mark.parametrize("shape", [(224, (16, 16, 16)), (448, (16, 8, 8))], ids=["224", "448"])
def test_conv3d_stride_1_2_2(ops, shape):
    C, (D, H, W) = shape
    B = 2
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(B, C, D, H, W, device="cuda", generator=g)
    w = torch.randn(C, C, 3, 3, 3, device="cuda", generator=g) / math.sqrt(C * 27)
    b = torch.randn(C, device="cuda", generator=g)
    ref = F.conv3d(_bf(x).float(), _bf(w).float(), b, stride=(1, 2, 2), padding=1)
    got = ops.conv3d(_cl(x), ops.pack_conv_weight(w), stride=(1, 2, 2), bias=b)
    torch.cuda.synchronize()
    _report("stride122", _ncdhw(got), ref, rtol=2 ** -7, atol=2e-3)


def test_conv3d_stride2_asymmetric_pad(ops):
    # VQ-VAE Downsample: pad (0,1) per dim then 3x3x3 stride 2 (vqvae_modules.py:54-58)
    B, C, R = 1, 64, 32
    g = torch.Generator(device="cuda").manual_seed(6)
    x = torch.randn(B, C, R, R, R, device="cuda", generator=g)
    w = torch.randn(C, C, 3, 3, 3, device="cuda", generator=g) / math.sqrt(C * 27)
    ref = F.conv3d(F.pad(_bf(x).float(), (0, 1, 0, 1, 0, 1)), _bf(w).float(), None, stride=2)
    got = ops.conv3d(_cl(x), ops.pack_conv_weight(w), stride=(2, 2, 2), pad=(0, 0, 0), pad_back=(1, 1, 1))
    torch.cuda.synchronize()
    _report("stride2_asym", _ncdhw(got), ref, rtol=2 ** -7, atol=2e-3)


def test_conv3d_head_small_cout_ncdhw_f32(ops):
    B, C, D, H, W = 2, 224, 16, 16, 16
    g = torch.Generator(device="cuda").manual_seed(8)
    x = torch.randn(B, C, D, H, W, device="cuda", generator=g)
    w = torch.randn(3, C, 3, 3, 3, device="cuda", generator=g) / math.sqrt(C * 27)
    b = torch.randn(3, device="cuda", generator=g)
    ref = F.conv3d(_bf(x).float(), _bf(w).float(), b, padding=1)
    from commonscenes_b200._lib import OUT_F32_NCDHW
    got = ops.conv3d(_cl(x), ops.pack_conv_weight(w), bias=b, out_mode=OUT_F32_NCDHW)
    torch.cuda.synchronize()
    _report("head", got, ref, rtol=1e-4, atol=1e-4)


def test_stem_im2col_gemm(ops):
    B, D = 4, 16
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn(2, 3, D, D, D, device="cuda", generator=g)
    w = torch.randn(224, 3, 3, 3, 3, device="cuda", generator=g) / math.sqrt(81)
    b = torch.randn(224, device="cuda", generator=g)
    wp, kp = ops.pack_patch_weight(w)                             # columns in (tap, c) order
    col = ops.im2col_small(x, batch=B, kp=kp)
    got = ops.linear_tokens(col, wp, bias=b)
    torch.cuda.synchronize()
    ref = F.conv3d(_bf(torch.cat([x, x])).float(), _bf(w).float(), b, padding=1)
    _report("stem", _ncdhw(got), ref, rtol=2 ** -7, atol=2e-3)


@pytest.mark.parametrize("C,eps,act", [(224, 1e-5, "silu"), (672, 1e-6, "none"), (64, 1e-6, "gelu")])
def test_groupnorm(ops, C, eps, act):
    B, D, H, W = 3, 16, 8, 8
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(B, C, D, H, W, device="cuda", generator=g) * 1.7 + 0.3
    ga = torch.randn(C, device="cuda", generator=g)
    be = torch.randn(C, device="cuda", generator=g)
    ref = F.group_norm(_bf(x).float(), 32, ga, be, eps)
    ref = {"silu": F.silu, "gelu": F.gelu, "none": lambda t: t}[act](ref)
    code = {"none": ops.ACT_NONE, "silu": ops.ACT_SILU, "gelu": ops.ACT_GELU}[act]
    got = ops.groupnorm(_cl(x), ga, be, eps=eps, act=code)
    torch.cuda.synchronize()
    _report(f"groupnorm{C}", _ncdhw(got), ref, rtol=2 ** -7, atol=2e-3)


def test_groupnorm_two_sources(ops):
    B, C1, C2, D, H, W = 2, 448, 224, 16, 8, 8
    g = torch.Generator(device="cuda").manual_seed(4)
    x1 = torch.randn(B, C1, D, H, W, device="cuda", generator=g)
    x2 = torch.randn(B, C2, D, H, W, device="cuda", generator=g) * 2 + 1
    ga = torch.randn(C1 + C2, device="cuda", generator=g)
    be = torch.randn(C1 + C2, device="cuda", generator=g)
    ref = F.silu(F.group_norm(_bf(torch.cat([x1, x2], 1)).float(), 32, ga, be, 1e-5))
    got = ops.groupnorm(_cl(x1), ga, be, act=ops.ACT_SILU, x2=_cl(x2))
    torch.cuda.synchronize()
    _report("groupnorm_cat", _ncdhw(got), ref, rtol=2 ** -7, atol=2e-3)


@pytest.mark.parametrize("C", [448, 672])
def test_layernorm(ops, C):
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(2, 16, 4, 4, C, device="cuda", generator=g) * 2 + 0.5
    ga = torch.randn(C, device="cuda", generator=g)
    be = torch.randn(C, device="cuda", generator=g)
    ref = F.layer_norm(_bf(x).float(), (C,), ga, be, 1e-5)
    got = ops.layernorm(_bf(x), ga, be)
    torch.cuda.synchronize()
    _report(f"layernorm{C}", got.float(), ref, rtol=2 ** -7, atol=2e-3)


@pytest.mark.parametrize("N,heads,d,dp", [(1024, 8, 56, 64), (256, 8, 84, 96), (4096, 1, 256, 256), (256, 8, 4, 32)],
                         ids=["n1024d56", "n256d84", "vq4096d256", "tiny"])
def test_attention(ops, N, heads, d, dp):
    B = 2
    g = torch.Generator(device="cuda").manual_seed(1)
    qkv = torch.zeros(B, N, 3 * heads * dp, device="cuda")
    for i in range(3):
        for h in range(heads):
            qkv[:, :, (i * heads + h) * dp:(i * heads + h) * dp + d] = torch.randn(B, N, d, device="cuda", generator=g)
    qkv = _bf(qkv)
    q, k, v = (qkv[:, :, i * heads * dp:(i + 1) * heads * dp] for i in range(3))
    scale = d ** -0.5
    got = ops.attention(q, k, v, heads=heads, head_dim=d, head_dim_padded=dp, scale=scale)
    torch.cuda.synchronize()

    def split(t):
        return t.float().reshape(B, N, heads, dp)[..., :d].permute(0, 2, 1, 3)
    qf, kf, vf = split(q), split(k), split(v)
    attn = torch.softmax(qf @ kf.transpose(-1, -2) * scale, dim=-1)
    ref = (attn @ vf).permute(0, 2, 1, 3).reshape(B, N, heads * d)
    _report("attention", got.float(), ref, rtol=2e-2, atol=1e-2)


def test_attention_two_sweep_kernel_many_items_and_lse(ops):
    """The two-sweep tcgen05 kernel (cs_attn_tc2.cu) with more work items than SMs (every persistent CTA walks ~5 (sample,
    head, 256-query block) items: barrier phases across items), peaked score rows (|s| up to ~40: the final-maximum
    subtraction must hold), the log-sum-exp rows of the training forward, and bit-reproducibility."""
    from commonscenes_b200 import ops_bwd
    B, N, heads, d, dp = 24, 1024, 8, 56, 64
    g = torch.Generator(device="cuda").manual_seed(3)
    qkv = torch.zeros(B, N, 3 * heads * dp, device="cuda")
    for i in range(3):
        for h in range(heads):
            sc = 3.0 if i < 2 else 1.0
            qkv[:, :, (i * heads + h) * dp:(i * heads + h) * dp + d] = sc * torch.randn(B, N, d, device="cuda", generator=g)
    qkv = _bf(qkv)
    q, k, v = (qkv[:, :, i * heads * dp:(i + 1) * heads * dp] for i in range(3))
    scale = d ** -0.5
    got = ops.attention(q, k, v, heads=heads, head_dim=d, head_dim_padded=dp, scale=scale)
    got2, lse = ops_bwd.attention_lse(q, k, v, heads=heads, head_dim=d, head_dim_padded=dp, scale=scale)
    torch.cuda.synchronize()
    assert torch.equal(got, got2)

    def split(t):
        return t.float().reshape(B, N, heads, dp)[..., :d].permute(0, 2, 1, 3)
    qf, kf, vf = split(q), split(k), split(v)
    sraw = qf @ kf.transpose(-1, -2) * scale
    ref = (torch.softmax(sraw, dim=-1) @ vf).permute(0, 2, 1, 3).reshape(B, N, heads * d)
    _report("attention_tc2", got.float(), ref, rtol=2e-2, atol=1e-2)
    lse_ref = torch.logsumexp(sraw, dim=-1) * 1.4426950408889634          # base 2, (B, heads, N)
    _report("attention_tc2_lse", lse, lse_ref, rtol=1e-3, atol=2e-2)
    # the default path is the third-generation kernel (cs_attn_tc3.cu): same arithmetic as the second generation (debug flag
    # 32768 selects that one), so the two must agree bit for bit
    from commonscenes_b200 import _lib
    try:
        _lib.load().cs_debug_set(32768)
        gen2 = ops.attention(q, k, v, heads=heads, head_dim=d, head_dim_padded=dp, scale=scale)
        torch.cuda.synchronize()
    finally:
        _lib.load().cs_debug_set(0)
    assert torch.equal(got, gen2), f"gen 3 vs gen 2: max |diff| {float((got.float() - gen2.float()).abs().max())}"


@pytest.mark.parametrize("B,N,heads", [(3, 384, 1), (1, 128, 1), (5, 640, 3), (2, 2048, 2)])
def test_attention_gen3_odd_tile_counts(ops, B, N, heads):
    """cs_attn_tc3.cu: an odd number of 128-query tiles (the second pipeline of the last CTA has one item less or none),
    one-tile sequences, more key tiles than K stages."""
    d, dp = 56, 64
    g = torch.Generator(device="cuda").manual_seed(N)
    qkv = torch.zeros(B, N, 3 * heads * dp, device="cuda")
    for i in range(3):
        for h in range(heads):
            qkv[:, :, (i * heads + h) * dp:(i * heads + h) * dp + d] = 1.5 * torch.randn(B, N, d, device="cuda", generator=g)
    qkv = _bf(qkv)
    q, k, v = (qkv[:, :, i * heads * dp:(i + 1) * heads * dp] for i in range(3))
    scale = d ** -0.5
    got = ops.attention(q, k, v, heads=heads, head_dim=d, head_dim_padded=dp, scale=scale)
    again = ops.attention(q, k, v, heads=heads, head_dim=d, head_dim_padded=dp, scale=scale)
    torch.cuda.synchronize()
    assert torch.equal(got, again)

    def split(t):
        return t.float().reshape(B, N, heads, dp)[..., :d].permute(0, 2, 1, 3)
    qf, kf, vf = split(q), split(k), split(v)
    ref = (torch.softmax(qf @ kf.transpose(-1, -2) * scale, dim=-1) @ vf).permute(0, 2, 1, 3).reshape(B, N, heads * d)
    _report("attention_tc3", got.float(), ref, rtol=2e-2, atol=1e-2)


def test_attention_cross_short_context(ops):
    # generic cross-attention with a ragged (non multiple of 64) context length
    B, Nq, Nk, heads, d, dp = 2, 256, 77, 8, 56, 64
    g = torch.Generator(device="cuda").manual_seed(12)
    q = _bf(torch.randn(B, Nq, heads * dp, device="cuda", generator=g))
    k = _bf(torch.randn(B, Nk, heads * dp, device="cuda", generator=g))
    v = _bf(torch.randn(B, Nk, heads * dp, device="cuda", generator=g))
    got = ops.attention(q, k, v, heads=heads, head_dim=d, head_dim_padded=dp, scale=d ** -0.5)
    torch.cuda.synchronize()
    qf = q.float().reshape(B, Nq, heads, dp).permute(0, 2, 1, 3)
    kf = k.float().reshape(B, Nk, heads, dp).permute(0, 2, 1, 3)
    vf = v.float().reshape(B, Nk, heads, dp).permute(0, 2, 1, 3)
    ref = (torch.softmax(qf @ kf.transpose(-1, -2) * d ** -0.5, -1) @ vf)[..., :d].permute(0, 2, 1, 3).reshape(B, Nq, heads * d)
    _report("cross_attention", got.float(), ref, rtol=2e-2, atol=1e-2)


def test_geglu_upsample_layout(ops):
    g = torch.Generator(device="cuda").manual_seed(13)
    x = _bf(torch.randn(2, 16, 4, 4, 2 * 1792, device="cuda", generator=g))
    a, gate = x.float().chunk(2, dim=-1)
    _report("geglu", ops.geglu(x).float(), a * F.gelu(gate), rtol=2 ** -7, atol=1e-3)
    y = _bf(torch.randn(2, 16, 4, 4, 672, device="cuda", generator=g))
    up = ops.upsample_nearest(y, (1, 2, 2))
    ref = F.interpolate(_ncdhw(y), (16, 8, 8), mode="nearest")
    assert torch.equal(_ncdhw(up), ref)
    z = torch.randn(3, 3, 16, 16, 16, device="cuda", generator=g)
    cl = ops.to_channels_last(z, 8)
    assert torch.equal(cl[..., :3].float(), _bf(z).float().permute(0, 2, 3, 4, 1))
    assert float(cl[..., 3:].abs().max()) == 0.0
    assert torch.equal(ops.to_ncdhw(cl, 3), _bf(z).float())


def test_timestep_embedding_and_small_linear(ops):
    t = torch.tensor([0, 1, 11, 500, 991, 999], device="cuda", dtype=torch.int64)
    dim, half = 224, 112
    freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32, device="cuda") / half)
    args = t[:, None].float() * freqs[None]
    ref = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    _report("timestep_embedding", ops.timestep_embedding(t, dim), ref, rtol=0, atol=2e-4)
    g = torch.Generator(device="cuda").manual_seed(14)
    x = torch.randn(13, 896, device="cuda", generator=g)
    w = torch.randn(672, 896, device="cuda", generator=g) / 30
    b = torch.randn(672, device="cuda", generator=g)
    got = ops.linear_small(x, w, b, act_in=ops.ACT_SILU)
    _report("linear_small", got, F.linear(F.silu(x), w, b), rtol=1e-4, atol=1e-4)


def test_ddim_step_and_q_sample(ops):
    g = torch.Generator(device="cuda").manual_seed(15)
    x = torch.randn(4, 3, 16, 16, 16, device="cuda", generator=g)
    eps = torch.randn(8, 3, 16, 16, 16, device="cuda", generator=g)
    a_t, a_prev, sigma, s1m = 0.37, 0.41, 0.0, math.sqrt(1 - 0.37)
    e = eps[:4] + 3.0 * (eps[4:] - eps[:4])
    pred = (x - s1m * e) / math.sqrt(a_t)
    ref = math.sqrt(a_prev) * pred + math.sqrt(1 - a_prev - sigma ** 2) * e
    xp, p0 = ops.ddim_step(x, eps, guided=True, scale=3.0, a_t=a_t, a_prev=a_prev, sigma=sigma, sqrt_one_minus_at=s1m)
    _report("ddim_x_prev", xp, ref, rtol=1e-5, atol=1e-5)
    _report("ddim_pred_x0", p0, pred, rtol=1e-5, atol=1e-5)
    ac = torch.linspace(0.99, 0.01, 1000, device="cuda")
    t = torch.tensor([0, 10, 500, 999], device="cuda")
    noise = torch.randn_like(x)
    ref = ac.sqrt()[t].view(-1, 1, 1, 1, 1) * x + (1 - ac).sqrt()[t].view(-1, 1, 1, 1, 1) * noise
    _report("q_sample", ops.q_sample(x, noise, t, ac.sqrt().contiguous(), (1 - ac).sqrt().contiguous()), ref, 1e-6, 1e-6)


def test_linear_geglu_fused_epilogue(ops):
    B, C = 2, 448
    g = torch.Generator(device="cuda").manual_seed(21)
    x = torch.randn(B, 16, 8, 8, C, device="cuda", generator=g)
    w = torch.randn(8 * C, C, device="cuda", generator=g) / math.sqrt(C)
    b = torch.randn(8 * C, device="cuda", generator=g)
    a, gate = F.linear(_bf(x).float(), _bf(w).float(), b).chunk(2, dim=-1)
    ref = a * F.gelu(gate)
    wp, bp = ops.pack_geglu_weight(w, b)
    got = ops.linear_tokens(_bf(x), wp, bias=bp, act=ops.ACT_GEGLU)
    torch.cuda.synchronize()
    assert got.shape[-1] == 4 * C
    _report("geglu_fused", got.float(), ref, rtol=2 ** -7, atol=4e-3)


def test_conv3d_residual_view_and_pitched_output(ops):
    """Output / residual may be channel slices of wider buffers (row pitch > C)."""
    B, C, D, H, W = 2, 224, 16, 8, 8
    g = torch.Generator(device="cuda").manual_seed(22)
    x = torch.randn(B, C, D, H, W, device="cuda", generator=g)
    w = torch.randn(C, C, 3, 3, 3, device="cuda", generator=g) / math.sqrt(C * 27)
    wide_res = _bf(torch.randn(B, D, H, W, 2 * C, device="cuda", generator=g))
    wide_out = torch.zeros(B, D, H, W, 3 * C, device="cuda", dtype=torch.bfloat16)
    res, out = wide_res[..., C:], wide_out[..., C:2 * C]
    ops.conv3d(_cl(x), ops.pack_conv_weight(w), residual=res, out=out)
    torch.cuda.synchronize()
    ref = F.conv3d(_bf(x).float(), _bf(w).float(), None, padding=1) + res.float().permute(0, 4, 1, 2, 3)
    _report("pitched", _ncdhw(out), ref, rtol=2 ** -7, atol=4e-3)
    assert float(wide_out[..., :C].abs().max()) == 0.0 and float(wide_out[..., 2 * C:].abs().max()) == 0.0


@pytest.mark.parametrize("cin,cout,res", [(224, 3, 16), (64, 1, 32), (256, 3, 16)], ids=["unet_head", "vq_dec_out", "vq_enc_out"])
def test_conv3d_small_cout_tap_gemm_gather(ops, cin, cout, res):
    B = 2
    g = torch.Generator(device="cuda").manual_seed(31)
    x = torch.randn(B, cin, res, res, res, device="cuda", generator=g)
    w = torch.randn(cout, cin, 3, 3, 3, device="cuda", generator=g) / math.sqrt(cin * 27)
    b = torch.randn(cout, device="cuda", generator=g)
    ref = F.conv3d(_bf(x).float(), _bf(w).float(), b, padding=1)
    got = ops.conv3d_small_cout(_cl(x), ops.pack_small_cout_conv(w), b, cout)
    torch.cuda.synchronize()
    _report("small_cout", got, ref, rtol=1e-4, atol=2e-4)
