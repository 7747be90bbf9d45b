"""Training path of the denoiser: explicit backward on the B200 kernels vs autograd through the fp32 oracle.

The oracle (CPU, fp32) restates UNet3DModel.forward; torch.autograd over it gives the reference gradients of
p_losses' MSE (sdfusion_txt2shape_model.py:311-345).  The CUDA path keeps activations and activation gradients in bf16,
so the bar follows SURVEY.md §8c: cosine >= 0.999 on the whole gradient, and per-tensor rel-L2 within bf16 noise.
"""
import pytest
import torch

from oracle import denoiser as D, weights as Wt

pytestmark = pytest.mark.gpu


def _build(cfg, seed):
    from commonscenes_b200.model.networks.diffusion_networks.network import DiffusionUNet
    params = dict(cfg, use_spatial_transformer=True, use_checkpoint=True, legacy=False)
    m = DiffusionUNet(params, conditioning_key="crossattn")
    Wt.fill_module_(m, seed)
    return m.cuda()


def _oracle_grads(cfg, seed, x, t, ctx, noise):
    sd = {k: v.clone().requires_grad_(True) for k, v in Wt.synth_state_dict(D.unet_param_shapes(cfg), seed).items()}
    ctx = ctx.clone().requires_grad_(True)
    eps = D.unet_forward(sd, cfg, x, t, ctx)
    loss = torch.nn.functional.mse_loss(eps, noise)
    keys = list(sd.keys())
    grads = torch.autograd.grad(loss, [sd[k] for k in keys] + [ctx], allow_unused=True)
    return eps.detach(), loss.item(), dict(zip(keys, grads[:-1])), grads[-1]


def test_unet_backward_matches_oracle_autograd():
    from commonscenes_b200 import ops_bwd
    from commonscenes_b200.model.networks.diffusion_networks.unet_train import UNetTrainer
    cfg = D.UNET_TINY
    seed = 31
    g = torch.Generator().manual_seed(8)
    B = 4
    x = torch.randn(B, 3, 8, 8, 8, generator=g)
    t = torch.tensor([999, 3, 421, 650])
    ctx = torch.randn(B, 1, cfg["context_dim"], generator=g)
    noise = torch.randn(B, 3, 8, 8, 8, generator=g)
    eps_ref, loss_ref, gref, gctx_ref = _oracle_grads(cfg, seed, x, t, ctx, noise)

    m = _build(cfg, seed)
    tr = UNetTrainer(m.diffusion_net)
    eps, tape = tr.forward_train(x.cuda(), t.cuda(), ctx.cuda())
    assert float((eps.cpu() - eps_ref).norm() / eps_ref.norm()) <= 3e-2
    loss = torch.zeros((), device="cuda")
    d_eps = ops_bwd.mse_loss_grad(eps, noise.cuda(), loss)
    assert abs(loss.item() - loss_ref) / loss_ref < 3e-2
    sink, dctx = tr.backward(tape, d_eps)

    named = dict(m.named_parameters())
    rows = []
    for k, gr in gref.items():
        got = sink.grads.get(named[k])
        if gr is None or float(gr.norm()) == 0.0:       # attn2.to_q / to_k, norm2: exactly zero for a single context token
            assert got is None or float(got.abs().max()) == 0.0, k
            continue
        assert got is not None, f"no gradient produced for {k}"
        rows.append((k, got.cpu(), gr))
    num = sum(float((g - r).pow(2).sum()) for _, g, r in rows)
    den = sum(float(r.pow(2).sum()) for _, _, r in rows)
    dot = sum(float((g * r).sum()) for _, g, r in rows)
    gg = sum(float(g.pow(2).sum()) for _, g, r in rows)
    rms = (den / sum(r.numel() for _, _, r in rows)) ** 0.5      # typical gradient magnitude of the whole model
    worst = []
    for k, g, r in rows:
        err, ref = float((g - r).norm()), float(r.norm())
        worst.append((err / ref, k))
        # per tensor: bf16-level relative error, with an absolute floor for gradients that are (near) zero by symmetry
        # (e.g. a conv bias feeding a GroupNorm whose groups hold a single channel in the tiny configuration)
        assert err <= 0.15 * ref + 0.05 * rms * r.numel() ** 0.5, f"{k}: err {err:.3e} vs ref norm {ref:.3e}"
    worst.sort(reverse=True)
    print("worst tensors:", [(f"{r:.3e}", k) for r, k in worst[:8]])
    total_rel = (num / den) ** 0.5
    cos = dot / (den * gg) ** 0.5
    print(f"whole-gradient rel-L2 {total_rel:.3e}, cosine {cos:.6f}")
    assert total_rel < 3e-2 and cos > 0.999
    rel_ctx = float((dctx.cpu() - gctx_ref).norm() / gctx_ref.norm())
    print(f"d_context rel-L2 {rel_ctx:.3e}")
    assert rel_ctx < 3e-2


def test_autograd_bridge_and_native_step_agree_with_torch_adamw():
    """loss.backward() through UNet3DModel.forward + clip_grad_norm_ + torch.optim.AdamW (what the reference's training
    loop does, train_3dfront.py:387-418) vs DenoiserTrainStep (flat buffers, cs_sumsq + cs_adamw)."""
    from commonscenes_b200.model.sdfusion_txt2shape_model import SDFusionText2ShapeModel, diffusion_schedule
    from commonscenes_b200.train import DenoiserTrainStep
    cfg = D.UNET_TINY

    class Stub:     # the members DenoiserTrainStep reads from SDFusionText2ShapeModel
        q_sample = SDFusionText2ShapeModel.q_sample

        def __init__(self, df):
            self.df, self.num_timesteps, self.device = df, 1000, "cuda"
            for k, v in diffusion_schedule(1000, 0.00085, 0.012).items():
                setattr(self, k, v.cuda())

    g = torch.Generator().manual_seed(9)
    B = 4
    z = torch.randn(B, 3, 8, 8, 8, generator=g).cuda()
    ctx = torch.randn(B, 1, cfg["context_dim"], generator=g).cuda()
    ts = [torch.tensor([999, 3, 421, 650]).cuda(), torch.tensor([5, 800, 77, 300]).cuda()]
    noises = [torch.randn(B, 3, 8, 8, 8, generator=g).cuda() for _ in range(2)]

    ma, mb = Stub(_build(cfg, 41)), Stub(_build(cfg, 41))
    before = {k: v.detach().clone() for k, v in ma.df.named_parameters()}
    opt = torch.optim.AdamW(ma.df.parameters(), lr=1e-4)
    losses_a = []
    for t, n in zip(ts, noises):
        x_t = ma.q_sample(z, t, n)
        eps = ma.df(x_t, t, c_crossattn=[ctx])
        assert eps.requires_grad
        loss = torch.nn.functional.mse_loss(eps, n)
        opt.zero_grad()
        (100.0 * loss).backward()
        torch.nn.utils.clip_grad_norm_(ma.df.parameters(), 5.0)
        opt.step()
        losses_a.append(loss.item())

    step = DenoiserTrainStep(mb, lr=1e-4)
    losses_b = []
    for t, n in zip(ts, noises):
        loss, dc = step.step(z, ctx, t=t, noise=n, need_dcond=True)
        losses_b.append(loss.item())
        assert dc.shape == ctx.shape and torch.isfinite(dc).all()
    print("losses", losses_a, losses_b)
    for a, b in zip(losses_a, losses_b):
        assert abs(a - b) / a < 2e-2
    num = den = 0.0
    pb = dict(mb.df.named_parameters())
    for k, pa in ma.df.named_parameters():
        da, db = pa.detach() - before[k], pb[k].detach() - before[k]
        num += float((da - db).pow(2).sum()); den += float(da.pow(2).sum())
    print(f"parameter-update rel-L2 (native step vs torch AdamW): {(num / den) ** 0.5:.3e}")
    # AdamW's first updates are ~lr * sign(g): elements whose gradient is at the bf16 noise floor flip between the two runs
    # (fp32 atomics order), so the bar on the UPDATE is loose; the gradients themselves are checked against the oracle above
    assert (num / den) ** 0.5 < 0.25
    # state-dict keys and shapes are untouched by the flat re-homing
    assert {k: tuple(v.shape) for k, v in ma.df.state_dict().items()} == {k: tuple(v.shape) for k, v in mb.df.state_dict().items()}


def test_unet_backward_full_config_vs_oracle_on_device():
    """Full-size denoiser (413.5 M parameters), B = 2: gradients vs autograd through the oracle evaluated in fp32 on the
    same GPU (TF32 off)."""
    from commonscenes_b200 import ops_bwd
    from commonscenes_b200.model.networks.diffusion_networks.unet_train import UNetTrainer
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = D.UNET_FULL
    seed = 51
    g = torch.Generator().manual_seed(10)
    B = 2
    x = torch.randn(B, 3, 16, 16, 16, generator=g).cuda()
    t = torch.tensor([900, 120]).cuda()
    ctx = torch.randn(B, 1, cfg["context_dim"], generator=g).cuda()
    noise = torch.randn(B, 3, 16, 16, 16, generator=g).cuda()
    sd = {k: v.cuda().requires_grad_(True) for k, v in Wt.synth_state_dict(D.unet_param_shapes(cfg), seed).items()}
    with torch.device("cuda"):      # the oracle builds its constants on the default device
        eps_ref = D.unet_forward(sd, cfg, x, t, ctx)
    loss = torch.nn.functional.mse_loss(eps_ref, noise)
    keys = list(sd.keys())
    gref = dict(zip(keys, torch.autograd.grad(loss, [sd[k] for k in keys], allow_unused=True)))
    del loss

    m = _build(cfg, seed)
    tr = UNetTrainer(m.diffusion_net)
    eps, tape = tr.forward_train(x, t, ctx)
    assert float((eps - eps_ref.detach()).norm() / eps_ref.detach().norm()) <= 3e-2
    lbuf = torch.zeros((), device="cuda")
    sink, _ = tr.backward(tape, ops_bwd.mse_loss_grad(eps, noise, lbuf), need_dcontext=False)
    named = dict(m.named_parameters())
    num = den = dot = gg = 0.0
    worst = []
    for k, gr in gref.items():
        got = sink.grads.get(named[k])
        if gr is None or float(gr.norm()) == 0.0:
            continue
        assert got is not None, k
        num += float((got - gr).pow(2).sum()); den += float(gr.pow(2).sum())
        dot += float((got * gr).sum()); gg += float(got.pow(2).sum())
        worst.append((float((got - gr).norm() / gr.norm()), k))
    worst.sort(reverse=True)
    print("worst tensors:", [(f"{r:.3e}", k) for r, k in worst[:6]])
    print(f"full config: whole-gradient rel-L2 {(num / den) ** 0.5:.3e}, cosine {dot / (den * gg) ** 0.5:.6f}")
    assert (num / den) ** 0.5 < 3e-2 and dot / (den * gg) ** 0.5 > 0.999


def test_graphed_train_step_matches_eager():
    from commonscenes_b200.model.sdfusion_txt2shape_model import SDFusionText2ShapeModel, diffusion_schedule
    from commonscenes_b200.train import DenoiserTrainStep
    cfg = D.UNET_TINY

    class Stub:
        q_sample = SDFusionText2ShapeModel.q_sample
        z_shape = (3, 8, 8, 8)

        def __init__(self, df):
            self.df, self.num_timesteps, self.device = df, 1000, "cuda"
            for k, v in diffusion_schedule(1000, 0.00085, 0.012).items():
                setattr(self, k, v.cuda())

    g = torch.Generator().manual_seed(12)
    B = 4
    z = torch.randn(B, 3, 8, 8, 8, generator=g).cuda()
    ctx = torch.randn(B, 1, cfg["context_dim"], generator=g).cuda()
    t = torch.tensor([10, 999, 500, 250]).cuda()
    noise = torch.randn(B, 3, 8, 8, 8, generator=g).cuda()
    ma, mb = Stub(_build(cfg, 61)), Stub(_build(cfg, 61))
    sa, sb = DenoiserTrainStep(ma), DenoiserTrainStep(mb)
    p0 = sa.flat_p.clone()
    sb.capture(B, cfg["context_dim"], need_dcond=True)
    assert torch.equal(sb.flat_p, p0) and sb.step_count == 0      # warm-up iterations were rolled back
    for _ in range(3):
        la, da = sa.step(z, ctx, t=t, noise=noise, need_dcond=True)
        lb, db = sb.step_graphed(z, ctx, t=t, noise=noise)
        assert abs(la.item() - lb.item()) / la.item() < 1e-2
    ua, ub = sa.flat_p - p0, sb.flat_p - p0
    rel = float((ua - ub).norm() / ua.norm())
    print(f"graphed vs eager parameter update rel-L2 after 3 steps: {rel:.3e}")
    assert rel < 0.25 and int(sb.step_dev.item()) == 3
    assert float((da - db).norm() / da.norm()) < 0.1


def test_evaluation_between_graphed_steps_sees_current_weights():
    """Mid-training validation: an evaluation (UNet forward / DDIM graph) between two replays of the captured training step
    must use the weights of that moment.  The AdamW kernel writes through raw pointers (no autograd version bump), so
    step_graphed() has to invalidate the module's packed-weight cache itself; compare every evaluation against a second
    module that was freshly given the same parameter values (an eager repack)."""
    from commonscenes_b200.model.sdfusion_txt2shape_model import SDFusionText2ShapeModel, diffusion_schedule
    from commonscenes_b200.train import DenoiserTrainStep
    cfg = D.UNET_TINY

    class Stub:
        q_sample = SDFusionText2ShapeModel.q_sample
        z_shape = (3, 8, 8, 8)

        def __init__(self, df):
            self.df, self.num_timesteps, self.device = df, 1000, "cuda"
            for k, v in diffusion_schedule(1000, 0.00085, 0.012).items():
                setattr(self, k, v.cuda())

    g = torch.Generator().manual_seed(13)
    B = 4
    z = torch.randn(B, 3, 8, 8, 8, generator=g).cuda()
    ctx = torch.randn(B, 1, cfg["context_dim"], generator=g).cuda()
    t = torch.tensor([10, 999, 500, 250]).cuda()
    noise = torch.randn(B, 3, 8, 8, 8, generator=g).cuda()
    m, fresh = _build(cfg, 62), _build(cfg, 62)
    step = DenoiserTrainStep(Stub(m), lr=5e-2)               # a large step: stale weights would be far off
    step.capture(B, cfg["context_dim"])
    outs = []
    for i in range(3):
        with torch.no_grad():
            got = m(z, t, c_crossattn=[ctx])
            fresh.load_state_dict(m.state_dict())            # eager repack of the same values (version counters bump)
            want = fresh(z, t, c_crossattn=[ctx])
        assert torch.equal(got, want), f"evaluation after {i} graphed steps used stale packed weights"
        outs.append(got)
        step.step_graphed(z, ctx, t=t, noise=noise)
    assert float((outs[1] - outs[0]).abs().max()) > 0 and float((outs[2] - outs[1]).abs().max()) > 0


def test_concat_unet_backward_matches_oracle_autograd():
    """Concat-conditioning denoiser (AttentionBlock variant): `loss.backward()` through
    DiffusionUNet(conditioning_key='concat') -- parameter gradients (incl. the Conv1d qkv / proj_out of every AttentionBlock in
    the reference's legacy head-major row order) and the gradient of the concatenated conditioning volume, which flows through
    the stem's input gradient -- vs autograd through the oracle."""
    from commonscenes_b200.model.networks.diffusion_networks.network import DiffusionUNet
    cfg, seed = D.UNET_CONCAT_TINY, 71
    g = torch.Generator().manual_seed(14)
    B = 4
    x = torch.randn(B, 3, 8, 8, 8, generator=g)
    cc = torch.randn(B, 1, 8, 8, 8, generator=g)
    t = torch.tensor([999, 3, 421, 650])
    noise = torch.randn(B, 3, 8, 8, 8, generator=g)
    sd = {k: v.clone().requires_grad_(True) for k, v in Wt.synth_state_dict(D.unet_param_shapes(cfg), seed).items()}
    cc_ref = cc.clone().requires_grad_(True)
    eps_ref = D.unet_forward(sd, cfg, x, t, c_concat=cc_ref)
    loss_ref = torch.nn.functional.mse_loss(eps_ref, noise)
    keys = list(sd.keys())
    grads = torch.autograd.grad(loss_ref, [sd[k] for k in keys] + [cc_ref], allow_unused=True)
    gref, gcc_ref = dict(zip(keys, grads[:-1])), grads[-1]

    m = DiffusionUNet(dict(cfg, use_checkpoint=True, legacy=False), conditioning_key="concat")
    Wt.fill_module_(m, seed)
    m = m.cuda()
    cc_dev = cc.cuda().requires_grad_(True)
    eps = m(x.cuda(), t.cuda(), c_concat=[cc_dev])
    assert eps.requires_grad and float((eps.detach().cpu() - eps_ref.detach()).norm() / eps_ref.detach().norm()) <= 3e-2
    torch.nn.functional.mse_loss(eps, noise.cuda()).backward()
    named = dict(m.named_parameters())
    num = den = dot = gg = 0.0
    worst = []
    n_el = 0
    for k, gr in gref.items():
        got = named[k].grad
        if gr is None or float(gr.norm()) == 0.0:
            continue
        assert got is not None, f"no gradient produced for {k}"
        got = got.cpu()
        num += float((got - gr).pow(2).sum()); den += float(gr.pow(2).sum())
        dot += float((got * gr).sum()); gg += float(got.pow(2).sum())
        n_el += gr.numel()
        worst.append((float((got - gr).norm() / gr.norm()), k))
    worst.sort(reverse=True)
    rms = (den / n_el) ** 0.5
    for k, gr in gref.items():
        if gr is None or float(gr.norm()) == 0.0:
            continue
        err = float((named[k].grad.cpu() - gr).norm())
        assert err <= 0.15 * float(gr.norm()) + 0.05 * rms * gr.numel() ** 0.5, f"{k}: err {err:.3e} vs ref norm {float(gr.norm()):.3e}"
    total_rel, cos = (num / den) ** 0.5, dot / (den * gg) ** 0.5
    rel_cc = float((cc_dev.grad.cpu() - gcc_ref).norm() / gcc_ref.norm())
    print("worst tensors:", [(f"{r:.3e}", k) for r, k in worst[:6]])
    print(f"concat variant: whole-gradient rel-L2 {total_rel:.3e}, cosine {cos:.6f}; d c_concat rel-L2 {rel_cc:.3e}")
    # measured 2.0e-2 / 0.9998 / 2.1e-2 (tiny configuration, every activation and activation gradient in bf16)
    assert total_rel < 4e-2 and cos > 0.999 and rel_cc < 4e-2


def test_fused_update_equals_the_separate_kernels():
    """cs_adamw_repack (packed-gradient slots -> clip + AdamW + both bf16 packs in one pass) against the un-pack / cs_adamw /
    cs_pack_weight sequence it replaces: same parameters and moments after two steps, and the packs it maintains equal a
    fresh pack of the updated weights (evaluation through them == evaluation of a module freshly loaded with those weights)."""
    from commonscenes_b200.model.sdfusion_txt2shape_model import SDFusionText2ShapeModel, diffusion_schedule
    from commonscenes_b200.train import DenoiserTrainStep
    cfg = D.UNET_TINY

    class Stub:
        q_sample = SDFusionText2ShapeModel.q_sample
        z_shape = (3, 8, 8, 8)

        def __init__(self, df):
            self.df, self.num_timesteps, self.device = df, 1000, "cuda"
            for k, v in diffusion_schedule(1000, 0.00085, 0.012).items():
                setattr(self, k, v.cuda())

    g = torch.Generator().manual_seed(14)
    B = 4
    z = torch.randn(B, 3, 8, 8, 8, generator=g).cuda()
    ctx = torch.randn(B, 1, cfg["context_dim"], generator=g).cuda()
    t = torch.tensor([10, 999, 500, 250]).cuda()
    noise = torch.randn(B, 3, 8, 8, 8, generator=g).cuda()
    ma, mb, fresh = _build(cfg, 64), _build(cfg, 64), _build(cfg, 64)
    p_start = {n: p.detach().clone() for n, p in ma.named_parameters()}
    sa, sb = DenoiserTrainStep(Stub(ma), lr=1e-3, fused_update=False), DenoiserTrainStep(Stub(mb), lr=1e-3)
    assert sb.fused and not sa.fused
    n_packed = sum(p.numel() for p in sb.packed_views)
    print(f"fused-update weights: {len(sb.packed_views)} tensors, {n_packed / sum(p.numel() for p in sb.params):.1%} of the parameters")
    assert len(sb.packed_views) >= 10
    for it in range(2):
        la, _ = sa.step(z, ctx, t=t, noise=noise)
        lb, _ = sb.step(z, ctx, t=t, noise=noise)
        assert abs(la.item() - lb.item()) <= 1e-3 * abs(la.item())
        if it == 0:
            # after ONE step the moments are the clipped gradients themselves ((1 - beta) g, (1 - beta2) g^2): a direct comparison
            # of the packed-slot gradients with the un-packed ones, parameter by parameter (two launches of the fp32-atomic
            # weight-gradient kernel differ by rounding only)
            oa, ob = sa.optimizer_state_dict(), sb.optimizer_state_dict()
            worst = 0.0
            for i in oa["state"]:
                for k in ("exp_avg", "exp_avg_sq"):
                    a, b = oa["state"][i][k], ob["state"][i][k]
                    worst = max(worst, float((a - b).norm()) / (float(a.norm()) + 1e-30))
            print(f"fused vs separate: worst per-parameter moment rel-L2 after 1 step {worst:.3e}")
            assert worst < 1e-4
    num = den = 0.0
    for (n, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        ua, ub = pa.detach() - p_start[n], pb.detach() - p_start[n]
        num += float((ua - ub).double().pow(2).sum()); den += float(ua.double().pow(2).sum())
    print(f"fused vs separate: parameter-update rel-L2 after 2 steps {(num / den) ** 0.5:.3e}")
    assert (num / den) ** 0.5 < 5e-2          # rounding noise through AdamW's sign-like first steps at lr 1e-3, not layout errors
    assert float(sb.flat_g[sb.n_plain:].abs().max()) == 0.0     # the packed slots were cleared by the update
    with torch.no_grad():
        got = mb(z, t, c_crossattn=[ctx])
        fresh.load_state_dict(mb.state_dict())
        want = fresh(z, t, c_crossattn=[ctx])
    assert torch.equal(got, want), "packs written by cs_adamw_repack differ from a fresh pack of the same weights"
    # parameters written through torch (a checkpoint load) are noticed and re-packed
    with torch.no_grad():
        mb.load_state_dict(ma.state_dict())
        fresh.load_state_dict(ma.state_dict())
        # ... even by an evaluation that never goes through the training step (the registrations carry the parameter versions)
        assert torch.equal(mb(z, t, c_crossattn=[ctx]), fresh(z, t, c_crossattn=[ctx]))
    sb.step(z, ctx, t=t, noise=noise); sa.step(z, ctx, t=t, noise=noise)
    with torch.no_grad():
        fresh.load_state_dict(mb.state_dict())
        assert torch.equal(mb(z, t, c_crossattn=[ctx]), fresh(z, t, c_crossattn=[ctx]))
