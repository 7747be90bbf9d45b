"""BASELINE cfg3: one v2_full shape-branch training iteration (scene graphs -> GCN-E2 + rel_mlp -> frozen VQ-VAE encode ->
denoiser forward / backward -> gradient back into the graph networks -> clip + AdamW on both groups).

Two routes through the product are compared on identical inputs: the reference-style loop (module calls +
`loss.backward()` through the autograd bridges + torch clip_grad_norm_ / AdamW, i.e. what train_3dfront.py:387-418 does
with the drop-in classes) and `ShapeBranchTrainStep` (flat buffers, one clip + AdamW launch per group).  The gradient
kernels themselves are pinned to the oracle's autograd in test_unet_train_gpu.py / test_gcn_gpu.py."""
import pytest
import torch

from oracle import weights as Wt

pytestmark = pytest.mark.gpu


def _scene_batch(n_scenes, objs_per_scene, triples_per_scene, n_classes, n_preds, seed):
    """Synthetic collate_fn_vaegan batch (threedfront_dataset.py:774-781): per-scene index offsets on the triples."""
    g = torch.Generator().manual_seed(seed)
    objs, triples, scene = [], [], []
    for s in range(n_scenes):
        off = s * objs_per_scene
        objs.append(torch.randint(1, n_classes, (objs_per_scene,), generator=g))
        sub = torch.randint(0, objs_per_scene, (triples_per_scene,), generator=g)
        ob = (sub + torch.randint(1, objs_per_scene, (triples_per_scene,), generator=g)) % objs_per_scene     # s != o
        triples.append(torch.stack([sub + off, torch.randint(1, n_preds, (triples_per_scene,), generator=g), ob + off], dim=1))
        scene.append(torch.full((objs_per_scene,), s))
    objs, triples = torch.cat(objs), torch.cat(triples)
    O, T = objs.shape[0], triples.shape[0]
    return dict(objs=objs, triples=triples, scene=torch.cat(scene), text=torch.randn(O, 512, generator=g),
                rel=torch.randn(T, 512, generator=g), z=torch.randn(O, 64, generator=g),
                sdfs=(torch.randn(O, 1, 64, 64, 64, generator=g) * 0.1).clamp(-0.2, 0.2))


def _model(seed):
    from commonscenes_b200.model.VAEGAN_V2FULL import Sg2ScVAEModel
    vocab = {"object_idx_to_name": [f"o{i}" for i in range(36)], "pred_idx_to_name": [f"p{i}" for i in range(16)]}
    m = Sg2ScVAEModel(vocab, embedding_dim=64, mlp_normalization="batch", residual=True, gconv_num_layers=5)
    for mod in (m.obj_embeddings_dc, m.pred_embeddings_dc, m.gconv_net_ec_rel, m.rel_mlp, m.Diff.df, m.Diff.vqvae):
        Wt.fill_module_(mod, seed)
    return m.cuda().train()


def test_shape_branch_step_matches_autograd_route():
    from commonscenes_b200.train import ShapeBranchTrainStep
    batch = {k: v.cuda() for k, v in _scene_batch(2, 4, 6, 36, 16, seed=3).items()}
    O = batch["objs"].shape[0]
    g = torch.Generator().manual_seed(4)
    t = torch.randint(0, 1000, (O,), generator=g).cuda()
    noise = torch.randn(O, 3, 16, 16, 16, generator=g).cuda()

    # --- route A: drop-in modules + autograd bridges + torch optimizers ---
    ma = _model(77)
    graph_params = ma._enc2_params()
    uc, c = ma.encoder_2(batch["z"], batch["objs"], batch["triples"], batch["text"], batch["rel"])
    assert c.requires_grad
    with torch.no_grad():
        lat = ma.Diff.vqvae(batch["sdfs"], forward_no_quant=True, encode_only=True)
    _, _, loss_a, _ = ma.Diff.p_losses(lat, c, t, noise=noise)
    (100.0 * loss_a).backward()
    # the last GCN layer's predicate projection feeds nothing: grad None there, exactly as under the reference's autograd
    assert [n for n, p in ma.named_parameters() if p.grad is None and not n.startswith("Diff")] == \
        ["gconv_net_ec_rel.gconvs.4.linear_projection_pred.weight", "gconv_net_ec_rel.gconvs.4.linear_projection_pred.bias"]
    ga = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).flatten() for p in graph_params])
    gda = torch.cat([p.grad.flatten() for p in ma.Diff.df.parameters() if p.grad is not None])
    assert float(ga.norm()) > 0

    # --- route B: the native step ---
    mb = _model(77)
    step = ShapeBranchTrainStep(mb)
    p0 = step.graph_params.flat_p.clone()
    loss_b, d_z = step.step(batch["z"], batch["objs"], batch["triples"], batch["text"], batch["rel"], batch["sdfs"], t=t, noise=noise)
    assert abs(loss_a.item() - loss_b.item()) / loss_a.item() < 1e-2
    gb = torch.cat([step.graph_params.views[p].flatten() for p in step.graph_params.params])
    rel = float((ga - gb).norm() / ga.norm())
    cos = float((ga * gb).sum() / (ga.norm() * gb.norm()))
    print(f"graph-side gradient, native vs autograd route: rel-L2 {rel:.3e}, cosine {cos:.6f}; |g| = {float(ga.norm()):.3e}")
    assert rel < 3e-2 and cos > 0.999          # same kernels; differences = bf16 atomics order inside the denoiser backward
    assert d_z.shape == batch["z"].shape and torch.isfinite(d_z).all()
    assert float((step.graph_params.flat_p - p0).abs().max()) > 0       # AdamW moved the graph parameters
    assert int(mb.rel_mlp[1].num_batches_tracked) == 2                     # BatchNorm saw c and uc, as in the reference
    # state-dict keys are the reference's after the flat re-homing
    assert {k: tuple(v.shape) for k, v in ma.state_dict().items()} == {k: tuple(v.shape) for k, v in mb.state_dict().items()}
    del gda


def test_shape_branch_loss_decreases_over_steps():
    """Overfit one tiny batch for a few iterations with fixed (t, noise): the loss must go down (end-to-end sanity of
    every gradient and both optimizers)."""
    from commonscenes_b200.train import ShapeBranchTrainStep
    batch = {k: v.cuda() for k, v in _scene_batch(1, 4, 6, 36, 16, seed=8).items()}
    g = torch.Generator().manual_seed(5)
    t = torch.randint(100, 900, (4,), generator=g).cuda()
    noise = torch.randn(4, 3, 16, 16, 16, generator=g).cuda()
    step = ShapeBranchTrainStep(_model(78), lr=2e-4)
    losses = []
    for _ in range(6):
        loss, _ = step.step(batch["z"], batch["objs"], batch["triples"], batch["text"], batch["rel"], batch["sdfs"], t=t, noise=noise)
        losses.append(loss.item())
    print("losses:", [f"{l:.4f}" for l in losses])
    assert losses[-1] < losses[0]


def test_graphed_whole_iteration_equals_eager():
    """ShapeBranchTrainStep.capture / step_graphed: the whole cfg3 iteration as ONE CUDA graph.  Two models with identical
    weights take two steps each on identical inputs, one eagerly and one through the graph: same losses (the forward is
    deterministic), parameters equal up to the order of the bf16 / fp32 atomics inside the denoiser's weight gradients, and
    the BatchNorm running statistics -- touched by the warm-up iterations, then rolled back -- identical."""
    from commonscenes_b200.train import ShapeBranchTrainStep
    batches = [{k: v.cuda() for k, v in _scene_batch(2, 4, 6, 36, 16, seed=s).items()} for s in (11, 12)]
    O, T = batches[0]["objs"].shape[0], batches[0]["triples"].shape[0]
    g = torch.Generator().manual_seed(9)
    ts = [torch.randint(0, 1000, (O,), generator=g).cuda() for _ in range(2)]
    noises = [torch.randn(O, 3, 16, 16, 16, generator=g).cuda() for _ in range(2)]
    me, mg = _model(55), _model(55)
    eager, graphed = ShapeBranchTrainStep(me), ShapeBranchTrainStep(mg)
    graphed.capture(O, T)
    assert torch.equal(eager.graph_params.flat_p, graphed.graph_params.flat_p)          # warm-up rolled back
    assert torch.equal(eager.denoiser.flat_p, graphed.denoiser.flat_p)
    for (n1, b1), (n2, b2) in zip(me.named_buffers(), mg.named_buffers()):
        assert n1 == n2 and torch.equal(b1, b2), n1
    for i, b in enumerate(batches):
        le, dze = eager.step(b["z"], b["objs"], b["triples"], b["text"], b["rel"], b["sdfs"], t=ts[i], noise=noises[i])
        lg, dzg = graphed.step_graphed(b["z"], b["objs"], b["triples"], b["text"], b["rel"], b["sdfs"], t=ts[i], noise=noises[i])
        le, lg = float(le), float(lg)
        print(f"step {i}: eager loss {le:.6f}, graphed loss {lg:.6f}")
        # step 0 starts from identical weights: same loss (deterministic forward); gradients differ only by the order of the
        # atomics in the denoiser's weight / context gradients (two EAGER runs differ by 1.7e-3 in d_z too, tools/dbg_graph.py).
        # Afterwards AdamW's normalised update amplifies those differences (measured: 10 % in d_z at step 1), so only the
        # loss is compared there.
        assert abs(le - lg) <= (1e-6 if i == 0 else 1e-2) * abs(le)
        if i == 0:
            assert float((dze - dzg).norm()) <= 1e-2 * float(dze.norm()) + 1e-8
            pe, pg = eager.graph_params.flat_p, graphed.graph_params.flat_p
            assert float((pe - pg).norm()) <= 1e-3 * float(pe.norm())
            de, dg = eager.denoiser.flat_p, graphed.denoiser.flat_p
            assert float((de - dg).norm()) <= 1e-3 * float(de.norm())
        assert torch.isfinite(dzg).all()
    assert int(eager.denoiser.step_dev) == int(graphed.denoiser.step_dev) == 2
    assert int(mg.rel_mlp[1].num_batches_tracked) == int(me.rel_mlp[1].num_batches_tracked) == 4
