"""Parity of the CUDA scene-graph conditioning path with the reference goldens.
All fp32: tolerance 1e-4 relative to the output range (summation-order differences only)."""
import os

import numpy as np
import pytest
import torch

from oracle import graph as G, weights as Wt

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


class E2(torch.nn.Module):
    """The encoder_2 members of Sg2ScVAEModel, built from the product's graph modules with the reference's key names."""

    def __init__(self, cfg):
        super().__init__()
        from commonscenes_b200.model.graph import GraphTripleConvNet2, make_mlp
        e, add = cfg["embedding_dim"], cfg["add_dim"]
        self.obj_embeddings_dc = torch.nn.Embedding(cfg["num_objs"] + 1, e)
        self.pred_embeddings_dc = torch.nn.Embedding(cfg["num_preds"], 2 * e)
        self.gconv_net_ec_rel = GraphTripleConvNet2(input_dim_obj=2 * e + add, input_dim_pred=2 * e + add, hidden_dim=4 * e,
                                                    pooling="avg", num_layers=cfg["num_layers"], mlp_normalization="batch", residual=True)
        self.rel_mlp = make_mlp([2 * e + add, cfg["rel_hidden"], cfg["rel_out"]], batch_norm="batch", norelu=True)
        self.embedding_dim, self.clip, self.use_E2 = e, True, True

    def __getattr__(self, name):      # encoder_2 and its helpers are the product's own (unbound from Sg2ScVAEModel)
        from commonscenes_b200.model.VAEGAN_V2FULL import Sg2ScVAEModel
        if name in ("encoder_2", "encoder_2_train", "encoder_2_backward", "_encoder_2_impl", "_enc2_params"):
            return getattr(Sg2ScVAEModel, name).__get__(self)
        return super().__getattr__(name)


@pytest.mark.parametrize("tag,cfg", [("tiny", G.GCN_TINY), ("full", G.GCN_FULL)])
def test_encoder2_matches_reference_golden(tag, cfg):
    from commonscenes_b200.model.VAEGAN_V2FULL import Sg2ScVAEModel
    g = np.load(os.path.join(GOLD, f"gcn_{tag}.npz"))
    m = E2(cfg)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == G.gcn_param_shapes(cfg)
    z, objs, triples, text, rel = (torch.tensor(g[k]).cuda() for k in ("z", "objs", "triples", "text", "rel"))
    for mode in ("eval", "train"):
        Wt.fill_module_(m, int(g["weight_seed"]))     # train mode updates running stats: reset before each pass
        m = m.cuda().train(mode == "train")
        with torch.no_grad():
            uc, c = m.encoder_2(z, objs, triples, text, rel)
        for got, key in ((c, f"c_{mode}"), (uc, f"uc_{mode}")):
            ref = torch.tensor(g[key])
            err = float((got.cpu() - ref).abs().max())
            print(f"encoder_2[{tag},{mode}] {key}: max abs err {err:.3e} (ref absmax {float(ref.abs().max()):.3f})")
            assert got.shape == ref.shape and err <= 1e-4 * max(1.0, float(ref.abs().max()))


def test_graph_conv_isolated_nodes_and_running_stats():
    from commonscenes_b200.model.graph import GraphTripleConv
    torch.manual_seed(0)
    layer = GraphTripleConv(32, 32, hidden_dim=16, mlp_normalization="batch", residual=True).cuda().train()
    obj = torch.randn(6, 32, device="cuda")
    pred = torch.randn(4, 32, device="cuda")
    edges = torch.tensor([[0, 1], [1, 2], [0, 2], [2, 0]], device="cuda")          # nodes 3..5 appear in no triple
    sd = {k: v.clone().cpu() for k, v in layer.state_dict().items()}
    new_obj, new_pred = layer(obj, pred, edges)
    o_ref, p_ref = G.graph_triple_conv({f"L.{k}": v for k, v in sd.items()}, "L", obj.cpu(), pred.cpu(), edges.cpu(), 16, True)
    assert torch.allclose(new_obj.cpu(), o_ref, atol=1e-4) and torch.allclose(new_pred.cpu(), p_ref, atol=1e-4)
    bn = layer.net1[1]
    assert int(bn.num_batches_tracked) == 1 and float((bn.running_mean.cpu() - sd["net1.1.running_mean"]).abs().max()) > 0
    with pytest.raises(ValueError):
        layer(obj[:2], pred[:1], edges[:1])                                           # BatchNorm1d needs > 1 row in training


@pytest.mark.parametrize("tag,cfg,mode", [("tiny", G.GCN_TINY, "train"), ("tiny", G.GCN_TINY, "eval"), ("full", G.GCN_FULL, "train")])
def test_encoder2_gradients_match_oracle_autograd(tag, cfg, mode):
    """`loss.backward()` through encoder_2 (explicit CUDA backward behind the autograd bridge) gives the gradients the
    reference's autograd computes: every parameter of rel_mlp / gconv_net_ec_rel / both embeddings, and z."""
    g = np.load(os.path.join(GOLD, f"gcn_{tag}.npz"))
    m = E2(cfg)
    Wt.fill_module_(m, int(g["weight_seed"]))
    pnames = {k for k, _ in m.named_parameters()}
    sd = {k: (v.clone().requires_grad_(True) if k in pnames else v.clone()) for k, v in m.state_dict().items()}
    m = m.cuda().train(mode == "train")
    z, objs, triples, text, rel = (torch.tensor(g[k]) for k in ("z", "objs", "triples", "text", "rel"))
    gen = torch.Generator().manual_seed(5)
    w_c = torch.randn(objs.shape[0], 1, cfg["rel_out"], generator=gen)
    w_uc = torch.randn(objs.shape[0], 1, cfg["rel_out"], generator=gen)
    for use_uc in (False, True):                  # training uses c only; the bridge must also handle a gradient into uc
        for t in sd.values():
            t.grad = None
        z_ref = z.clone().requires_grad_(True)
        uc_ref, c_ref = G.encoder_2(sd, cfg, z_ref, objs, triples, text, rel, training=(mode == "train"))
        ((c_ref * w_c).sum() + ((uc_ref * w_uc).sum() if use_uc else 0.0)).backward()
        m.zero_grad(set_to_none=True)
        z_dev = z.cuda().requires_grad_(True)
        uc, c = m.encoder_2(z_dev, objs.cuda(), triples.cuda(), text.cuda(), rel.cuda())
        assert c.requires_grad and float((c.detach().cpu() - c_ref.detach()).abs().max()) <= 1e-4 * max(1.0, float(c_ref.abs().max()))
        ((c * w_c.cuda()).sum() + ((uc * w_uc.cuda()).sum() if use_uc else 0.0)).backward()
        worst = 0.0
        refs = [t.grad for t in sd.values() if t.requires_grad and t.grad is not None]
        rms = (sum(float(r.pow(2).sum()) for r in refs) / sum(r.numel() for r in refs)) ** 0.5
        for name, p in m.named_parameters():
            ref = sd[name].grad
            if ref is None or float(ref.norm()) == 0.0:
                assert p.grad is None or float(p.grad.abs().max()) <= 1e-6, name
                continue
            assert p.grad is not None, f"no gradient produced for {name}"
            err, rn = float((p.grad.cpu() - ref).norm()), float(ref.norm())
            # absolute floor: a Linear bias in front of a train-mode BatchNorm has a mathematically zero gradient (fp32 noise)
            floor = 1e-4 * rms * ref.numel() ** 0.5
            if rn > floor:
                worst = max(worst, err / rn)
            assert err <= 1e-3 * rn + floor, f"{name}: err {err:.3e} vs ref norm {rn:.3e}"
        ez = float((z_dev.grad.cpu() - z_ref.grad).norm() / z_ref.grad.norm())
        print(f"encoder_2 backward [{tag},{mode},uc={use_uc}]: worst parameter-gradient rel-L2 {worst:.2e}, d_z rel-L2 {ez:.2e}")
        assert ez <= 1e-3


def test_gcn_backward_kernels_edge_cases():
    """Isolated nodes (no incident triple), repeated indices in the embedding scatter, a self-loop edge."""
    from commonscenes_b200 import ops, ops_bwd
    torch.manual_seed(3)
    O, T, H, Dm = 6, 5, 8, 4
    edges = torch.tensor([[0, 1], [1, 2], [0, 2], [2, 0], [4, 4]])
    tv = torch.randn(T, 2 * H + Dm, requires_grad=True)
    s_idx, o_idx = edges[:, 0], edges[:, 1]
    pooled = torch.zeros(O, H).index_add(0, s_idx, tv[:, :H]).index_add(0, o_idx, tv[:, H + Dm:])
    cnt = torch.zeros(O).index_add(0, s_idx, torch.ones(T)).index_add(0, o_idx, torch.ones(T)).clamp(min=1)
    w = torch.randn(O, H)
    mid = torch.randn(T, Dm)
    ((pooled / cnt[:, None] * w).sum() + (tv[:, H:H + Dm] * mid).sum()).backward()
    got = ops_bwd.gcn_scatter_mean_bwd(w.cuda(), edges.cuda(), H, mid.cuda())
    assert torch.allclose(got.cpu(), tv.grad, atol=1e-6)
    # gather backward, accumulating on top of existing gradients
    Do, Dp = 3, 2
    obj, pred = torch.randn(O, Do, requires_grad=True), torch.randn(T, Dp, requires_grad=True)
    cat = torch.cat([obj[s_idx], pred, obj[o_idx]], dim=1)
    d_in = torch.randn(T, 2 * Do + Dp)
    (cat * d_in).sum().backward()
    d_obj, d_pred = torch.ones(O, Do).cuda(), torch.ones(T, Dp).cuda()
    ops_bwd.gcn_gather_triples_bwd(d_in.cuda(), edges.cuda(), Do, Dp, d_obj, d_pred, accumulate=True)
    assert torch.allclose(d_obj.cpu(), obj.grad + 1, atol=1e-6) and torch.allclose(d_pred.cpu(), pred.grad + 1, atol=1e-6)
    # embedding scatter with repeats and an unused row
    idx = torch.tensor([2, 0, 2, 2, 5])
    rows = torch.randn(5, 7)
    dW = torch.zeros(6, 4).cuda()
    ops_bwd.embedding_bwd(rows.cuda(), 2, idx.cuda(), dW)
    assert torch.allclose(dW.cpu(), torch.zeros(6, 4).index_add(0, idx, rows[:, 2:6]), atol=1e-6)


@pytest.mark.parametrize("tag", ["tiny", "full"])
def test_layout_branch_forward_matches_reference_class_golden(tag, tmp_path):
    """encoder / manipulate / decoder of Sg2ScVAEModel(layout_branch=True) vs what the reference's REAL class computed
    (tests/golden/layout_*.npz), eval- and train-mode BatchNorm; fp32 kernels: 1e-4 relative to the output range."""
    import yaml
    from oracle import layout as Lo
    from commonscenes_b200.model.VAEGAN_V2FULL import Sg2ScVAEModel
    from commonscenes_b200.model.sdfusion_txt2shape_model import default_opt
    cfg = Lo.LAYOUT_TINY if tag == "tiny" else Lo.LAYOUT_FULL
    g = np.load(os.path.join(GOLD, f"layout_{tag}.npz"))
    df = dict(model=dict(params=dict(linear_start=0.00085, linear_end=0.012, conditioning_key="crossattn", timesteps=1000)),
              unet=dict(params=dict(image_size=8, in_channels=3, out_channels=3, model_channels=32, num_res_blocks=1,
                                    attention_resolutions=[4, 2], channel_mult=[1, 2, 3], num_heads=4, dims=3, use_spatial_transformer=True,
                                    transformer_depth=1, context_dim=1280, use_checkpoint=True, legacy=False)))
    vq = dict(model=dict(params=dict(embed_dim=3, n_embed=64, ddconfig=dict(double_z=False, z_channels=3, resolution=16, in_channels=1, out_ch=1,
                                                                           ch=16, ch_mult=[1, 2], num_res_blocks=1, attn_resolutions=[], dropout=0.0))))
    (tmp_path / "df.yaml").write_text(yaml.safe_dump(df)); (tmp_path / "vq.yaml").write_text(yaml.safe_dump(vq))
    vocab = {"object_idx_to_name": [f"o{i}" for i in range(cfg["num_objs"])], "pred_idx_to_name": [f"p{i}" for i in range(cfg["num_preds"])]}
    m = Sg2ScVAEModel(vocab, diff_opt=default_opt(device="cuda", df_cfg=str(tmp_path / "df.yaml"), vq_cfg=str(tmp_path / "vq.yaml")),
                      embedding_dim=cfg["embedding_dim"], mlp_normalization="batch", residual=True, gconv_num_layers=cfg["num_layers"],
                      layout_branch=True).cuda()
    shapes = Lo.layout_param_shapes(cfg)
    z, objs, triples, text, rel, boxes, angles, zz = (torch.tensor(g[k]).cuda() for k in ("z", "objs", "triples", "text", "rel", "boxes", "angles", "zz"))

    def reset():        # train mode updates the running statistics: restore the seeded state before every call
        m.load_state_dict({k: Wt.synth_tensor(int(g["weight_seed"]), k, tuple(s)) for k, s in shapes.items()}, strict=False)

    def check(got, key):
        ref = torch.tensor(g[key])
        err = float((got.cpu() - ref).abs().max())
        print(f"layout[{tag}] {key}: max abs err {err:.3e} (ref absmax {float(ref.abs().max()):.3f})")
        assert got.shape == ref.shape and err <= 1e-4 * max(1.0, float(ref.abs().max()))
    for mode in ("eval", "train"):
        m.train(mode == "train")
        reset(); mu, logvar = m.encoder(objs, triples, boxes, None, text, rel, angles)
        check(mu, f"mu_{mode}"); check(logvar, f"logvar_{mode}")
        reset(); check(m.manipulate(zz, objs, triples, text, rel), f"man_{mode}")
        reset(); b, a = m.decoder(z, objs, triples, text, rel)
        check(b, f"boxes_{mode}"); check(a, f"angle_logp_{mode}")


def test_layout_branch_gradients_match_oracle_autograd(tmp_path):
    """`loss.backward()` through encoder -> reparameterise -> decoder (+ manipulate) of Sg2ScVAEModel(layout_branch=True) on the
    GPU vs autograd through the oracle (pinned to the real class): every layout parameter gradient, train-mode BatchNorm."""
    import yaml
    from oracle import layout as Lo
    from commonscenes_b200.model.VAEGAN_V2FULL import Sg2ScVAEModel
    from commonscenes_b200.model.sdfusion_txt2shape_model import default_opt
    cfg = Lo.LAYOUT_TINY
    g = np.load(os.path.join(GOLD, "layout_tiny.npz"))
    df = dict(model=dict(params=dict(linear_start=0.00085, linear_end=0.012, conditioning_key="crossattn", timesteps=1000)),
              unet=dict(params=dict(image_size=8, in_channels=3, out_channels=3, model_channels=32, num_res_blocks=1,
                                    attention_resolutions=[4, 2], channel_mult=[1, 2, 3], num_heads=4, dims=3, use_spatial_transformer=True,
                                    transformer_depth=1, context_dim=1280, use_checkpoint=True, legacy=False)))
    vq = dict(model=dict(params=dict(embed_dim=3, n_embed=64, ddconfig=dict(double_z=False, z_channels=3, resolution=16, in_channels=1, out_ch=1,
                                                                           ch=16, ch_mult=[1, 2], num_res_blocks=1, attn_resolutions=[], dropout=0.0))))
    (tmp_path / "df.yaml").write_text(yaml.safe_dump(df)); (tmp_path / "vq.yaml").write_text(yaml.safe_dump(vq))
    vocab = {"object_idx_to_name": [f"o{i}" for i in range(cfg["num_objs"])], "pred_idx_to_name": [f"p{i}" for i in range(cfg["num_preds"])]}
    m = Sg2ScVAEModel(vocab, diff_opt=default_opt(device="cuda", df_cfg=str(tmp_path / "df.yaml"), vq_cfg=str(tmp_path / "vq.yaml")),
                      embedding_dim=64, mlp_normalization="batch", residual=True, gconv_num_layers=cfg["num_layers"], layout_branch=True)
    shapes = Lo.layout_param_shapes(cfg)
    sd0 = Wt.synth_state_dict(shapes, 55)
    m.load_state_dict(sd0, strict=False)
    m = m.cuda().train()
    pnames = {k for k, _ in m.named_parameters()}
    sd = {k: (v.clone().requires_grad_(True) if k in pnames else v.clone()) for k, v in sd0.items()}
    z, objs, triples, text, rel, boxes, angles, zz = (torch.tensor(g[k]) for k in ("z", "objs", "triples", "text", "rel", "boxes", "angles", "zz"))
    eps = torch.randn(objs.shape[0], 64, generator=torch.Generator().manual_seed(1))

    def loss_of(enc, dec, man, dev):
        mu, logvar = enc()
        b, a = dec(eps.to(dev) * torch.exp(0.5 * logvar) + mu)
        tot, _ = Lo.layout_losses(b, boxes.to(dev), a, angles.to(dev), mu, logvar, 0.1)
        return tot + man().pow(2).mean()
    ref = loss_of(lambda: Lo.encoder(sd, cfg, objs, triples, boxes, text, rel, angles, True),
                  lambda zs: Lo.decoder(sd, cfg, zs, objs, triples, text, rel, True),
                  lambda: Lo.manipulate(sd, cfg, zz, objs, triples, text, rel, True), "cpu")
    ref.backward()
    c = lambda t: t.cuda()
    got = loss_of(lambda: m.encoder(c(objs), c(triples), c(boxes), None, c(text), c(rel), c(angles)),
                  lambda zs: m.decoder(zs, c(objs), c(triples), c(text), c(rel)),
                  lambda: m.manipulate(c(zz), c(objs), c(triples), c(text), c(rel)), "cuda")
    assert abs(float(got.detach()) - float(ref.detach())) <= 1e-4 * abs(float(ref.detach()))
    got.backward()
    named = dict(m.named_parameters())
    refs = [t.grad for k, t in sd.items() if k in pnames and t.grad is not None]
    rms = (sum(float(r.pow(2).sum()) for r in refs) / sum(r.numel() for r in refs)) ** 0.5
    worst = 0.0
    for k in shapes:
        if k not in pnames or sd[k].grad is None or float(sd[k].grad.norm()) == 0.0:
            continue
        r, p = sd[k].grad, named[k].grad
        assert p is not None, f"no gradient for {k}"
        err, rn = float((p.cpu() - r).norm()), float(r.norm())
        floor = 1e-4 * rms * r.numel() ** 0.5
        if rn > floor:
            worst = max(worst, err / rn)
        assert err <= 1e-3 * rn + floor, f"{k}: err {err:.3e} vs ref norm {rn:.3e}"
    print(f"layout backward on the GPU: worst parameter-gradient rel-L2 {worst:.2e}")


def test_box_discriminator_on_gpu_matches_reference_golden():
    from commonscenes_b200.model.discriminators import BoxDiscriminator
    g = np.load(os.path.join(GOLD, "box_discriminator.npz"))
    d = BoxDiscriminator(6, 16, 36)
    objs, triples, boxes, keep = (torch.tensor(g[k]).cuda() for k in ("objs", "triples", "boxes", "keep"))
    modes = {"plain": dict(), "keeps": dict(keeps=keep), "real": dict(with_grad=True, is_real=True), "fake_keeps": dict(keeps=keep, with_grad=True, is_real=False)}
    for name, kw in modes.items():
        Wt.fill_module_(d, int(g["weight_seed"]))
        d = d.cuda().train()
        d.zero_grad()
        y, reg = d(objs, triples, boxes.clone(), **kw)
        (y.mean() + (reg.mean() if reg is not None else 0.0)).backward()
        assert np.allclose(y.detach().cpu().numpy(), g[f"{name}_y"], atol=1e-5)
        if reg is not None:
            assert np.allclose(reg.detach().cpu().numpy(), g[f"{name}_reg"], rtol=1e-3, atol=1e-6)
        for k, p in d.named_parameters():
            if f"{name}_grad_{k}" in g.files:
                ref = g[f"{name}_grad_{k}"]
                assert float(np.linalg.norm(p.grad.cpu().numpy() - ref)) <= 1e-3 * float(np.linalg.norm(ref)) + 1e-5 * ref.size ** 0.5, (name, k)
            else:
                nrm = float(g[f"{name}_gsum_{k}"][0])
                assert abs(float(p.grad.norm()) - nrm) <= 1e-3 * nrm, (name, k)
