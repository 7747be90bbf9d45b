"""Pin the oracle against the reference's own modules (run in the BUILD container only: it imports
/root/reference, which does not exist on the GPU box).

    python oracle/validate_against_reference.py [--full]

For each sub-system it (1) checks that the oracle's state-dict key/shape inventory equals the reference
module's state_dict() exactly — that inventory is the drop-in contract of SURVEY.md §8b — and (2) runs the
reference module and the oracle on identical seeded weights/inputs and reports the max abs difference
(expected ~1e-6: same fp32 torch ops, possibly different association).  --full adds the full-size UNet.
Exit code 0 iff everything is within 2e-5 relative.
"""
from __future__ import annotations

import argparse
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("CS_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)


def import_reference():
    """Make `model.*` of the reference importable (one shim: a stub omegaconf.listconfig)."""
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} not found: this script only runs where the reference is mounted")
    if "omegaconf" not in sys.modules:
        oc = types.ModuleType("omegaconf")
        lc = types.ModuleType("omegaconf.listconfig")

        class ListConfig(list):
            pass
        lc.ListConfig = ListConfig
        oc.listconfig = lc
        sys.modules["omegaconf"] = oc
        sys.modules["omegaconf.listconfig"] = lc
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from model.networks.diffusion_networks.network import DiffusionUNet
    from model.networks.vqvae_networks.network import VQVAE
    from model import graph as ref_graph
    return DiffusionUNet, VQVAE, ref_graph


def ref_unet(cfg, seed):
    from oracle import weights
    DiffusionUNet, _, _ = import_reference()
    params = dict(cfg)
    params["attention_resolutions"] = list(cfg["attention_resolutions"])
    params["channel_mult"] = list(cfg["channel_mult"])
    params.update(use_spatial_transformer=True, use_checkpoint=False, legacy=False)
    m = DiffusionUNet(params, conditioning_key="crossattn").eval()
    weights.fill_module_(m, seed)
    return m


def ref_unet_concat(cfg, seed):
    """config/sdfusion-txt2shape_concat.yaml: AttentionBlock variant, conditioning_key='concat' (SURVEY.md §8f rank 1)."""
    from oracle import weights
    DiffusionUNet, _, _ = import_reference()
    params = {k: v for k, v in cfg.items()}
    params["attention_resolutions"] = list(cfg["attention_resolutions"])
    params["channel_mult"] = list(cfg["channel_mult"])
    params.update(use_spatial_transformer=False, context_dim=None, use_checkpoint=False, legacy=False)
    m = DiffusionUNet(params, conditioning_key="concat").eval()
    weights.fill_module_(m, seed)
    return m


def ref_vqvae(cfg, seed):
    from oracle import weights
    _, VQVAE, _ = import_reference()
    dd = dict(double_z=False, z_channels=cfg["z_channels"], resolution=cfg["resolution"], in_channels=cfg["in_channels"],
              out_ch=cfg["out_ch"], ch=cfg["ch"], ch_mult=list(cfg["ch_mult"]), num_res_blocks=cfg["num_res_blocks"],
              attn_resolutions=[], dropout=0.0)
    m = VQVAE(dd, cfg["n_embed"], cfg["embed_dim"]).eval()
    weights.fill_module_(m, seed)
    return m


class RefE2(torch.nn.Module):
    """The encoder_2 slice of Sg2ScVAEModel (VAEGAN_V2FULL.py:69-75, 128-155, 220-242) built from the
    reference's own GraphTripleConvNet2 / make_mlp (the full class needs omegaconf/pytorch3d to import)."""

    def __init__(self, cfg):
        super().__init__()
        _, _, g = import_reference()
        e, add = cfg["embedding_dim"], cfg["add_dim"]
        self.obj_embeddings_dc = torch.nn.Embedding(cfg["num_objs"] + 1, e)
        self.pred_embeddings_dc = torch.nn.Embedding(cfg["num_preds"], 2 * e)
        self.gconv_net_ec_rel = g.GraphTripleConvNet2(input_dim_obj=2 * e + add, input_dim_pred=2 * e + add, hidden_dim=4 * e,
                                                      pooling="avg", num_layers=cfg["num_layers"], mlp_normalization="batch",
                                                      residual=True)
        self.rel_mlp = g.make_mlp([2 * e + add, cfg["rel_hidden"], cfg["rel_out"]], batch_norm="batch", norelu=True)

    def forward(self, z, objs, triples, text_feat, rel_feat):
        s, p, o = [x.squeeze(1) for x in triples.chunk(3, dim=1)]
        edges = torch.stack([s, o], dim=1)
        obj_vecs_ = torch.cat([text_feat, self.obj_embeddings_dc(objs)], dim=1)
        pred_vecs_ = torch.cat([rel_feat, self.pred_embeddings_dc(p)], dim=1)
        rel_vecs_ = torch.cat([obj_vecs_, z], dim=1)
        rel2, _ = self.gconv_net_ec_rel(rel_vecs_, pred_vecs_, edges)
        return self.rel_mlp(rel_vecs_).unsqueeze(1), self.rel_mlp(rel2).unsqueeze(1)


def synth_graph(cfg, n_obj, n_tri, seed):
    g = torch.Generator().manual_seed(seed)
    objs = torch.randint(1, cfg["num_objs"], (n_obj,), generator=g)
    s = torch.randint(0, n_obj, (n_tri,), generator=g)
    o = (s + 1 + torch.randint(0, n_obj - 1, (n_tri,), generator=g)) % n_obj
    p = torch.randint(1, cfg["num_preds"], (n_tri,), generator=g)
    triples = torch.stack([s, p, o], dim=1)
    text = torch.randn(n_obj, cfg["add_dim"], generator=g)
    rel = torch.randn(n_tri, cfg["add_dim"], generator=g)
    z = torch.randn(n_obj, cfg["embedding_dim"], generator=g)
    return z, objs, triples, text, rel


def _cmp(name, a, b, tol=2e-5):
    err = (a - b).abs().max().item()
    scale = b.abs().max().item()
    ok = err <= tol * max(scale, 1.0)
    print(f"{'ok ' if ok else 'BAD'} {name}: max|oracle-ref|={err:.3e} (ref absmax {scale:.3e})")
    return ok


def _keys(name, shapes, module):
    ref = {k: tuple(v.shape) for k, v in module.state_dict().items()}
    ok = ref == dict(shapes)
    if not ok:
        missing = sorted(set(ref) - set(shapes))[:5]
        extra = sorted(set(shapes) - set(ref))[:5]
        diff = [k for k in ref if k in shapes and ref[k] != tuple(shapes[k])][:5]
        print(f"BAD {name} keys: missing {missing} extra {extra} shape-mismatch {diff}")
    else:
        print(f"ok  {name}: {len(ref)} state-dict keys/shapes identical to the reference")
    return ok


@torch.no_grad()
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="also run the full-size UNet / VQ-VAE (slow)")
    args = ap.parse_args()
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    from oracle import denoiser as D, graph as G, vqvae as V, weights as Wt
    ok = True

    for tag, cfg in [("tiny", D.UNET_TINY)] + ([("full", D.UNET_FULL)] if args.full else []):
        m = ref_unet(cfg, seed=1)
        shapes = D.unet_param_shapes(cfg)
        ok &= _keys(f"unet[{tag}]", shapes, m)
        sd = Wt.synth_state_dict(shapes, seed=1)
        r = cfg["image_size"]
        g = torch.Generator().manual_seed(2)
        x = torch.randn(2, cfg["in_channels"], r, r, r, generator=g)
        t = torch.tensor([500, 37])
        ctx = torch.randn(2, 1, cfg["context_dim"], generator=g)
        ok &= _cmp(f"unet_forward[{tag}]", D.unet_forward(sd, cfg, x, t, ctx), m(x, t, c_crossattn=[ctx]))
        if tag == "tiny":
            sys.path.insert(0, REF)
            from model.networks.diffusion_networks import ldm_diffusion_util as U
            sched = D.register_schedule(**D.DIFFUSION)
            betas = U.make_beta_schedule("linear", 1000, linear_start=0.00085, linear_end=0.012)
            ok &= _cmp("betas", sched["betas"], torch.tensor(betas, dtype=torch.float32), 0)
            dd = D.ddim_schedule(sched, 100)
            ts = U.make_ddim_timesteps("uniform", 100, 1000, verbose=False)
            sig, al, alp = U.make_ddim_sampling_parameters(sched["alphas_cumprod"], ts, 0.0, verbose=False)
            ok &= bool((ts == dd["timesteps"]).all())
            ok &= _cmp("ddim_alphas", torch.tensor(dd["alphas"]), torch.as_tensor(al), 0)
            ok &= _cmp("ddim_alphas_prev", torch.tensor(dd["alphas_prev"]), torch.as_tensor(alp).float(), 0)
            ok &= _cmp("timestep_embedding", D.timestep_embedding(t, 224), U.timestep_embedding(t, 224), 0)

    for tag, cfg in [("tiny", D.UNET_CONCAT_TINY)] + ([("full", D.UNET_CONCAT_FULL)] if args.full else []):
        m = ref_unet_concat(cfg, seed=7)
        shapes = D.unet_param_shapes(cfg)
        ok &= _keys(f"unet_concat[{tag}]", shapes, m)
        sd = Wt.synth_state_dict(shapes, seed=7)
        r = cfg["image_size"]
        g = torch.Generator().manual_seed(8)
        x = torch.randn(2, 3, r, r, r, generator=g)
        cc = torch.randn(2, cfg["in_channels"] - 3, r, r, r, generator=g)
        t = torch.tensor([900, 4])
        ok &= _cmp(f"unet_forward_concat[{tag}]", D.unet_forward(sd, cfg, x, t, c_concat=cc), m(x, t, c_concat=[cc]))

    for tag, cfg in [("tiny", V.VQ_TINY)] + ([("full", V.VQ_FULL)] if args.full else []):
        m = ref_vqvae(cfg, seed=3)
        shapes = V.vq_param_shapes(cfg)
        ok &= _keys(f"vqvae[{tag}]", shapes, m)
        sd = Wt.synth_state_dict(shapes, seed=3)
        g = torch.Generator().manual_seed(4)
        r = cfg["resolution"]
        x = (torch.randn(1, 1, r, r, r, generator=g) * 0.1).clamp(-0.2, 0.2)
        z_ref = m(x, forward_no_quant=True, encode_only=True)
        z = V.encode_no_quant(sd, cfg, x)
        ok &= _cmp(f"vq_encode_no_quant[{tag}]", z, z_ref)
        ok &= _cmp(f"vq_decode_no_quant[{tag}]", V.decode_no_quant(sd, cfg, z_ref), m.decode_no_quant(z_ref))
        zq_ref, _, (_, _, idx_ref) = m.quantize(z_ref, is_voxel=True)
        zq, idx = V.quantize(sd, z_ref)
        ok &= bool((idx == idx_ref).all()) and _cmp(f"vq_quantize[{tag}]", zq, zq_ref, 0)

    for tag, cfg in [("tiny", G.GCN_TINY), ("full", G.GCN_FULL)]:
        m = RefE2(cfg)
        shapes = G.gcn_param_shapes(cfg)
        ok &= _keys(f"encoder_2[{tag}]", shapes, m)
        Wt.fill_module_(m, seed=5)
        sd = Wt.synth_state_dict(shapes, seed=5)
        inp = synth_graph(cfg, 9, 20, seed=6)
        for training in (False, True):
            m.train(training)
            uc_ref, c_ref = m(*inp)
            uc, c = G.encoder_2(sd, cfg, *inp, training=training)
            ok &= _cmp(f"encoder_2.c[{tag},train={training}]", c, c_ref) and _cmp(f"encoder_2.uc[{tag},train={training}]", uc, uc_ref)
    # ---- the REAL Sg2ScVAEModel class (stubbed third-party imports, oracle/reference_scene_model.py): encoder_2 wiring ----
    from oracle import reference_scene_model as RS
    for tag, cfg in [("tiny", dict(G.GCN_TINY, add_dim=512, rel_hidden=960, rel_out=1280)), ("full", G.GCN_FULL)]:
        # (the class hard-codes CLIP width 512 and rel_mlp 960 -> 1280, VAEGAN_V2FULL.py:64-66,152)
        real = RS.build(cfg, seed=9)
        shapes = G.gcn_param_shapes(cfg)
        rsd = {k: tuple(v.shape) for k, v in RS.module_state_dict(real).items()}
        sub = {k: rsd.get(k) for k in shapes}
        if sub != {k: tuple(v) for k, v in shapes.items()}:
            bad = [k for k in shapes if rsd.get(k) != tuple(shapes[k])][:5]
            print(f"BAD Sg2ScVAEModel[{tag}] shape-branch keys: {bad}")
            ok = False
        else:
            print(f"ok  Sg2ScVAEModel[{tag}]: all {len(shapes)} shape-branch keys/shapes present in the real class ({len(rsd)} keys in total)")
        sd = Wt.synth_state_dict(shapes, seed=10)
        torch.nn.Module.load_state_dict(real, sd, strict=False)
        inp = synth_graph(cfg, 9, 20, seed=11)
        for training in (False, True):
            real.train(training)
            uc_ref, c_ref = real.encoder_2(*inp, None)
            uc, c = G.encoder_2(sd, cfg, *inp, training=training)
            ok &= _cmp(f"Sg2ScVAEModel.encoder_2.c[{tag},train={training}]", c, c_ref)
            ok &= _cmp(f"Sg2ScVAEModel.encoder_2.uc[{tag},train={training}]", uc, uc_ref)
    # ---- layout branch of the REAL class (oracle/layout.py): encoder, manipulate, decoder, losses ----
    from oracle import layout as Lo
    sys.path.insert(0, REF)
    for tag, lcfg in [("tiny", Lo.LAYOUT_TINY), ("full", Lo.LAYOUT_FULL)]:
        real = RS.build(dict(lcfg, rel_hidden=960, rel_out=1280), seed=12)
        shapes = Lo.layout_param_shapes(lcfg)
        rsd = {k: tuple(v.shape) for k, v in RS.module_state_dict(real).items()}
        bad = [k for k in shapes if rsd.get(k) != tuple(shapes[k])]
        extra = sorted(k for k in rsd if k not in shapes and k not in G.gcn_param_shapes(dict(lcfg, rel_hidden=960, rel_out=1280)))
        if bad or extra:
            print(f"BAD layout[{tag}] keys: mismatch {bad[:5]} unaccounted {extra[:5]}")
            ok = False
        else:
            print(f"ok  layout[{tag}]: {len(shapes)} layout-branch keys/shapes identical; with the shape branch they account for all {len(rsd)} keys of the class")
        lsd = Wt.synth_state_dict(shapes, seed=13)
        torch.nn.Module.load_state_dict(real, lsd, strict=False)
        z_, objs_, triples_, text_, rel_ = synth_graph(dict(lcfg), 9, 20, seed=14)
        gg = torch.Generator().manual_seed(15)
        boxes = torch.randn(9, 6, generator=gg)
        angles = torch.randint(0, 24, (9,), generator=gg)
        zz = torch.randn(9, 2 * lcfg["embedding_dim"], generator=gg)
        for training in (False, True):
            real.train(training)
            mu_r, lv_r = real.encoder(objs_, triples_, boxes, None, text_, rel_, angles)
            mu_o, lv_o = Lo.encoder(lsd, lcfg, objs_, triples_, boxes, text_, rel_, angles, training)
            ok &= _cmp(f"layout.encoder.mu[{tag},train={training}]", mu_o, mu_r) and _cmp("  logvar", lv_o, lv_r)
            ok &= _cmp(f"layout.manipulate[{tag},train={training}]", Lo.manipulate(lsd, lcfg, zz, objs_, triples_, text_, rel_, training),
                       real.manipulate(zz, objs_, triples_, text_, rel_, None))
            b_r, a_r = real.decoder(z_, objs_, triples_, text_, rel_, None)
            b_o, a_o = Lo.decoder(lsd, lcfg, z_, objs_, triples_, text_, rel_, training)
            ok &= _cmp(f"layout.decoder.boxes[{tag},train={training}]", b_o, b_r) and _cmp("  angle log-probs", a_o, a_r)
        from model.losses import calculate_model_losses

        class _W:
            def add_scalar(self, *a, **k):
                pass
        tot_r, _ = calculate_model_losses(None, b_r, boxes, "box", angles=angles, angles_pred=a_r, mu=mu_r, logvar=lv_r, KL_weight=0.1,
                                          writer=_W(), counter=0, withangles=True)
        tot_o, _ = Lo.layout_losses(b_o, boxes, a_o, angles, mu_o, lv_o, 0.1)
        ok &= _cmp(f"layout.losses[{tag}]", tot_o, tot_r)

    # ---- the REAL SDFusionText2ShapeModel class on CPU (oracle/reference_diffusion_model.py): schedule, q_sample, p_losses,
    #      forward() = frozen VQ-VAE encode -> randint t -> randn noise -> p_losses (sdfusion_txt2shape_model.py:184-365) ----
    from oracle import reference_diffusion_model as RD
    real = RD.build(D.UNET_TINY, V.VQ_TINY, seed_unet=21, seed_vq=22)
    sd = Wt.synth_state_dict(D.unet_param_shapes(D.UNET_TINY), seed=21)
    vsd = Wt.synth_state_dict(V.vq_param_shapes(V.VQ_TINY), seed=22)
    sched = D.register_schedule(**D.DIFFUSION)
    for k in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
              "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_variance",
              "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2", "lvlb_weights"):
        ok &= _cmp(f"SDFusionText2ShapeModel.{k}", sched[k], getattr(real, k), 0)
    assert tuple(real.z_shape) == (3, 4, 4, 4) and real.num_timesteps == 1000
    g = torch.Generator().manual_seed(23)
    x0 = torch.randn(3, 3, 4, 4, 4, generator=g)            # (the tiny VQ-VAE has 16^3 -> 4^3 latents; the tiny UNet runs any size)
    x0 = torch.randn(3, 3, 8, 8, 8, generator=g)
    noise = torch.randn(3, 3, 8, 8, 8, generator=g)
    cond = torch.randn(3, 1, D.UNET_TINY["context_dim"], generator=g)
    t = torch.tensor([0, 517, 999])
    ok &= _cmp("SDFusionText2ShapeModel.q_sample", D.q_sample(sched, x0, t, noise), real.q_sample(x0, t, noise), 0)
    xr, tr, lr, ldr = real.p_losses(x0, cond, t, noise=noise)
    xo, to, lo, ldo = D.p_losses(sd, D.UNET_TINY, sched, x0, cond, t, noise)
    ok &= _cmp("SDFusionText2ShapeModel.p_losses.x_noisy", xo, xr, 0) and _cmp("p_losses.loss", lo, lr)
    for k in ("loss_simple", "loss_vlb", "loss_total"):
        ok &= _cmp(f"p_losses.loss_dict[{k}]", ldo[k], ldr[k])
    # loss.backward() through the real class (what train_3dfront.py:390 does) vs autograd through the oracle: the gradients
    # the CUDA backward kernels are tested against (tests/test_unet_train_gpu.py) are the reference's own
    for p_ in real.df.parameters():
        p_.grad = None
    with torch.enable_grad():
        _, _, lr2, _ = real.p_losses(x0, cond, t, noise=noise)
        lr2.backward()
        sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        _, _, lo2, _ = D.p_losses(sdg, D.UNET_TINY, sched, x0, cond, t, noise)
        og = dict(zip(sdg.keys(), torch.autograd.grad(lo2, list(sdg.values()), allow_unused=True)))
    worst = 0.0
    for k, p_ in real.df.named_parameters():
        if p_.grad is None or og[k] is None:
            ok &= (p_.grad is None or float(p_.grad.abs().max()) == 0.0) and (og[k] is None or float(og[k].abs().max()) == 0.0)
            continue
        worst = max(worst, float((p_.grad - og[k]).abs().max()) / max(1e-12, float(p_.grad.abs().max())))
    print(f"{'ok ' if worst <= 1e-4 else 'BAD'} SDFusionText2ShapeModel loss.backward(): worst relative gradient difference oracle vs class {worst:.3e} over {len(og)} tensors")
    ok &= worst <= 1e-4
    # forward(): same RNG draws in the same order
    sdf = (torch.randn(2, 1, 16, 16, 16, generator=g) * 0.1).clamp(-0.2, 0.2)
    rel = torch.randn(2, 1, D.UNET_TINY["context_dim"], generator=g)
    torch.Tensor.cuda = lambda self, *a, **k: self          # BaseModel.tocuda (base_model.py:113-118); stay on the CPU
    real.set_input({"sdf": sdf, "rel": rel, "uc": rel})
    torch.manual_seed(24)
    real.forward()
    torch.manual_seed(24)
    z = V.encode_no_quant(vsd, V.VQ_TINY, sdf)
    tt = torch.randint(0, 1000, (2,)).long()
    _, _, lo, _ = D.p_losses(sd, D.UNET_TINY, sched, z, rel, tt, torch.randn_like(z))
    ok &= _cmp("SDFusionText2ShapeModel.forward().loss_df", lo, real.loss_df)
    # rel2shape (sdfusion_txt2shape_model.py:459-516): one shared x_T for all objects (seeded from time.time()), DDIM with
    # CFG in mini-batches of 7, decode_no_quant.  The class hard-codes device='cuda' (:485,490; ddim.py:22-26): patched to CPU.
    import time as _time
    import numpy as np
    from model.networks.diffusion_networks.samplers.ddim import DDIMSampler
    DDIMSampler.register_buffer = lambda self, name, attr: setattr(self, name, attr)
    _randn, _now = torch.randn, _time.time
    torch.randn = lambda *a, **k: _randn(*a, **{kk: vv for kk, vv in k.items() if kk != "device"})
    _time.time = lambda: 1234567.0
    try:
        nobj = 9                                           # 7 + 2: exercises the mini-batch loop
        data = {"sdf": torch.zeros(nobj, 1, 16, 16, 16), "rel": torch.randn(nobj, 1, D.UNET_TINY["context_dim"], generator=g),
                "uc": torch.randn(nobj, 1, D.UNET_TINY["context_dim"], generator=g)}
        with torch.no_grad():
            ref_sdf = real.rel2shape(data, ddim_steps=20, ddim_eta=0.0, uc_scale=3.0)
            torch.manual_seed(1234567)
            x_T = _randn((1, 3, 4, 4, 4)).repeat(nobj, 1, 1, 1, 1)
            z0, _ = D.ddim_sample(sd, D.UNET_TINY, sched, data["rel"], data["uc"], x_T, S=20, eta=0.0, scale=3.0)
            got = V.decode_no_quant(vsd, V.VQ_TINY, z0)
        ok &= _cmp("SDFusionText2ShapeModel.rel2shape (9 objects, 20 DDIM steps, CFG 3, decode)", got, ref_sdf, 1e-3)
    finally:
        torch.randn, _time.time = _randn, _now
    # concat-conditioning variant through the real wrapper: set_input views the 4096-d rel_mlp output as one 16^3 latent channel
    # (:246-248, hard-coded 16), apply_model routes it to c_concat (:281-283), p_losses as above
    realc = RD.build(D.UNET_CONCAT_TINY, V.VQ_TINY, seed_unet=31, seed_vq=32, conditioning_key="concat")
    csd = Wt.synth_state_dict(D.unet_param_shapes(D.UNET_CONCAT_TINY), seed=31)
    relc = torch.randn(2, 1, 4096, generator=g)
    x0c, noisec, tc = torch.randn(2, 3, 16, 16, 16, generator=g), torch.randn(2, 3, 16, 16, 16, generator=g), torch.tensor([12, 801])
    realc.set_input({"sdf": torch.zeros(2, 1, 16, 16, 16), "rel": relc, "uc": relc})
    ok &= tuple(realc.rel.shape) == (2, 1, 16, 16, 16)
    _, _, lr, ldr = realc.p_losses(x0c, realc.rel, tc, noise=noisec)
    _, _, lo, ldo = D.p_losses(csd, D.UNET_CONCAT_TINY, sched, x0c, relc.view(2, 1, 16, 16, 16), tc, noisec, concat=True)
    ok &= _cmp("SDFusionText2ShapeModel[concat].p_losses.loss", lo, lr) and _cmp("p_losses[concat].loss_vlb", ldo["loss_vlb"], ldr["loss_vlb"])
    # guided DDIM step of the concat variant through the reference's own sampler + wrapper (c_concat routing inside apply_model)
    samp = DDIMSampler(realc)
    samp.make_schedule(ddim_num_steps=100, ddim_eta=0.0, verbose=False)
    ddc = D.ddim_schedule(sched, 100)
    cvol, ucvol = relc.view(2, 1, 16, 16, 16), torch.randn(2, 1, 16, 16, 16, generator=g)
    with torch.no_grad():
        for index in (99, 37):
            step = int(ddc["timesteps"][index])
            xr, p0r = samp.p_sample_ddim(x0c, cvol, torch.full((2,), step, dtype=torch.long), index=index,
                                         unconditional_guidance_scale=3.0, unconditional_conditioning=ucvol)
            xo, p0o, _ = D.p_sample_ddim(csd, D.UNET_CONCAT_TINY, ddc, x0c, cvol, step, index, 3.0, ucvol, concat=True)
            ok &= _cmp(f"DDIMSampler.p_sample_ddim[concat, index {index}].x_prev", xo, xr) and _cmp("  pred_x0", p0o, p0r)
    # helpers/util.py:31-45 sample_points: the file imports pytorch3d (absent), so only that function's source is executed
    import ast
    src = open(os.path.join(REF, "helpers", "util.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "sample_points")
    ns = {"torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "helpers/util.py", "exec"), ns)
    from oracle import mesh as M
    for n_pts in (12000, 700):
        pts = torch.randn(n_pts, 3, generator=g)
        torch.manual_seed(5)
        (r,) = ns["sample_points"]([pts], 5000)
        torch.manual_seed(5)
        o = pts[M.sample_points_indices(n_pts, 5000)]
        ok &= _cmp(f"helpers.util.sample_points[{n_pts} -> 5000]", o, r, tol=0.0)
    print("sdf_to_mesh: PyMCubes / pytorch3d are absent here -- oracle/mesh.py is pinned by properties only (parity unpinned)")
    print("ORACLE PINNED" if ok else "ORACLE MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    raise SystemExit(main())
