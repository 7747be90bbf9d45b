"""CPU suite, part 1: the oracle against the golden fixtures that the REFERENCE's own modules produced
(tests/golden/make_golden.py).  fp32 CPU torch on both sides -> tolerance 2e-5 relative to the output range
(differences come only from thread-count dependent reduction order)."""
import os

import numpy as np
import pytest
import torch

from oracle import denoiser as D, graph as G, vqvae as V, weights as Wt

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name))


def _close(got, ref, tol=2e-5):
    got, ref = torch.as_tensor(got), torch.as_tensor(ref)
    err = (got - ref).abs().max().item()
    assert err <= tol * max(1.0, ref.abs().max().item()), f"max err {err}"


@pytest.mark.parametrize("tag,cfg", [("tiny", D.UNET_TINY), ("full", D.UNET_FULL)])
def test_unet_forward_matches_reference_golden(tag, cfg):
    g = _load(f"unet_{tag}.npz")
    sd = Wt.synth_state_dict(D.unet_param_shapes(cfg), int(g["weight_seed"]))
    with torch.no_grad():
        eps = D.unet_forward(sd, cfg, torch.tensor(g["x"]), torch.tensor(g["t"]), torch.tensor(g["ctx"]))
    _close(eps, g["eps"])


@pytest.mark.parametrize("tag,cfg", [("tiny", D.UNET_CONCAT_TINY), ("full", D.UNET_CONCAT_FULL)])
def test_concat_unet_forward_matches_reference_golden(tag, cfg):
    """AttentionBlock / concat-conditioning denoiser vs the reference module's own output."""
    g = _load(f"unet_concat_{tag}.npz")
    sd = Wt.synth_state_dict(D.unet_param_shapes(cfg), int(g["weight_seed"]))
    with torch.no_grad():
        eps = D.unet_forward(sd, cfg, torch.tensor(g["x"]), torch.tensor(g["t"]), c_concat=torch.tensor(g["c_concat"]))
    _close(eps, g["eps"])


def test_unet_inventory_counts():
    shapes = D.unet_param_shapes(D.UNET_FULL)
    assert len(shapes) == 496                                            # SURVEY.md §8b [probe]
    assert sum(int(np.prod(s)) for s in shapes.values()) == 413_540_739  # SURVEY.md §3a [probe]
    assert sum(int(np.prod(s)) for s in V.vq_param_shapes(V.VQ_FULL).values()) == 26_349_788


def test_schedule_known_values():
    s = D.register_schedule(**D.DIFFUSION)
    assert abs(float(s["betas"][0]) - 8.5e-4) < 1e-9 and abs(float(s["betas"][-1]) - 1.2e-2) < 1e-8
    assert abs(float(s["alphas_cumprod"][-1]) - 0.0046601) < 1e-6       # SURVEY.md §8 a12 [probe]
    dd = D.ddim_schedule(s, 100)
    assert dd["timesteps"][0] == 1 and dd["timesteps"][-1] == 991 and len(dd["timesteps"]) == 100
    assert float(dd["alphas_prev"][0]) == float(s["alphas_cumprod"][0])  # ldm_diffusion_util.py:88
    with pytest.raises(IndexError):                                      # S=1000 overruns the table
        D.ddim_schedule(s, 1000)


def test_ddim_guided_steps_match_reference_sampler():
    g = _load("ddim_tiny.npz")
    cfg = D.UNET_TINY
    sd = Wt.synth_state_dict(D.unet_param_shapes(cfg), int(g["weight_seed"]))
    sched = D.register_schedule(**D.DIFFUSION)
    dd = D.ddim_schedule(sched, 100)
    assert (dd["timesteps"] == g["ddim_timesteps"]).all()
    np.testing.assert_array_equal(dd["alphas"], g["ddim_alphas"])
    with torch.no_grad():
        x, trace = D.ddim_sample(sd, cfg, sched, torch.tensor(g["c"]), torch.tensor(g["uc"]), torch.tensor(g["x_T"]),
                                 S=100, scale=3.0, max_steps=4)
    for i, (xp, p0, _) in enumerate(trace):
        _close(xp, g["x_steps"][i], 5e-5)
        _close(p0, g["pred_x0_steps"][i], 5e-5)


def test_p_losses_properties():
    cfg = D.UNET_TINY
    sd = Wt.synth_state_dict(D.unet_param_shapes(cfg), 11)
    sched = D.register_schedule(**D.DIFFUSION)
    g = torch.Generator().manual_seed(0)
    x0 = torch.randn(2, 3, 8, 8, 8, generator=g)
    noise = torch.randn(2, 3, 8, 8, 8, generator=g)
    ctx = torch.randn(2, 1, 64, generator=g)
    t = torch.tensor([10, 900])
    with torch.no_grad():
        x_noisy, target, loss, ld = D.p_losses(sd, cfg, sched, x0, ctx, t, noise)
        eps = D.unet_forward(sd, cfg, x_noisy, t, ctx)
    assert torch.equal(target, noise)
    assert abs(float(loss) - float(((eps - noise) ** 2).mean())) < 1e-6
    assert set(ld) == {"loss_simple", "loss_vlb", "loss_total"}          # sdfusion_txt2shape_model.py:329-343


@pytest.mark.parametrize("tag,cfg", [("tiny", V.VQ_TINY), ("full", V.VQ_FULL)])
def test_vqvae_matches_reference_golden(tag, cfg):
    g = _load(f"vqvae_{tag}.npz")
    sd = Wt.synth_state_dict(V.vq_param_shapes(cfg), int(g["weight_seed"]))
    gen = torch.Generator().manual_seed(int(g["input_seed"]))
    r = cfg["resolution"]
    x = (torch.randn(1, 1, r, r, r, generator=gen) * 0.1).clamp(-0.2, 0.2)
    with torch.no_grad():
        z = V.encode_no_quant(sd, cfg, x)
        _close(z, g["z"], 5e-5)
        zq, idx = V.quantize(sd, torch.tensor(g["z"]))
        assert (idx.numpy() == g["idx"]).all()
        dec = V.decode_no_quant(sd, cfg, torch.tensor(g["z"]))
    sub = dec[:, :, ::4, ::4, ::4] if tag == "full" else dec
    _close(sub, g["dec_sub"], 5e-5)
    assert abs(float(dec.double().abs().sum()) - float(g["dec_abs_sum"])) <= 1e-4 * float(g["dec_abs_sum"])


@pytest.mark.parametrize("tag,cfg", [("tiny", G.GCN_TINY), ("full", G.GCN_FULL)])
def test_encoder2_matches_reference_golden(tag, cfg):
    g = _load(f"gcn_{tag}.npz")
    sd = Wt.synth_state_dict(G.gcn_param_shapes(cfg), int(g["weight_seed"]))
    args = [torch.tensor(g[k]) for k in ("z", "objs", "triples", "text", "rel")]
    for mode in ("eval", "train"):
        with torch.no_grad():
            uc, c = G.encoder_2(sd, cfg, *args, training=(mode == "train"))
        _close(c, g[f"c_{mode}"], 5e-5)
        _close(uc, g[f"uc_{mode}"], 5e-5)


def test_ddpm_posterior_equals_eta1_ddim_form():
    """The ancestral sampler feeds cs_ddim_step with (abar_t, abar_{t-1}, sigma_t = sqrt(beta~_t)); that DDIM-form update must
    equal the posterior-mean form the oracle restates from the reference's registered buffers, for every t."""
    from oracle import denoiser as D
    sched = D.register_schedule(**D.DIFFUSION)
    ac = sched["alphas_cumprod"].double().numpy()
    ac_prev = np.append(1.0, ac[:-1])
    betas = 1.0 - ac / ac_prev
    sig = np.sqrt(betas * (1.0 - ac_prev) / (1.0 - ac))
    assert sig[0] == 0.0
    g = torch.Generator().manual_seed(0)
    x, e = torch.randn(64, generator=g).double(), torch.randn(64, generator=g).double()
    for t in (0, 1, 2, 10, 500, 998, 999):
        x0 = (x - np.sqrt(1.0 - ac[t]) * e) / np.sqrt(ac[t])
        ddim_form = np.sqrt(ac_prev[t]) * x0 + np.sqrt(max(1.0 - ac_prev[t] - sig[t] ** 2, 0.0)) * e
        x0_ref = sched["sqrt_recip_alphas_cumprod"][t].double() * x - sched["sqrt_recipm1_alphas_cumprod"][t].double() * e
        post = sched["posterior_mean_coef1"][t].double() * x0_ref + sched["posterior_mean_coef2"][t].double() * x
        assert float((ddim_form - post).abs().max()) < 2e-5, t           # fp32 tables vs float64 recomputation
        assert abs(sig[t] - float(torch.exp(0.5 * sched["posterior_log_variance_clipped"][t]))) < 1e-6 or t == 0


def test_rel2shape_chain_matches_the_reference_class_golden():
    """The oracle's free-running chain (shared x_T -> 20 guided DDIM steps on all 9 objects at once -> decode) reproduces what
    the reference's REAL SDFusionText2ShapeModel.rel2shape computed in mini-batches of 7 (tests/golden/rel2shape_tiny.npz)."""
    g = _load("rel2shape_tiny.npz")
    sd = Wt.synth_state_dict(D.unet_param_shapes(D.UNET_TINY), int(g["weight_seed_unet"]))
    vcfg = dict(V.VQ_TINY, resolution=int(g["resolution"]))
    vsd = Wt.synth_state_dict(V.vq_param_shapes(vcfg), int(g["weight_seed_vq"]))
    sched = D.register_schedule(**D.DIFFUSION)
    rel, uc = torch.tensor(g["rel"]), torch.tensor(g["uc"])
    x_T = torch.tensor(g["x_T"]).repeat(rel.shape[0], 1, 1, 1, 1)
    with torch.no_grad():
        z0, _ = D.ddim_sample(sd, D.UNET_TINY, sched, rel, uc, x_T, S=int(g["steps"]), eta=0.0, scale=3.0)
        sdf = V.decode_no_quant(vsd, vcfg, z0[torch.tensor(g["rows"])])
    _close(sdf, g["sdf"], tol=1e-3)


@pytest.mark.parametrize("tag", ["tiny", "full"])
def test_layout_branch_oracle_matches_reference_class_golden(tag):
    """SURVEY.md §8f rank 2, first gate: the layout-branch oracle (encoder / manipulate / decoder / losses) vs what the
    reference's REAL Sg2ScVAEModel computed (tests/golden/layout_*.npz), eval- and train-mode BatchNorm."""
    from oracle import layout as Lo
    cfg = Lo.LAYOUT_TINY if tag == "tiny" else Lo.LAYOUT_FULL
    g = _load(f"layout_{tag}.npz")
    sd = Wt.synth_state_dict(Lo.layout_param_shapes(cfg), int(g["weight_seed"]))
    z, objs, triples, text, rel, boxes, angles, zz = (torch.tensor(g[k]) for k in ("z", "objs", "triples", "text", "rel", "boxes", "angles", "zz"))
    with torch.no_grad():
        for mode in ("eval", "train"):
            tr = mode == "train"
            mu, logvar = Lo.encoder(sd, cfg, objs, triples, boxes, text, rel, angles, tr)
            _close(mu, g[f"mu_{mode}"]); _close(logvar, g[f"logvar_{mode}"])
            _close(Lo.manipulate(sd, cfg, zz, objs, triples, text, rel, tr), g[f"man_{mode}"])
            b, a = Lo.decoder(sd, cfg, z, objs, triples, text, rel, tr)
            _close(b, g[f"boxes_{mode}"]); _close(a, g[f"angle_logp_{mode}"])
            tot, terms = Lo.layout_losses(b, boxes, a, angles, mu, logvar, 0.1)
            assert abs(float(tot) - float(g[f"loss_{mode}"])) <= 2e-5 * max(1.0, abs(float(tot)))
            assert set(terms) == {"box", "angle_pred", "KLD_Gauss"}
