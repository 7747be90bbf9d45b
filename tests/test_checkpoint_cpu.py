"""CPU suite: the reference's `model{epoch}.pth` checkpoint layout (VAEGAN_V2FULL.py:687-699; loaded by VAE.load_networks,
VAE.py:120-158) round-trips through the shape-branch Sg2ScVAEModel mirror — including a checkpoint that also carries the
layout branch's tensors and an optimizer state built over `params + df_params` (VAEGAN_V2FULL.py:636-645)."""
import os

import pytest
import torch
import yaml

TINY_DF = dict(model=dict(params=dict(linear_start=0.00085, linear_end=0.012, conditioning_key="crossattn", timesteps=1000)),
               unet=dict(params=dict(image_size=8, in_channels=3, out_channels=3, model_channels=32, num_res_blocks=1,
                                     attention_resolutions=[4, 2], channel_mult=[1, 2, 3], num_heads=4, dims=3,
                                     use_spatial_transformer=True, transformer_depth=1, context_dim=1280, use_checkpoint=True,
                                     legacy=False)))
TINY_VQ = dict(model=dict(params=dict(embed_dim=3, n_embed=64, ddconfig=dict(
    double_z=False, z_channels=3, resolution=16, in_channels=1, out_ch=1, ch=16, ch_mult=[1, 2], num_res_blocks=1,
    attn_resolutions=[], dropout=0.0))))


def _model(tmp_path, seed):
    from commonscenes_b200.model.VAEGAN_V2FULL import Sg2ScVAEModel
    from commonscenes_b200.model.sdfusion_txt2shape_model import default_opt
    df, vq = tmp_path / "df.yaml", tmp_path / "vq.yaml"
    df.write_text(yaml.safe_dump(TINY_DF)); vq.write_text(yaml.safe_dump(TINY_VQ))
    torch.manual_seed(seed)
    vocab = {"object_idx_to_name": [f"o{i}" for i in range(12)], "pred_idx_to_name": [f"p{i}" for i in range(6)]}
    return Sg2ScVAEModel(vocab, diff_opt=default_opt(device="cpu", df_cfg=str(df), vq_cfg=str(vq)), embedding_dim=64,
                         mlp_normalization="batch", residual=True, gconv_num_layers=2)


def test_reference_checkpoint_layout_round_trip(tmp_path):
    a, b = _model(tmp_path, 1), _model(tmp_path, 2)
    with torch.no_grad():                                   # make every tensor of `a` distinctive (incl. zero-init convs)
        for p in list(a.parameters()) + list(a.Diff.df.parameters()) + list(a.Diff.vqvae.parameters()):
            p.add_(torch.randn_like(p) * 0.01)
    assert not torch.equal(a.rel_mlp[0].weight, b.rel_mlp[0].weight)
    # a reference optimizerFULL state: AdamW over [graph-side params] + df params, after one step
    params = [p for p in a.parameters() if p.requires_grad] + a.Diff.trainable_params
    opt = torch.optim.AdamW(params, lr=1e-4)
    for p in params:
        p.grad = torch.randn_like(p) * 1e-3
    opt.step()
    path = a.save_checkpoint(os.path.join(tmp_path, "model100.pth"), epoch=100, counter=4321, optimizer_state=opt.state_dict())
    ck = torch.load(path)
    assert {"epoch", "counter", "vqvae", "df", "opt"} <= set(ck) and ck["epoch"] == 100
    assert any(k.startswith("gconv_net_ec_rel.gconvs.0.net1.0") for k in ck) and "rel_mlp.0.weight" in ck
    # a full v2_full checkpoint also carries the layout branch: those keys are reported, not an error
    ck["gconv_net_ec_box.gconvs.0.net1.0.weight"] = torch.zeros(3, 3)
    ck["d3_net.0.weight"] = torch.zeros(2, 2)
    info = b.load_checkpoint(ck)
    assert info["epoch"] == 100 and info["counter"] == 4321
    assert info["ignored_keys"] == ["d3_net.0.weight", "gconv_net_ec_box.gconvs.0.net1.0.weight"]
    for (k, x), (_, y) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(x, y), k
    for mod in ("df", "vqvae"):
        for (k, x), (_, y) in zip(getattr(a.Diff, mod).state_dict().items(), getattr(b.Diff, mod).state_dict().items()):
            assert torch.equal(x, y), (mod, k)
    with pytest.raises(KeyError):
        b.load_checkpoint(ck, strict=True)
    ck2 = dict(ck); ck2.pop("rel_mlp.0.weight")
    with pytest.raises(KeyError):
        b.load_checkpoint(ck2)
    # the denoiser's slice of the optimizer state resumes the native step
    from commonscenes_b200.train import DenoiserTrainStep
    den = b.denoiser_optimizer_state(info["opt"])
    n_graph = len([p for p in a.parameters() if p.requires_grad])
    assert len(den["state"]) == len(a.Diff.trainable_params)
    step = DenoiserTrainStep(b.Diff)
    step.load_optimizer_state_dict(den)
    assert step.step_count == 1
    for j, p in enumerate(step.params):
        off, n = step.offsets[p], p.numel()
        assert torch.equal(step.flat_m[off:off + n].view(p.shape), info["opt"]["state"][n_graph + j]["exp_avg"])


def test_full_reference_checkpoint_inventory_loads(tmp_path):
    """A checkpoint holding EVERY tensor of the reference's real Sg2ScVAEModel in the v2_full wiring (inventory recorded
    from the class itself: tests/golden/sg2sc_v2full_state_dict_keys.json) loads into the shape-branch mirror: all of the
    mirror's keys are found with the right shapes, the layout-branch tensors are reported as ignored."""
    import json
    from commonscenes_b200.model.VAEGAN_V2FULL import Sg2ScVAEModel
    from commonscenes_b200.model.sdfusion_txt2shape_model import default_opt
    inv = dict(json.load(open(os.path.join(os.path.dirname(__file__), "golden", "sg2sc_v2full_state_dict_keys.json"))))
    df, vq = tmp_path / "df.yaml", tmp_path / "vq.yaml"
    df.write_text(yaml.safe_dump(TINY_DF)); vq.write_text(yaml.safe_dump(TINY_VQ))
    vocab = {"object_idx_to_name": [f"o{i}" for i in range(36)], "pred_idx_to_name": [f"p{i}" for i in range(16)]}
    m = Sg2ScVAEModel(vocab, diff_opt=default_opt(device="cpu", df_cfg=str(df), vq_cfg=str(vq)), embedding_dim=64,
                      mlp_normalization="batch", residual=True, gconv_num_layers=5)
    own = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert len(own) == 171 and all(inv.get(k) == s for k, s in own.items())      # drop-in contract: same names, same shapes
    g = torch.Generator().manual_seed(0)
    ck = {k: (torch.randn(s, generator=g) if "num_batches_tracked" not in k else torch.tensor(3)) for k, s in inv.items()}
    ck.update(epoch=5, counter=99, opt={}, vqvae=m.Diff.vqvae.state_dict(), df=m.Diff.df.state_dict())
    info = m.load_checkpoint(ck)
    assert info["epoch"] == 5 and len(info["ignored_keys"]) == len(inv) - 171
    assert all(not k.startswith(("gconv_net_ec_rel", "rel_mlp", "obj_embeddings_dc", "pred_embeddings_dc")) for k in info["ignored_keys"])
    assert torch.equal(m.rel_mlp[0].weight, ck["rel_mlp.0.weight"])


def test_layout_branch_mirror_has_the_real_class_inventory_in_order(tmp_path):
    """With layout_branch=True the mirror registers exactly the modules of the reference's real Sg2ScVAEModel, in the same
    order: 711 state-dict keys / shapes, so `parameters()` — and therefore the indices of an optimizerFULL state dict —
    line up with reference checkpoints."""
    import json
    from commonscenes_b200.model.VAEGAN_V2FULL import Sg2ScVAEModel
    from commonscenes_b200.model.sdfusion_txt2shape_model import default_opt
    inv = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "sg2sc_v2full_state_dict_keys.json")))
    df, vq = tmp_path / "df.yaml", tmp_path / "vq.yaml"
    df.write_text(yaml.safe_dump(TINY_DF)); vq.write_text(yaml.safe_dump(TINY_VQ))
    vocab = {"object_idx_to_name": [f"o{i}" for i in range(36)], "pred_idx_to_name": [f"p{i}" for i in range(16)]}
    m = Sg2ScVAEModel(vocab, diff_opt=default_opt(device="cpu", df_cfg=str(df), vq_cfg=str(vq)), embedding_dim=64,
                      mlp_normalization="batch", residual=True, gconv_num_layers=5, layout_branch=True)
    own = [[k, list(v.shape)] for k, v in m.state_dict().items()]
    assert len(own) == 711 and own == inv
    ck = {k: torch.zeros(s) if "num_batches_tracked" not in k else torch.tensor(0) for k, s in inv}
    ck.update(epoch=1, counter=2, opt={}, vqvae=m.Diff.vqvae.state_dict(), df=m.Diff.df.state_dict())
    assert m.load_checkpoint(ck, strict=True)["ignored_keys"] == []
    with pytest.raises(RuntimeError):
        Sg2ScVAEModel(vocab, diff_opt=default_opt(device="cpu", df_cfg=str(df), vq_cfg=str(vq)), embedding_dim=64,
                      mlp_normalization="batch", residual=True).encoder(None, None, None, None, None, None)
