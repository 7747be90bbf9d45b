"""Oracle (test infrastructure, BUILD CONTAINER ONLY): construct the reference's REAL `SDFusionText2ShapeModel`
(model/sdfusion_txt2shape_model.py:51-703) on the CPU, so that its schedule buffers, q_sample, p_losses / forward (the
training loss wiring) are pinned against the class itself.

Absent third-party imports are replaced by inert stubs in sys.modules (mcubes, termcolor, fvcore, pytorch3d via
model.diff_utils.util_3d, helpers.util); OmegaConf.load becomes a yaml loader with attribute access.  The networks
(DiffusionUNet, VQVAE), the schedule, the loss and the sampler are the reference's own code, imported from /root/reference.
The VQ-VAE checkpoint the constructor insists on (model_utils.py:8) is written here from the seeded synthetic weights.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

import torch
import yaml

REF = os.environ.get("CS_REFERENCE", "/root/reference")


class Cfg(dict):
    __getattr__ = dict.get

    @staticmethod
    def wrap(x):
        return Cfg({k: Cfg.wrap(v) for k, v in x.items()}) if isinstance(x, dict) else x


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference_diffusion_class():
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} not found: this module only works where the reference is mounted")
    if REF not in sys.path:
        sys.path.insert(0, REF)

    class _OC:
        @staticmethod
        def load(path):
            with open(path) as f:
                return Cfg.wrap(yaml.safe_load(f))

    lc = _stub("omegaconf.listconfig", ListConfig=type("ListConfig", (list,), {}))
    _stub("omegaconf", OmegaConf=_OC, listconfig=lc)
    _stub("mcubes")
    _stub("termcolor", colored=lambda s, *a, **k: s, cprint=lambda *a, **k: None)
    _stub("helpers.util", bool_flag=lambda s: bool(s), _CustomDataParallel=torch.nn.DataParallel)
    _stub("helpers.lr_scheduler")
    _stub("fvcore"); _stub("fvcore.common"); _stub("fvcore.common.param_scheduler", MultiStepParamScheduler=object)
    _stub("model.diff_utils.util_3d", init_mesh_renderer=lambda **k: None, render_sdf=lambda *a, **k: None)
    _stub("model.diff_utils.util")
    sys.modules.pop("model.sdfusion_txt2shape_model", None)
    import model.sdfusion_txt2shape_model as S
    return S.SDFusionText2ShapeModel


def build(unet_cfg: dict, vq_cfg: dict, seed_unet: int, seed_vq: int, workdir: str | None = None,
          conditioning_key: str = "crossattn"):
    """The real class on CPU with the oracle's seeded synthetic weights (oracle.weights) in both networks.
    conditioning_key='concat' builds the AttentionBlock variant of config/sdfusion-txt2shape_concat.yaml."""
    from oracle import weights as Wt
    cls = import_reference_diffusion_class()
    workdir = workdir or tempfile.mkdtemp(prefix="cs_ref_diff_")
    unet = dict(unet_cfg)
    unet["attention_resolutions"] = list(unet["attention_resolutions"]); unet["channel_mult"] = list(unet["channel_mult"])
    if conditioning_key == "concat":
        unet.update(use_spatial_transformer=False, context_dim=None, use_checkpoint=False, legacy=False)
    else:
        unet.update(use_spatial_transformer=True, use_checkpoint=False, legacy=False)
    df = dict(model=dict(params=dict(linear_start=0.00085, linear_end=0.012, conditioning_key=conditioning_key, timesteps=1000,
                                     scale_factor=0.18215)), unet=dict(params=unet))
    dd = dict(double_z=False, z_channels=vq_cfg["z_channels"], resolution=vq_cfg["resolution"], in_channels=vq_cfg["in_channels"],
              out_ch=vq_cfg["out_ch"], ch=vq_cfg["ch"], ch_mult=list(vq_cfg["ch_mult"]), num_res_blocks=vq_cfg["num_res_blocks"],
              attn_resolutions=[], dropout=0.0)
    vq = dict(model=dict(params=dict(embed_dim=vq_cfg["embed_dim"], n_embed=vq_cfg["n_embed"], ddconfig=dd)))
    df_path, vq_path, ck_path = (os.path.join(workdir, n) for n in ("df.yaml", "vq.yaml", "vqvae.pth"))
    with open(df_path, "w") as f:
        yaml.safe_dump(df, f)
    with open(vq_path, "w") as f:
        yaml.safe_dump(vq, f)
    from model.networks.vqvae_networks.network import VQVAE
    v = VQVAE(Cfg.wrap(dd), vq_cfg["n_embed"], vq_cfg["embed_dim"])
    Wt.fill_module_(v, seed_vq)
    torch.save(v.state_dict(), ck_path)
    opt = Cfg.wrap(dict(hyper=dict(isTrain=True, device="cpu", batch_size=4, gpu_ids=[0]),
                        network=dict(df_cfg=df_path, vq_cfg=vq_path, vq_ckpt=ck_path), misc=dict(debug=0, local_rank=0)))
    m = cls(opt)
    Wt.fill_module_(m.df, seed_unet)
    return m
