"""Oracle (test infrastructure, BUILD CONTAINER ONLY): construct the reference's REAL `Sg2ScVAEModel`
(model/VAEGAN_V2FULL.py:17-760) so that the shape-branch wiring — encoder_2, select_sdfs, balance_objects, the checkpoint
dictionary — is pinned against the class itself and not only against the modules it calls.

The class imports packages that are absent here (omegaconf, fvcore, termcolor, mcubes, pytorch3d via its SDFusion wrapper).
None of them takes part in the arithmetic being pinned, so they are replaced by inert stubs in sys.modules: OmegaConf.load
returns the two fields the constructor reads, SDFusionText2ShapeModel becomes an empty holder (the denoiser itself is pinned
module by module in validate_against_reference.py), the Visualizer does nothing.  Everything else — the embeddings, every
GraphTripleConvNet, rel_mlp, the selection logic — is the reference's own code, unmodified, imported from /root/reference.
"""
from __future__ import annotations

import os
import sys
import types

import torch

REF = os.environ.get("CS_REFERENCE", "/root/reference")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference_scene_class(diffusion_bs_from_cfg=None, conditioning_key="crossattn"):
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} not found: this module only works where the reference is mounted")
    if REF not in sys.path:
        sys.path.insert(0, REF)

    class _OC:
        @staticmethod
        def load(path):
            return types.SimpleNamespace(hyper=types.SimpleNamespace(batch_size=diffusion_bs_from_cfg, distributed=0))

    class _Diff:      # SDFusionText2ShapeModel stand-in: the constructor only reads df.conditioning_key / trainable_params
        def __init__(self, opt):
            self.opt = opt
            self.df = types.SimpleNamespace(conditioning_key=conditioning_key)
            self.trainable_params = []

    class _Vis:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, n):
            return lambda *a, **k: None

    _stub("omegaconf", OmegaConf=_OC)
    _stub("model.sdfusion_txt2shape_model", SDFusionText2ShapeModel=_Diff)
    _stub("model.diff_utils.visualizer", Visualizer=_Vis)
    _stub("helpers.util", bool_flag=lambda s: bool(s), _CustomDataParallel=torch.nn.DataParallel)
    _stub("helpers.lr_scheduler")
    _stub("fvcore"); _stub("fvcore.common"); _stub("fvcore.common.param_scheduler", MultiStepParamScheduler=object)
    sys.modules.pop("model.VAEGAN_V2FULL", None)
    import model.VAEGAN_V2FULL as V
    return V.Sg2ScVAEModel


def vocab(num_objs: int, num_preds: int):
    return {"object_idx_to_name": [f"o{i}" for i in range(num_objs)], "object_idx_to_name_grained": [f"g{i}" for i in range(2 * num_objs)],
            "pred_idx_to_name": [f"p{i}" for i in range(num_preds)]}


def build(cfg: dict, diffusion_bs: int = 8, seed: int = 0, conditioning_key: str = "crossattn"):
    """The reference Sg2ScVAEModel in the v2_full wiring of model/VAE.py:57-63 for an oracle.graph config."""
    cls = import_reference_scene_class(conditioning_key=conditioning_key)
    torch.manual_seed(seed)
    return cls(vocab(cfg["num_objs"], cfg["num_preds"]), diff_opt="unused.yaml", diffusion_bs=diffusion_bs,
               embedding_dim=cfg["embedding_dim"], batch_size=4, decoder_cat=True, gconv_num_layers=cfg["num_layers"],
               mlp_normalization="batch", use_E2=True, residual=True, use_angles=True, clip=True)


def module_state_dict(m) -> dict:
    """nn.Module.state_dict of the reference class (its own state_dict(epoch, counter) override builds the checkpoint dict)."""
    return torch.nn.Module.state_dict(m)
