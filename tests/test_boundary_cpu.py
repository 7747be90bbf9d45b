"""CPU suite, part 2: the drop-in boundary — the C-ABI library loads and exports every symbol the header
declares, and the nn.Module mirrors expose exactly the reference's state-dict keys.
No compute call is made (there is no GPU here and no CPU fallback)."""
import os
import re

import pytest
import torch

from oracle import denoiser as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from commonscenes_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "cs_b200.h")).read()
    declared = set(re.findall(r"\b(cs_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/cs_b200.h but not exported"
    assert lib.cs_abi_version() == 1


def test_compute_fails_loudly_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from commonscenes_b200 import _lib
    with pytest.raises(_lib.CsError):
        _lib.require_device()
    from commonscenes_b200.model.networks.diffusion_networks.network import DiffusionUNet
    m = DiffusionUNet(dict(D.UNET_TINY, use_spatial_transformer=True, legacy=False), conditioning_key="crossattn")
    with pytest.raises(Exception):
        m(torch.zeros(1, 3, 8, 8, 8), torch.zeros(1, dtype=torch.long), c_crossattn=[torch.zeros(1, 1, 64)])


@pytest.mark.parametrize("cfg", [D.UNET_TINY, D.UNET_FULL], ids=["tiny", "full"])
def test_unet_state_dict_keys_equal_reference_inventory(cfg):
    from commonscenes_b200.model.networks.diffusion_networks.network import DiffusionUNet
    with torch.device("meta"):
        m = DiffusionUNet(dict(cfg, use_spatial_transformer=True, use_checkpoint=True, legacy=False), conditioning_key="crossattn")
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == D.unet_param_shapes(cfg)      # the inventory validate_against_reference.py pinned to the reference


@pytest.mark.parametrize("cfg", [D.UNET_CONCAT_TINY, D.UNET_CONCAT_FULL], ids=["tiny", "full"])
def test_concat_unet_state_dict_keys_equal_reference_inventory(cfg):
    """The AttentionBlock variant (config/sdfusion-txt2shape_concat.yaml): norm / qkv (Conv1d) / proj_out keys."""
    from commonscenes_b200.model.networks.diffusion_networks.network import DiffusionUNet
    with torch.device("meta"):
        m = DiffusionUNet(dict(cfg, use_checkpoint=True, legacy=False), conditioning_key="concat")
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == D.unet_param_shapes(cfg)
