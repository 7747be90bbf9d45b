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


def test_select_sdfs_and_balance_objects_match_the_reference_class(monkeypatch):
    """Host-side object selection: same picks, in the same order, as the reference's REAL
    Sg2ScVAEModel.select_sdfs / balance_objects under the same `random` / torch seeds (index work: bit-exact).
    Golden: tests/golden/select_sdfs.npz (tests/golden/make_golden_scene.py)."""
    import random
    import numpy as np
    from commonscenes_b200.model.VAEGAN_V2FULL import Sg2ScVAEModel
    g = np.load(os.path.join(ROOT, "tests", "golden", "select_sdfs.npz"))
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)        # select_sdfs ends with .cuda() like the reference
    m = object.__new__(Sg2ScVAEModel)                                            # the methods only read diffusion_bs
    scene, objs, grained, sdfs, uc, c = (torch.tensor(g[k]) for k in ("scene", "objs", "grained", "sdfs", "uc", "c"))
    for bs in (8, 4, 12):
        m.diffusion_bs = bs
        random.seed(1000 + bs)
        cats, d = m.select_sdfs(scene, objs, grained, sdfs, uc, c, random=False)
        assert np.array_equal(cats.numpy(), g[f"balanced_bs{bs}_cats"])
        for k in ("sdf", "uc", "rel"):
            assert np.array_equal(d[k].numpy(), g[f"balanced_bs{bs}_{k}"]), (bs, k)
        torch.manual_seed(2000 + bs)
        cats, d = m.select_sdfs(scene, objs, grained, sdfs, uc, c, random=True)
        assert np.array_equal(cats.numpy(), g[f"random_bs{bs}_cats"]) and np.array_equal(d["sdf"].numpy(), g[f"random_bs{bs}_sdf"])
    random.seed(77)
    ids = torch.tensor(g["balance_ids"])
    assert np.array_equal(m.balance_objects(ids, ids, 3).numpy(), g["balance_n3"])
    assert np.array_equal(m.balance_objects(ids, ids, 6).numpy(), g["balance_n6"])
    with pytest.raises(AssertionError):
        m.balance_objects(ids, ids[:3], 2)
