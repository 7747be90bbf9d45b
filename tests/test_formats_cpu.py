"""On-disk formats of the data side (commonscenes_b200/dataset/formats.py) against the reference's own statements, executed
verbatim here on the same files (threedfront_dataset.py:387-391 and :399-409, :503-508)."""
import os
import pickle

import numpy as np
import pytest
import torch

from commonscenes_b200.dataset import formats as F


def test_sdf_grid_matches_the_reference_statements(tmp_path):
    rng = np.random.default_rng(0)
    raw = (rng.standard_normal(64 ** 3) * 0.3).astype(np.float64)            # stored wider than fp32, values beyond +-0.2
    np.save(tmp_path / "ori_sample_grid.npy", raw)
    got = F.load_sdf_grid(str(tmp_path / "ori_sample_grid.h5"))               # .h5 named, .npy twin read (no h5py in this image)
    # reference: obj_sdf = h5_f['pc_sdf_sample'][:].astype(np.float32); sdf = torch.Tensor(obj_sdf).view(1, 64, 64, 64); clamp
    ref = torch.clamp(torch.Tensor(raw.astype(np.float32)).view(1, 64, 64, 64), min=-0.2, max=0.2)
    assert got.dtype == torch.float32 and torch.equal(got, ref)
    assert torch.equal(F.load_sdf_grid(None), torch.zeros((1, 64, 64, 64)))   # floor / _scene_ node
    with pytest.raises(ValueError):
        np.save(tmp_path / "bad.npy", raw[:100])
        F.load_sdf_grid(str(tmp_path / "bad.npy"))
    with pytest.raises(FileNotFoundError):
        F.load_sdf_grid(str(tmp_path / "missing" / "ori_sample_grid.h5"))     # neither the .h5 nor an exported twin: loud
    assert F.sdf_path_for_model("/d/3D-FUTURE-model/abc/raw_model.obj") == "/d/3D-FUTURE-SDF/abc/ori_sample_grid.h5"


def test_prefetcher_delivers_scenes_in_order(tmp_path):
    paths = []
    for i in range(5):
        np.save(tmp_path / f"g{i}.npy", np.full(64 ** 3, 0.01 * (i + 1), np.float32))
        paths.append(str(tmp_path / f"g{i}.npy"))
    scenes = [[paths[0], None, paths[1]], [paths[2]], [paths[3], paths[4], None, None]]
    got = list(F.SdfPrefetcher(scenes, device="cpu", max_objects=4))
    assert [g.shape[0] for g in got] == [3, 1, 4]
    assert float(got[0][0].mean()) == pytest.approx(0.01) and float(got[0][1].abs().max()) == 0 and float(got[2][1].mean()) == pytest.approx(0.05)
    with pytest.raises(ValueError):
        list(F.SdfPrefetcher([[None] * 5], device="cpu", max_objects=4))


def test_clip_cache_round_trip_and_reordering(tmp_path):
    rng = np.random.default_rng(1)
    order = [7, 3, 12, 5]                                     # order in which the features were computed
    feats = rng.standard_normal((len(order) + 1, 512)).astype(np.float32)      # + the room's feature, last
    words = ["chair left table", "table in room", "chair left table"]
    rel = {w: rng.standard_normal(512).astype(np.float32) for w in set(words)}
    path = F.clip_feats_path(str(tmp_path), "scan0", large=False)
    os.makedirs(os.path.dirname(path))
    assert path.endswith("scan0/CLIP_small_scan0.pkl") and F.clip_feats_path("r", "s", True, True).endswith("s/CLIP_s.pkltmp")
    F.write_clip_feats(path, feats, order, rel)
    instances_order = [5, 12, 7, 3]
    text_feats, rel_feats = F.read_clip_feats(path, instances_order, words)
    # the reference's statements (threedfront_dataset.py:400-409, 503-508) on the same file
    dic = pickle.load(open(path, "rb"))
    ins, ordr = dic["instance_feats"], np.asarray(dic["instance_order"])
    ordered = [ins[:-1][inst == ordr] for inst in instances_order]
    ordered.append(ins[-1][np.newaxis, :])
    ref_text = list(np.concatenate(ordered, axis=0))
    ref_rel = [dic["rel_feats"][w] for w in words]
    assert len(text_feats) == 5 and all(np.array_equal(a, b) for a, b in zip(text_feats, ref_text))
    assert all(np.array_equal(a, b) for a, b in zip(rel_feats, ref_rel))
    assert np.array_equal(text_feats[0], feats[3]) and np.array_equal(text_feats[-1], feats[-1])
