"""GPU parity of the SDF -> surface-points step, through the C ABI (cs_surface_count / cs_surface_emit):

* against the numpy oracle (oracle/mesh.py) on seeded grids: BIT-EXACT vertices (fp32 after the reference's
  `verts / n_cell - .5` in float64) and identical triangle index lists -- ragged batches, empty surfaces, values equal to the
  level, non-cubic grids;
* at the evaluation size (32 objects x 64^3) through size-independent properties: vertex count = number of crossing grid
  edges (counted with torch), every vertex on one grid edge, closed surface for a shape away from the border;
* the chain the evaluation runs: sdf_to_mesh(...).verts_list() -> sample_points -> Chamfer distance.
"""
from collections import Counter

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _u3d():
    from commonscenes_b200 import _lib
    from commonscenes_b200.model.diff_utils import util_3d
    _lib.require_device()
    return util_3d


def _shapes(n, seed, count):
    rng = np.random.default_rng(seed)
    g = np.stack(np.meshgrid(*[np.arange(n)] * 3, indexing="ij"), -1).astype(np.float32)
    out = []
    for i in range(count):
        c = rng.uniform(0.35, 0.65, 3) * n
        r = rng.uniform(0.15, 0.3) * n
        a = rng.uniform(0.6, 1.4, 3)
        out.append((np.linalg.norm((g - c) / a, axis=-1) - r).astype(np.float32) / n)      # an ellipsoid-like field, O(0.2) values
    return np.stack(out)


@pytest.mark.parametrize("n,count,level", [(16, 5, 0.02), (24, 3, 0.0), (33, 2, -0.01)])
def test_vertices_and_triangles_equal_the_oracle(n, count, level):
    from oracle import mesh
    u = _u3d()
    grids = _shapes(n, 3 + n, count)
    grids[1] = 1.0                                            # an object without a surface in the middle of the batch
    verts, faces, tot = u.surface_extract(torch.from_numpy(grids).cuda(), level)
    vb = tb = 0
    for i in range(count):
        v_o, f_o = mesh.marching_cubes(grids[i], level)
        nv, nt = int(tot[i, 0]), int(tot[i, 1])
        assert (nv, nt) == (len(v_o), len(f_o))
        want = (v_o / n - .5).astype(np.float32)
        assert np.array_equal(verts[vb:vb + nv].cpu().numpy(), want)        # bit-exact
        assert np.array_equal(faces[tb:tb + nt].cpu().numpy(), f_o)
        vb, tb = vb + nv, tb + nt
    assert vb == verts.shape[0] and tb == faces.shape[0] and int(tot[1, 0]) == 0


def test_noise_grid_non_cubic_and_level_ties():
    from oracle import mesh
    u = _u3d()
    rng = np.random.default_rng(1)
    vol = rng.standard_normal((2, 9, 20, 13)).astype(np.float32)
    vol[0, 3:5, 4:9, 2:6] = 0.02                              # exact ties: value == level counts as inside
    verts, faces, tot = u.surface_extract(torch.from_numpy(vol).cuda(), float(np.float32(0.02)), n_cell=13)
    vb = tb = 0
    for i in range(2):
        v_o, f_o = mesh.marching_cubes(vol[i], float(np.float32(0.02)))
        nv, nt = int(tot[i, 0]), int(tot[i, 1])
        assert np.array_equal(verts[vb:vb + nv].cpu().numpy(), (v_o / 13 - .5).astype(np.float32))
        assert np.array_equal(faces[tb:tb + nt].cpu().numpy(), f_o)
        vb, tb = vb + nv, tb + nt


def test_evaluation_size_properties():
    u = _u3d()
    grids = torch.from_numpy(_shapes(64, 11, 32)).cuda()
    level = 0.02
    verts, faces, tot = u.surface_extract(grids, level)
    inside = grids.double() <= level
    cross = ((inside[:, :-1] != inside[:, 1:]).flatten(1).sum(1) + (inside[:, :, :-1] != inside[:, :, 1:]).flatten(1).sum(1)
             + (inside[:, :, :, :-1] != inside[:, :, :, 1:]).flatten(1).sum(1)).cpu()
    assert torch.equal(tot[:, 0], cross)
    idx = (verts.double() + .5) * 64
    frac = idx - idx.floor()
    assert bool(((frac > 1e-6).sum(1) <= 1).all()) and float(idx.min()) >= 0 and float(idx.max()) <= 63
    # object 0 does not touch the border: closed, oriented, genus 0
    f = faces[:int(tot[0, 1])].cpu().numpy()
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    und = Counter(map(tuple, np.sort(e, 1)))
    assert set(und.values()) == {2} and int(tot[0, 0]) - len(und) + len(f) == 2
    # repeat launch: identical output (integer scans, no atomics)
    v2, f2, _ = u.surface_extract(grids, level)
    assert torch.equal(verts, v2) and torch.equal(faces, f2)


def test_sdf_to_mesh_sample_points_chamfer_chain():
    u = _u3d()
    from commonscenes_b200.helpers.util import sample_points
    import commonscenes_b200.extension.dist_chamfer as ext
    sdf = torch.from_numpy(_shapes(64, 5, 18)).cuda()[:, None]
    m = u.sdf_to_mesh(sdf)                       # util_3d.py:205-209: 16 meshes unless render_all
    assert len(m) == 16 and len(u.sdf_to_mesh(sdf, render_all=True)) == 18
    vl = m.verts_list()
    assert all(v.is_cuda and v.dtype == torch.float32 and v.shape[1] == 3 and float(v.abs().max()) <= 0.5 for v in vl)
    assert all(int(f.max()) < v.shape[0] for v, f in zip(vl, m.faces_list()))
    torch.manual_seed(0)
    pts = torch.stack(sample_points(vl, 5000))   # eval_3dfront.py:589-592
    assert pts.shape == (16, 5000, 3)
    d1, d2 = ext.chamferDist()(pts, pts.roll(1, 0))
    assert bool(torch.isfinite(d1).all()) and float(ext.chamferDist()(pts, pts)[0].max()) == 0.0


def test_empty_batch_and_all_inside_grids():
    u = _u3d()
    verts, faces, tot = u.surface_extract(torch.zeros((0, 8, 8, 8), device="cuda"), 0.02)
    assert verts.shape == (0, 3) and faces.shape == (0, 3) and tot.shape == (0, 2)
    m = u.sdf_to_mesh(torch.zeros((0, 1, 8, 8, 8), device="cuda"))
    assert len(m) == 0 and m.verts_list() == []
    # a grid entirely below the level has no crossing edge either
    verts, faces, tot = u.surface_extract(torch.full((2, 8, 8, 8), -1.0, device="cuda"), 0.02)
    assert verts.shape == (0, 3) and int(tot.sum()) == 0
    with pytest.raises(Exception):
        u.surface_extract(torch.zeros((1, 8, 8, 8)), 0.02)        # CPU tensor: no CPU path
