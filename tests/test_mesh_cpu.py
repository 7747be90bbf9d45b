"""CPU tests of the SDF -> surface-points step: the numpy oracle (oracle/mesh.py) against size-independent
properties of marching cubes -- PyMCubes itself is absent, so these are what pins it -- the baked triangle table of the product
against the oracle's construction, and the sample_points mirror against the reference function's own arithmetic."""
import importlib.util
import os
from collections import Counter

import numpy as np
import pytest
import torch

from oracle import mesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _grid(n):
    g = np.stack(np.meshgrid(*[np.arange(n)] * 3, indexing="ij"), -1).astype(np.float64)
    return g, (n - 1) / 2


def _edges(f):
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    return Counter(map(tuple, np.sort(e, 1))), Counter(map(tuple, e))


def test_sphere_is_closed_oriented_and_on_the_level_set():
    g, c = _grid(32)
    sdf = np.linalg.norm(g - c, axis=-1) - 10.3
    v, f = mesh.marching_cubes(sdf, 0.02)
    und, dirc = _edges(f)
    assert set(und.values()) == {2}                      # watertight, manifold
    assert max(dirc.values()) == 1 and all((b, a) in dirc for (a, b) in dirc)     # consistently oriented
    assert len(v) - len(und) + len(f) == 2               # Euler characteristic of a sphere
    assert np.abs(np.linalg.norm(v - c, axis=1) - 10.32).max() < 0.02     # linear interpolation of a distance field
    p0, p1, p2 = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    vol = np.einsum("ij,ij->i", p0, np.cross(p1, p2)).sum() / 6
    assert 0.98 < vol / (4 / 3 * np.pi * 10.32 ** 3) < 1.0        # normals point towards increasing values; inscribed polyhedron


def test_torus_has_genus_one():
    g, c = _grid(40)
    q = np.sqrt((g[..., 0] - c) ** 2 + (g[..., 1] - c) ** 2) - 11.0
    sdf = np.sqrt(q ** 2 + (g[..., 2] - c) ** 2) - 4.2
    v, f = mesh.marching_cubes(sdf, 0.0)
    und, _ = _edges(f)
    assert set(und.values()) == {2}
    assert len(v) - len(und) + len(f) == 0


def test_noise_field_stays_manifold_and_vertices_are_the_crossing_edges():
    rng = np.random.default_rng(0)
    vol = rng.standard_normal((14, 17, 20))
    v, f = mesh.marching_cubes(vol, 0.02)
    und, dirc = _edges(f)
    assert set(und.values()) <= {1, 2} and max(dirc.values()) == 1      # 1 = open at the grid boundary only
    inside = vol <= 0.02
    n_cross = int((inside[:-1] != inside[1:]).sum() + (inside[:, :-1] != inside[:, 1:]).sum() + (inside[:, :, :-1] != inside[:, :, 1:]).sum())
    assert len(v) == n_cross and f.min() == 0 and f.max() == len(v) - 1
    # every vertex lies on exactly one grid edge, strictly between its ends or on one
    frac = v - np.floor(v)
    assert ((frac > 0).sum(1) <= 1).all()


def test_value_equal_to_the_level_counts_as_inside_and_empty_surfaces_are_empty():
    vol = np.full((4, 4, 4), 1.0)
    assert mesh.surface_vertices(vol, 0.02).shape == (0, 3)
    vol[1, 1, 1] = 0.02                                   # <= level: inside
    v, f = mesh.marching_cubes(vol, 0.02)
    assert len(v) == 6 and len(f) == 8                    # an octahedron collapsed onto the grid point
    assert np.allclose(v, 1.0)


def test_baked_table_of_the_product_equals_the_construction():
    spec = importlib.util.spec_from_file_location("_mc_table", os.path.join(ROOT, "commonscenes_b200", "model", "diff_utils", "_mc_table.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    cnt, tab = mesh.triangle_table()
    assert m.MAX_TRIS == tab.shape[1]
    assert m.TRI_COUNT == cnt.tobytes() and m.TRI_TABLE == tab.tobytes()
    assert int(cnt.sum()) == 820 and cnt[0] == 0 and cnt[255] == 0


def test_sdf_to_verts_scaling_follows_the_reference():
    g, c = _grid(16)
    sdf = (np.linalg.norm(g - c, axis=-1) - 5.0)[None, None].astype(np.float32)
    (v,) = mesh.sdf_to_verts(sdf, 0.02)
    assert v.dtype == np.float32 and np.abs(v).max() < 0.5
    assert np.allclose(np.linalg.norm(v * 16 + 8 - c, axis=1), 5.02, atol=0.03)     # verts / n_cell - .5  (util_3d.py:221)


@pytest.mark.parametrize("n_points", [12000, 5000, 700])
def test_sample_points_mirror_draws_the_reference_indices(n_points):
    from commonscenes_b200.helpers.util import sample_points
    pts = torch.arange(n_points * 3, dtype=torch.float32).view(n_points, 3)
    torch.manual_seed(7)
    (got,) = sample_points([pts], 5000)
    torch.manual_seed(7)      # helpers/util.py:31-45: randperm(n)[:num] when n >= num, randint(n, (num,)) otherwise
    idx = torch.randperm(n_points)[:5000] if n_points >= 5000 else torch.randint(n_points, size=(5000,))
    assert got.shape == (5000, 3) and torch.equal(got, pts[idx])
