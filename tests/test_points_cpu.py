"""CPU checks of the point-cloud path's test infrastructure: the C oracle (oracle/points.c) against brute-force
definitions and against the golden vectors produced by the reference's own CUDA kernels on a B200
(tests/golden/points_ref.npz, tests/golden/make_golden_points.py), and the host-side argument checks of the product."""
import os

import numpy as np
import pytest
import torch

from oracle import points as P

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "points_ref.npz")


def test_oracle_nn_distance_is_the_lowest_index_argmin():
    rng = np.random.default_rng(0)
    for b, n, m in ((2, 70, 45), (1, 5, 600)):
        a, c = rng.random((b, n, 3), dtype=np.float32), rng.random((b, m, 3), dtype=np.float32)
        d1, i1, d2, i2 = P.nn_distance(a, c)
        D = ((a[:, :, None, :].astype(np.float64) - c[:, None, :, :]) ** 2).sum(-1)
        assert np.array_equal(i1, D.argmin(2)) and np.array_equal(i2, D.argmin(1))
        np.testing.assert_allclose(d1, D.min(2), rtol=1e-6, atol=1e-9)
    # exact ties on a lattice: the first index wins (chamfer.cu:30-33 strict '<', :127 strict '>')
    a = (rng.integers(0, 3, (1, 50, 3)) / 4).astype(np.float32)
    c = (rng.integers(0, 3, (1, 90, 3)) / 4).astype(np.float32)
    d1, i1, _, _ = P.nn_distance(a, c)
    D = ((a[:, :, None, :] - c[:, None, :, :]) ** 2).sum(-1)
    assert np.array_equal(i1, D.argmin(2)) and np.array_equal(d1, D.min(2))
    d, i, _, _ = P.nn_distance(a, np.zeros((1, 0, 3), np.float32))
    assert not d.any() and not i.any()


def test_oracle_transport_plan_properties_and_cost_gradient():
    rng = np.random.default_rng(1)
    a, c = rng.random((2, 96, 3), dtype=np.float32), rng.random((2, 96, 3), dtype=np.float32)
    m = P.approx_match(a, c)
    assert m.min() >= 0 and m.sum(1).max() <= 1 + 1e-4 and m.sum(2).max() <= 1 + 1e-4 and m.sum((1, 2)).min() > 0.99 * 96
    cost = P.match_cost(a, c, m)
    dist = np.sqrt(((c[:, :, None, :] - a[:, None, :, :]) ** 2).sum(-1))          # (b, m, n)
    np.testing.assert_allclose(cost, (m * dist).sum((1, 2)), rtol=1e-5)
    g1, g2 = P.match_cost_grad(a, c, m)
    eps = 1e-3
    for (k, u) in ((3, 0), (50, 2)):
        a2 = a.copy(); a2[0, k, u] += eps
        fd = ((m * np.sqrt(((c[:, :, None, :] - a2[:, None, :, :]) ** 2).sum(-1))).sum((1, 2))[0] - (m * dist).sum((1, 2))[0]) / eps
        assert abs(fd - g1[0, k, u]) < 5e-3 * max(1.0, abs(fd))
        c2 = c.copy(); c2[1, k, u] += eps
        fd = ((m * np.sqrt(((c2[:, :, None, :] - a[:, None, :, :]) ** 2).sum(-1))).sum((1, 2))[1] - (m * dist).sum((1, 2))[1]) / eps
        assert abs(fd - g2[1, k, u]) < 5e-3 * max(1.0, abs(fd))
    # unequal sizes: the smaller side carries integer-quotient weights (approxmatch.cu:5-11)
    a, c = rng.random((1, 20, 3), dtype=np.float32), rng.random((1, 65, 3), dtype=np.float32)
    m = P.approx_match(a, c)
    assert m.sum(1).max() <= 3 + 1e-3 and m.sum(2).max() <= 1 + 1e-4


@pytest.mark.skipif(not os.path.exists(GOLDEN), reason="tests/golden/points_ref.npz not generated yet")
def test_oracle_is_pinned_to_the_reference_kernels_golden_vectors():
    z = np.load(GOLDEN)
    for name in sorted({k.split(".")[0] for k in z.files}):
        a, c = z[f"{name}.xyz1"], z[f"{name}.xyz2"]
        d1, i1, d2, i2 = P.nn_distance(a, c)
        assert np.array_equal(i1, z[f"{name}.idx1"]) and np.array_equal(i2, z[f"{name}.idx2"]), name
        assert np.array_equal(d1, z[f"{name}.dist1"]) and np.array_equal(d2, z[f"{name}.dist2"]), name
        g1, g2 = P.nn_distance_grad(a, c, i1, i2, z[f"{name}.gd1"], z[f"{name}.gd2"])
        np.testing.assert_allclose(g1, z[f"{name}.gxyz1"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(g2, z[f"{name}.gxyz2"], rtol=1e-5, atol=1e-6)
        m = P.approx_match(a, c)
        ref = z[f"{name}.match"]
        assert np.abs(m - ref).max() <= 2e-4 * ref.max(), name
        np.testing.assert_allclose(P.match_cost(a, c, ref), z[f"{name}.cost"], rtol=1e-5)
        q1, q2 = P.match_cost_grad(a, c, ref)
        np.testing.assert_allclose(q1, z[f"{name}.mgrad1"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(q2, z[f"{name}.mgrad2"], rtol=1e-4, atol=1e-5)


def test_product_rejects_cpu_tensors_and_bad_shapes():
    from commonscenes_b200 import _lib, ops_points
    a = torch.rand(2, 8, 3)
    for fn in (ops_points.nn_distance, ops_points.approx_match):
        with pytest.raises(_lib.CsError):
            fn(a, a)                                         # no CPU path
    import commonscenes_b200.extension.dist_chamfer as ext
    from commonscenes_b200.scripts.pytorch_structural_losses import match_cost, nn_distance
    with pytest.raises(_lib.CsError):
        ext.chamferDist()(a, a)
    with pytest.raises(_lib.CsError):
        nn_distance(a, a)
    with pytest.raises(_lib.CsError):
        match_cost(a, a)
