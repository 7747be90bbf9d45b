"""GPU parity of the point-cloud kernels, through the C ABI:

* against the reference's OWN CUDA kernels (oracle/_ref/libref_points.so, built unmodified from
  scripts/pytorch_structural_losses/src/*.cu): BIT-EXACT -- distances, indices, the match matrix, cost and cost gradients
  (the nearest-neighbour gradient uses float atomics in the reference too: tolerance 1e-6);
* against the C oracle (oracle/points.c) and the committed golden vectors of the reference kernels
  (tests/golden/points_ref.npz);
* size-independent properties at the evaluation size (2048 x 2048 points).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "points_ref.npz")


def _ops():
    from commonscenes_b200 import _lib, ops_points
    _lib.require_device()
    return ops_points


def _sets(b, n, m, seed, kind="uniform"):
    g = torch.Generator().manual_seed(seed)
    if kind == "lattice":
        return (torch.randint(0, 5, (b, n, 3), generator=g).float() / 8).cuda(), (torch.randint(0, 5, (b, m, 3), generator=g).float() / 8).cuda()
    return torch.rand(b, n, 3, generator=g).cuda(), torch.rand(b, m, 3, generator=g).cuda()


SHAPES = [(1, 1, 1), (2, 37, 5), (3, 160, 100), (1, 64, 700), (2, 300, 260), (2, 2048, 2048), (5, 1000, 2500), (1, 4100, 2049),
          (33, 256, 256)]


@pytest.mark.parametrize("b,n,m", SHAPES)
@pytest.mark.parametrize("kind", ["uniform", "lattice"])
def test_nn_distance_bit_exact_vs_reference_kernels(b, n, m, kind):
    from oracle import points as P
    if not P.reference_available():
        pytest.skip("oracle/_ref/libref_points.so not built")
    ops = _ops()
    a, c = _sets(b, n, m, 11 + n, kind)
    d1, i1, d2, i2 = ops.nn_distance(a, c)
    r1, j1, r2, j2 = P.ref_nn_distance(a, c)
    assert torch.equal(i1, j1) and torch.equal(i2, j2)
    assert torch.equal(d1, r1) and torch.equal(d2, r2)
    g = torch.Generator().manual_seed(3)
    gd1, gd2 = torch.randn(b, n, generator=g).cuda(), torch.randn(b, m, generator=g).cuda()
    ga, gc = ops.nn_distance_grad(a, c, i1, i2, gd1, gd2)
    ra, rc = P.ref_nn_distance_grad(a, c, j1, j2, gd1, gd2)
    torch.testing.assert_close(ga, ra, rtol=1e-5, atol=1e-6)      # float atomics on both sides
    torch.testing.assert_close(gc, rc, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("b,n,m", [(2, 37, 5), (3, 160, 100), (1, 64, 700), (2, 300, 260), (2, 2048, 2048), (3, 1000, 2500),
                                   (1, 2500, 1000), (33, 256, 256), (20, 1024, 1024)])
def test_approx_match_bit_exact_vs_reference_kernels(b, n, m):
    from oracle import points as P
    if not P.reference_available():
        pytest.skip("oracle/_ref/libref_points.so not built")
    ops = _ops()
    a, c = _sets(b, n, m, 5 + m)
    match, _ = ops.approx_match(a, c)
    ref = P.ref_approx_match(a, c)
    assert torch.equal(match, ref), f"max |diff| {float((match - ref).abs().max())}"
    cost = ops.match_cost(a, c, match)
    assert torch.equal(cost, P.ref_match_cost(a, c, ref))
    g1, g2 = ops.match_cost_grad(a, c, match)
    r1, r2 = P.ref_match_cost_grad(a, c, ref)
    assert torch.equal(g1, r1) and torch.equal(g2, r2)


@pytest.mark.parametrize("b,n,m", [(2, 37, 5), (3, 160, 100), (1, 64, 700), (2, 512, 640)])
def test_against_c_oracle(b, n, m):
    from oracle import points as P
    ops = _ops()
    a, c = _sets(b, n, m, 21)
    d1, i1, d2, i2 = ops.nn_distance(a, c)
    o1, p1, o2, p2 = P.nn_distance(a.cpu().numpy(), c.cpu().numpy())
    assert np.array_equal(i1.cpu().numpy(), p1) and np.array_equal(i2.cpu().numpy(), p2)        # index work: bit-exact
    assert np.array_equal(d1.cpu().numpy(), o1) and np.array_equal(d2.cpu().numpy(), o2)        # same fma sequence
    match, _ = ops.approx_match(a, c)
    om = P.approx_match(a.cpu().numpy(), c.cpu().numpy())
    # __expf (2 ulp, GPU) vs expf feeding 27 dependent passes: tolerance 2e-4 of the largest entry
    assert np.abs(match.cpu().numpy() - om).max() <= 2e-4 * om.max()
    cost = ops.match_cost(a, c, match).cpu().numpy()
    np.testing.assert_allclose(cost, P.match_cost(a.cpu().numpy(), c.cpu().numpy(), match.cpu().numpy()), rtol=1e-5)
    g1, g2 = ops.match_cost_grad(a, c, match)
    q1, q2 = P.match_cost_grad(a.cpu().numpy(), c.cpu().numpy(), match.cpu().numpy())
    np.testing.assert_allclose(g1.cpu().numpy(), q1, rtol=1e-4, atol=1e-5)     # rsqrtf approximation on the GPU
    np.testing.assert_allclose(g2.cpu().numpy(), q2, rtol=1e-4, atol=1e-5)


def test_against_golden_vectors_of_the_reference_kernels():
    if not os.path.exists(GOLDEN):
        pytest.skip("tests/golden/points_ref.npz not generated yet")
    ops = _ops()
    z = np.load(GOLDEN)
    for name in sorted({k.split(".")[0] for k in z.files}):
        a, c = torch.from_numpy(z[f"{name}.xyz1"]).cuda(), torch.from_numpy(z[f"{name}.xyz2"]).cuda()
        d1, i1, d2, i2 = ops.nn_distance(a, c)
        for got, key in ((d1, "dist1"), (i1, "idx1"), (d2, "dist2"), (i2, "idx2")):
            assert np.array_equal(got.cpu().numpy(), z[f"{name}.{key}"]), (name, key)
        match, _ = ops.approx_match(a, c)
        assert np.array_equal(match.cpu().numpy(), z[f"{name}.match"]), name
        assert np.array_equal(ops.match_cost(a, c, match).cpu().numpy(), z[f"{name}.cost"]), name
        g1, g2 = ops.match_cost_grad(a, c, match)
        assert np.array_equal(g1.cpu().numpy(), z[f"{name}.mgrad1"]) and np.array_equal(g2.cpu().numpy(), z[f"{name}.mgrad2"]), name
        ga, gc = ops.nn_distance_grad(a, c, i1, i2, torch.from_numpy(z[f"{name}.gd1"]).cuda(), torch.from_numpy(z[f"{name}.gd2"]).cuda())
        np.testing.assert_allclose(ga.cpu().numpy(), z[f"{name}.gxyz1"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(gc.cpu().numpy(), z[f"{name}.gxyz2"], rtol=1e-5, atol=1e-6)


def test_properties_at_evaluation_size():
    """2048 x 2048 points, batch 8 (scripts/compute_mmd_cov_1nn.py works on 2048-point clouds): properties that do not need
    a reference -- symmetry under swapping the sets, zero self-distance, transport plan marginals, permutation covariance."""
    ops = _ops()
    a, c = _sets(8, 2048, 2048, 99)
    d1, i1, d2, i2 = ops.nn_distance(a, c)
    e2, k2, e1, k1 = ops.nn_distance(c, a)
    assert torch.equal(d1, e1) and torch.equal(i1, k1) and torch.equal(d2, e2) and torch.equal(i2, k2)
    s1, t1, _, _ = ops.nn_distance(a, a)
    assert float(s1.abs().max()) == 0.0 and torch.equal(t1, torch.arange(2048, device="cuda", dtype=torch.int32).expand(8, -1))
    # gathered distance equals the reported one
    near = torch.gather(c, 1, i1.long().unsqueeze(-1).expand(-1, -1, 3))
    torch.testing.assert_close(((a - near) ** 2).sum(-1), d1, rtol=1e-5, atol=1e-7)
    match, _ = ops.approx_match(a, c)
    assert float(match.min()) >= 0.0
    assert float(match.sum(1).max()) <= 1.0 + 1e-4 and float(match.sum(2).max()) <= 1.0 + 1e-4      # nobody ships more than it has
    assert float(match.sum((1, 2)).min()) >= 0.99 * 2048                                           # (almost) everything is matched
    cost = ops.match_cost(a, c, match)
    perm = torch.randperm(2048, generator=torch.Generator().manual_seed(1)).cuda()
    match_p, _ = ops.approx_match(a[:, perm].contiguous(), c)
    # relabelling the left set permutes the plan; the fp32 sums over the left points change their order, and 27 dependent
    # passes amplify that: entries (<= 1) agree to 2e-3, the cost to 1e-3
    assert float((match_p - match[:, :, perm]).abs().max()) <= 2e-3
    torch.testing.assert_close(ops.match_cost(a[:, perm].contiguous(), c, match_p), cost, rtol=1e-3, atol=0)
    # EMD between a cloud and itself is ~0 and far below the EMD to another cloud
    self_cost = ops.match_cost(a, a, ops.approx_match(a, a)[0])
    assert float(self_cost.max()) < 0.05 * float(cost.min())


def test_edge_cases_and_errors():
    from commonscenes_b200 import _lib
    ops = _ops()
    e = torch.empty(0, 10, 3, device="cuda")
    d1, i1, d2, i2 = ops.nn_distance(e, torch.empty(0, 7, 3, device="cuda"))
    assert d1.shape == (0, 10) and i2.shape == (0, 7)
    a = torch.rand(2, 9, 3, device="cuda")
    d1, i1, d2, i2 = ops.nn_distance(a, torch.empty(2, 0, 3, device="cuda"))
    assert float(d1.abs().max()) == 0 and int(i1.abs().max()) == 0 and d2.shape == (2, 0)       # dist_chamfer.py's zero-initialised outputs
    with pytest.raises(_lib.CsError):
        ops.nn_distance(a.cpu(), a.cpu())
    with pytest.raises(_lib.CsError):
        ops.nn_distance(a, torch.rand(3, 9, 3, device="cuda"))
    with pytest.raises(_lib.CsError):
        ops.nn_distance(a.double(), a.double())
    with pytest.raises(_lib.CsError):
        ops.approx_match(a, torch.empty(2, 0, 3, device="cuda"))
    with pytest.raises(_lib.CsError):
        ops.nn_distance(torch.rand(2, 9, 6, device="cuda"), a)


def test_reference_api_mirrors_and_autograd():
    """extension/dist_chamfer.py and scripts/pytorch_structural_losses/{nn_distance,match_cost}.py as the scripts use them
    (eval_3dfront.py:395-397, compute_mmd_cov_1nn.py:25-62), gradients against the C oracle."""
    from oracle import points as P
    _ops()
    import commonscenes_b200.extension.dist_chamfer as ext
    from commonscenes_b200.scripts.pytorch_structural_losses import match_cost, nn_distance
    a, c = _sets(2, 150, 130, 8)
    a.requires_grad_(True); c.requires_grad_(True)
    chamfer = ext.chamferDist()
    dist1, dist2 = chamfer(a, c)
    (torch.mean(dist1) + torch.mean(dist2)).backward()
    o1, p1, o2, p2 = P.nn_distance(a.detach().cpu().numpy(), c.detach().cpu().numpy())
    g1, g2 = P.nn_distance_grad(a.detach().cpu().numpy(), c.detach().cpu().numpy(), p1, p2, np.full_like(o1, 1 / o1.size),
                                np.full_like(o2, 1 / o2.size))
    np.testing.assert_allclose(a.grad.cpu().numpy(), g1, rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(c.grad.cpu().numpy(), g2, rtol=1e-5, atol=1e-8)
    dl, dr = nn_distance(a.detach(), c.detach())
    assert torch.equal(dl, dist1.detach()) and torch.equal(dr, dist2.detach())
    a.grad = None; c.grad = None
    a2, c2 = _sets(2, 128, 128, 9)
    a2.requires_grad_(True); c2.requires_grad_(True)
    emd = match_cost(a2, c2)
    (emd / 128.0).sum().backward()
    mt = P.approx_match(a2.detach().cpu().numpy(), c2.detach().cpu().numpy())
    q1, q2 = P.match_cost_grad(a2.detach().cpu().numpy(), c2.detach().cpu().numpy(), mt)
    np.testing.assert_allclose(a2.grad.cpu().numpy(), q1 / 128.0, rtol=2e-3, atol=2e-6)
    np.testing.assert_allclose(c2.grad.cpu().numpy(), q2 / 128.0, rtol=2e-3, atol=2e-6)
