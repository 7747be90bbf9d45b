"""Golden vectors of the point-cloud path, produced by the REFERENCE'S OWN CUDA kernels (oracle/_ref/libref_points.so =
scripts/pytorch_structural_losses/src/{approxmatch,nndistance}.cu compiled unmodified, oracle/build_ref.py).  CUDA code
only runs on the GPU box:

    gpurun -- 'python tests/golden/make_golden_points.py gpurun_out/points_ref.npz'      # then copy to tests/golden/

Cases: ragged sizes (n != m, not multiples of the block sizes, integer-quotient weights m / n = 11), a lattice case full
of exact ties (tie-break rule), one 256 x 256 case.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import points as P  # noqa: E402

CASES = {"ragged": (3, 160, 100, "uniform"), "wide": (1, 64, 700, "normal"), "square": (2, 256, 256, "uniform"),
         "lattice": (2, 300, 260, "lattice")}


def make_inputs(name):
    b, n, m, kind = CASES[name]
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    if kind == "uniform":
        a, c = torch.rand(b, n, 3, generator=g), torch.rand(b, m, 3, generator=g)
    elif kind == "normal":
        a, c = torch.randn(b, n, 3, generator=g) * 0.3, torch.randn(b, m, 3, generator=g) * 0.3
    else:   # coordinates k/8: every product is exact in fp32, many equal distances
        a = torch.randint(0, 5, (b, n, 3), generator=g).float() / 8
        c = torch.randint(0, 5, (b, m, 3), generator=g).float() / 8
    return a.contiguous(), c.contiguous()


def main(out):
    res = {}
    for name in CASES:
        a, c = make_inputs(name)
        A, Cc = a.cuda(), c.cuda()
        d1, i1, d2, i2 = P.ref_nn_distance(A, Cc)
        g = torch.Generator().manual_seed(5)
        gd1, gd2 = torch.randn(d1.shape, generator=g).cuda(), torch.randn(d2.shape, generator=g).cuda()
        ga, gc = P.ref_nn_distance_grad(A, Cc, i1, i2, gd1, gd2)
        match = P.ref_approx_match(A, Cc)
        cost = P.ref_match_cost(A, Cc, match)
        m1, m2 = P.ref_match_cost_grad(A, Cc, match)
        torch.cuda.synchronize()
        for k, v in dict(xyz1=a, xyz2=c, dist1=d1, idx1=i1, dist2=d2, idx2=i2, gd1=gd1, gd2=gd2, gxyz1=ga, gxyz2=gc, match=match,
                         cost=cost, mgrad1=m1, mgrad2=m2).items():
            res[f"{name}.{k}"] = v.cpu().numpy()
        print(name, "cost", cost.tolist())
    np.savez_compressed(out, **res)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/points_ref.npz")
