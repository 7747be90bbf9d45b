"""Point-cloud kernels on one B200: libcsb200.so next to the reference's own CUDA kernels
(oracle/_ref/libref_points.so, unmodified sources) on the sizes of scripts/compute_mmd_cov_1nn.py (2048-point clouds).
CUDA events on the current stream, 3 warm-ups, mean of 10; inputs resident in HBM."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from commonscenes_b200 import ops_points as ops
from oracle import points as P


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    have_ref = P.reference_available()
    rows = []
    for b, n in ((32, 2048), (200, 2048), (8, 2048), (64, 1024)):
        g = torch.Generator().manual_seed(b)
        a, c = torch.rand(b, n, 3, generator=g).cuda(), torch.rand(b, n, 3, generator=g).cuda()
        match, _ = ops.approx_match(a, c)
        row = {"batch": b, "points": n}
        row["nn_distance_ms"] = timeit(lambda: ops.nn_distance(a, c))
        row["approx_match_ms"] = timeit(lambda: ops.approx_match(a, c), reps=3, warm=1)
        row["match_cost_ms"] = timeit(lambda: ops.match_cost(a, c, match))
        row["match_cost_grad_ms"] = timeit(lambda: ops.match_cost_grad(a, c, match))
        # 27 passes x n x m pair evaluations (one exponential each) per batch element
        row["approx_match_Gpairs_per_s"] = 27.0 * n * n * b / row["approx_match_ms"] / 1e6
        row["nn_Gpairs_per_s"] = 2.0 * n * n * b / row["nn_distance_ms"] / 1e6
        if have_ref:
            row["ref_nn_distance_ms"] = timeit(lambda: P.ref_nn_distance(a, c))
            row["ref_approx_match_ms"] = timeit(lambda: P.ref_approx_match(a, c), reps=2, warm=1)
            row["ref_match_cost_ms"] = timeit(lambda: P.ref_match_cost(a, c, match))
            row["ref_match_cost_grad_ms"] = timeit(lambda: P.ref_match_cost_grad(a, c, match), reps=3, warm=1)
            for k in ("nn_distance", "approx_match", "match_cost", "match_cost_grad"):
                row[f"speedup_{k}"] = row[f"ref_{k}_ms"] / row[f"{k}_ms"]
        rows.append(row)
        print(json.dumps(row))
    return rows


if __name__ == "__main__":
    main()
