"""SDF -> surface points for a batch of decoded 64^3 grids: the device extraction (cs_surface_count / cs_surface_emit, incl. the
host read of the totals) next to the numpy oracle on the host (what the reference does per object with PyMCubes on one core)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from commonscenes_b200.model.diff_utils import util_3d

B = int(sys.argv[sys.argv.index("--b") + 1]) if "--b" in sys.argv else 32
n = 64
g = np.stack(np.meshgrid(*[np.arange(n)] * 3, indexing="ij"), -1).astype(np.float32)
rng = np.random.default_rng(0)
grids = np.stack([(np.linalg.norm((g - rng.uniform(0.4, 0.6, 3) * n) / rng.uniform(0.6, 1.4, 3), axis=-1) - rng.uniform(0.15, 0.3) * n) / n
                  for _ in range(B)]).astype(np.float32)
d = torch.from_numpy(grids).cuda()
for _ in range(3):
    util_3d.surface_extract(d, 0.02)
torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 20
for _ in range(reps):
    verts, faces, tot = util_3d.surface_extract(d, 0.02)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / reps * 1e3
print(f"device: {B} grids of 64^3 -> {int(tot[:, 0].sum())} vertices, {int(tot[:, 1].sum())} triangles in {ms:.3f} ms "
      f"(wall, incl. the totals read-back; {B * n ** 3 * 4 / ms / 1e6:.1f} GB/s of grid)")
if "--cpu" in sys.argv:
    from oracle import mesh
    t0 = time.perf_counter()
    k = min(B, 4)
    for i in range(k):
        mesh.marching_cubes(grids[i], 0.02)
    print(f"numpy oracle on the host (1 core): {(time.perf_counter() - t0) / k * 1e3:.1f} ms per grid + "
          f"a {n ** 3 * 4 / 1e6:.1f} MB device->host copy each")
