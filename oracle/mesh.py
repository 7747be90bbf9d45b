"""TEST INFRASTRUCTURE -- CPU restatement (numpy) of the SDF -> surface-points step of the evaluation chain.

Reference call sites (the path SURVEY.md 8(f)-3 names):
  * model/diff_utils/util_3d.py:194-235  sdf_to_mesh: per object `mcubes.marching_cubes(sdf_i, level)` on the CPU,
    `verts / n_cell - .5`, pytorch3d Meshes;
  * scripts/eval_3dfront.py:313-317, 589-592: ONLY `.verts_list()` of that mesh is consumed -- the vertices are
    re-sampled to 5000 points by helpers/util.py:31-45 `sample_points` (randperm / randint over the VERTICES, no
    surface-area sampling) and go to the Chamfer distance.

The algorithm lives in a third-party dependency that is absent from /root/reference and from this image: **PyMCubes**
(`import mcubes`, un-pinned in the reference's environment; pytorch3d likewise).  **Parity unpinned**: what is restated
here is its published algorithm (Lorensen-Cline marching cubes with vertices shared per grid edge):
  * a grid corner is "inside" when value <= isovalue;
  * every grid edge whose two corners differ gets exactly ONE vertex, at x1 + (isovalue - f1) / (f2 - f1) along the edge
    (double precision; coordinates in index units, array axis order (x, y, z));
  * triangles connect the edge vertices of each cell.
The vertex SET is fully determined by that description and is what the evaluation consumes; the vertex ORDER (PyMCubes
walks x-slabs and appends three edges per cell) and its 256-case triangle table could not be checked against the
package, so: vertices are emitted in (cell linear index, axis) order, and the triangles come from a table GENERATED here
(closed loops of the edge crossings on the six cell faces, ambiguous faces resolved by separating the inside corners --
the same rule on both sides of a shared face, so the mesh is watertight -- fan-triangulated, oriented towards
increasing values).  Same vertices, a valid surface, not necessarily PyMCubes' diagonal choices.

Only tests/, smoke() and bench legs may import this.
"""
from __future__ import annotations

import numpy as np

__all__ = ["edge_tables", "triangle_table", "marching_cubes", "surface_vertices", "sdf_to_verts", "sample_points_indices"]


# ----------------------------------------------------------------------------------------------
# cell topology: corner c = dx + 2 dy + 4 dz; edge e = 4 * axis + u + 2 v with (u, v) the offsets along the two other
# axes (in increasing axis order); the edge is owned by the voxel at its lower end.
# ----------------------------------------------------------------------------------------------
def edge_tables():
    """-> (edge_corners (12, 2), edge_owner_offset (12, 3), edge_axis (12,))."""
    corners = np.zeros((12, 2), np.int64)
    owner = np.zeros((12, 3), np.int64)
    axis = np.zeros(12, np.int64)
    for a in range(3):
        others = [b for b in range(3) if b != a]
        for v in range(2):
            for u in range(2):
                e = 4 * a + u + 2 * v
                o = [0, 0, 0]
                o[others[0]], o[others[1]] = u, v
                lo = o[0] + 2 * o[1] + 4 * o[2]
                corners[e] = (lo, lo + (1 << a))
                owner[e] = o
                axis[e] = a
    return corners, owner, axis


def _face_cycles():
    """The 6 faces as 4 corners in counter-clockwise order seen from OUTSIDE the cell."""
    faces = []
    for n in range(3):
        p, q = (n + 1) % 3, (n + 2) % 3          # e_p x e_q = e_n
        for s in range(2):
            ring = [(0, 0), (1, 0), (1, 1), (0, 1)]
            if s == 0:
                ring = ring[::-1]                # outside is -n: reverse to stay counter-clockwise
            cyc = []
            for (a, b) in ring:
                o = [0, 0, 0]
                o[n], o[p], o[q] = s, a, b
                cyc.append(o[0] + 2 * o[1] + 4 * o[2])
            faces.append(cyc)
    return faces


_TABLE = None


def triangle_table():
    """-> (count (256,) uint8, tris (256, T, 3) uint8 edge ids, padded with 255).  Bit c of the case index = corner c inside."""
    global _TABLE
    if _TABLE is not None:
        return _TABLE
    corners, _, _ = edge_tables()
    edge_of = {}
    for e, (c0, c1) in enumerate(corners):
        edge_of[(int(c0), int(c1))] = e
        edge_of[(int(c1), int(c0))] = e
    faces = _face_cycles()
    all_tris = []
    for case in range(256):
        inside = [(case >> c) & 1 for c in range(8)]
        nxt = {}
        for cyc in faces:
            for j in range(4):
                a, b = cyc[j], cyc[(j + 1) % 4]
                if inside[a] and not inside[b]:              # a segment starts where the ring leaves the inside ...
                    i = j
                    while True:                              # ... and ends at the nearest earlier edge that enters it
                        i = (i - 1) % 4
                        c, d = cyc[i], cyc[(i + 1) % 4]
                        if not inside[c] and inside[d]:
                            break
                    nxt[edge_of[(a, b)]] = edge_of[(c, d)]
        tris, seen = [], set()
        for e0 in sorted(nxt):
            if e0 in seen:
                continue
            loop, e = [], e0
            while e not in seen:
                seen.add(e)
                loop.append(e)
                e = nxt[e]
            assert e == e0 and len(loop) >= 3
            # fan apex: the rotation whose diagonals avoid chords lying IN a cell face (two loop vertices on one face):
            # the neighbouring cell may draw the same chord, which would make that edge non-manifold
            def in_face_chords(rot):
                n = len(rot)
                return sum(1 for k in range(2, n - 1) if any(set(corners[rot[0]]) | set(corners[rot[k]]) <= set(cyc) for cyc in faces))
            rots = [loop[i:] + loop[:i] for i in range(len(loop))]
            loop = min(rots, key=in_face_chords)                  # min() keeps the first of equals: deterministic
            for k in range(1, len(loop) - 1):
                tris.append((loop[0], loop[k + 1], loop[k]))     # orientation: normals towards increasing values (tested)
        all_tris.append(tris)
    T = max(len(t) for t in all_tris)
    count = np.array([len(t) for t in all_tris], np.uint8)
    table = np.full((256, T, 3), 255, np.uint8)
    for case, tris in enumerate(all_tris):
        for k, t in enumerate(tris):
            table[case, k] = t
    _TABLE = (count, table)
    return _TABLE


# ----------------------------------------------------------------------------------------------
def _edge_flags(vol: np.ndarray, level: float):
    inside = vol <= level
    flags = np.zeros(vol.shape + (3,), bool)
    flags[:-1, :, :, 0] = inside[:-1] != inside[1:]
    flags[:, :-1, :, 1] = inside[:, :-1] != inside[:, 1:]
    flags[:, :, :-1, 2] = inside[:, :, :-1] != inside[:, :, 1:]
    return inside, flags


def surface_vertices(vol: np.ndarray, level: float) -> np.ndarray:
    """(V, 3) float64 edge-crossing vertices in index units, ordered by (owner voxel linear index, axis)."""
    vol = np.asarray(vol, np.float64)
    _, flags = _edge_flags(vol, level)
    x, y, z, a = np.nonzero(flags)                   # C order = voxel linear index, then axis
    p = np.stack([x, y, z], 1)
    q = p.copy()
    q[np.arange(len(a)), a] += 1
    f1 = vol[p[:, 0], p[:, 1], p[:, 2]]
    f2 = vol[q[:, 0], q[:, 1], q[:, 2]]
    t = (level - f1) / (f2 - f1)
    v = p.astype(np.float64)
    v[np.arange(len(a)), a] = p[np.arange(len(a)), a] + t
    return v


def marching_cubes(vol: np.ndarray, level: float):
    """-> (verts (V, 3) float64 in index units, faces (F, 3) int64)."""
    vol = np.asarray(vol, np.float64)
    inside, flags = _edge_flags(vol, level)
    verts = surface_vertices(vol, level)
    index = np.full(flags.shape, -1, np.int64)
    index[flags] = np.arange(int(flags.sum()))
    count, table = triangle_table()
    _, owner, axis = edge_tables()
    i = inside.astype(np.int64)
    case = np.zeros(tuple(s - 1 for s in vol.shape), np.int64)
    for c in range(8):
        dx, dy, dz = c & 1, (c >> 1) & 1, (c >> 2) & 1
        case |= i[dx:vol.shape[0] - 1 + dx, dy:vol.shape[1] - 1 + dy, dz:vol.shape[2] - 1 + dz] << c
    cx, cy, cz = np.nonzero(count[case] > 0)
    faces = []
    for x, y, z in zip(cx, cy, cz):
        cs = case[x, y, z]
        for k in range(count[cs]):
            tri = []
            for e in table[cs, k]:
                o = owner[e]
                tri.append(index[x + o[0], y + o[1], z + o[2], axis[e]])
            faces.append(tri)
    return verts, np.asarray(faces, np.int64).reshape(-1, 3)


def sdf_to_verts(sdf: np.ndarray, level: float = 0.02):
    """util_3d.py:194-235 up to `.verts_list()`: list of (V_i, 3) float32 arrays, `verts / n_cell - .5`."""
    n_cell = sdf.shape[-1]
    out = []
    for i in range(sdf.shape[0]):
        v = surface_vertices(sdf[i, 0], level)
        out.append((v / n_cell - .5).astype(np.float32))
    return out


def sample_points_indices(n_points: int, num: int, generator=None):
    """helpers/util.py:31-45: the index vector `sample_points` draws for one cloud (torch CPU generator)."""
    import torch
    if n_points >= num:
        return torch.randperm(n_points, generator=generator)[:num]
    return torch.randint(n_points, size=(num,), generator=generator)
