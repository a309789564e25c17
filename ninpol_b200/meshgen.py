"""Seeded synthetic meshes and fields of the shapes BASELINE.json names (SURVEY.md §8 d).

The reference's own test meshes (tests/mesh/*.msh) are git-ignored upstream and not shipped, and
`meshio` is not installed here, so every workload is generated: a structured hexahedral unit box, a
Kuhn 6-tetrahedra-per-cube box with perturbed interior nodes, and a conforming mixed
hexahedron / wedge / pyramid / tetra box.  `SimpleMesh` / `CellBlock` are duck-typed to what
`Interpolator.load_mesh(mesh_obj=...)` reads from a meshio.Mesh (reference interpolator.pyx:255-369,
428-454): `.points`, `.cells[i].type`, `.cells[i].data`, `.cell_data`, `.cell_data_dict`,
`.point_data`.
"""
import numpy as np

_KUHN = ((0, 1, 2, 6), (0, 2, 3, 6), (0, 3, 7, 6), (0, 7, 4, 6), (0, 4, 5, 6), (0, 5, 1, 6))
_WEDGES = ((0, 1, 5, 3, 2, 6), (0, 5, 4, 3, 6, 7))
_TRANS_PYRAMID = (0, 4, 7, 3)
_TRANS_TETS = ((1, 2, 6), (1, 6, 5), (0, 1, 2), (0, 2, 3), (4, 5, 6), (4, 6, 7), (0, 1, 5), (0, 5, 4),
               (3, 2, 6), (3, 6, 7))


class CellBlock:
    def __init__(self, cell_type, data):
        self.type = cell_type
        self.data = np.ascontiguousarray(data)

    def __len__(self):
        return len(self.data)


class SimpleMesh:
    """Container with the attributes of meshio.Mesh that ninpol reads."""

    def __init__(self, points, cells, point_data=None, cell_data=None):
        self.points = np.ascontiguousarray(points)
        self.cells = [c if hasattr(c, "type") else CellBlock(c[0], c[1]) for c in cells]
        self.point_data = {} if point_data is None else point_data
        self.cell_data = {} if cell_data is None else cell_data

    @property
    def cells_dict(self):
        out = {}
        for blk in self.cells:
            out.setdefault(blk.type, []).append(blk.data)
        return {k: np.concatenate(v) for k, v in out.items()}

    @property
    def cell_data_dict(self):
        out = {}
        for name, per_block in self.cell_data.items():
            by_type = {}
            for values, blk in zip(per_block, self.cells):
                by_type.setdefault(blk.type, []).append(np.asarray(values))
            out[name] = {t: np.concatenate(v) for t, v in by_type.items()}
        return out

    @property
    def n_cells(self):
        return sum(len(b) for b in self.cells)


# ------------------------------------------------------------------------------------------------
# lattices
# ------------------------------------------------------------------------------------------------
def _lattice_points(n, perturb, seed, frozen=None):
    """(n+1)^3 lattice nodes of the unit cube; id = i + (n+1)*(j + (n+1)*k).  Interior nodes are
    displaced by U(-perturb*h, perturb*h) per coordinate (PCG64, `seed`)."""
    g = np.arange(n + 1, dtype=np.float64) / n
    k, j, i = np.meshgrid(g, g, g, indexing="ij")
    pts = np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1)
    if perturb > 0.0:
        rng = np.random.default_rng(seed)
        h = 1.0 / n
        d = rng.uniform(-perturb * h, perturb * h, size=pts.shape)
        idx = np.arange(n + 1)
        inner = (idx > 0) & (idx < n)
        kk, jj, ii = np.meshgrid(inner, inner, inner, indexing="ij")
        mask = (ii & jj & kk).ravel()
        if frozen is not None:
            mask &= ~frozen
        pts[mask] += d[mask]
    return pts


def _hull_mask(n):
    idx = np.arange(n + 1)
    edge = (idx == 0) | (idx == n)
    kk, jj, ii = np.meshgrid(edge, edge, edge, indexing="ij")
    return (ii | jj | kk).ravel()


def _cube_corners(n, dtype=np.int64, i_range=None):
    """corner node ids [n_cubes, 8] in meshio hexahedron order; cube id = i + n*(j + n*k)."""
    N1 = n + 1
    ii = np.arange(n, dtype=dtype) if i_range is None else np.arange(i_range[0], i_range[1], dtype=dtype)
    r = np.arange(n, dtype=dtype)
    k, j, i = np.meshgrid(r, r, ii, indexing="ij")
    base = (i + N1 * (j + N1 * k)).ravel()
    off = np.array([0, 1, 1 + N1, N1, N1 * N1, 1 + N1 * N1, 1 + N1 + N1 * N1, N1 + N1 * N1], dtype=dtype)
    return base[:, None] + off[None, :]


def hex_box(n, perturb=0.0, seed=0):
    """Structured n^3 hexahedra on the unit cube (BASELINE config C3 at n = 200)."""
    pts = _lattice_points(n, perturb, seed)
    return SimpleMesh(pts, [CellBlock("hexahedron", _cube_corners(n))])


def kuhn_tet_box(n, perturb=0.25, seed=0):
    """6 n^3 tetrahedra (Kuhn split of every cube along the 0-6 diagonal), interior nodes perturbed
    by 0.25 h (BASELINE configs C1 n = 7, C2 n = 69, C4 n = 203)."""
    pts = _lattice_points(n, perturb, seed)
    cc = _cube_corners(n)
    sel = np.array(_KUHN, dtype=np.int64)          # [6, 4]
    tets = cc[:, sel].reshape(-1, 4)
    return SimpleMesh(pts, [CellBlock("tetra", tets)])


def mixed_box(n, a, b, perturb=0.25, seed=0):
    """Conforming mixed box on an n^3 lattice (BASELINE config C5): x-index slabs
    i < a hexahedra | a <= i < b wedges (2 per cube) | i == b one transition layer (1 pyramid + 10
    tets around a cube-centre node) | i > b Kuhn tets.  One CellBlock per type in the order
    hexahedron, wedge, pyramid, tetra.  Lattice nodes touching the hex / wedge / transition slabs are
    kept unperturbed so quadrilateral faces stay planar."""
    assert 0 < a < b < n - 1
    N1 = n + 1
    # nodes with lattice i-index <= b+1 touch non-tet cells: freeze them
    idx = np.arange(N1)
    kk, jj, ii = np.meshgrid(idx, idx, idx, indexing="ij")
    frozen = (ii <= b + 1).ravel()
    pts = _lattice_points(n, perturb, seed, frozen=frozen)
    hexes = _cube_corners(n, i_range=(0, a))
    wc = _cube_corners(n, i_range=(a, b))
    wedges = wc[:, np.array(_WEDGES, dtype=np.int64)].reshape(-1, 6)
    tc = _cube_corners(n, i_range=(b, b + 1))
    centre_ids = N1 ** 3 + np.arange(len(tc), dtype=np.int64)
    centres = pts[tc].mean(axis=1)
    pyr = np.concatenate([tc[:, np.array(_TRANS_PYRAMID)], centre_ids[:, None]], axis=1)
    tt = tc[:, np.array(_TRANS_TETS, dtype=np.int64)]                      # [nc, 10, 3]
    tt = np.concatenate([tt, np.broadcast_to(centre_ids[:, None, None], (len(tc), 10, 1))], axis=2)
    kc = _cube_corners(n, i_range=(b + 1, n))
    ktets = kc[:, np.array(_KUHN, dtype=np.int64)].reshape(-1, 4)
    tets = np.concatenate([tt.reshape(-1, 4), ktets], axis=0)
    points = np.concatenate([pts, centres], axis=0)
    return SimpleMesh(points, [CellBlock("hexahedron", hexes), CellBlock("wedge", wedges),
                               CellBlock("pyramid", pyr), CellBlock("tetra", tets)])


# ------------------------------------------------------------------------------------------------
# fields
# ------------------------------------------------------------------------------------------------
def _spd_chunk(args):
    ss, count, lam_decades = args
    rng = np.random.default_rng(ss)
    q = rng.standard_normal((4, count))
    q /= np.sqrt((q * q).sum(axis=0))
    w, x, y, z = q
    R = ((1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)),
         (2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)),
         (2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)))
    lam = 10.0 ** rng.uniform(0.0, lam_decades, size=(3, count))
    out = np.empty((count, 9), dtype=np.float64)
    for a in range(3):
        for b in range(a, 3):
            v = lam[0] * R[a][0] * R[b][0] + lam[1] * R[a][1] * R[b][1] + lam[2] * R[a][2] * R[b][2]
            out[:, 3 * a + b] = v
            out[:, 3 * b + a] = v
    return out


def random_spd_permeability(n, seed=2, lam_decades=2.0, chunk=1 << 20, threads=None):
    """Heterogeneous anisotropic SPD tensors K = R diag(lambda) R^T flattened row-major to [n, 9]:
    R a uniformly random rotation (unit quaternion), lambda ~ 10^U(0, lam_decades), so tr K >= 3 and
    the release-build diff_mag exponent eta = (1 - 3/tr K)^2 stays in [0, 1).
    Deterministic for a given (n, seed, chunk): one PCG64 stream per chunk, spawned from `seed`."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    nchunks = max(1, (n + chunk - 1) // chunk)
    seeds = np.random.SeedSequence(seed).spawn(nchunks)
    jobs = [(seeds[i], min(chunk, n - i * chunk), lam_decades) for i in range(nchunks)]
    if threads is None:
        threads = min(16, os.cpu_count() or 1)
    if nchunks == 1 or threads == 1:
        parts = [_spd_chunk(j) for j in jobs]
    else:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            parts = list(ex.map(_spd_chunk, jobs))
    return parts[0] if len(parts) == 1 else np.concatenate(parts, axis=0)


def attach_fields(mesh, variable="u", seed=2, neumann_rate=0.5, hull_nodes=None, permeability=True):
    """Adds the data the three methods read (reference idw.pyx:27, ls.pyx:28, gls.pyx:47-50):
    cell scalar `variable` (x+y+z at the vertex-mean centroid; its values never enter the weights),
    cell tensor `permeability` [n,9], point data `neumann_flag_<variable>` (1 on a random
    `neumann_rate` fraction of hull nodes, else 0) and `neumann_<variable>` (N(0,1) on flagged
    nodes — a dead input of the reference)."""
    rng = np.random.default_rng(seed)
    npts = len(mesh.points)
    if hull_nodes is None:
        p = mesh.points
        lo, hi = p.min(axis=0), p.max(axis=0)
        live = hi > lo                      # a flat axis (2-D meshes with z = 0) is not a hull
        hull_nodes = np.any(((p == lo) | (p == hi)) & live[None, :], axis=1)
    flag = np.zeros(npts, dtype=np.float64)
    pick = rng.random(npts) < neumann_rate
    flag[hull_nodes & pick] = 1.0
    nval = np.where(flag > 0, rng.standard_normal(npts), 0.0)
    mesh.point_data = {f"neumann_flag_{variable}": flag, f"neumann_{variable}": nval}
    cell_data = {variable: []}
    if permeability:
        cell_data["permeability"] = []
    for bi, blk in enumerate(mesh.cells):
        cen = mesh.points[blk.data].mean(axis=1)
        cell_data[variable].append(cen.sum(axis=1))
        if permeability:
            cell_data["permeability"].append(random_spd_permeability(len(blk), seed=seed + 100 + bi))
    mesh.cell_data = cell_data
    return mesh


def quad_plane(n, perturb=0.0, seed=0, triangles=False):
    """2-D unit square: n^2 quadrilaterals, or 2 n^2 triangles (each quad split along its 0-2 diagonal), with
    3-column coordinates (z = 0) as meshio delivers them.  Exercises the reference's dim == 2 branches
    (faces are edges: interpolator.pyx:296-298, grid.pyx:787-806, ls.pyx:79-80,105-106)."""
    g = np.arange(n + 1, dtype=np.float64) / n
    j, i = np.meshgrid(g, g, indexing="ij")
    pts = np.stack([i.ravel(), j.ravel(), np.zeros((n + 1) ** 2)], axis=1)
    if perturb > 0.0:
        rng = np.random.default_rng(seed)
        idx = np.arange(n + 1)
        inner = (idx > 0) & (idx < n)
        jj, ii = np.meshgrid(inner, inner, indexing="ij")
        mask = (ii & jj).ravel()
        d = rng.uniform(-perturb / n, perturb / n, size=(len(pts), 2))
        pts[mask, :2] += d[mask]
    N1 = n + 1
    r = np.arange(n, dtype=np.int64)
    jj, ii = np.meshgrid(r, r, indexing="ij")
    base = (ii + N1 * jj).ravel()
    quads = base[:, None] + np.array([0, 1, 1 + N1, N1], dtype=np.int64)[None, :]
    if not triangles:
        return SimpleMesh(pts, [CellBlock("quad", quads)])
    tris = quads[:, np.array([[0, 1, 2], [0, 2, 3]])].reshape(-1, 3)
    return SimpleMesh(pts, [CellBlock("triangle", tris)])


def scramble(mesh, seed=0):
    """Random renumbering of nodes and (per block) of elements, plus a random rotation of every tet's
    local node order that keeps its orientation: same geometry, but irregular esup / fsup orderings as an
    unstructured mesh generator would produce.  Fields must be attached afterwards."""
    rng = np.random.default_rng(seed)
    npts = len(mesh.points)
    perm = rng.permutation(npts)            # new id of old node i
    inv = np.empty(npts, dtype=np.int64)
    inv[perm] = np.arange(npts)
    pts = mesh.points[inv]
    cells = []
    for blk in mesh.cells:
        data = perm[blk.data]
        data = data[rng.permutation(len(data))]
        if blk.type == "tetra":             # even permutations of (0,1,2,3) keep the orientation
            even = np.array([[0, 1, 2, 3], [1, 2, 0, 3], [2, 0, 1, 3], [0, 3, 1, 2], [1, 0, 3, 2], [3, 2, 1, 0]])
            pick = even[rng.integers(0, len(even), size=len(data))]
            data = np.take_along_axis(data, pick, axis=1)
        cells.append(CellBlock(blk.type, data))
    return SimpleMesh(pts, cells)


def make_case(kind, n, variable="u", seed=0, neumann_rate=0.5, perturb=None, **kw):
    """Convenience: mesh + fields.  kind in {'hex', 'tet', 'mixed'}."""
    if kind == "hex":
        mesh = hex_box(n, perturb=0.0 if perturb is None else perturb, seed=seed)
    elif kind == "tet":
        mesh = kuhn_tet_box(n, perturb=0.25 if perturb is None else perturb, seed=seed)
    elif kind in ("quad2d", "tri2d"):
        mesh = quad_plane(n, perturb=0.0 if perturb is None else perturb, seed=seed, triangles=(kind == "tri2d"))
    elif kind == "mixed":
        a = kw.get("a", max(1, n // 4))
        b = kw.get("b", max(a + 1, n // 2))
        mesh = mixed_box(n, a, b, perturb=0.25 if perturb is None else perturb, seed=seed)
    else:
        raise ValueError(kind)
    if kw.get("scramble"):
        mesh = scramble(mesh, seed=seed + 11)
    return attach_fields(mesh, variable=variable, seed=seed + 2, neumann_rate=neumann_rate)
