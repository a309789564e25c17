"""CPU oracle for the ninpol hot path — TEST INFRASTRUCTURE ONLY.

Nothing in the product package `ninpol_b200` imports this package.  It may be used by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs, as the checker.

Contents
  ninpol_oracle.c   plain-C serial restatement of grid.pyx / idw.pyx / ls.pyx / gls.pyx (cited per
                    function), compiled here with gcc into oracle/_build/.
  OracleInterpolator  Python restatement of Interpolator.load_mesh / interpolate
                    (interpolator.pyx:168-252, 255-369, 372-454, 501-509, 549-629) on top of it.
  build_ref.py      recipe that compiles the unmodified reference into oracle/_ref/ (git-ignored).
  load_reference()  imports that compiled reference (with the meshio shim on sys.path).

Parity status: pinned against the compiled reference (tests/test_oracle.py) and the
fixtures under tests/golden/.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(HERE, "_build")
_LIB = os.path.join(_BUILD, "libninpol_oracle.so")
_SRC = os.path.join(HERE, "ninpol_oracle.c")

MX_PE, MX_FE, MX_PF, N_TYPES = 8, 6, 4, 8

# element tables: the reference's utils/point_ordering.yaml:6-53, restated (meshio ordering)
ELEMENTS = {
    "vertex": dict(element_type=0, number_of_points=1, edges=[], faces=[]),
    "line": dict(element_type=1, number_of_points=2, edges=[[0, 1]], faces=[]),
    "triangle": dict(element_type=2, number_of_points=3, edges=[[0, 1], [1, 2], [2, 0]], faces=[]),
    "quad": dict(element_type=3, number_of_points=4, edges=[[0, 1], [1, 2], [2, 3], [3, 0]], faces=[]),
    "tetra": dict(element_type=4, number_of_points=4,
                  edges=[[0, 1], [1, 2], [2, 0], [0, 3], [1, 3], [2, 3]],
                  faces=[[0, 2, 1], [0, 1, 3], [1, 2, 3], [0, 3, 2]]),
    "hexahedron": dict(element_type=5, number_of_points=8,
                       edges=[[0, 1], [1, 2], [2, 3], [3, 0], [4, 5], [5, 6], [6, 7], [7, 4], [0, 4], [1, 5], [2, 6], [3, 7]],
                       faces=[[0, 3, 2, 1], [4, 5, 6, 7], [0, 1, 5, 4], [1, 2, 6, 5], [2, 3, 7, 6], [3, 0, 4, 7]]),
    "wedge": dict(element_type=6, number_of_points=6,
                  edges=[[0, 1], [1, 2], [2, 0], [3, 4], [4, 5], [5, 3], [0, 3], [1, 4], [2, 5]],
                  faces=[[0, 2, 1], [3, 4, 5], [0, 1, 4, 3], [1, 2, 5, 4], [0, 3, 5, 2]]),
    "pyramid": dict(element_type=7, number_of_points=5,
                    edges=[[0, 1], [1, 2], [2, 3], [3, 0], [0, 4], [1, 4], [2, 4], [3, 4]],
                    faces=[[0, 3, 2, 1], [0, 1, 4], [1, 2, 4], [2, 3, 4], [3, 0, 4]]),
}
TYPES_PER_DIM = {0: ["vertex"], 1: ["line"], 2: ["triangle", "quad"], 3: ["tetra", "hexahedron", "wedge", "pyramid"]}


def build_c_oracle(force=False):
    """gcc-compile ninpol_oracle.c (no FMA contraction, no fast-math)."""
    if (not force) and os.path.exists(_LIB) and os.path.getmtime(_LIB) >= os.path.getmtime(_SRC):
        return _LIB
    os.makedirs(_BUILD, exist_ok=True)
    cmd = ["/usr/bin/gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", _LIB, _SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_c_oracle())
        _lib.orc_build_esup.restype = ctypes.c_longlong
        _lib.orc_build_psup.restype = ctypes.c_longlong
        _lib.orc_build_infael.restype = ctypes.c_longlong
        _lib.orc_build_fsup.restype = ctypes.c_longlong
        _lib.orc_build_esuf.restype = ctypes.c_longlong
        _lib.orc_gls.restype = ctypes.c_int
        _lib.orc_build_inedel.restype = ctypes.c_longlong
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _ll(v):
    return ctypes.c_longlong(int(v))


def _blas_lapack_pointers():
    """Function pointers of the dgels/dgemv the reference binds (gls.pyx:26-27)."""
    import scipy.linalg.cython_blas as cb
    import scipy.linalg.cython_lapack as cl
    get = ctypes.pythonapi.PyCapsule_GetPointer
    get.restype = ctypes.c_void_p
    get.argtypes = [ctypes.py_object, ctypes.c_char_p]
    name = ctypes.pythonapi.PyCapsule_GetName
    name.restype = ctypes.c_char_p
    name.argtypes = [ctypes.py_object]
    out = []
    for mod, fn in ((cl, "dgels"), (cb, "dgemv")):
        cap = mod.__pyx_capi__[fn]
        out.append(ctypes.c_void_p(get(cap, name(cap))))
    return out


def element_tables(dim):
    """npoel, nfael, lnofa, lpofa, nedel, lpoed exactly as process_mesh fills them
    (interpolator.pyx:274-330): -1 everywhere except the entries of the mesh dimension."""
    npoel = -np.ones(N_TYPES, dtype=np.int64)
    nfael = -np.ones(N_TYPES, dtype=np.int64)
    lnofa = -np.ones((N_TYPES, MX_FE), dtype=np.int64)
    lpofa = -np.ones((N_TYPES, MX_FE, MX_PF), dtype=np.int64)
    nedel = -np.ones(N_TYPES, dtype=np.int64)
    lpoed = -np.ones((N_TYPES, 12, 2), dtype=np.int64)
    faces_key = "edges" if dim == 2 else "faces"
    for name, el in ELEMENTS.items():
        t = el["element_type"]
        npoel[t] = el["number_of_points"]
        if name not in TYPES_PER_DIM[dim]:
            continue
        fl = el.get(faces_key, [])
        nfael[t] = len(fl)
        # the reference only copies the tables `if "faces" in elem_type` (always true in the yaml)
        for i, face in enumerate(fl):
            lnofa[t, i] = len(face)
            for j, p in enumerate(face):
                lpofa[t, i, j] = p
        nedel[t] = len(el["edges"])
        for i, e in enumerate(el["edges"]):
            lpoed[t, i, 0], lpoed[t, i, 1] = e
    return npoel, nfael, lnofa, lpofa, nedel, lpoed


def process_mesh(mesh):
    """interpolator.pyx:255-369 — dim, padded int64 connectivity, element types."""
    dim = 1
    for blk in mesh.cells:
        for d, names in TYPES_PER_DIM.items():
            if blk.type in names:
                dim = max(dim, d)
    blocks = [b for b in mesh.cells if b.type in TYPES_PER_DIM[dim]]
    n_elems = sum(len(b.data) for b in blocks)
    conn = -np.ones((n_elems, MX_PE), dtype=np.int64)
    etype = -np.ones(n_elems, dtype=np.int64)
    at = 0
    for b in blocks:
        d = np.asarray(b.data)
        conn[at:at + len(d), :d.shape[1]] = d
        etype[at:at + len(d)] = ELEMENTS[b.type]["element_type"]
        at += len(d)
    return dim, n_elems, len(mesh.points), conn, etype


class OracleGrid:
    """Grid.build + load_point_coords + calculate_centroids + calculate_normal_faces
    (grid.pyx:142-231, 661-809) through the C restatement; attribute names follow grid.pxd:128-187."""

    def __init__(self, dim, n_elems, n_points, conn, etype, coords, build_psup=True, build_edges=False):
        L = lib()
        self.dim, self.n_elems, self.n_points = dim, n_elems, n_points
        self.inpoel = np.ascontiguousarray(conn, dtype=np.int64)
        self.element_types = np.ascontiguousarray(etype, dtype=np.int64)
        self.npoel, self.nfael, self.lnofa, self.lpofa, self.nedel, self.lpoed = element_tables(dim)
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        self.point_coords = coords
        nsum = int(self.npoel[self.element_types].sum())
        self.esup_ptr = np.zeros(n_points + 1, dtype=np.int64)
        self.esup = np.zeros(nsum, dtype=np.int64)
        self.MX_ELEMENTS_PER_POINT = int(L.orc_build_esup(_ll(n_elems), _ll(n_points), _p(self.inpoel), _p(self.element_types),
                                                          _p(self.npoel), _p(self.esup_ptr), _p(self.esup)))
        if build_psup:
            self.psup_ptr = np.zeros(n_points + 1, dtype=np.int64)
            psup = np.zeros(nsum * (MX_PE - 1), dtype=np.int64)
            mx = ctypes.c_longlong(0)
            used = L.orc_build_psup(_ll(n_points), _p(self.inpoel), _p(self.element_types), _p(self.npoel),
                                    _p(self.esup_ptr), _p(self.esup), _p(self.psup_ptr), _p(psup), ctypes.byref(mx))
            self.psup = psup[:used].copy()
            self.MX_POINTS_PER_POINT = int(mx.value)
        self.esuel = np.empty((n_elems, MX_FE), dtype=np.int64)
        L.orc_build_esuel(_ll(n_elems), _p(self.inpoel), _p(self.element_types), _p(self.nfael), _p(self.lnofa),
                          _p(self.lpofa), _p(self.esup_ptr), _p(self.esup), _p(self.esuel))
        self.infael = np.empty((n_elems, MX_FE), dtype=np.int64)
        f2e = np.empty((n_elems * MX_FE, 2), dtype=np.int64)
        self.n_faces = int(L.orc_build_infael(_ll(n_elems), _p(self.element_types), _p(self.nfael), _p(self.esuel),
                                              _p(self.infael), _p(f2e)))
        self.inpofa = np.empty((self.n_faces, MX_PF), dtype=np.int64)
        L.orc_fill_inpofa(_ll(self.n_faces), _p(self.inpoel), _p(self.element_types), _p(self.lnofa), _p(self.lpofa),
                          _p(f2e), _p(self.inpofa))
        del f2e
        self.fsup_ptr = np.zeros(n_points + 1, dtype=np.int64)
        mx = ctypes.c_longlong(0)
        total = L.orc_build_fsup(_ll(self.n_faces), _ll(n_points), _p(self.inpofa), _p(self.fsup_ptr), None, ctypes.byref(mx))
        self.MX_FACES_PER_POINT = int(mx.value)
        self.fsup = np.zeros(total, dtype=np.int64)
        L.orc_build_fsup(_ll(self.n_faces), _ll(n_points), _p(self.inpofa), _p(self.fsup_ptr), _p(self.fsup), ctypes.byref(mx))
        nfsum = int(self.nfael[self.element_types].sum())
        self.esuf_ptr = np.zeros(self.n_faces + 1, dtype=np.int64)
        self.esuf = np.zeros(nfsum, dtype=np.int64)
        self.boundary_faces = np.zeros(self.n_faces, dtype=np.int64)
        self.boundary_points = np.zeros(n_points, dtype=np.int64)
        self.MX_ELEMENTS_PER_FACE = int(L.orc_build_esuf(
            _ll(n_elems), _ll(self.n_faces), _ll(n_points), _p(self.inpoel), _p(self.element_types), _p(self.nfael),
            _p(self.lnofa), _p(self.lpofa), _p(self.infael), _p(self.esuf_ptr), _p(self.esuf), _p(self.inpofa),
            _p(self.boundary_faces), _p(self.boundary_points)))
        self.n_edges = 0
        if build_edges:                                     # Grid.build_inedel, grid.pyx:527-580
            self.inedel = np.empty((n_elems, 12), dtype=np.int64)
            inpoed = np.empty((n_elems * 12, 2), dtype=np.int64)
            self.n_edges = int(L.orc_build_inedel(_ll(n_elems), _p(self.inpoel), _p(self.element_types), _p(self.nedel),
                                                  _p(self.lpoed), _p(self.inedel), _p(inpoed)))
            self.inpoed = inpoed[:self.n_edges].copy()
        self.centroids = np.zeros((n_elems, 3), dtype=np.float64)
        self.faces_centers = np.zeros((self.n_faces, 3), dtype=np.float64)
        L.orc_centroids(_ll(dim), _ll(n_elems), _ll(self.n_faces), _p(self.inpoel), _p(self.element_types), _p(self.npoel),
                        _p(self.inpofa), _p(coords), _p(self.centroids), _p(self.faces_centers))
        self.normal_faces = np.zeros((self.n_faces, 3), dtype=np.float64)
        self.faces_areas = np.zeros(self.n_faces, dtype=np.float64)
        L.orc_normals(_ll(dim), _ll(self.n_faces), _p(self.inpofa), _p(coords), _p(self.normal_faces), _p(self.faces_areas))


def diffusion_magnitude(perm):
    """interpolator.pyx:501-509 in the release build: `1 / 3` is C integer division (cdivision=True),
    so the exponent is 0 and diff_mag = (1 - 3/tr K)^2."""
    Ks = np.reshape(perm, (len(perm), 3, 3))
    det = np.linalg.det(Ks)
    tr = np.trace(Ks, axis1=1, axis2=2)
    return (1 - (3 * (det ** 0) / tr)) ** 2


class OracleInterpolator:
    """Interpolator.load_mesh / interpolate restated (interpolator.pyx:168-252, 549-629)."""

    supported = ("gls", "idw", "ls")

    def __init__(self):
        self.grid = None

    def load_mesh(self, mesh_obj, build_psup=True, build_edges=False):
        dim, n_elems, n_points, conn, etype = process_mesh(mesh_obj)
        self.grid = OracleGrid(dim, n_elems, n_points, conn, etype, np.asarray(mesh_obj.points, dtype=np.float64),
                               build_psup=build_psup, build_edges=build_edges)
        self.cells = {}
        cdd = mesh_obj.cell_data_dict if mesh_obj.cell_data else {}
        for var, by_type in cdd.items():                    # load_cell_data, :428-451
            parts = [np.asarray(v) for t, v in by_type.items() if t in TYPES_PER_DIM[dim]]
            arr = np.concatenate(parts) if parts else np.zeros(0)
            self.cells[var] = arr
            if var == "permeability":
                self.cells["diff_mag"] = diffusion_magnitude(arr)
        self.points = {k: np.asarray(v) for k, v in (mesh_obj.point_data or {}).items()}
        return self

    def weights_dense(self, variable, method):
        """prepare_interpolator (interpolator.pyx:631-670) + the plug-in: dense [n_points, MXE] weights
        and neumann_ws."""
        g, L = self.grid, lib()
        ncol = g.MX_ELEMENTS_PER_POINT
        weights = np.zeros((g.n_points, ncol), dtype=np.float64)
        neumann_ws = np.zeros(g.n_points, dtype=np.float64)
        flag = np.ascontiguousarray(np.asarray(self.points["neumann_flag_" + variable]).astype(np.int64))
        if method == "idw":
            L.orc_idw(_ll(g.dim), _ll(g.n_points), _ll(ncol), _p(g.esup_ptr), _p(g.esup), _p(g.boundary_points), _p(flag),
                      _p(g.point_coords), _p(g.centroids), _p(weights))
        elif method == "ls":
            L.orc_ls(_ll(g.n_points), _ll(ncol), _p(g.esup_ptr), _p(g.esup), _p(g.boundary_points), _p(flag),
                     _p(g.point_coords), _p(g.centroids), _p(weights))
        elif method == "gls":
            perm = np.ascontiguousarray(np.reshape(self.cells["permeability"], (g.n_elems, 9)), dtype=np.float64)
            dm = np.ascontiguousarray(self.cells["diff_mag"], dtype=np.float64)
            nval = np.ascontiguousarray(self.points["neumann_" + variable], dtype=np.float64)
            dgels, dgemv = _blas_lapack_pointers()
            rc = L.orc_gls(_ll(g.n_points), _ll(ncol), _ll(g.MX_FACES_PER_POINT), _p(g.esup_ptr), _p(g.esup), _p(g.fsup_ptr),
                           _p(g.fsup), _p(g.esuf_ptr), _p(g.esuf), _p(g.inpofa), _p(g.boundary_faces), _p(g.boundary_points),
                           _p(flag), _p(nval), _p(g.point_coords), _p(g.centroids), _p(g.faces_centers), _p(g.normal_faces),
                           _p(perm), _p(dm), dgels, dgemv, _p(weights), _p(neumann_ws), _ll(-1), None, None)
            if rc != 0:
                raise MemoryError("oracle gls")
        else:
            raise ValueError(method)
        return weights, neumann_ws

    def gls_system(self, variable, point):
        """Dense GLS system M (m x n) of one node, for conditioning checks in tests."""
        g, L = self.grid, lib()
        ncol = g.MX_ELEMENTS_PER_POINT
        flag = np.ascontiguousarray(np.asarray(self.points["neumann_flag_" + variable]).astype(np.int64))
        perm = np.ascontiguousarray(np.reshape(self.cells["permeability"], (g.n_elems, 9)), dtype=np.float64)
        dm = np.ascontiguousarray(self.cells["diff_mag"], dtype=np.float64)
        nval = np.ascontiguousarray(self.points["neumann_" + variable], dtype=np.float64)
        # run only this node: fake a one-node range by masking all others as Dirichlet boundary
        bp = np.ones(g.n_points, dtype=np.int64)
        fl = np.zeros(g.n_points, dtype=np.int64)
        bp[point] = g.boundary_points[point]
        fl[point] = flag[point]
        weights = np.zeros((g.n_points, ncol))
        nws = np.zeros(g.n_points)
        M = np.zeros((ncol + 4 * g.MX_FACES_PER_POINT) * (3 * ncol + 1))
        mn = np.zeros(2, dtype=np.int64)
        dgels, dgemv = _blas_lapack_pointers()
        L.orc_gls(_ll(g.n_points), _ll(ncol), _ll(g.MX_FACES_PER_POINT), _p(g.esup_ptr), _p(g.esup), _p(g.fsup_ptr),
                  _p(g.fsup), _p(g.esuf_ptr), _p(g.esuf), _p(g.inpofa), _p(g.boundary_faces), _p(bp), _p(fl), _p(nval),
                  _p(g.point_coords), _p(g.centroids), _p(g.faces_centers), _p(g.normal_faces), _p(perm), _p(dm),
                  dgels, dgemv, _p(weights), _p(nws), _ll(point), _p(M), _p(mn))
        m, n = int(mn[0]), int(mn[1])
        return M[:m * n].reshape(m, n), weights[point], nws[point]

    def interpolate(self, variable, method):
        """interpolator.pyx:598-629: COO fill `weights + neumann_ws`, scipy COO->CSR, eliminate_zeros."""
        import scipy.sparse as sp
        g = self.grid
        weights, nws = self.weights_dense(variable, method)
        ptr, esup = g.esup_ptr, g.esup
        cnt = np.diff(ptr)
        rows = np.repeat(np.arange(g.n_points, dtype=np.int64), cnt)
        local = np.arange(len(esup), dtype=np.int64) - np.repeat(ptr[:-1], cnt)
        data = weights[rows, local] + nws[rows]
        W = sp.csr_matrix((data, (rows, esup.copy())), shape=(g.n_points, g.n_elems))
        W.eliminate_zeros()
        return W, nws


def sample_rows(grid, method, nodes, flags, permeability=None, diff_mag=None, neumann_val=None):
    """The reference's per-node loops (idw.pyx:57-84, ls.pyx:56-135, gls.pyx:161-219) over a SAMPLE of target
    nodes.  `grid` is anything with the reference's Grid attributes in the reference's layouts (an
    OracleGrid, or the arrays exported by the CUDA path — the at-size parity tests use that: the connectivity
    arrays are checked separately, and a node's weights depend only on its own esup / fsup rows).
    Returns dense weights [len(nodes), MX_ELEMENTS_PER_POINT] and neumann_ws [len(nodes)]."""
    L = lib()
    i8 = lambda a: np.ascontiguousarray(a, dtype=np.int64)
    f8 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    nodes = i8(nodes)
    ncol = int(grid.MX_ELEMENTS_PER_POINT)
    weights = np.zeros((len(nodes), ncol), dtype=np.float64)
    nws = np.zeros(len(nodes), dtype=np.float64)
    esup_ptr, esup, bp, fl = i8(grid.esup_ptr), i8(grid.esup), i8(grid.boundary_points), i8(flags)
    coords, cent = f8(grid.point_coords), f8(grid.centroids)
    if method == "idw":
        L.orc_idw_nodes(_ll(grid.dim), _ll(len(nodes)), _p(nodes), _ll(ncol), _p(esup_ptr), _p(esup), _p(bp), _p(fl),
                        _p(coords), _p(cent), _p(weights))
    elif method == "ls":
        L.orc_ls_nodes(_ll(len(nodes)), _p(nodes), _ll(ncol), _p(esup_ptr), _p(esup), _p(bp), _p(fl), _p(coords), _p(cent),
                       _p(weights))
    elif method == "gls":
        L.orc_gls_nodes.restype = ctypes.c_int
        fsup_ptr, fsup, esuf_ptr, esuf = i8(grid.fsup_ptr), i8(grid.fsup), i8(grid.esuf_ptr), i8(grid.esuf)
        inpofa, bf = i8(grid.inpofa), i8(grid.boundary_faces)
        fcent, fnorm = f8(grid.faces_centers), f8(grid.normal_faces)
        perm = f8(np.reshape(permeability, (-1, 9)))
        dm = f8(diff_mag)
        nval = f8(neumann_val) if neumann_val is not None else np.zeros(grid.n_points)
        dgels, dgemv = _blas_lapack_pointers()
        rc = L.orc_gls_nodes(_ll(len(nodes)), _p(nodes), _ll(ncol), _ll(grid.MX_FACES_PER_POINT), _p(esup_ptr), _p(esup),
                             _p(fsup_ptr), _p(fsup), _p(esuf_ptr), _p(esuf), _p(inpofa), _p(bf), _p(bp), _p(fl), _p(nval),
                             _p(coords), _p(cent), _p(fcent), _p(fnorm), _p(perm), _p(dm), dgels, dgemv, _p(weights), _p(nws),
                             _ll(-1), None, None)
        if rc != 0:
            raise MemoryError("oracle gls")
    else:
        raise ValueError(method)
    return weights, nws


def gls_system_of(grid, point, flags, permeability, diff_mag, neumann_val=None):
    """Dense GLS system [A | c] (m x (3E+1)) of ONE node exactly as build_ls_matrices / set_neumann_rows fill it
    (gls.pyx:252-416), from any grid-like object in the reference's layouts."""
    L = lib()
    i8 = lambda a: np.ascontiguousarray(a, dtype=np.int64)
    f8 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    ncol, mxf = int(grid.MX_ELEMENTS_PER_POINT), int(grid.MX_FACES_PER_POINT)
    node = np.array([int(point)], dtype=np.int64)
    w = np.zeros((1, ncol))
    nw = np.zeros(1)
    M = np.zeros((ncol + 4 * mxf) * (3 * ncol + 1))
    mn = np.zeros(2, dtype=np.int64)
    nval = f8(neumann_val) if neumann_val is not None else np.zeros(grid.n_points)
    dgels, dgemv = _blas_lapack_pointers()
    L.orc_gls_nodes.restype = ctypes.c_int
    # the big arrays are passed as they are when already in the reference's dtypes (no copy per call)
    L.orc_gls_nodes(_ll(1), _p(node), _ll(ncol), _ll(mxf), _p(i8(grid.esup_ptr)), _p(i8(grid.esup)), _p(i8(grid.fsup_ptr)),
                    _p(i8(grid.fsup)), _p(i8(grid.esuf_ptr)), _p(i8(grid.esuf)), _p(i8(grid.inpofa)), _p(i8(grid.boundary_faces)),
                    _p(i8(grid.boundary_points)), _p(i8(flags)), _p(nval), _p(f8(grid.point_coords)), _p(f8(grid.centroids)),
                    _p(f8(grid.faces_centers)), _p(f8(grid.normal_faces)), _p(f8(np.reshape(permeability, (-1, 9)))),
                    _p(f8(diff_mag)), dgels, dgemv, _p(w), _p(nw), _ll(int(point)), _p(M), _p(mn))
    m, n = int(mn[0]), int(mn[1])
    return M[:m * n].reshape(m, n), w[0], float(nw[0])


def gls_exact_row(M, n_elem, is_neumann):
    """The CSR row the reference's algebra defines for the system M = [A | c], evaluated in extended precision:
    weights_i = r_i / sum_j r_j with r = c - A argmin|A g - c| (the last row of DGELS's X), plus the
    neumann value (weight of the last element, Q3) added to every entry (Q4).  float64 QR as the preconditioner of an
    iterative refinement whose residuals are accumulated in x87 long double (64-bit mantissa): converges to the exact
    least-squares solution of the float64 system as long as cond(A) * 2^-53 << 1.  Used to tell which of two float64
    answers that differ by more than the parity bar is the accurate one."""
    A64, c64 = M[:, :-1], M[:, -1]
    A, c = A64.astype(np.longdouble), c64.astype(np.longdouble)
    Q, R = np.linalg.qr(A64)
    g = np.zeros(A.shape[1], dtype=np.longdouble)
    for _ in range(8):
        r = c - A @ g
        g = g + np.linalg.solve(R, Q.T @ r.astype(np.float64)).astype(np.longdouble)
    r = (c - A @ g)[:n_elem]
    w = r / r.sum()
    nv = w[n_elem - 1] if is_neumann else np.longdouble(0)
    return (w + nv).astype(np.float64), float(nv)


def sample_csr_rows(grid, nodes, weights, nws):
    """interpolator.pyx:598-624 for the sampled rows: data = weights + neumann_ws in esup order, exact zeros
    dropped.  Returns a list of (indices int64, data float64) per sampled node."""
    esup_ptr, esup = np.asarray(grid.esup_ptr), np.asarray(grid.esup)
    out = []
    for i, p in enumerate(np.asarray(nodes)):
        b, e = int(esup_ptr[p]), int(esup_ptr[p + 1])
        d = weights[i, :e - b] + nws[i]
        keep = d != 0.0
        out.append((np.asarray(esup[b:e])[keep].astype(np.int64), d[keep]))
    return out


def load_reference():
    """Import the compiled reference from oracle/_ref (None if it has not been built)."""
    ref = os.path.join(HERE, "_ref")
    if not os.path.isdir(os.path.join(ref, "ninpol")):
        return None
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    for p in (os.path.join(HERE, "shim"), ref):
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        import ninpol  # noqa
        return ninpol
    except Exception:
        return None


def to_reference_mesh(mesh):
    """Wrap a duck-typed mesh into the shim's meshio.Mesh for the compiled reference."""
    sys.path.insert(0, os.path.join(HERE, "shim")) if os.path.join(HERE, "shim") not in sys.path else None
    import meshio
    return meshio.Mesh(mesh.points, [(c.type, c.data) for c in mesh.cells], mesh.point_data, mesh.cell_data)
