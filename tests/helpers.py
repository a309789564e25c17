import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GRID_ARRAYS = ("esup", "esup_ptr", "psup", "psup_ptr", "esuel", "infael", "inpofa", "fsup", "fsup_ptr", "esuf", "esuf_ptr",
               "boundary_faces", "boundary_points", "centroids", "faces_centers", "normal_faces", "faces_areas")
GRID_SCALARS = ("n_elems", "n_points", "n_faces", "MX_ELEMENTS_PER_POINT", "MX_POINTS_PER_POINT", "MX_ELEMENTS_PER_FACE",
                "MX_FACES_PER_POINT")
GOLDEN_CASES = ("tet_n4", "hex_n4", "hex_n3_perturbed", "mixed_n6")
GLS_TOL = 1e-12   # BASELINE.json north_star: <= 1e-12 relative weight error in FP64


def load_golden(name):
    sys.path.insert(0, GOLDEN)
    import make_golden
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    return make_golden.arrays_to_mesh(d), d


def row_normwise_error(W, indptr, data_ref):
    """max |w - w_ref| / max_row |w_ref| (SURVEY.md 8d parity metric); structures must be equal."""
    if len(data_ref) == 0:
        return 0.0
    rows = np.repeat(np.arange(len(indptr) - 1), np.diff(indptr))
    scale = np.zeros(len(indptr) - 1)
    np.maximum.at(scale, rows, np.abs(np.nan_to_num(data_ref)))
    scale[scale == 0] = 1.0
    return float(np.nanmax(np.abs(W.data - data_ref) / scale[rows]))


def check_against_golden(I, d, gls_tol=GLS_TOL, exact_gls=False):
    """I: anything with .grid and .interpolate (oracle, reference or the CUDA Interpolator)."""
    g = I.grid
    for s, v in zip(GRID_SCALARS, d["grid_scalars"]):
        assert getattr(g, s) == int(v), s
    for a in GRID_ARRAYS:
        got, want = np.asarray(getattr(g, a)), d["grid_" + a]
        assert got.shape == want.shape and got.dtype == want.dtype, a
        assert np.array_equal(got, want), a
    for method in ("idw", "ls", "gls"):
        W, nv = I.interpolate("u", method)
        assert np.array_equal(W.indptr, d[method + "_indptr"]), method
        assert np.array_equal(W.indices, d[method + "_indices"]), method
        if method == "gls" and not exact_gls:
            assert np.array_equal(np.isnan(W.data), np.isnan(d["gls_data"]))
            assert row_normwise_error(W, W.indptr, d["gls_data"]) <= gls_tol
            scale = max(1.0, float(np.max(np.abs(d["gls_neumann"]))))
            assert np.max(np.abs(np.asarray(nv) - d["gls_neumann"])) <= gls_tol * scale
        else:
            assert np.array_equal(W.data, d[method + "_data"], equal_nan=True), method
            assert np.array_equal(np.asarray(nv), d[method + "_neumann"]), method
