"""CUDA path against the committed golden fixtures (outputs of the compiled reference) and, when
oracle/_ref travelled to the box, against the compiled reference itself on a larger mesh."""
import numpy as np
import pytest

from helpers import GLS_TOL, GOLDEN_CASES, GRID_ARRAYS, check_against_golden, load_golden, row_normwise_error

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_cuda_matches_golden_fixtures(name):
    import ninpol_b200
    mesh, d = load_golden(name)
    I = ninpol_b200.Interpolator()
    I.load_mesh(mesh_obj=mesh)
    check_against_golden(I, d)


def test_cuda_reproduces_published_accuracy():
    """accuracy.yaml hexa rows (QUAD, FAN, ALH; n = 4, 8, 16) through the CUDA path."""
    import json
    import os
    import sys
    from helpers import GOLDEN
    sys.path.insert(0, GOLDEN)
    import make_golden
    import ninpol_b200
    table = json.load(open(os.path.join(GOLDEN, "accuracy_hexa.json")))["table"]
    for case in ("QUAD", "FAN", "ALH"):
        for n_i, n in enumerate((4, 8, 16)):
            mesh, sol = make_golden.analytic_mesh(case, n)
            I = ninpol_b200.Interpolator()
            I.load_mesh(mesh_obj=mesh)
            for method in ("gls", "idw", "ls"):
                W, _ = I.interpolate(case, method)
                err = make_golden.interior_l2(W, mesh, sol, case, I.grid.boundary_points)
                pub = table[case][method]["published"][n_i]
                assert abs(err - pub) <= 1e-11 * max(1.0, abs(pub)), (case, method, n, err, pub)


def test_cuda_matches_compiled_reference_when_available():
    import oracle
    ref = oracle.load_reference()
    if ref is None:
        pytest.skip("oracle/_ref not present on this box")
    import ninpol_b200
    from ninpol_b200 import meshgen
    mesh = meshgen.make_case("tet", 40)       # 384,000 cells / 68,921 nodes, 50 % Neumann hull nodes
    R = ref.Interpolator()
    R.load_mesh(mesh_obj=oracle.to_reference_mesh(mesh))
    I = ninpol_b200.Interpolator()
    I.load_mesh(mesh_obj=mesh)
    for a in GRID_ARRAYS:
        assert np.array_equal(np.asarray(getattr(I.grid, a)), np.asarray(getattr(R.grid, a))), a
    for method in ("idw", "ls", "gls"):
        W, nv = I.interpolate("u", method)
        Wr, nvr = R.interpolate("u", method)
        assert np.array_equal(W.indptr, Wr.indptr) and np.array_equal(W.indices, Wr.indices)
        if method == "gls":
            assert row_normwise_error(W, W.indptr, Wr.data) <= GLS_TOL
        else:
            assert np.array_equal(W.data, Wr.data, equal_nan=True)
            assert np.array_equal(nv, np.asarray(nvr))
