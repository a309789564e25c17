"""CPU tests of the oracle (test infrastructure): it must agree with the reference's golden vectors,
with the fixtures generated from the compiled reference, and — where oracle/_ref is present — with the
compiled reference itself on fresh meshes."""
import json
import os
import sys

import numpy as np
import pytest

from helpers import GOLDEN, GOLDEN_CASES, GRID_ARRAYS, GRID_SCALARS, check_against_golden, load_golden

import oracle
from ninpol_b200 import meshgen

sys.path.insert(0, GOLDEN)
import make_golden  # noqa: E402


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_golden_fixtures(name):
    """Fixtures = outputs of the compiled reference; the oracle calls the same scipy DGELS/DGEMV, so
    even GLS is compared bit-for-bit."""
    mesh, d = load_golden(name)
    O = oracle.OracleInterpolator().load_mesh(mesh)
    check_against_golden(O, d, exact_gls=True)


@pytest.mark.parametrize("case", ["QUAD", "FAN", "ALH"])
def test_oracle_reproduces_published_accuracy(case):
    """Known-answer test from the reference's own results (tests/results/yaml/accuracy.yaml, hexa rows)."""
    table = json.load(open(os.path.join(GOLDEN, "accuracy_hexa.json")))["table"][case]
    for method in ("gls", "idw", "ls"):
        for n, published in zip(table[method]["n"][:2], table[method]["published"][:2]):
            mesh, sol = make_golden.analytic_mesh(case, n)
            O = oracle.OracleInterpolator().load_mesh(mesh)
            W, _ = O.interpolate(case, method)
            err = make_golden.interior_l2(W, mesh, sol, case, O.grid.boundary_points)
            assert abs(err - published) <= 5e-13 * max(1.0, abs(published)) + 1e-15, (case, method, n, err, published)


def test_oracle_linear_exactness():
    """LIN rows of accuracy.yaml (:293-438): LS and GLS reproduce a linear field to rounding."""
    mesh, sol = make_golden.analytic_mesh("LIN", 6)
    O = oracle.OracleInterpolator().load_mesh(mesh)
    for method in ("ls", "gls"):
        W, _ = O.interpolate("LIN", method)
        assert make_golden.interior_l2(W, mesh, sol, "LIN", O.grid.boundary_points) < 1e-14


REF = oracle.load_reference()


@pytest.mark.skipif(REF is None, reason="compiled reference (oracle/_ref) not built in this checkout")
@pytest.mark.parametrize("kind,n,kw", [("tet", 6, {}), ("hex", 7, {}), ("mixed", 8, {"a": 2, "b": 4}), ("hex", 5, {"perturb": 0.2}),
                                       ("tet", 1, {}), ("hex", 1, {}), ("quad2d", 6, {}), ("tri2d", 6, {"perturb": 0.2}),
                                       ("tet", 6, {"scramble": True}), ("mixed", 8, {"a": 2, "b": 4, "scramble": True})])
def test_oracle_matches_compiled_reference(kind, n, kw):
    mesh = meshgen.make_case(kind, n, **kw)
    I = REF.Interpolator()
    I.load_mesh(mesh_obj=oracle.to_reference_mesh(mesh))
    O = oracle.OracleInterpolator().load_mesh(mesh)
    for s in GRID_SCALARS:
        assert getattr(I.grid, s) == getattr(O.grid, s), s
    for a in GRID_ARRAYS:
        assert np.array_equal(np.asarray(getattr(I.grid, a)), getattr(O.grid, a)), a
    for method in ("idw", "ls", "gls"):
        W, nv = I.interpolate("u", method)
        Wo, nvo = O.interpolate("u", method)
        assert np.array_equal(W.indptr, Wo.indptr) and np.array_equal(W.indices, Wo.indices)
        assert np.array_equal(W.data, Wo.data, equal_nan=True), method
        assert np.array_equal(np.asarray(nv), nvo)


def test_release_build_diffusion_magnitude():
    """(1 - 3/tr K)^2, independent of det K."""
    K = meshgen.random_spd_permeability(50, seed=7)
    dm = oracle.diffusion_magnitude(K)
    tr = K[:, 0] + K[:, 4] + K[:, 8]
    assert np.array_equal(dm, (1 - 3 / tr) ** 2)


@pytest.mark.skipif(REF is None, reason="compiled reference (oracle/_ref) not built in this checkout")
@pytest.mark.parametrize("kind,n,kw", [("tet", 5, {}), ("mixed", 8, {"a": 2, "b": 4}), ("tri2d", 7, {})])
def test_oracle_edges_match_compiled_reference(kind, n, kw):
    mesh = meshgen.make_case(kind, n, **kw)
    I = REF.Interpolator(build_edges=True)
    I.load_mesh(mesh_obj=oracle.to_reference_mesh(mesh))
    O = oracle.OracleInterpolator().load_mesh(mesh, build_edges=True)
    assert I.grid.n_edges == O.grid.n_edges
    assert np.array_equal(np.asarray(I.grid.inedel), O.grid.inedel)
    assert np.array_equal(np.asarray(I.grid.inpoed), O.grid.inpoed)
