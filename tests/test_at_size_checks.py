"""The at-size checker (tests/at_size_checks.py) must accept the oracle's own grid and reject corrupted
arrays — it is the proof of the 50 M-cell connectivity on the GPU box, so it is tested here on the CPU."""
import copy

import numpy as np
import pytest

import oracle
from at_size_checks import check_connectivity_and_geometry
from ninpol_b200 import meshgen


@pytest.mark.parametrize("kind,n,kw", [("tet", 6, {}), ("mixed", 8, {"a": 2, "b": 4}), ("hex", 5, {})])
def test_checker_accepts_the_oracle_grid(kind, n, kw):
    mesh = meshgen.make_case(kind, n, **kw)
    O = oracle.OracleInterpolator().load_mesh(mesh)
    check_connectivity_and_geometry(O.grid, mesh, np.random.default_rng(0), oracle)


@pytest.mark.parametrize("what", ["esup_swap", "fsup_entry", "face_order", "esuel", "bface", "normal"])
def test_checker_rejects_corruption(what):
    mesh = meshgen.make_case("tet", 5)
    g = copy.copy(oracle.OracleInterpolator().load_mesh(mesh).grid)
    if what == "esup_swap":       # a row no longer ascending
        g.esup = g.esup.copy()
        b = int(g.esup_ptr[40])
        g.esup[b], g.esup[b + 1] = g.esup[b + 1], g.esup[b]
    elif what == "fsup_entry":    # a face that does not contain the node
        g.fsup = g.fsup.copy()
        b = int(g.fsup_ptr[17])
        g.fsup[b + 1] = g.fsup[b] + 1 if g.fsup[b] + 1 != g.fsup[b + 1] else g.fsup[b] + 2
    elif what == "face_order":    # two face ids exchanged: numbering no longer first-encounter
        g.infael = g.infael.copy()
        m = g.infael.copy()
        g.infael[m == 3], g.infael[m == 4] = 4, 3
    elif what == "esuel":
        g.esuel = g.esuel.copy()
        e = int(np.nonzero((g.esuel[:, 0] >= 0))[0][5])
        g.esuel[e, 0] = g.esuel[e, 1] if g.esuel[e, 1] >= 0 else g.esuel[e, 2]
    elif what == "bface":
        g.boundary_faces = g.boundary_faces.copy()
        g.boundary_faces[0] ^= 1
    elif what == "normal":
        g.normal_faces = g.normal_faces.copy()
        g.normal_faces[:, 0] = np.nextafter(g.normal_faces[:, 0], 2.0)
    with pytest.raises(AssertionError):
        check_connectivity_and_geometry(g, mesh, np.random.default_rng(0), oracle)
