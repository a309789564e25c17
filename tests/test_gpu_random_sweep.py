"""Seeded sweep over mesh families, sizes, numberings, perturbations and Neumann rates: every case compares
all three methods with the oracle (bit-exact IDW / LS and CSR structure, GLS <= 1e-12 row-normwise).
Covers the star shapes the fixed cases do not: pyramid apexes next to leaf fronts, Neumann nodes with one,
two or three boundary faces per element, scrambled numberings that change the elimination order."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GLS_TOL = 1e-12


def _cases():
    rng = np.random.default_rng(20261018)
    out = []
    for i in range(28):
        kind = ("tet", "hex", "mixed")[i % 3]
        n = int(rng.integers(2, 9)) if kind != "mixed" else int(rng.integers(5, 10))
        kw = {"seed": int(rng.integers(0, 1000)), "neumann_rate": float(rng.choice([0.0, 0.3, 0.5, 1.0])),
              "perturb": float(rng.choice([0.0, 0.1, 0.25, 0.3])), "scramble": bool(rng.integers(0, 2))}
        if kind == "mixed":
            a = int(rng.integers(1, n - 2))
            kw.update(a=a, b=int(rng.integers(a + 1, n - 1)))
        out.append((kind, n, kw))
    return out


@pytest.mark.parametrize("kind,n,kw", _cases())
def test_random_case_matches_oracle(kind, n, kw):
    import ninpol_b200
    import oracle
    from ninpol_b200 import meshgen
    mesh = meshgen.make_case(kind, n, **kw)
    I = ninpol_b200.Interpolator()
    I.load_mesh(mesh_obj=mesh)
    O = oracle.OracleInterpolator().load_mesh(mesh)
    for method in ("idw", "ls", "gls"):
        W, nv = I.interpolate("u", method)
        Wo, nvo = O.interpolate("u", method)
        assert np.array_equal(W.indptr, Wo.indptr) and np.array_equal(W.indices, Wo.indices), method
        if method == "gls":
            rows = np.repeat(np.arange(W.shape[0]), np.diff(W.indptr))
            scale = np.zeros(W.shape[0])
            np.maximum.at(scale, rows, np.abs(Wo.data))
            scale[scale == 0] = 1.0
            err = float(np.max(np.abs(W.data - Wo.data) / scale[rows])) if W.nnz else 0.0
            assert err <= GLS_TOL, (method, err)
            assert np.allclose(nv, nvo, rtol=0, atol=GLS_TOL * max(1.0, float(np.abs(nvo).max())))
        else:
            assert np.array_equal(W.data, Wo.data, equal_nan=True), method
            assert np.array_equal(nv, nvo)
