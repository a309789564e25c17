"""Where does the GLS difference against the reference come from?  For a few meshes and the three kernel variants
(default multifrontal, general fronts only, dense Householder) prints the row-normwise difference against the oracle
(max, p99.9, rows above 1e-12) and, for the worst rows, the distance of BOTH answers from the extended-precision
solution of the same float64 system (oracle.gls_exact_row).  usage: python tools/gls_accuracy_probe.py [big]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ninpol_b200
import oracle
from ninpol_b200 import meshgen

CASES = [("tri2d", 6, {"perturb": 0.2}), ("quad2d", 9, {"perturb": 0.2}), ("tri2d", 12, {}), ("hex", 24, {}), ("tet", 24, {}),
         ("mixed", 16, {"a": 4, "b": 8})]
if "big" in sys.argv:
    CASES += [("hex", 64, {}), ("tet", 48, {})]
VARIANTS = [("default", {}), ("no_leaf", {"NPB_GLS_NO_LEAF": "1"}), ("dense", {"NPB_FORCE_GLS_DENSE": "1"})]

for kind, n, kw in CASES:
    mesh = meshgen.make_case(kind, n, **kw)
    O = oracle.OracleInterpolator().load_mesh(mesh, build_psup=False)
    g = O.grid
    Wo, nvo = O.interpolate("u", "gls")
    flags = np.asarray(O.points["neumann_flag_u"]).astype(np.int64)
    perm, dm, nval = O.cells["permeability"], O.cells["diff_mag"], O.points["neumann_u"]
    rows = np.repeat(np.arange(g.n_points), np.diff(Wo.indptr))
    scale = np.zeros(g.n_points)
    np.maximum.at(scale, rows, np.abs(Wo.data))
    scale[scale == 0] = 1.0
    I = ninpol_b200.Interpolator()
    I.load_mesh(mesh_obj=mesh)
    for name, env in VARIANTS:
        for k in ("NPB_GLS_NO_LEAF", "NPB_FORCE_GLS_DENSE"):
            os.environ.pop(k, None)
        os.environ.update(env)
        I.invalidate_inputs()
        W, nv = I.interpolate("u", "gls")
        same = np.array_equal(W.indptr, Wo.indptr) and np.array_equal(W.indices, Wo.indices)
        if not same:
            print(f"{kind}{n} {name}: STRUCTURE DIFFERS nnz {W.nnz} vs {Wo.nnz}")
            continue
        rel = np.abs(W.data - Wo.data) / scale[rows]
        row_err = np.zeros(g.n_points)
        np.maximum.at(row_err, rows, rel)
        worst = np.argsort(row_err)[::-1][:6]
        line = []
        for p in worst:
            p = int(p)
            if row_err[p] == 0:
                continue
            M, _w, _n = oracle.gls_system_of(g, p, flags, perm, dm, nval)
            E = int(g.esup_ptr[p + 1] - g.esup_ptr[p])
            if M.shape[0] == 0:
                continue
            ex, _ = oracle.gls_exact_row(M, E, bool(flags[p]) and bool(g.boundary_points[p]))
            a, b = Wo.indptr[p], Wo.indptr[p + 1]
            sc = np.max(np.abs(ex))
            sv = np.linalg.svd(M[:, :-1], compute_uv=False)
            line.append(f"p{p}{'b' if g.boundary_points[p] else 'i'} E{E} diff {row_err[p]:.1e} ours-exact {np.max(np.abs(W.data[a:b] - ex)) / sc:.1e} "
                        f"ref-exact {np.max(np.abs(Wo.data[a:b] - ex)) / sc:.1e} cond {sv[0] / sv[-1]:.1e}")
        print(f"{kind}{n} {name}: max {row_err.max():.2e} p99.9 {np.quantile(row_err, 0.999):.2e} rows>1e-12 {int((row_err > 1e-12).sum())}/{g.n_points}"
              f" neumann {np.max(np.abs(nv - nvo)):.1e}")
        for l in line[:4]:
            print("     ", l)
