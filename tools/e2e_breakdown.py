"""Wall-clock breakdown of one end-to-end interpolate() (inputs re-uploaded), for bench.py's e2e figure.
usage: python tools/e2e_breakdown.py KIND N METHOD"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ninpol_b200
from ninpol_b200 import meshgen

kind, n, method = sys.argv[1], int(sys.argv[2]), sys.argv[3]
mesh = meshgen.make_case(kind, n)
I = ninpol_b200.Interpolator(pinned_outputs=True, pin_inputs=True)
I.load_mesh(mesh_obj=mesh)
for rep in range(3):
    I.invalidate_inputs()
    I._ctx.synchronize()
    t0 = time.perf_counter()
    I._stage_inputs(method, "u", I.variable_to_index, I._rows["cells"], I._rows["points"])
    I._ctx.synchronize()
    t1 = time.perf_counter()
    nnz = I._ctx.interpolate_count(method)
    I._ctx.synchronize()
    t2 = time.perf_counter()
    g = I.grid
    indptr = I._out("indptr", g.n_points + 1, np.int32)
    indices = I._out("indices", nnz, np.int32)
    data = I._out("data", nnz, np.float64)
    neumann = I._out("neumann", g.n_points, np.float64)
    t3 = time.perf_counter()
    I._ctx.interpolate_fetch(indptr, indices, data, neumann)
    I._ctx.synchronize()
    t4 = time.perf_counter()
    h2d = 8 * g.n_points + (80 * g.n_elems if method == "gls" else 0)
    d2h = indptr.nbytes + indices.nbytes + data.nbytes + neumann.nbytes
    print(f"rep {rep}: stage {1e3 * (t1 - t0):.1f} ms ({h2d / 1e9 / (t1 - t0):.1f} GB/s)  count(K2+K3a) {1e3 * (t2 - t1):.1f} ms  "
          f"alloc {1e3 * (t3 - t2):.1f} ms  fetch(K3b+D2H) {1e3 * (t4 - t3):.1f} ms ({d2h / 1e9 / (t4 - t3):.1f} GB/s)  "
          f"timers {dict((k, round(v, 2)) for k, v in I.last_timings.items())}")
    names = ["h2d_flags", "h2d_field", "k2", "k3_count", "k3_fill", "d2h_csr"]
    print("   ", {k: round(I._ctx.timing_or(k, -1), 2) for k in names})
