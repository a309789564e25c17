"""Where the end-to-end time of the streamed interpolate() goes (wall clock per phase + device timers).
usage: python tools/stream_probe.py KIND N [CHUNKS]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
import ninpol_b200
from ninpol_b200 import meshgen
kind, n = sys.argv[1], int(sys.argv[2])
chunks = int(sys.argv[3]) if len(sys.argv) > 3 else 8
mesh = meshgen.make_case(kind, n)
I = ninpol_b200.Interpolator(pinned_outputs=True, pin_inputs=True, stream_chunks=chunks)
I.load_mesh(mesh_obj=mesh)
g = I.grid
for rep in range(5):
    if rep < 3:
        I.invalidate_inputs()
    I._ctx.synchronize()
    t0 = time.perf_counter()
    fields = I._stage_inputs("gls", "u", I.variable_to_index, I._rows["cells"], I._rows["points"], defer_fields=True)
    I._ctx.synchronize()
    t1 = time.perf_counter()
    indptr, indices, data, neumann = I._run_streamed("gls", fields)
    t2 = time.perf_counter()
    W = sp.csr_matrix((data, indices, indptr), shape=(g.n_points, g.n_elems), copy=False)
    t3 = time.perf_counter()
    t = I._ctx.timing_or
    names = ["k2", "k2_gls_c1", "k2_gls_c2", "k2_gls_c4", "k2_gls_dense", "k3_fill"]
    print("   last chunk:", {k: round(t(k), 2) for k in names})
    print(f"rep {rep} ({'upload' if rep < 3 else 'resident'}): stage(flags) {1e3*(t1-t0):.1f} ms | streamed call {1e3*(t2-t1):.1f} ms (device timer {t('streamed'):.1f}, last chunk k2 {t('k2_main'):.1f}) | csr wrap {1e3*(t3-t2):.1f} ms")
