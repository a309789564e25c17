import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ninpol_b200
from ninpol_b200 import meshgen
mesh = meshgen.make_case("tet", int(sys.argv[1]) if len(sys.argv) > 1 else 100)
I = ninpol_b200.Interpolator(pinned_outputs=True, pin_inputs=True)
I.load_mesh(mesh_obj=mesh)
g = I.grid
v2i = I.variable_to_index
cells, points = I._rows["cells"], I._rows["points"]
def T(label, f):
    I._ctx.synchronize(); t0 = time.perf_counter(); r = f(); I._ctx.synchronize()
    print(f"   {label}: {1e3 * (time.perf_counter() - t0):.2f} ms"); return r
for rep in range(3):
    print("rep", rep)
    flags = T("flags astype", lambda: np.asarray(points[v2i["points"]["neumann_flag_u"]])[:g.n_points].astype(np.int64))
    T("set_point_flags", lambda: I._ctx.set_point_flags(flags))
    perm = T("perm view", lambda: np.ascontiguousarray(np.asarray(cells[v2i["cells"]["permeability"]])[:g.n_elems * 9], dtype=np.float64))
    dm = T("dm view", lambda: np.ascontiguousarray(np.asarray(cells[v2i["cells"]["diff_mag"]])[:g.n_elems], dtype=np.float64))
    T("pin perm", lambda: I._maybe_pin("permeability", perm))
    T("set perm", lambda: I._ctx.set_cell_field("permeability", perm))
    T("pin dm", lambda: I._maybe_pin("diff_mag", dm))
    T("set dm", lambda: I._ctx.set_cell_field("diff_mag", dm))
    nnz = T("count", lambda: I._ctx.interpolate_count("gls"))
    T("run", lambda: I._run("gls"))
