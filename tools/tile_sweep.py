"""Device time of the IDW / LS step for the plain tile kernels and every shape of the pipelined (TMA bulk + cp.async)
variant, on the BASELINE meshes.  usage: python tools/tile_sweep.py [tet203 hex200 ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import ninpol_b200

for wl in (sys.argv[1:] or ["tet203", "hex200"]):
    kind, n, desc = bench.WORKLOADS[wl]
    mesh = bench.make_mesh(kind, n, 0.0 if kind == "hex" else 0.5)
    I = ninpol_b200.Interpolator()
    I.load_mesh(mesh_obj=mesh)
    ctx = I._ctx
    for method in ("idw", "ls"):
        W, _ = I.interpolate("u", method)
        del W
        nbytes, _f, _p = bench.algorithmic_model(I, method)
        for shape in ("0", "A", "B", "C", "D"):
            os.environ["NPB_TILE_PIPE"] = shape
            ts, ks = [], []
            for it in range(6):
                ctx.timer_start()
                nnz, fb = ctx.interpolate_run(method, 1)
                ms = ctx.timer_stop()
                if it >= 2:
                    ts.append(ms)
                    ks.append(ctx.timing_or("k2_main"))
            k = float(np.median(ks))
            print(f"{wl} {method} shape {shape}: step {np.median(ts):.3f} ms kernel {k:.3f} ms -> {nbytes / (k * 1e-3) / 1e9 / 6541.1:.3f} of HBM roofline"
                  f"{' (fell back)' if fb else ''}", flush=True)
    os.environ["NPB_TILE_PIPE"] = "0"
    del I
