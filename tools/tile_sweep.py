"""Device time of the IDW / LS step for the variants of the plain tile kernels (NPB_TILE_VARIANT: resident CTAs x load
batch) and every shape of the pipelined (TMA bulk + cp.async) variant (NPB_TILE_PIPE), on the BASELINE meshes.
usage: python tools/tile_sweep.py [tet203 hex200 ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import ninpol_b200

SETTINGS = {"idw": [("plain v0", {"NPB_TILE_VARIANT": "00"}), ("plain v1", {"NPB_TILE_VARIANT": "10"}), ("plain v2", {"NPB_TILE_VARIANT": "20"}),
                    ("plain v3", {"NPB_TILE_VARIANT": "30"}), ("plain v4", {"NPB_TILE_VARIANT": "40"}),
                    ("pipe A", {"NPB_TILE_PIPE": "A"}), ("pipe B", {"NPB_TILE_PIPE": "B"}), ("pipe C", {"NPB_TILE_PIPE": "C"}), ("pipe D", {"NPB_TILE_PIPE": "D"})],
            "ls": [("plain v0", {"NPB_TILE_VARIANT": "00"}), ("plain v1", {"NPB_TILE_VARIANT": "01"}), ("plain v2", {"NPB_TILE_VARIANT": "02"}),
                   ("plain v3", {"NPB_TILE_VARIANT": "03"}),
                   ("pipe A", {"NPB_TILE_PIPE": "A"}), ("pipe B", {"NPB_TILE_PIPE": "B"}), ("pipe C", {"NPB_TILE_PIPE": "C"}), ("pipe D", {"NPB_TILE_PIPE": "D"})]}

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
        for name, env in SETTINGS[method]:
            for k in ("NPB_TILE_VARIANT", "NPB_TILE_PIPE"):
                os.environ.pop(k, None)
            os.environ.update(env)
            ts, ks = [], []
            for it in range(6):
                ctx.timer_start()
                nnz, fb = ctx.interpolate_run(method, 1)
                ms = ctx.timer_stop()
                if it >= 2:
                    ts.append(ms)
                    ks.append(ctx.timing_or("k2_main"))
            k = float(np.median(ks))
            print(f"{wl} {method} {name}: step {np.median(ts):.3f} ms kernel {k:.3f} ms -> {nbytes / (k * 1e-3) / 1e9 / 6541.1:.3f} of HBM roofline"
                  f"{' (fell back)' if fb else ''}", flush=True)
    for k in ("NPB_TILE_VARIANT", "NPB_TILE_PIPE"):
        os.environ.pop(k, None)
    del I
