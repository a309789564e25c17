"""A/B timing of the GLS multifrontal kernel's launch variants on one mesh (device time of the interior-node class).
usage: python tools/gls_variant_probe.py [n]   (Kuhn tets, default n = 100: 1.03M nodes)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ninpol_b200
from ninpol_b200 import meshgen

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
mesh = meshgen.make_case("tet", n)
I = ninpol_b200.Interpolator()
I.load_mesh(mesh_obj=mesh)
W, _ = I.interpolate("u", "gls")
ref = W.data.copy()
del W
ctx = I._ctx
SETTINGS = [("default (12 CTAs/SM, 168 regs, front 1664)", {}),
            ("16 CTAs/SM, 128 regs, front 1280", {"NPB_GLS_VARIANT": "16", "NPB_GLS_FCAP": "4:1280"}),
            ("16 CTAs/SM, 128 regs, front 1152", {"NPB_GLS_VARIANT": "16", "NPB_GLS_FCAP": "4:1152"}),
            ("12 CTAs/SM, front 1280", {"NPB_GLS_FCAP": "4:1280"}),
            ("12 CTAs/SM, front 2048 (11 fit)", {"NPB_GLS_FCAP": "4:2048"}),
            ("10 CTAs/SM", {"NPB_GLS_CTAS_PER_SM": "10"}),
            ("8 CTAs/SM", {"NPB_GLS_CTAS_PER_SM": "8"})]
for name, env in SETTINGS:
    for k in ("NPB_GLS_VARIANT", "NPB_GLS_FCAP", "NPB_GLS_CTAS_PER_SM"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ts = []
    for it in range(4):
        ctx.timer_start()
        nnz, fb = ctx.interpolate_run("gls", 1)
        ms = ctx.timer_stop()
        if it >= 1:
            ts.append((ms, ctx.timing_or("k2_main"), ctx.timing_or("gls_dense_nodes")))
    ms, k, dn = np.median(np.array(ts), axis=0)
    print(f"{name}: step {ms:.2f} ms, class-4 kernel {k:.2f} ms, {len(mesh.points) / ms / 1e3:.2f} M nodes/s, dense-path nodes {int(dn)}", flush=True)
for k in ("NPB_GLS_VARIANT", "NPB_GLS_FCAP", "NPB_GLS_CTAS_PER_SM"):
    os.environ.pop(k, None)
