"""Per-phase wall clock of one multi-GPU interpolate() on every rank (torchrun), for gather in host/root.
usage: torchrun ... tools/mgpu_probe.py KIND N"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ninpol_b200
from ninpol_b200 import dist, meshgen
kind, n = sys.argv[1], int(sys.argv[2])
comm = dist.init_from_env()
mesh = meshgen.make_case(kind, n)
I = ninpol_b200.Interpolator(comm=comm, pinned_outputs=True, pin_inputs=True, gather="host")
I.load_mesh(mesh_obj=mesh)
ctx = I._ctx
for gather in ("host", "root", "host"):
    I.set_gather(gather)
    for rep in range(4):
        I.invalidate_inputs()
        ctx.comm_barrier()
        t = [time.perf_counter()]
        I._stage_inputs("gls", "u", I.variable_to_index, I._rows["cells"], I._rows["points"]); ctx.synchronize(); t.append(time.perf_counter())
        if gather == "host":
            so = I._shared_outputs(); t.append(time.perf_counter())
            ctx.comm_barrier(); t.append(time.perf_counter())
            nnz = ctx.interpolate_count("gls"); ctx.synchronize(); t.append(time.perf_counter())
            ctx.interpolate_fetch(so.indptr, so.indices, so.data, so.neumann); t.append(time.perf_counter())
            ctx.comm_barrier(); t.append(time.perf_counter())
            names = ["stage", "shared()", "barrier1", "count", "fetch", "barrier2"]
        else:
            out = I._run("gls"); t.append(time.perf_counter())
            names = ["stage", "_run"]
        if rep >= 2:
            print(f"rank {comm.rank} {gather} rep {rep}: " + "  ".join(f"{nm} {1e3 * (b - a):.2f}" for nm, a, b in zip(names, t, t[1:])) +
                  f"  total {1e3 * (t[-1] - t[0]):.2f} ms; timers k2 {ctx.timing_or('k2'):.2f} d2h {ctx.timing_or('d2h_csr'):.2f}", flush=True)
