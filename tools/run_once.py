"""One load_mesh + interpolate pass on a synthetic mesh (profiling harness for ncu / compute-sanitizer).
usage: python tools/run_once.py KIND N METHOD [REPEAT]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ninpol_b200
from ninpol_b200 import meshgen

kind, n, method = sys.argv[1], int(sys.argv[2]), sys.argv[3]
rep = int(sys.argv[4]) if len(sys.argv) > 4 else 1
mesh = meshgen.make_case(kind, n)
I = ninpol_b200.Interpolator(stream_chunks=int(os.environ.get("RUN_ONCE_CHUNKS", "8")))   # 1: one launch per kernel over all nodes
I.load_mesh(mesh_obj=mesh)
for meth in method.split(","):
    for _ in range(rep):
        W, nv = I.interpolate("u", meth)
    print(kind, n, meth, "nnz", W.nnz, "sum", repr(float(W.data.sum())), "neumann", repr(float(nv.sum())), {k: round(v, 3) for k, v in I.last_timings.items()})
names = ["k2", "k2_main", "k2_gls_c1", "k2_gls_c2", "k2_gls_c3", "k2_gls_c4", "k2_gls_c5", "k2_gls_c6", "k2_gls_c7", "k2_gls_dense", "gls_dense_nodes"]
print({n: round(I._ctx.timing_or(n, -1), 3) for n in names})
