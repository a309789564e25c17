"""Multi-GPU parity check, to be launched with torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/check_multi_gpu.py
Every rank holds the whole mesh, computes the CSR rows of its node range, the row blocks are all-gathered
over NCCL, and every rank compares the full result with the oracle (test infrastructure)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ninpol_b200
import oracle
from ninpol_b200 import dist, meshgen

comm = dist.init_from_env()
ok = True
I = ninpol_b200.Interpolator(comm=comm)      # one NCCL communicator per unique id: reuse the object
for kind, n, kw in (("tet", 10, {}), ("mixed", 10, {"a": 2, "b": 5}), ("hex", 12, {})):
    mesh = meshgen.make_case(kind, n, **kw)
    I.load_mesh(mesh_obj=mesh)
    O = oracle.OracleInterpolator().load_mesh(mesh)
    for method in ("idw", "ls", "gls"):
        W, nv = I.interpolate("u", method)
        Wo, nvo = O.interpolate("u", method)
        same = np.array_equal(W.indptr, Wo.indptr) and np.array_equal(W.indices, Wo.indices)
        if method == "gls":
            rows = np.repeat(np.arange(W.shape[0]), np.diff(Wo.indptr))
            scale = np.zeros(W.shape[0]); np.maximum.at(scale, rows, np.abs(Wo.data)); scale[scale == 0] = 1
            err = float(np.max(np.abs(W.data - Wo.data) / scale[rows])) if same and W.nnz else 0.0
            good = same and err <= 1e-12 and np.allclose(nv, nvo, rtol=0, atol=1e-12 * max(1.0, np.abs(nvo).max()))
        else:
            good = same and np.array_equal(W.data, Wo.data, equal_nan=True) and np.array_equal(nv, nvo)
        bounds = getattr(I, "partition_bounds", None)
        print(f"rank {comm.rank}/{comm.world} {kind}{n} {method}: {'OK' if good else 'MISMATCH'} nnz {W.nnz} bounds {None if bounds is None else list(map(int, bounds))}", flush=True)
        ok = ok and good
sys.exit(0 if ok else 1)
