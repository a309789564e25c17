"""Multi-GPU parity check, to be launched with torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/check_multi_gpu.py
Every rank holds the whole mesh and computes the CSR rows of its node range (GLS ranks upload only the
slice of the cell fields their nodes read).  gather="all": the row blocks are all-gathered over NCCL and
every rank compares the full result with the oracle (test infrastructure).  gather="root": the blocks go
to rank 0 only; rank 0 compares everything, the other ranks compare the rows they own and check that the
rest of their matrix is empty.  gather="host": every rank copies its rows into one shared host mapping and
every rank compares the full result."""
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
I.min_chunk_nodes = 16                       # cut even these small meshes into several chunks per rank
big = "--big" in sys.argv                    # adds a > 100k-node mesh (Kuhn n = 48: 117,649 nodes)


def rows_of(W, lo, hi):
    a, b = W.indptr[lo], W.indptr[hi]
    return W.indptr[lo:hi + 1] - a, W.indices[a:b], W.data[a:b]


CASES = [("tet", 10, {}), ("mixed", 10, {"a": 2, "b": 5}), ("hex", 12, {}), ("tet", 8, {"scramble": True})]
if big:
    CASES.append(("tet", 48, {}))
for kind, n, kw in CASES:
    mesh = meshgen.make_case(kind, n, **kw)
    I.load_mesh(mesh_obj=mesh)
    O = oracle.OracleInterpolator().load_mesh(mesh, build_psup=False)
    ref = {m: O.interpolate("u", m) for m in ("idw", "ls", "gls")}
    # chunks = 0: plain count + fetch (two-pass, counts all-gathered); 3 / 8: the planned chunk pipeline
    for gather, chunks in (("all", 8), ("all", 0), ("root", 8), ("host", 3), ("host", 0)):
        I.set_gather(gather)
        I.stream_chunks = chunks
        for method in ("idw", "ls", "gls"):
            I.invalidate_inputs()
            W, nv = I.interpolate("u", method)
            Wo, nvo = ref[method]
            bounds = [int(b) for b in I.partition_bounds]
            lo, hi = (0, W.shape[0]) if (gather in ("all", "host") or comm.rank == 0) else (bounds[comm.rank], bounds[comm.rank + 1])
            ip, ix, dt = rows_of(W, lo, hi)
            ipo, ixo, dto = rows_of(Wo, lo, hi)
            same = W.shape == Wo.shape and np.array_equal(ip, ipo) and np.array_equal(ix, ixo)
            same = same and W.nnz == len(dto) and W.indptr[lo] == 0 and W.indptr[-1] == W.indptr[hi]   # other rows empty
            if method == "gls":
                rows = np.repeat(np.arange(hi - lo), np.diff(ipo))
                scale = np.zeros(hi - lo); np.maximum.at(scale, rows, np.abs(dto)); scale[scale == 0] = 1
                err = float(np.max(np.abs(dt - dto) / scale[rows])) if same and len(dto) else 0.0
                good = same and err <= 1e-12 and np.allclose(nv, nvo, rtol=0, atol=1e-12 * max(1.0, np.abs(nvo).max()))
            else:
                good = same and np.array_equal(dt, dto, equal_nan=True) and np.array_equal(nv, nvo)
            print(f"rank {comm.rank}/{comm.world} {kind}{n}{'s' if kw.get('scramble') else ''} gather={gather} chunks={chunks} {method}: "
                  f"{'OK' if good else 'MISMATCH'} nnz {W.nnz} rows [{lo},{hi}) bounds {bounds}", flush=True)
            ok = ok and good
    # ---- flags re-staged slice by slice (npb_set_point_flags_f64_range) ----
    # unchanged row: only this rank's slice is uploaded; a row changed anywhere is caught by the summed checksum, the
    # whole row is uploaded again and the node ranges are cut anew
    I.set_gather("all")
    I.stream_chunks = 8
    for method in ("idw", "gls"):
        I.invalidate_inputs()
        I.interpolate("u", method)
        I.invalidate_inputs()
        W, nv = I.interpolate("u", method)
        b = [int(x) for x in I.partition_bounds]
        flag_bytes = 8 * (b[comm.rank + 1] - b[comm.rank])
        moved = I.last_timings["h2d_input_bytes"] - (80 * max(0, I._elem_range[1] - I._elem_range[0] + 1) if method == "gls" else 0)
        Wo, nvo = ref[method]
        good = moved == flag_bytes and np.array_equal(W.indptr, Wo.indptr) and np.array_equal(W.indices, Wo.indices)
        print(f"rank {comm.rank}/{comm.world} {kind}{n} flag slice, unchanged row, {method}: {'OK' if good else 'MISMATCH'} "
              f"(flag bytes moved {moved}, slice {flag_bytes}, row {8 * W.shape[0]})", flush=True)
        ok = ok and good
    name = "neumann_flag_u"
    flag = np.array(mesh.point_data[name])
    p = mesh.points
    live = p.max(axis=0) > p.min(axis=0)
    hull = np.any(((p == p.min(axis=0)) | (p == p.max(axis=0))) & live[None, :], axis=1)
    idx = np.flatnonzero(hull)
    flip = np.concatenate([idx[:3], idx[-3:]])          # nodes of the first and of the last rank
    flag[flip] = 1.0 - flag[flip]
    mesh.point_data[name] = flag
    I.load_point_data()
    O2 = oracle.OracleInterpolator().load_mesh(mesh, build_psup=False)
    for method in ("idw", "gls"):
        W, nv = I.interpolate("u", method)
        Wo, nvo = O2.interpolate("u", method)
        good = np.array_equal(W.indptr, Wo.indptr) and np.array_equal(W.indices, Wo.indices)
        if method == "idw":
            good = good and np.array_equal(W.data, Wo.data, equal_nan=True) and np.array_equal(nv, nvo)
        else:
            good = good and np.allclose(W.data, Wo.data, rtol=0, atol=1e-11) and np.allclose(nv, nvo, rtol=0, atol=1e-11)
        print(f"rank {comm.rank}/{comm.world} {kind}{n} flag slice, CHANGED row, {method}: {'OK' if good else 'MISMATCH'} nnz {W.nnz}", flush=True)
        ok = ok and good
sys.exit(0 if ok else 1)
