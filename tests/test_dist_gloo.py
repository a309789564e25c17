"""World-size-2 test of the N>1 host path on CPU (gloo): every rank takes the node range the product's
partitioner gives it, produces that range's CSR row block (the oracle stands in for the GPU kernels —
test infrastructure, not a product fallback), the blocks are exchanged, and the assembled matrix must
equal the single-rank result.  This covers partition_nodes / node_cost / assemble_row_blocks /
init_from_env, i.e. everything of K4 that is not NCCL itself."""
import os
import socket
import sys

import numpy as np
import pytest

from helpers import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update({"RANK": str(rank), "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank), "MASTER_ADDR": "127.0.0.1",
                       "MASTER_PORT": str(port), "OPENBLAS_NUM_THREADS": "1"})
    import torch.distributed as td
    import oracle
    from ninpol_b200 import dist, meshgen
    td.init_process_group("gloo", rank=rank, world_size=world)
    comm = dist.init_from_env(get_unique_id=lambda: bytes(range(128)))
    assert comm.rank == rank and comm.world == world and comm.unique_id == bytes(range(128))
    mesh = meshgen.make_case("tet", 5)
    O = oracle.OracleInterpolator().load_mesh(mesh)
    g = O.grid
    flags = np.asarray(mesh.point_data["neumann_flag_u"]).astype(np.int64)
    processed = ~((g.boundary_points != 0) & (flags == 0))
    results = {}
    for method in ("idw", "gls"):
        bounds = dist.partition_nodes(dist.node_cost(method, np.diff(g.esup_ptr), processed), world)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        W, nv = O.interpolate("u", method)
        s, e = W.indptr[lo], W.indptr[hi]
        mine = {"lo": lo, "hi": hi, "counts": np.diff(W.indptr[lo:hi + 1]), "indices": W.indices[s:e].copy(),
                "data": W.data[s:e].copy(), "neumann": nv[lo:hi].copy()}
        blocks = [None] * world
        td.all_gather_object(blocks, mine)
        indptr, indices, data, neumann = dist.assemble_row_blocks(blocks, g.n_points)
        ok = (np.array_equal(indptr, W.indptr) and np.array_equal(indices, W.indices) and
              np.array_equal(data, W.data, equal_nan=True) and np.array_equal(neumann, nv))
        results[method] = (ok, [int(b) for b in bounds])
    td.barrier()
    td.destroy_process_group()
    q.put((rank, results))


def test_two_rank_row_block_gather_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = [q.get(timeout=240) for _ in procs]
    [p.join(60) for p in procs]
    assert sorted(r for r, _ in got) == [0, 1]
    for _, res in got:
        for method, (ok, bounds) in res.items():
            assert ok, method
            assert bounds[0] == 0 and 0 < bounds[1] < bounds[2]
    assert got[0][1]["gls"][1] == got[1][1]["gls"][1]
