"""CPU tests of the host side: the C-ABI library loads and exports every symbol the header declares,
the product fails loudly without a GPU (no CPU fallback), and the vectorised ingest / partition logic
matches literal restatements of the reference's Python loops."""
import os
import re
import threading

import numpy as np
import pytest

from helpers import ROOT

from ninpol_b200 import _capi, dist, element_tables as et, meshgen


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ninpol_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(npb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_library):
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(built_library, s), f"{s} declared in include/ninpol_b200.h but not exported"
    for s in _capi.EXPORTED_SYMBOLS:
        assert s in syms, f"{s} bound in _capi.py but not declared in the header"
    assert built_library.npb_version() >= 100


def test_no_cpu_fallback(built_library):
    """Without a usable sm_100 device every entry point that needs one fails; nothing is computed on
    the host instead."""
    try:
        n = _capi.device_count()
    except _capi.NinpolB200Error:
        n = 0
    if n > 0:
        pytest.skip("a GPU is visible: covered by the -m gpu suite")
    import ninpol_b200
    with pytest.raises(_capi.NinpolB200Error):
        ninpol_b200.Interpolator()
    with pytest.raises(_capi.NinpolB200Error):
        _capi.Context(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ninpol_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def literal_process_mesh(mesh):
    """reference interpolator.pyx:333-361, loops as written."""
    dim = 1
    for blk in mesh.cells:
        for d, names in et.TYPES_PER_DIMENSION.items():
            if blk.type in names:
                dim = max(dim, d)
    n_elems = sum(len(b.data) for b in mesh.cells if b.type in et.TYPES_PER_DIMENSION[dim])
    conn = -np.ones((n_elems, 8), dtype=np.int64)
    types = -np.ones(n_elems, dtype=np.int64)
    ci = 0
    for blk in mesh.cells:
        if blk.type not in et.TYPES_PER_DIMENSION[dim]:
            continue
        tid = et.POINT_ORDERING["elements"][blk.type]["element_type"]
        for cell in blk.data:
            for j, p in enumerate(cell):
                conn[ci, j] = p
            types[ci] = tid
            ci += 1
    return dim, n_elems, conn, types


class _HostOnly:
    """Interpolator's host methods without a device context."""

    def __new__(cls):
        from ninpol_b200.interpolator import Interpolator
        obj = object.__new__(Interpolator)
        obj.point_ordering = et.POINT_ORDERING
        obj.types_per_dimension = {k: list(v) for k, v in et.TYPES_PER_DIMENSION.items()}
        obj.logging, obj.build_edges = False, False
        obj.variable_to_index = {"points": {}, "cells": {}, "faces": {}}
        obj._rows = {"cells": [], "points": []}
        obj._dense = {"cells": None, "points": None}
        obj._data_version, obj._staged = 0, None
        return obj


@pytest.mark.parametrize("kind,n,kw", [("tet", 3, {}), ("mixed", 6, {"a": 1, "b": 3})])
def test_vectorised_process_mesh_equals_reference_loops(kind, n, kw):
    mesh = meshgen.make_case(kind, n, **kw)
    # a lower-dimensional block must be ignored (interpolator.pyx:335-336)
    mesh.cells.append(meshgen.CellBlock("triangle", np.array([[0, 1, 2]])))
    I = _HostOnly()
    args = I.process_mesh(mesh)
    dim, n_elems, conn, types = literal_process_mesh(mesh)
    assert args[0] == dim and args[1] == n_elems and args[2] == len(mesh.points)
    assert np.array_equal(args[9], conn) and np.array_equal(args[10], types)
    npoel, nfael, lnofa, lpofa, nedel, lpoed = args[3:9]
    assert npoel[4] == 4 and nfael[4] == 4 and nfael[2] == -1 and lnofa[5, 0] == 4 and lnofa[4, 0] == 3
    assert list(lpofa[4, 0]) == [0, 2, 1, -1] and list(lpofa[7, 0]) == [0, 3, 2, 1]


def test_vectorised_load_data_equals_reference_loops():
    class G:
        n_elems, n_points, dim = 5, 4, 3
    I = _HostOnly()
    I.grid = G()
    rng = np.random.default_rng(0)
    data = {"s1": rng.random(5), "s2": rng.random((5, 1)), "v": rng.random((5, 9))}
    I.load_data(data, "cells")
    want = np.zeros((3, 45))
    for idx, (name, a) in enumerate(data.items()):      # interpolator.pyx:403-419
        cur = a.shape[1] if a.ndim > 1 else 1
        for e in range(5):
            if cur == 1:
                want[idx, e] = a[e] if a.ndim == 1 else a[e][0]
            else:
                for j in range(cur):
                    want[idx, e * cur + j] = a[e][j]
    assert np.array_equal(I.cells_data, want)
    assert list(I.cells_data_dimensions) == [1, 1, 9]
    assert I.variable_to_index["cells"] == {"s1": 0, "s2": 1, "v": 2}


def test_diffusion_magnitude_release_semantics():
    from ninpol_b200.interpolator import Interpolator
    K = meshgen.random_spd_permeability(64, seed=3)
    dm = Interpolator.compute_diffusion_magnitude(K)
    Ks = K.reshape(-1, 3, 3)
    literal = (1 - (3 * (np.linalg.det(Ks) ** (1 // 3)) / np.trace(Ks, axis1=1, axis2=2))) ** 2   # 1/3 == 0 under cdivision
    assert np.array_equal(dm, literal)


def test_partition_is_contiguous_balanced_and_complete():
    rng = np.random.default_rng(1)
    E = rng.integers(1, 40, size=10000)
    processed = rng.random(10000) > 0.1
    for method in ("idw", "gls"):
        cost = dist.node_cost(method, E, processed)
        for world in (1, 2, 3, 8):
            b = dist.partition_nodes(cost, world)
            assert b[0] == 0 and b[-1] == len(cost) and len(b) == world + 1 and np.all(np.diff(b) >= 0)
            if world > 1:
                sums = np.add.reduceat(cost, b[:-1])
                assert sums.max() <= cost.sum() / world + cost.max() * 2
    assert list(dist.partition_nodes(np.ones(3), 8)) == sorted(dist.partition_nodes(np.ones(3), 8))


def test_unique_id_exchange_over_tcp():
    payload = bytes(range(128))
    out = {}

    def run(rank):
        out[rank] = dist.exchange_bytes(payload if rank == 0 else b"", rank, 3, "127.0.0.1", 29731)

    ts = [threading.Thread(target=run, args=(r,)) for r in range(3)]
    [t.start() for t in ts]
    [t.join(30) for t in ts]
    assert out == {0: payload, 1: payload, 2: payload}


def test_assemble_row_blocks_roundtrip():
    import scipy.sparse as sp
    rng = np.random.default_rng(2)
    W = sp.random(50, 30, density=0.2, random_state=3, format="csr")
    nv = rng.random(50)
    bounds = [0, 17, 17, 41, 50]
    blocks = []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        s, e = W.indptr[lo], W.indptr[hi]
        blocks.append({"lo": lo, "hi": hi, "counts": np.diff(W.indptr[lo:hi + 1]), "indices": W.indices[s:e], "data": W.data[s:e],
                       "neumann": nv[lo:hi]})
    indptr, indices, data, neumann = dist.assemble_row_blocks(blocks, 50)
    assert np.array_equal(indptr, W.indptr) and np.array_equal(indices, W.indices) and np.array_equal(data, W.data)
    assert np.array_equal(neumann, nv)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` runs on host cores only (compiled reference when oracle/_ref exists,
    else the oracle port) and prints ONE JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-workload", "tet7"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "nodes/s" and d["higher_is_better"] is True
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data",
                "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]


def test_shared_outputs_mapping_is_shared_and_disappears(tmp_path, monkeypatch):
    """gather="host" plumbing: two mappings of one segment see each other's writes; after unlink the name
    is gone but the mappings live on; the four arrays do not overlap."""
    monkeypatch.setattr(dist.SharedOutputs, "DIR", str(tmp_path))
    a = dist.SharedOutputs(b"token-1", n_points=1000, cap=5000)
    b = dist.SharedOutputs(b"token-1", n_points=1000, cap=5000)
    assert a.path == b.path and dist.SharedOutputs(b"token-2", 1000, 5000).path != a.path
    assert dist.SharedOutputs.fits(a.nbytes)
    a.create()
    b.attach()
    a.unlink()
    assert not os.path.exists(a.path)
    assert a.indptr.shape == (1001,) and a.indices.shape == (5000,) and a.data.shape == (5000,) and a.neumann.shape == (1000,)
    a.indptr[:] = np.arange(1001)
    a.indices[:] = 7
    b.data[:] = 0.5
    b.neumann[:] = -1.0
    assert np.array_equal(b.indptr, np.arange(1001)) and np.all(b.indices == 7)
    assert np.all(a.data == 0.5) and np.all(a.neumann == -1.0)
    import scipy.sparse as sp
    a.indptr[:] = 0
    a.indptr[1:] = 3                      # one row of three entries
    a.indices[:3] = [0, 1, 2]
    W = sp.csr_matrix((a.exact("data", 3), a.exact("indices", 3), a.exact("indptr", 1001)), shape=(1000, 10), copy=False)
    b.data[0] = 42.0                      # written through the other mapping: visible if scipy kept the view
    assert W.data[0] == 42.0 and W.data.shape == (3,)
    offs = [a.off_indptr, a.off_neumann, a.off_indices, a.off_data, a.nbytes]
    assert offs == sorted(offs) and all(o % 4096 == 0 for o in offs)


def test_pinned_pool_reuses_a_block_only_when_unreferenced(monkeypatch):
    """_PinnedPool (the default output buffers of interpolate()): drop-in semantics need that a result the caller
    still holds is never overwritten; a steady loop must not allocate."""
    from ninpol_b200 import _capi, interpolator
    allocs = []

    def fake_pinned_empty(n, dtype):
        allocs.append(n)
        return np.empty(n, dtype)

    monkeypatch.setattr(_capi, "pinned_empty", fake_pinned_empty)
    P = interpolator._PinnedPool()
    a = P.take("data", 100, np.float64)
    a[:] = 1.0
    b = P.take("data", 100, np.float64)          # a is alive: a second block
    b[:] = 2.0
    assert len(allocs) == 2 and a[0] == 1.0
    del a
    c = P.take("data", 90, np.float64)           # a's block is free again
    assert len(allocs) == 2
    v = c[3:7]                                   # a view of a view keeps the block busy
    del c
    d = P.take("data", 100, np.float64)
    assert len(allocs) == 3 and v.base is not None
    del b, d, v
    for _ in range(5):                           # steady state: no growth
        e = P.take("data", 100, np.float64)
        del e
    assert len(allocs) == 3
    small = P.take("data", 10, np.float64)       # a block more than twice the request is never handed out:
    assert len(allocs) == 4 and small.base.size <= 2 * 10 + 64   # scipy would copy such a view


def test_flag_row_is_restaged_by_slices_only_when_unchanged():
    """Multi-GPU re-staging (interpolator._stage_inputs): with a resident flag row and a current partition a rank uploads
    only the slice of its own nodes; the summed slice checksums decide whether that was enough.  Host logic only: the
    device context is a stand-in that records what it is asked to do."""
    calls = []

    def checksum(flags):
        return int(np.flatnonzero(np.asarray(flags) != 0).sum() * 2654435761 % (1 << 62))

    class Ctx:
        resident = None

        def set_point_flags(self, flags):
            calls.append(("full", flags.size))
            Ctx.resident = checksum(flags)

        def set_point_flags_slice(self, flags, first, count):
            calls.append(("slice", first, count))
            return checksum(flags)               # what the all-reduce of the ranks' slice checksums yields

        def scalar(self, name):
            assert name == "flags_checksum"
            return Ctx.resident

        def set_partition(self, bounds):
            calls.append(("partition", tuple(int(b) for b in bounds)))

    class Comm:
        world, rank = 2, 1

    class G:
        n_points, n_elems, dim = 1000, 10, 3
        esup_ptr = np.arange(0, 4 * 1001, 4)
        boundary_points = np.zeros(1000, dtype=np.int64)

    G.boundary_points[:100] = 1
    I = _HostOnly()
    I.grid, I.comm, I._ctx = G(), Comm(), Ctx()
    I.pin_inputs, I._registered, I.last_timings = False, {}, {}
    I._flag_mask, I._flag_version, I._partition_key, I._mesh_serial = None, 0, None, 1
    I._pending_key = None
    flags = np.zeros(1000)
    flags[:50] = 1.0
    I.variable_to_index = {"points": {"neumann_flag_u": 0}, "cells": {"u": 0}, "faces": {}}
    I._rows = {"cells": [np.zeros(10)], "points": [flags]}
    stage = lambda: I._stage_inputs("idw", "u", I.variable_to_index, I._rows["cells"], I._rows["points"])
    stage()                                                     # first time: the whole row, then the node ranges
    assert [c[0] for c in calls] == ["full", "partition"] and I.last_timings["h2d_input_bytes"] == 8000
    bounds = calls[1][1]
    calls.clear()
    stage()                                                     # resident and keyed: nothing moves
    assert calls == []
    I.invalidate_inputs()
    stage()                                                     # re-staged, unchanged: this rank's slice only
    lo, hi = bounds[1], bounds[2]
    assert calls == [("slice", lo, hi - lo)] and I.last_timings["h2d_input_bytes"] == 8 * (hi - lo)
    calls.clear()
    flags2 = flags.copy()
    flags2[60:80] = 1.0                                         # changed outside this rank's slice too
    I._rows["points"][0] = flags2
    I._data_version += 1
    stage()                                                     # the slice checksums do not add up: whole row, new cut
    assert [c[0] for c in calls] == ["slice", "full", "partition"] and I.last_timings["h2d_input_bytes"] == 8000
    assert I._flag_version == 2
