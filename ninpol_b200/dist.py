"""Multi-GPU wiring: one process per GPU, node-partitioned rows (kernel group K4, host side).

The device side (NCCL all-gather of the CSR row blocks) lives in csrc/k4_shard.cu; this module holds
the host logic that is independent of CUDA and therefore testable on CPU: reading the launcher's
environment, shipping the 128-byte NCCL unique id from rank 0 to the other ranks over a plain TCP
socket, and cutting the node range into contiguous per-rank blocks balanced on a cost prefix sum.
"""
import hashlib
import mmap
import os
import socket
import struct
import time

import numpy as np


class Comm:
    """Rank / world size / NCCL unique id of this process."""

    def __init__(self, rank=0, world=1, local_rank=None, unique_id=None):
        self.rank, self.world = int(rank), int(world)
        self.local_rank = self.rank if local_rank is None else int(local_rank)
        self.unique_id = unique_id


def exchange_bytes(payload, rank, world, addr="127.0.0.1", port=29517, timeout=300.0):
    """Rank 0 serves `payload` to the world-1 other ranks; they return what they received."""
    if world == 1:
        return payload
    if rank == 0:
        srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        srv.bind((addr, port))
        srv.listen(world)
        srv.settimeout(timeout)
        for _ in range(world - 1):
            conn, _a = srv.accept()
            conn.sendall(struct.pack("<I", len(payload)) + payload)
            conn.close()
        srv.close()
        return payload
    deadline = time.time() + timeout
    while True:
        try:
            s = socket.create_connection((addr, port), timeout=5.0)
            break
        except OSError:
            if time.time() > deadline:
                raise
            time.sleep(0.05)
    buf = b""
    while len(buf) < 4:
        chunk = s.recv(4 - len(buf))
        if not chunk:
            raise ConnectionError("unique-id exchange: peer closed before the header arrived")
        buf += chunk
    n = struct.unpack("<I", buf)[0]
    out = b""
    while len(out) < n:
        chunk = s.recv(n - len(out))
        if not chunk:
            raise ConnectionError("unique-id exchange: peer closed")
        out += chunk
    s.close()
    return out


def init_from_env(get_unique_id=None, port_offset=17):
    """Build a Comm from torchrun-style variables (RANK, WORLD_SIZE, LOCAL_RANK, MASTER_ADDR,
    MASTER_PORT).  `get_unique_id` is called on rank 0 only (defaults to the library's NCCL id)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world == 1:
        return Comm(0, 1, local, None)
    addr = os.environ.get("MASTER_ADDR", "127.0.0.1")
    port = int(os.environ.get("MASTER_PORT", "29500")) + port_offset
    payload = b""
    if rank == 0:
        if get_unique_id is None:
            from . import _capi
            get_unique_id = _capi.comm_unique_id
        payload = get_unique_id()
    uid = exchange_bytes(payload, rank, world, addr, port)
    return Comm(rank, world, local, uid)


def node_cost(method, n_elem_per_node, processed):
    """Relative cost of one node: IDW/LS stream ~40 E + 42 bytes; GLS factors a
    ~5.5E x 3E system, ~E^3 flops.  Skipped (Dirichlet) nodes cost a constant."""
    E = np.asarray(n_elem_per_node, dtype=np.float64)
    if method == "gls":
        c = E ** 3 + 50.0 * E + 20.0
    else:
        c = 40.0 * E + 42.0
    return np.where(processed, c, 18.0)


def partition_nodes(cost, world):
    """bounds[world+1]: contiguous node ranges with (nearly) equal summed cost."""
    cost = np.asarray(cost, dtype=np.float64)
    n = len(cost)
    if world == 1:
        return np.array([0, n], dtype=np.int64)
    cum = np.cumsum(cost)
    total = cum[-1] if n else 0.0
    targets = total * np.arange(1, world) / world
    cuts = np.searchsorted(cum, targets, side="left") + 1
    bounds = np.concatenate([[0], np.minimum(cuts, n), [n]]).astype(np.int64)
    return np.maximum.accumulate(bounds)


def assemble_row_blocks(blocks, n_points):
    """Host-side statement of what K4 does on the device: concatenate per-rank CSR row blocks
    (row counts, indices, data, neumann) of contiguous node ranges into the global arrays.
    `blocks` = list of dicts with keys lo, hi, counts, indices, data, neumann, ordered by rank."""
    counts = np.zeros(n_points, dtype=np.int32)
    neumann = np.zeros(n_points, dtype=np.float64)
    for b in blocks:
        counts[b["lo"]:b["hi"]] = b["counts"]
        neumann[b["lo"]:b["hi"]] = b["neumann"]
    indptr = np.zeros(n_points + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    nnz = int(indptr[-1])
    indices = np.empty(nnz, dtype=np.int32)
    data = np.empty(nnz, dtype=np.float64)
    for b in blocks:
        s, e = int(indptr[b["lo"]]), int(indptr[b["hi"]])
        indices[s:e] = b["indices"]
        data[s:e] = b["data"]
    return indptr.astype(np.int32), indices, data, neumann


class _ExactView:
    """Array-interface owner of exactly n items inside a mapping.  scipy copies an index / data array that is
    a view of a base more than twice its size (`_prune_array`); a view that owns its exact extent is kept."""

    def __init__(self, address, n, dtype, keepalive):
        self.size = int(n)
        self._keepalive = keepalive
        self.__array_interface__ = {"shape": (int(n),), "typestr": np.dtype(dtype).str, "data": (int(address), False),
                                    "version": 3}


class SharedOutputs:
    """One host mapping shared by the ranks of a box, holding the CSR arrays of gather="host":
    indptr [n_points+1] int32 | neumann [n_points] f64 | indices [cap] int32 | data [cap] f64 (4 KiB aligned).

    Rank 0 creates `/dev/shm/<name>` (create()), the others open it (attach()) once rank 0 is known to be
    done - the caller separates the two with a barrier - and rank 0 unlinks the name after a second barrier,
    so the segment disappears with the last process.  Every rank writes only its own rows."""

    DIR = "/dev/shm"

    def __init__(self, token, n_points, cap):
        self.n_points, self.cap = int(n_points), max(1, int(cap))
        al = lambda x: (x + 4095) & ~4095
        self.off_indptr = 0
        self.off_neumann = al(4 * (self.n_points + 1))
        self.off_indices = self.off_neumann + al(8 * self.n_points)
        self.off_data = self.off_indices + al(4 * self.cap)
        self.nbytes = self.off_data + al(8 * self.cap)
        self.path = os.path.join(self.DIR, "npb_" + hashlib.sha1(token).hexdigest()[:24])
        self.mm = None

    @classmethod
    def fits(cls, nbytes):
        try:
            st = os.statvfs(cls.DIR)
        except OSError:
            return False
        return st.f_bavail * st.f_frsize > nbytes + (64 << 20)

    def _map(self, fd):
        self.mm = mmap.mmap(fd, self.nbytes, mmap.MAP_SHARED, mmap.PROT_READ | mmap.PROT_WRITE)
        os.close(fd)
        self.buf = np.frombuffer(self.mm, dtype=np.uint8)
        self.indptr = self.buf[self.off_indptr:self.off_indptr + 4 * (self.n_points + 1)].view(np.int32)
        self.neumann = self.buf[self.off_neumann:self.off_neumann + 8 * self.n_points].view(np.float64)
        self.indices = self.buf[self.off_indices:self.off_indices + 4 * self.cap].view(np.int32)
        self.data = self.buf[self.off_data:self.off_data + 8 * self.cap].view(np.float64)

    def exact(self, name, n):
        """The first n items of indptr / indices / data / neumann as an array that is not a view of a larger one."""
        full = getattr(self, name)
        if n == full.size:
            n = full.size
        return np.asarray(_ExactView(full.ctypes.data, n, full.dtype, self.mm))

    def create(self):
        try:
            os.unlink(self.path)
        except OSError:
            pass
        fd = os.open(self.path, os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
        os.ftruncate(fd, self.nbytes)
        self._map(fd)

    def attach(self):
        self._map(os.open(self.path, os.O_RDWR))

    def unlink(self):
        try:
            os.unlink(self.path)
        except OSError:
            pass
