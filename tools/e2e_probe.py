"""End-to-end time of interpolate() (host buffers, inputs re-staged every call) against the chunking rule.
usage: python tools/e2e_probe.py KIND N METHOD[,METHOD] [MIN_CHUNK_NODES ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import ninpol_b200
from ninpol_b200 import meshgen

kind, n, methods = sys.argv[1], int(sys.argv[2]), sys.argv[3].split(",")
rules = [int(x) for x in sys.argv[4:]] or [200_000, 100_000, 50_000, 25_000]
mesh = meshgen.make_case(kind, n)
I = ninpol_b200.Interpolator()
I.load_mesh(mesh_obj=mesh)
caps = [int(x) for x in os.environ.get("PROBE_CHUNKS", "8").split(",")]   # the constructor's stream_chunks
for method in methods:
  for cap in caps:
    I.stream_chunks = cap
    for rule in rules:
        I.min_chunk_nodes = I.min_chunk_nodes_gls = rule
        ts = []
        for rep in range(9):
            I.invalidate_inputs()
            I._ctx.synchronize()
            t0 = time.perf_counter()
            W, nv = I.interpolate("u", method)
            I._ctx.synchronize()
            ts.append(time.perf_counter() - t0)
        ts = np.array(ts[3:]) * 1e3
        chunks = max(1, min(I.stream_chunks, I.grid.n_points // max(1, rule)))
        print(f"{kind}{n} {method} stream_chunks {cap} min_chunk_nodes {rule:7d} -> {chunks} chunks: e2e median {np.median(ts):8.3f} ms  best {ts.min():8.3f} ms  "
              f"({I.grid.n_points / np.median(ts) / 1e3:.2f} M nodes/s), device pipeline {I.last_timings.get('streamed_ms', -1):.3f} ms", flush=True)
