#!/usr/bin/env python
"""bench.py — node weights/sec of the ninpol hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the unmodified reference, host cores

A step is one pass of `interpolate` (K2 weights + K3 CSR emit [+ K4 NCCL gather]) over every node of
one synthetic mesh.  Default workload = BASELINE config C4: Kuhn tetrahedra n = 203 (50.2 M cells,
8.49 M nodes), perturbed 0.25 h, heterogeneous anisotropic K, 50 % Neumann hull nodes, method GLS.
For N > 1 (launched by torchrun, one rank per GPU) the SAME mesh is split by nodes over the ranks
(strong scaling); value = n_nodes / max-over-ranks device time per step.

One JSON line on stdout (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` goes
through the public `Interpolator.interpolate` with host buffers (H2D of the per-variable inputs and D2H
of the CSR inside the timed region); `roofline` is the dominant kernel against the measured HBM peak
(plus an FP64 view, because GLS is FMA-bound); `cpu_baseline` is the compiled reference (oracle/_ref)
on the box's host cores over a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (kind, n, description)
    "tet203": ("tet", 203, "C4: Kuhn tets n=203 (50,192,562 cells / 8,489,664 nodes), perturb 0.25h, heterogeneous anisotropic K, 50% Neumann hull nodes"),
    "tet69": ("tet", 69, "C2: Kuhn tets n=69 (1,971,054 cells), perturb 0.25h, heterogeneous anisotropic K, 50% Neumann hull nodes"),
    "tet40": ("tet", 40, "Kuhn tets n=40 (384,000 cells / 68,921 nodes)"),
    "tet7": ("tet", 7, "C1: Kuhn tets n=7 (2,058 cells / 512 nodes)"),
    "hex200": ("hex", 200, "C3: structured hex box 200^3 (8,000,000 cells / 8,120,601 nodes)"),
    "hex128": ("hex", 128, "C2: structured hex box 128^3 (2,097,152 cells)"),
    "mixed170": ("mixed", 170, "C5: conforming mixed hex/wedge/pyramid/tet box n=170, a=40, b=80 (19.2M cells), 50% Neumann hull nodes"),
    "mixed60": ("mixed", 60, "mixed hex/wedge/pyramid/tet box n=60 (a=15, b=30)"),
}
CPU_SAMPLE = {"tet": ("tet", 40), "hex": ("hex", 64), "mixed": ("mixed", 40)}
VARIABLE = "u"


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def make_mesh(kind, n):
    from ninpol_b200 import meshgen
    kw = {"a": (n * 40) // 170, "b": (n * 80) // 170} if kind == "mixed" else {}
    return meshgen.make_case(kind, n, variable=VARIABLE, **kw)


# ------------------------------------------------------------------------------------------------
# clocks: nvidia-smi sampled DURING the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# algorithmic bytes / flops (stated again in DESIGN.md)
# ------------------------------------------------------------------------------------------------
def algorithmic_model(I, method):
    g = I.grid
    E = np.diff(np.asarray(g.esup_ptr)).astype(np.float64)
    F = np.diff(np.asarray(g.fsup_ptr)).astype(np.float64)
    flags = I._flags_host
    bpts = np.asarray(g.boundary_points) != 0
    processed = ~(bpts & (flags == 0))
    if method == "gls":
        per = 120.0 * E + 60.0 * F + 42.0
    else:
        per = 40.0 * E + 42.0
    nbytes = float(np.where(processed, per, 18.0).sum())
    flops = 0.0
    if method == "gls":
        # B (boundary faces at the node) only enters m for Neumann nodes; interior nodes have B = 0
        fs, fp = np.asarray(g.fsup), np.asarray(g.fsup_ptr)
        bf = np.asarray(g.boundary_faces)[fs]
        B = np.add.reduceat(bf, fp[:-1].clip(max=max(len(bf) - 1, 0))) if len(bf) else np.zeros_like(E)
        B = np.where(np.diff(fp) > 0, B, 0).astype(np.float64)
        m = E + 3.0 * F + np.where(flags != 0, B, 0.0)
        n = 3.0 * E + 1.0
        per_f = 2.0 * m * n * n - (2.0 / 3.0) * n ** 3 + 4.0 * m * n
        live = processed & ~(B >= F)
        flops = float(np.where(live, per_f, 0.0).sum())
    return nbytes, flops, int(processed.sum())


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (driver-measured copy bandwidth, burst)"
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def profiled_traffic(workload, method):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return d.get(f"{workload}:{method}")
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference (oracle/_ref) on the host cores
# ------------------------------------------------------------------------------------------------
def reference_interpolator(kind, n):
    import oracle
    ninpol = oracle.load_reference()
    mesh = make_mesh(kind, n)
    if ninpol is not None:
        I = ninpol.Interpolator()
        t0 = time.time()
        I.load_mesh(mesh_obj=oracle.to_reference_mesh(mesh))
        return I, "reference", time.time() - t0, len(mesh.points), mesh.n_cells
    O = oracle.OracleInterpolator()
    t0 = time.time()
    O.load_mesh(mesh)
    return O, "port", time.time() - t0, len(mesh.points), mesh.n_cells


def time_reference(kind, n, method, steps, warmup):
    I, how, t_load, n_points, n_cells = reference_interpolator(kind, n)
    for _ in range(warmup):
        I.interpolate(VARIABLE, method)
    ts = []
    for _ in range(steps):
        t0 = time.time()
        I.interpolate(VARIABLE, method)
        ts.append(time.time() - t0)
    t = float(np.mean(ts))
    threads = min(16, os.cpu_count() or 1) if how == "reference" else 1
    return {"value": n_points / t, "unit": "nodes/s", "cores": threads, "kind": how,
            "sample": f"{kind} n={n}: {n_cells} cells / {n_points} nodes, interpolate('{VARIABLE}','{method}') mean of {steps} "
                      f"(OPENBLAS_NUM_THREADS=1, host has {os.cpu_count()} logical cores; load_mesh {t_load:.2f}s not counted)",
            "seconds_per_step": t, "load_mesh_s": t_load}


def run_reference(args, rank):
    if rank != 0:
        return
    kind = WORKLOADS[args.workload][0]
    skind, sn = CPU_SAMPLE[kind]
    if args.ref_n:
        sn = args.ref_n
    base = time_reference(skind, sn, args.method, args.steps, args.warmup)
    line = {"impl": "reference", "metric": f"node weights/sec ({args.method.upper()})", "value": base["value"], "unit": "nodes/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["seconds_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload][2], "method": args.method,
                       "note": "reference CPU path timed on a bounded sample of the workload (same generator, smaller n)"},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": "nodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------
# this repository's arm
# ------------------------------------------------------------------------------------------------
class Plumbing:
    """barrier + max-over-ranks; torch.distributed (gloo) only when world > 1."""

    def __init__(self, rank, world):
        self.rank, self.world, self.dist = rank, world, None
        if world > 1:
            import torch
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("gloo", rank=rank, world_size=world)
            self.dist, self.torch = dist, torch

    def barrier(self):
        if self.dist:
            self.dist.barrier()

    def max(self, v):
        if not self.dist:
            return float(v)
        t = self.torch.tensor([float(v)], dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, v):
        if not self.dist:
            return float(v)
        t = self.torch.tensor([float(v)], dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.dist:
            self.dist.destroy_process_group()


def device_step(ctx, method):
    ctx.interpolate_count(method)
    ctx.interpolate_fetch(None, None, None, None)   # K3 fill (+ K4 gather), no D2H


def timed_device_steps(I, plumb, method, steps, warmup, sampler=None):
    ctx = I._ctx
    for _ in range(warmup):
        device_step(ctx, method)
    k2_ms, main_ms = [], []
    plumb.barrier()
    ctx.synchronize()
    if sampler:
        sampler.start()
    l0 = ctx.launch_count()
    ctx.timer_start()
    for _ in range(steps):
        device_step(ctx, method)
        k2_ms.append(ctx.timing_or("k2"))
        main_ms.append(ctx.timing_or("k2_main", ctx.timing_or("k2")))
    ms = ctx.timer_stop()
    ctx.synchronize()
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    plumb.barrier()
    return plumb.max(ms) / steps, float(np.mean(k2_ms)), float(np.mean(main_ms)), launches, clocks


def timed_e2e_steps(I, plumb, method, steps, warmup):
    def one():
        I.invalidate_inputs()
        W, nv = I.interpolate(VARIABLE, method)
        return W, nv
    for _ in range(warmup):
        one()
    plumb.barrier()
    I._ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        W, nv = one()
    I._ctx.synchronize()
    dt = time.perf_counter() - t0
    plumb.barrier()
    # bytes actually copied per step, summed over the ranks (a GLS rank uploads the slice of the cell
    # fields its nodes read; with gather="root" only rank 0 downloads the whole CSR)
    h2d = plumb.sum(I.last_timings["h2d_input_bytes"])
    d2h = plumb.sum(I.last_timings["d2h_bytes"])
    return plumb.max(dt) / steps, int(h2d), int(d2h)


def run_ours(args, rank, world):
    import ninpol_b200
    from ninpol_b200 import dist as nd
    kind, n, desc = WORKLOADS[args.workload]
    if args.n:
        n = args.n
        desc = f"{kind} n={n} (override)"
    plumb = Plumbing(rank, world)
    comm = nd.init_from_env() if world > 1 else nd.Comm(0, 1)
    t0 = time.time()
    mesh = make_mesh(kind, n)
    if rank == 0:
        log(f"mesh {desc}: generated in {time.time() - t0:.1f}s")
    t0 = time.time()
    I = ninpol_b200.Interpolator(comm=comm, pinned_outputs=True, pin_inputs=True, gather=args.gather,
                                 stream_chunks=args.stream_chunks if world == 1 else 0)
    I.load_mesh(mesh_obj=mesh)
    t_load = time.time() - t0
    ctx = I._ctx
    g = I.grid
    n_points, n_elems = g.n_points, g.n_elems
    k1 = {k: ctx.timing_or(k) for k in ("k1", "k1_esup", "k1_esuel", "k1_faces", "k1_fsup", "k1_geom", "h2d_mesh")}
    if rank == 0:
        log(f"load_mesh {t_load:.2f}s wall; device K1 {k1['k1']:.1f} ms (esup {k1['k1_esup']:.1f}, esuel {k1['k1_esuel']:.1f}, "
            f"faces {k1['k1_faces']:.1f}, fsup {k1['k1_fsup']:.1f}, geom {k1['k1_geom']:.1f}); H2D {k1['h2d_mesh']:.1f} ms")
    method = args.method
    try:
        W0, _ = I.interpolate(VARIABLE, method)      # stages the inputs, sets the partition
    except Exception as e:                          # /dev/shm too small on this box (every rank sees the same)
        if world > 1 and args.gather == "host" and "gather='host'" in str(e):
            args.gather = "root"
            I.set_gather("root")
            W0, _ = I.interpolate(VARIABLE, method)
        else:
            raise
    nnz = W0.nnz   # only rank 0 prints it (gather="root": the other ranks hold their own block)
    del W0
    sampler = ClockSampler(ctx.device) if rank == 0 else None
    ms_step, k2_ms, main_ms, launches, clocks = timed_device_steps(I, plumb, method, args.steps, max(args.warmup, 3), sampler)
    e2e_s, h2d, d2h = timed_e2e_steps(I, plumb, method, args.steps, 1)
    value = n_points / (ms_step * 1e-3)
    nbytes, flops, n_proc = algorithmic_model(I, method)
    # this rank's share of the algorithmic work (contiguous node range) for the per-launch figure
    lo, hi = ctx.scalar("row_lo"), ctx.scalar("row_hi")
    share = 1.0
    if world > 1:
        Eall = np.diff(np.asarray(g.esup_ptr)).astype(np.float64)
        share = float(Eall[lo:hi].sum() / max(Eall.sum(), 1.0))
    peak, peak_src = measured_peaks()
    kern_ms = main_ms if main_ms > 0 else k2_ms
    achieved = nbytes * share / (kern_ms * 1e-3) / 1e9
    line = {
        "metric": f"node weights/sec ({method.upper()})", "value": value, "unit": "nodes/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "method": method, "n_nodes": n_points, "n_cells": n_elems, "nnz": nnz,
                   "processed_nodes": n_proc, "partition": f"nodes split into {world} contiguous cost-balanced ranges; mesh replicated; e2e gather='{args.gather}'",
                   "cache": "inputs larger than L2 (working set >> 126 MB); no flush needed" if n_elems > 2_000_000 else
                            "small workload: L2-resident between iterations"},
        "e2e": {"value": n_points / e2e_s, "unit": "nodes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s * 1e3, "api": f"Interpolator(pinned_outputs=True, pin_inputs=True, gather='{args.gather}', stream_chunks={args.stream_chunks if world == 1 else 0}).interpolate(variable, method) after "
                "invalidate_inputs(): H2D of flags (+ the slice of permeability / diff_mag this rank's nodes read) from page-locked "
                "host arrays + K2 + K3 + (N > 1) K4: row counts all-gathered over NCCL, then gather='host': every rank copies its CSR rows "
                "into one page-locked host mapping shared by the ranks (/dev/shm), NCCL barrier; gather='root': row blocks to rank 0 over "
                "NCCL, rank 0 downloads the CSR; scipy.csr_matrix wrap; byte counts are sums over ranks; stream_chunks > 0 (1 GPU): the three legs run as a pipeline over node chunks"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_gls_mf (largest size class)" if method == "gls" else f"k_{method}",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": profiled_traffic(args.workload, method), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": nbytes * share, "kernel_ms": kern_ms, "k2_ms": k2_ms},
        "load_mesh": {"k1_device_ms": k1["k1"], "h2d_ms": k1["h2d_mesh"], "cells_per_s_device": n_elems / (k1["k1"] * 1e-3) if k1["k1"] else None,
                      "wall_s": t_load, "breakdown_ms": k1},
    }
    if method == "gls":
        fp64_peak = ctx.measure_fp64_peak() if rank == 0 else 0.0
        ach_tf = flops * share / (kern_ms * 1e-3) / 1e12
        line["roofline"]["fp64"] = {"achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_tf / fp64_peak if fp64_peak else None,
                                    "flop_model": "sum over processed nodes of 2mn^2 - 2/3 n^3 + 4mn, m=E+3F+B, n=3E+1 (one-RHS Householder QR)",
                                    "peak_source": "measured live: register-resident DFMA loop (npb_measure_fp64_peak)",
                                    "note": "GLS is FP64-FMA bound, not HBM bound (SURVEY.md Q13); the HBM fraction is reported because the metric asks for it"}
    # secondary methods on the same mesh (device-timed)
    if args.also and rank == 0 or (args.also and world > 1):
        also = {}
        for m2 in args.also.split(","):
            if m2 == method or m2 not in ("idw", "ls", "gls"):
                continue
            I.interpolate(VARIABLE, m2)
            ms2, k2b, mainb, _l, _c = timed_device_steps(I, plumb, m2, args.steps, 3, None)
            nb2, _f, _p = algorithmic_model(I, m2)
            kk = mainb if mainb > 0 else k2b
            also[m2] = {"value": n_points / (ms2 * 1e-3), "unit": "nodes/s", "ms_per_step": ms2, "k2_ms": k2b,
                        "roofline": {"bound": "hbm", "achieved": nb2 * share / (kk * 1e-3) / 1e9, "peak": peak,
                                     "frac": nb2 * share / (kk * 1e-3) / 1e9 / peak, "unit": "GB/s"}}
        line["also"] = also
    if rank == 0 and world == 1 and not args.no_cpu:
        skind, sn = CPU_SAMPLE[kind]
        if args.ref_n:
            sn = args.ref_n
        try:
            base = time_reference(skind, sn, method, 2, 1)
            line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:  # the checker must never take the bench down
            line["cpu_baseline"] = {"value": None, "unit": "nodes/s", "cores": 0, "kind": "unavailable", "sample": repr(e)}
    if rank == 0:
        emit(line)
    plumb.close()


_JSON_OUT = None


def emit(line):
    """The one JSON line goes to the real stdout; everything else a library prints on fd 1 (NCCL's version
    banner, for one) was redirected to stderr in main()."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _JSON_OUT
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="tet203", choices=sorted(WORKLOADS))
    ap.add_argument("--method", default="gls", choices=["gls", "idw", "ls"])
    ap.add_argument("--also", default="idw", help="comma list of extra methods to time on the same mesh")
    ap.add_argument("--n", type=int, default=0, help="override the lattice size of the workload (debug)")
    ap.add_argument("--ref-n", type=int, default=0, help="lattice size of the CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--gather", default="host", choices=["host", "root", "all"],
                    help="N > 1: where the row blocks meet (host: shared host mapping; root: rank 0's GPU; all: every GPU)")
    ap.add_argument("--stream-chunks", type=int, default=8,
                    help="e2e at 1 GPU: node chunks of the upload / compute / download pipeline (0 = plain count + fetch)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            sys.exit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, world)


if __name__ == "__main__":
    main()
