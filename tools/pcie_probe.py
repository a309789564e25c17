"""Measure pinned host<->device copy bandwidth on this box (explains the e2e - device gap of bench.py).
usage: python tools/pcie_probe.py [GiB]"""
import ctypes
import glob
import os
import sys
import time

cands = glob.glob("/usr/local/cuda/lib64/libcudart.so*")
rt = ctypes.CDLL(sorted(cands)[0])
n = int(float(sys.argv[1]) * (1 << 30)) if len(sys.argv) > 1 else (1 << 30)
h = ctypes.c_void_p()
d = ctypes.c_void_p()
assert rt.cudaSetDevice(0) == 0
assert rt.cudaHostAlloc(ctypes.byref(h), ctypes.c_size_t(n), 0) == 0
assert rt.cudaMalloc(ctypes.byref(d), ctypes.c_size_t(n)) == 0
ctypes.memset(h, 1, n)
for name, kind, dst, src in (("H2D", 1, d, h), ("D2H", 2, h, d)):
    best = 0.0
    for _ in range(4):
        rt.cudaDeviceSynchronize()
        t0 = time.perf_counter()
        assert rt.cudaMemcpy(dst, src, ctypes.c_size_t(n), kind) == 0
        rt.cudaDeviceSynchronize()
        best = max(best, n / (time.perf_counter() - t0) / 1e9)
    print(f"{name} pinned {n / 1e9:.2f} GB: {best:.1f} GB/s")
