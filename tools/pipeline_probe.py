"""Probe: cost of column-block assembly calls and of the overlapped copy-back (N=20 280 PEC icosphere)."""
import os, sys, time, importlib.util
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
def load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, rel)); m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m); return m
ps = load("pyscuff", "scuff-em_b200/pyscuff.py"); mg = load("meshgen", "scuff-em_b200/meshgen.py")
import torch
nu = int(sys.argv[1]) if len(sys.argv) > 1 else 26
d = "/tmp/probe"; os.makedirs(d, exist_ok=True)
geo = mg.write_sphere_geo(d, nu, sys.argv[2] if len(sys.argv) > 2 else None)
G = ps.RWGGeometry(geo); N = G.TotalBFs
M = G.AllocateBEMMatrix()
G.AssembleBEMMatrix(1.0, M=M, to_host=False)
for K in ((1, 4) if len(sys.argv) > 2 else (1, 2, 4, 8)):
    c = [int(N * np.sqrt(j / K) + 0.5) for j in range(K + 1)]
    tt = []
    for rep in range(2):
        tt = []
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for j in range(K):
            t1 = time.perf_counter()
            G.AssembleBEMMatrixColumns(1.0 + 0.01 * rep, M, c[j], c[j + 1])
            tt.append((1e3 * (time.perf_counter() - t1), ps.last_timing()[:3]))
        torch.cuda.synchronize(); tot = 1e3 * (time.perf_counter() - t0)
    print("K", K, "total ms", round(tot, 2), [(round(a, 2), [round(x, 2) for x in b]) for a, b in tt], flush=True)
for K in (1, 4):
    os.environ["SCUFF_CUDA_PIPELINE_BLOCKS"] = str(K)
    for rep in range(3):
        t0 = time.perf_counter(); G.AssembleBEMMatrix(1.0 + 0.01 * rep, M=M, to_host=True); t = 1e3 * (time.perf_counter() - t0)
    print("to_host K", K, "ms", round(t, 2), flush=True)
# pure download into page-locked memory: contiguous vs strided regions
import ctypes as C
L, ctx = ps.cuda_lib(), ps.context()
L.scuffcuda_matrix_download_region_async.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int]
L.scuffcuda_matrix_download_wait.argtypes = [C.c_void_p, C.c_int]
L.scuffcuda_host_alloc.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
buf = C.c_void_p(); assert L.scuffcuda_host_alloc(ctx, N * N * 16, C.byref(buf)) == 0
def timed(name, regions):
    for rep in range(2):
        t0 = time.perf_counter()
        nb = 0
        for (r0, nr, c0, nc) in regions:
            L.scuffcuda_matrix_download_region_async(ctx, M.device_handle, buf, N, r0, nr, c0, nc); nb += nr * nc * 16
        L.scuffcuda_matrix_download_wait(ctx, 0)
        t = time.perf_counter() - t0
    print(name, "ms", round(1e3 * t, 2), "GB/s", round(nb / t / 1e9, 1), flush=True)
timed("full contiguous", [(0, N, 0, N)])
timed("half columns contiguous", [(0, N, 0, N // 2)])
timed("half rows strided", [(0, N // 2, 0, N)])
timed("quarter rows strided", [(0, N // 4, 0, N)])
c = [int(N * np.sqrt(j / 4) + 0.5) for j in range(5)]
timed("4 trapezoids", [(0, c[j + 1], c[j], c[j + 1] - c[j]) for j in range(4)])
timed("3 lower parts", [(c[j + 1], N - c[j + 1], c[j], c[j + 1] - c[j]) for j in range(3)])
