"""Developer driver: assemble one synthetic icosphere BEM matrix on the GPU and print kernel timings.
   python tools/run_assemble.py --nu 26 --omega 1.0 [--far-only] [--material CONST_EPS_10] [--reps 3]"""
import argparse, importlib.util, os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scuff-em_b200"))
import meshgen
spec = importlib.util.spec_from_file_location("pyscuff", os.path.join(ROOT, "scuff-em_b200", "pyscuff.py"))
ps = importlib.util.module_from_spec(spec); spec.loader.exec_module(ps)

ap = argparse.ArgumentParser()
ap.add_argument("--nu", type=int, default=26)
ap.add_argument("--omega", type=complex, default=1.0)
ap.add_argument("--far-only", action="store_true")
ap.add_argument("--material", default=None)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--imag", action="store_true")
ap.add_argument("--pft", action="store_true", help="skip the assembly: RHS of a plane wave + EMTPFT sums with that vector as the solution (kernel timing only)")
a = ap.parse_args()
if a.far_only:
    os.environ["SCUFF_CUDA_FAR_ONLY"] = "1"
geo = meshgen.write_sphere_geo("/tmp/scuff_b200_mesh", a.nu, a.material)
G = ps.RWGGeometry(geo)
M = G.AllocateBEMMatrix(a.imag)
peak = ps.measure_fp64_peak()
P = G.TotalPanels
if a.pft:
    import numpy as np
    IF = ps.PlaneWave((1.0, 0.0, 0.0), (0.0, 0.0, 1.0))
    KN = G.AssembleRHSVector(a.omega, IF)
    for r in range(a.reps):
        t0 = time.time(); Pm = G.GetPFTMatrix(KN, a.omega, IF, EMTPFTIMethod=0); wall = time.time() - t0
        npt = 6.0 * P
        print(json.dumps(dict(N=G.TotalBFs, panels=P, pft_wall_s=wall, point_pairs=npt * npt, gpairs_per_s=npt * npt / wall / 1e9, pft=[float(v) for v in Pm[0]])))
    sys.exit(0)
for r in range(a.reps):
    t0 = time.time()
    G.AssembleBEMMatrix(a.omega, M=M, to_host=False)
    wall = time.time() - t0
    t = ps.last_timing()
    farflop = 3492.0 * P * (P + 1) / 2
    print(json.dumps(dict(N=G.TotalBFs, panels=P, wall_s=wall, ms_total=t[0], ms_far=t[1], ms_near=t[2], ms_gather=t[4],
                          far_tflops=farflop / (t[1] * 1e-3) / 1e12, fp64_peak=peak, counts=G.ppi_counts(),
                          gelem_s=G.TotalBFs ** 2 / (t[0] * 1e-3) / 1e9)))
