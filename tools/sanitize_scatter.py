"""compute-sanitizer target: the kernels of rhs_pft.cu on small geometries (no matrix assembly: a seeded vector stands in for the solution).
   compute-sanitizer --tool memcheck python tools/sanitize_scatter.py ;  compute-sanitizer --tool racecheck python tools/sanitize_scatter.py"""
import importlib.util, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("pyscuff", os.path.join(ROOT, "scuff-em_b200", "pyscuff.py"))
ps = importlib.util.module_from_spec(spec); spec.loader.exec_module(ps)
GEODIR = os.path.join(ROOT, "tests", "golden", "geometries")
rng = np.random.default_rng(0)
for name, omega in (("PECSphere_255.scuffgeo", 1.0), ("SiSpheres_255.scuffgeo", 0.7), ("PECSpheres_255.scuffgeo", 0.7), ("OpenPlate_40.scuffgeo", 1.0)):
    G = ps.RWGGeometry(os.path.join(GEODIR, name))
    IF = ps.PlaneWave((1, 0, 0), (0, 0, 1)).chain(ps.PointSource((0.3, 0.2, 7.0), (0, 1, 0)))
    RHS = G.AssembleRHSVector(omega, IF)
    KN = G.AllocateRHSVector(); KN.set(rng.standard_normal(G.TotalBFs) + 1j * rng.standard_normal(G.TotalBFs))
    P = G.GetPFTMatrix(KN, omega, IF, EMTPFTIMethod=0)
    closed = bool(G.AllAllSurfacesClosed) if hasattr(G, "AllAllSurfacesClosed") else bool(G.AllSurfacesClosed)
    if closed:
        Pi = G.GetPFTMatrix(KN, omega, IF, Interior=True, EMTPFTIMethod=0)
        X = np.array([[0.0, 0.0, 9.0], [4.0, 1.0, 2.0], [0.0, 0.0, 0.0]])
        try:
            F = G.GetFields(IF, KN, omega, X)
            RF = G.GetRFMatrix(omega, X[:2]).array
        except ps.ScuffError as e:
            print('   fields refused:', str(e)[:80])
        try:
            D = G.GetDSIPFTMatrix(KN, omega, IF, DSIRadius=1.45, DSIPoints=50)
        except ps.ScuffError as e:
            print('   DSI refused:', str(e)[:80])
    print(name, "ok", float(np.abs(RHS.array).max()), P[0, :2])
