"""Full BEM matrix parity against the UNMODIFIED reference (oracle/_ref) — relative Frobenius error per
block <= 1e-10 (BASELINE.json north_star).  Here the FIPPI table is preloaded from the reference's own
cache file (StoreCache -> PreloadCache round trip through the reference's binary format), so these
tests pin everything that depends on frequency; tests/test_fippi_gpu.py pins the device FIPPI builder."""
import os
import numpy as np
import pytest
from conftest import GEODIR, ref_matrix

pytestmark = pytest.mark.gpu
TOL = 1e-10

CASES = [("PECSphere_501", 1.0, False), ("PECSphere_501", 0.01, False), ("PECSphere_255", 0.1j, False),
         ("PECSphere_255", 0.1j, True), ("SiO2Sphere_501", 1.0, False), ("SiSphere_255", 0.01, False),
         ("PECSpheres_255", 0.1, False), ("SiSpheres_255", 0.1, False), ("LossySphere_327", 1.0, False),
         ("PECSpheres_327", 1.0j, True), ("E10Spheres_327", 0.5j, False), ("GoldSphere_501", 0.3, False),
         ("OpenPlate_40", 1.0, False), ("OpenPlate_40", 2.0j, True)]   # open PEC surface: exterior edges carry no basis function


def block_errors(G, M, Mref):
    errs = {}
    offs = [G.surface(ns)["BFOffset"] for ns in range(G.NumSurfaces)] + [G.TotalBFs]
    for a in range(G.NumSurfaces):
        for b in range(G.NumSurfaces):
            A = M[offs[a]:offs[a + 1], offs[b]:offs[b + 1]]
            B = Mref[offs[a]:offs[a + 1], offs[b]:offs[b + 1]]
            nb = np.linalg.norm(B)
            errs[(a, b)] = np.linalg.norm(A - B) / nb if nb > 0 else np.linalg.norm(A)
    return errs


@pytest.mark.parametrize("name,omega,imag", CASES)
def test_full_matrix(ps, name, omega, imag, tmp_path):
    geo = os.path.join(GEODIR, name + ".scuffgeo")
    cache = str(tmp_path / "ref.cache")
    # On structured meshes (all panels congruent) the multi-threaded reference is not reproducible against ITSELF to 1e-10:
    # threads that miss the translation-invariant FIPPI cache simultaneously each use their own record and only one is
    # stored (measured run-to-run 1.6e-10 .. 4.9e-10 on OpenPlate_40, 1.5e-8 vs its own single-threaded run).  The
    # single-threaded reference is deterministic and is what the stored cache describes.
    env = {"OMP_NUM_THREADS": "1"} if name.startswith("OpenPlate") else None
    Mref, info = ref_matrix(geo, omega, imag=imag, cache_out=cache, env=env)
    ps.ClearCache()            # pin the FIPPI table to THIS reference run's records (the reference is not reproducible
    ps.PreloadCache(cache)     # run to run below ~1e-10: racing threads store only one of their records per key)
    G = ps.RWGGeometry(geo)
    Mh = G.AllocateBEMMatrix(imag)
    assert Mh.is_real == bool(info["real"])
    M = G.AssembleBEMMatrix(omega, M=Mh).array
    errs = block_errors(G, M, Mref)
    print(name, omega, "real" if imag else "complex", "block errs", {k: float("%.2e" % v) for k, v in errs.items()},
          "counts", G.ppi_counts())
    assert max(errs.values()) <= TOL
    assert np.linalg.norm(M - Mref) / np.linalg.norm(Mref) <= TOL


def test_lu_solve_downstream(ps, tmp_path):
    """downstream consumer: on-device LU (cuSOLVER) of the assembled matrix vs numpy solve of the reference matrix"""
    geo = os.path.join(GEODIR, "PECSphere_501.scuffgeo")
    cache = str(tmp_path / "ref.cache")
    Mref, _ = ref_matrix(geo, 1.0, cache_out=cache)
    ps.ClearCache()            # pin the FIPPI table to THIS reference run's records (the reference is not reproducible
    ps.PreloadCache(cache)     # run to run below ~1e-10: racing threads store only one of their records per key)
    G = ps.RWGGeometry(geo)
    Mh = G.AssembleBEMMatrix(1.0, to_host=False)
    rng = np.random.default_rng(1)
    rhs = rng.standard_normal(G.TotalBFs) + 1j * rng.standard_normal(G.TotalBFs)
    X = ps.HMatrix.new(G.TotalBFs, 1, False)
    import ctypes as C
    host = ps.host_lib()
    buf = (C.c_double * (2 * G.TotalBFs)).from_address(host.scuffb200_matrix_data(X.ptr))
    np.frombuffer(buf, dtype=np.complex128)[:] = rhs
    Mh.lu_factorize()
    Mh.lu_solve(X)
    x = X.array[:, 0]
    xref = np.linalg.solve(Mref, rhs)
    assert np.linalg.norm(x - xref) / np.linalg.norm(xref) <= 1e-8


def test_casimir_energy_integrand_downstream(ps, tmp_path):
    """downstream quantity of the imaginary-frequency path (scuff-cas3D, CasimirIntegrand.cc:296-312,428-451): the energy
    integrand log det(M) - log det(M_infinity) with M_infinity = the diagonal blocks; <= 1e-8 relative (north_star)"""
    geo = os.path.join(GEODIR, "PECSpheres_327.scuffgeo")
    vals = []
    for xi in (0.3, 1.0):
        cache = str(tmp_path / ("ref%g.cache" % xi))
        Mref, info = ref_matrix(geo, 1j * xi, imag=True, cache_out=cache)
        ps.ClearCache()
        ps.PreloadCache(cache)
        G = ps.RWGGeometry(geo)
        M = G.AssembleBEMMatrix(1j * xi, M=G.AllocateBEMMatrix(True)).array
        n = G.surface(1)["BFOffset"]

        def integrand(A):
            s, ld = np.linalg.slogdet(A)
            s1, l1 = np.linalg.slogdet(A[:n, :n])
            s2, l2 = np.linalg.slogdet(A[n:, n:])
            return ld - l1 - l2
        mine, ref = integrand(M), integrand(Mref)
        vals.append((xi, mine, ref))
        assert abs(mine - ref) <= 1e-8 * abs(ref), (xi, mine, ref)
    print("Casimir energy integrand (xi, mine, reference):", vals)
