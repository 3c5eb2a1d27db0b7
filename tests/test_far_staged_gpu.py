"""The three scatter paths of the far kernel (far_kernel.cu) against each other: "scalar" (18 RED.ADD.F64 per panel pair),
"staged" (shared-memory window per warp + cp.reduce.async.bulk, far_staged_kernel; used for column-block assembly) and
"twopass" (plain stores to the panel-pair array + far_gather_kernel, the default for whole PEC/PEC blocks): same far-pair
matrices, same reference PPI counts, for complex and real (imaginary-frequency) matrices, symmetric diagonal blocks,
off-diagonal blocks of a two-body geometry and column-block assembly."""
import os
import sys
import numpy as np
import pytest
from conftest import GEODIR, ROOT

pytestmark = pytest.mark.gpu


def far_matrix(ps, geo, omega, mode, monkeypatch, small=None, real=False):
    mode = {False: "scalar", True: "staged"}.get(mode, mode)
    monkeypatch.setenv("SCUFF_CUDA_FAR_ONLY", "1")
    monkeypatch.setenv("SCUFF_CUDA_FAR_TWOPASS", "1" if mode == "twopass" else "0")
    monkeypatch.setenv("SCUFF_CUDA_FAR_STAGED", "0" if mode == "scalar" else "1")
    if small is not None:
        monkeypatch.setenv("SCUFF_CUDA_FAR_SMALL", str(small))
    G = ps.RWGGeometry(geo)
    M = G.AssembleBEMMatrix(omega, M=G.AllocateBEMMatrix(real)).array.copy()
    return M, G.ppi_counts(), ps.pair_counts()


@pytest.mark.parametrize("name,omega", [("PECSphere_501", 1.0), ("PECSphere_501", 0.05), ("PECSphere_255", 1.0j), ("PECSphere_255", 0.02j),
                                        ("PECSphere_501", 0.7 + 0.3j), ("PECSpheres_327", 0.3), ("PECSpheres_327", 0.5j),
                                        ("PECSphere_R0P75_414", 2.0)])
def test_staged_equals_scalar(ps, name, omega, monkeypatch):
    geo = os.path.join(GEODIR, name + ".scuffgeo")
    for real in ((False, True) if np.real(omega) == 0 else (False,)):      # LHM_REAL matrices at imaginary frequency (8-byte entries)
        M0, c0, p0 = far_matrix(ps, geo, omega, "scalar", monkeypatch, real=real)
        for mode in ("staged", "twopass"):
            M1, c1, p1 = far_matrix(ps, geo, omega, mode, monkeypatch, real=real)
            err = np.linalg.norm(M1 - M0) / np.linalg.norm(M0)
            print(name, omega, M0.dtype, mode, "vs scalar rel diff", err, "counts", c1, "pairs", p1)
            assert M0.dtype == M1.dtype and (M0.dtype == np.float64) == real
            assert err <= 1e-13
            assert list(c0) == list(c1) and list(p0) == list(p1)
            assert np.array_equal(M1, M1.T) == np.array_equal(M0, M0.T)


def test_staged_small_phase_variants(ps, monkeypatch):
    """the small-phase polynomial variants (|k| (R_a+R_b) <= 1/8, 1/16) against the general range-reduced path"""
    geo = os.path.join(GEODIR, "PECSphere_501.scuffgeo")
    for omega in (0.05, 0.15, 0.05j, 0.15j):
        for mode in ("staged", "twopass"):
            Mg, _, _ = far_matrix(ps, geo, omega, mode, monkeypatch, small=0)
            Ms, _, _ = far_matrix(ps, geo, omega, mode, monkeypatch, small=-1)
            err = np.linalg.norm(Ms - Mg) / np.linalg.norm(Mg)
            print("omega", omega, mode, "small-phase vs general rel diff", err)
            assert err <= 2e-14


def test_staged_icosphere_tiles_and_columns(ps, monkeypatch, tmp_path):
    """a mesh with several CTA tiles per side (5120 panels): whole matrix and two column blocks, staged vs scalar"""
    sys.path.insert(0, os.path.join(ROOT, "scuff-em_b200"))
    import meshgen
    geo = meshgen.write_sphere_geo(str(tmp_path), 16, None)
    omega = 0.8
    M0, c0, _ = far_matrix(ps, geo, omega, "scalar", monkeypatch)
    for mode in ("staged", "twopass"):
        M1, c1, _ = far_matrix(ps, geo, omega, mode, monkeypatch)
        err = np.linalg.norm(M1 - M0) / np.linalg.norm(M0)
        print("icosphere nu=16 N", M0.shape[0], mode, "vs scalar", err, c1, ps.last_timing())
        assert err <= 1e-13 and list(c0) == list(c1)
        assert np.array_equal(M1, M1.T)
    G = ps.RWGGeometry(geo)
    N = G.TotalBFs
    cut = N // 3 + 1
    iu = np.triu_indices(N)
    for (a, b) in ((0, cut), (cut, N)):
        Mc = G.AllocateBEMMatrix()
        G.AssembleBEMMatrixColumns(omega, Mc, a, b)
        A = Mc.array
        U0 = np.triu(M0)[:, a:b]
        e = np.linalg.norm(np.triu(A)[:, a:b] - U0) / np.linalg.norm(U0)
        print("columns", a, b, "rel diff", e)
        assert e <= 1e-13


MIXED_GEO = """OBJECT PECBall
	MESHFILE SSphere_255.msh
ENDOBJECT

OBJECT GlassBall
	MESHFILE Sphere_327.msh
	DISPLACED 0 0 3
	MATERIAL CONST_EPS_10
ENDOBJECT

OBJECT LossyBall
	MESHFILE SSphere_R0P25_255.msh
	DISPLACED 2.5 0 0
	MATERIAL CONST_EPS_10_SIGMA_1
ENDOBJECT
"""


@pytest.mark.parametrize("name,omega", [("SiO2Sphere_501", 1.0), ("E10Spheres_327", 0.3), ("SiSpheres_255", 0.25j), ("LossySphere_327", 0.8),
                                        ("GoldSphere_501", 0.05), ("MIXED", 0.7), ("MIXED", 0.4j)])
def test_twopass_dielectric(ps, name, omega, monkeypatch, tmp_path):
    """blocks with a dielectric side (two values per (Q_a,Q_b) combination, 2x2 / 1x2 / 2x1 stamps, two region passes, complex k):
    the two-pass path against the scalar-atomic scatter, incl. a geometry that mixes a PEC body with two dielectrics"""
    if name == "MIXED":
        import shutil
        for f in ("SSphere_255.msh", "Sphere_327.msh", "SSphere_R0P25_255.msh"):
            shutil.copy(os.path.join(GEODIR, f), tmp_path)
        geo = str(tmp_path / "Mixed.scuffgeo")
        open(geo, "w").write(MIXED_GEO)
    else:
        geo = os.path.join(GEODIR, name + ".scuffgeo")
    for real in ((False, True) if np.real(omega) == 0 else (False,)):
        M0, c0, p0 = far_matrix(ps, geo, omega, "scalar", monkeypatch, real=real)
        M1, c1, p1 = far_matrix(ps, geo, omega, "twopass", monkeypatch, real=real)
        err = np.linalg.norm(M1 - M0) / np.linalg.norm(M0)
        print(name, omega, M0.dtype, "twopass vs scalar rel diff", err, "counts", c1, "pairs", p1, ps.last_timing())
        assert err <= 1e-13
        assert list(c0) == list(c1) and list(p0) == list(p1)
        assert np.array_equal(M1, M1.T) == np.array_equal(M0, M0.T)
