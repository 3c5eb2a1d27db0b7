"""Frequency sweep through the sweep driver (config "tests/Mie": SiO2Sphere_501, PMCHWT, frequencies of tests/Mie/OmegaFile):
geometry and FIPPI table stay resident, every point is assembled, factored and solved on the device, and the solution of
M x = b (fixed b) is compared with the solve of the reference's matrix at the same frequency (<= 1e-8, north_star)."""
import ctypes as C
import importlib.util
import os
import numpy as np
import pytest
from conftest import GEODIR, ROOT, ref_matrix

pytestmark = pytest.mark.gpu


def _sweep():
    spec = importlib.util.spec_from_file_location("sweep", os.path.join(ROOT, "scuff-em_b200", "sweep.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    return mod


def test_mie_frequency_sweep(ps, tmp_path):
    geo = os.path.join(GEODIR, "SiO2Sphere_501.scuffgeo")
    omegas = [float(l.split()[0]) for l in open(os.path.join(GEODIR, "OmegaFile")) if l.strip()][::3]   # every third point
    G = ps.RWGGeometry(geo)
    N = G.TotalBFs
    rng = np.random.default_rng(7)
    rhs = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    M = G.AllocateBEMMatrix()
    X = ps.HMatrix.new(N, 1, False)
    xbuf = np.frombuffer((C.c_double * (2 * N)).from_address(ps.host_lib().scuffb200_matrix_data(X.ptr)), dtype=np.complex128)
    ps.ClearCache()      # the FIPPI table is built on the device at the first point and reused by all others

    def evaluate(omega):
        G.AssembleBEMMatrix(omega, M=M, to_host=False)
        xbuf[:] = rhs
        M.lu_factorize()
        M.lu_solve(X)
        return np.array(xbuf)

    sols = _sweep().run_sweep(omegas, evaluate)
    assert len(sols) == len(omegas)
    for omega, x in zip(omegas, sols):
        Mref, _ = ref_matrix(geo, omega)
        xref = np.linalg.solve(Mref, rhs)
        err = np.linalg.norm(x - xref) / np.linalg.norm(xref)
        print("omega", omega, "solution rel err", err)
        assert err <= 1e-8
