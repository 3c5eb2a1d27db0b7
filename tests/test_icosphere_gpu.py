"""Synthetic geodesic icospheres from scuff-em_b200/meshgen.py (the benchmark's mesh family) at a size the
reference finishes in seconds: full-matrix parity, PEC and eps=10 dielectric, real and imaginary frequency."""
import importlib.util
import os
import numpy as np
import pytest
from conftest import ROOT, ref_matrix
from test_full_gpu import block_errors

pytestmark = pytest.mark.gpu


def _meshgen():
    spec = importlib.util.spec_from_file_location("meshgen", os.path.join(ROOT, "scuff-em_b200", "meshgen.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("nu,material,omega,imag", [(6, None, 2.0, False), (6, None, 1.5j, True), (4, "CONST_EPS_10", 1.0, False)])
def test_icosphere_matrix(ps, nu, material, omega, imag, tmp_path):
    geo = _meshgen().write_sphere_geo(str(tmp_path), nu, material)
    cache = str(tmp_path / "ref.cache")
    Mref, info = ref_matrix(geo, omega, imag=imag, cache_out=cache, tmpdir=str(tmp_path / "refcache"))
    ps.ClearCache()
    ps.PreloadCache(cache)
    G = ps.RWGGeometry(geo)
    assert G.TotalEdges == 30 * nu * nu and G.TotalPanels == 20 * nu * nu
    M = G.AssembleBEMMatrix(omega, M=G.AllocateBEMMatrix(imag)).array
    errs = block_errors(G, M, Mref)
    print("icosphere nu", nu, material, omega, {k: float("%.2e" % v) for k, v in errs.items()}, G.ppi_counts())
    assert max(errs.values()) <= 1e-10
    # and with the FIPPI table built on the device instead of preloaded (the reference's adaptive decisions replicated)
    ps.ClearCache()
    M2 = G.AssembleBEMMatrix(omega, M=G.AllocateBEMMatrix(imag)).array
    err2 = np.linalg.norm(M2 - Mref) / np.linalg.norm(Mref)
    print("   device-built FIPPI table: rel err", err2)
    # this compares with ONE multi-threaded reference run without sharing its FIPPI records: on meshes with many congruent
    # panels the reference differs from itself by up to ~5e-10 run to run (reproducibility note)
    assert err2 <= 1e-9
    # against the DETERMINISTIC reference (one thread: no race on the float-keyed FIPPI cache) the north_star tolerance holds for
    # the device-built table too (one case: the single-threaded reference takes a minute)
    if nu == 6 and not imag:
        Mref1, _ = ref_matrix(geo, omega, imag=imag, tmpdir=str(tmp_path / "refcache1"), env={"OMP_NUM_THREADS": "1"})
        errs1 = block_errors(G, M2, Mref1)
        print("   device-built FIPPI table vs single-threaded reference:", {k: float("%.2e" % v) for k, v in errs1.items()})
        assert max(errs1.values()) <= 1e-10
