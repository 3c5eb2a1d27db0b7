"""Short-wavelength regime (|k| R_panel > pi, PanelPanelInteractions.cc:426-491): 78x78-point cubature for near pairs
without common vertices, Taylor-Duffy for common-vertex / common-edge pairs (3-D / 2-D p-adaptive Clenshaw-Curtis with
the reference's MaxEval=1000 budget), and the SCUFF_DISABLE_CACHE switch (Taylor-Duffy for every pair with common
vertices).  Full matrices against the UNMODIFIED reference, <= 1e-10 relative Frobenius per block."""
import os
import numpy as np
import pytest
from conftest import GEODIR, ref_matrix
from test_full_gpu import block_errors

pytestmark = pytest.mark.gpu
TOL = 1e-10

# (geometry, omega, real matrix) — |k| R straddles pi for the first three (per-pair mixture of algorithms)
CASES = [("PECSphere_501", 14.0, False), ("PECSpheres_327", 12.0j, True), ("E10Spheres_327", 4.0j, False),
         ("GoldSphere_501", 0.3, False), ("SiO2Sphere_501", 16.0, False)]


@pytest.mark.parametrize("name,omega,imag", CASES)
def test_short_wavelength_matrix(ps, name, omega, imag, tmp_path):
    geo = os.path.join(GEODIR, name + ".scuffgeo")
    cache = str(tmp_path / "ref.cache")
    Mref, info = ref_matrix(geo, omega, imag=imag, cache_out=cache)
    ps.ClearCache()            # pin the FIPPI table to THIS reference run's records (the reference is not reproducible
    ps.PreloadCache(cache)     # run to run below ~1e-10: racing threads store only one of their records per key)
    G = ps.RWGGeometry(geo)
    M = G.AssembleBEMMatrix(omega, M=G.AllocateBEMMatrix(imag)).array
    counts = G.ppi_counts()
    errs = block_errors(G, M, Mref)
    print(name, omega, "block errs", {k: float("%.2e" % v) for k, v in errs.items()}, "counts", counts, "timing", ps.last_timing())
    assert counts[1] > 0 and counts[2] > 0          # high-order cubature and Taylor-Duffy were exercised
    assert max(errs.values()) <= TOL


# |k| R > 20 for (some) panels: high-k Taylor-Duffy kernels (PPIALG_HKTD), twice-integrable reduction
HK_CASES = [("PECSphere_255", 90.0, False), ("PECSpheres_327", 100.0j, True), ("SiSphere_255", 30.0, False)]


@pytest.mark.parametrize("name,omega,imag", HK_CASES)
def test_very_short_wavelength_matrix(ps, name, omega, imag, tmp_path):
    geo = os.path.join(GEODIR, name + ".scuffgeo")
    cache = str(tmp_path / "ref.cache")
    Mref, info = ref_matrix(geo, omega, imag=imag, cache_out=cache)
    ps.ClearCache()            # pin the FIPPI table to THIS reference run's records (the reference is not reproducible
    ps.PreloadCache(cache)     # run to run below ~1e-10: racing threads store only one of their records per key)
    G = ps.RWGGeometry(geo)
    M = G.AssembleBEMMatrix(omega, M=G.AllocateBEMMatrix(imag)).array
    counts = G.ppi_counts()
    errs = block_errors(G, M, Mref)
    print(name, omega, "block errs", {k: float("%.2e" % v) for k, v in errs.items()}, "counts", counts, "timing", ps.last_timing())
    assert counts[3] > 0                             # high-k Taylor-Duffy was exercised
    assert max(errs.values()) <= TOL


def test_disable_cache_forces_taylor_duffy(ps, monkeypatch):
    """RWGGeometry::DisableCache (env SCUFF_DISABLE_CACHE in the reference, RWGGeometry.cc:402)"""
    geo = os.path.join(GEODIR, "PECSphere_255.scuffgeo")
    omega = 1.0
    Mref, _ = ref_matrix(geo, omega, env={"SCUFF_DISABLE_CACHE": "1"})
    monkeypatch.setenv("SCUFF_DISABLE_CACHE", "1")       # read by the RWGGeometry constructor, as in the reference
    try:
        G = ps.RWGGeometry(geo)
        M = G.AssembleBEMMatrix(omega).array
        counts = G.ppi_counts()
    finally:
        ps.host_lib().scuffb200_set_disable_cache(0)
    err = np.linalg.norm(M - Mref) / np.linalg.norm(Mref)
    print("DisableCache rel err", err, "counts", counts)
    assert counts[2] > 1428          # more than the common-triangle pairs went through Taylor-Duffy
    assert err <= TOL
