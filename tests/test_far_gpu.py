"""Far-field kernel parity: entries of the BEM matrix whose four panel pairs are all far (rRel>4,
the reference's PPIALG_LOCUBATURE branch) must match the UNMODIFIED reference to <=1e-12."""
import os
import numpy as np
import pytest
from conftest import GEODIR, ref_matrix, rel_fro

pytestmark = pytest.mark.gpu


def far_edge_mask(s):
    """edge pairs all of whose (up to 4) panel pairs have rRel > 4 (AssessPanelPair.cc:161-166)"""
    PC, PR, ET = s["PC"], s["PR"], s["ET"]
    D = PC[:, None, :] - PC[None, :, :]
    r2 = (D[..., 0] * D[..., 0] + D[..., 1] * D[..., 1]) + D[..., 2] * D[..., 2]
    rrel = np.sqrt(r2) / np.maximum(PR[:, None], PR[None, :])
    far = rrel > 4.0
    P, M = ET[:, 4], ET[:, 5]
    mask = far[np.ix_(P, P)] & far[np.ix_(P, M)] & far[np.ix_(M, P)] & far[np.ix_(M, M)]
    return mask


@pytest.mark.parametrize("name,omega", [("PECSphere_501", 1.0), ("PECSphere_501", 0.1), ("PECSphere_255", 1.0j),
                                        ("PECSphere_501", 0.7 + 0.3j)])
def test_far_entries_pec(ps, name, omega, monkeypatch):
    geo = os.path.join(GEODIR, name + ".scuffgeo")
    Mref, _ = ref_matrix(geo, omega)
    monkeypatch.setenv("SCUFF_CUDA_FAR_ONLY", "1")
    G = ps.RWGGeometry(geo)
    M = G.AssembleBEMMatrix(omega).array
    mask = far_edge_mask(G.surface(0))
    assert mask.sum() > 0.5 * mask.size
    err = np.linalg.norm((M - Mref)[mask]) / np.linalg.norm(Mref[mask])
    print(name, omega, "far entries", mask.sum(), "rel err", err, "counts", G.ppi_counts(), "timing", ps.last_timing())
    assert err <= 1e-12
    assert np.allclose(M, M.T, rtol=0, atol=0)


def test_far_entries_dielectric(ps, monkeypatch):
    geo = os.path.join(GEODIR, "SiO2Sphere_501.scuffgeo")
    omega = 1.0
    Mref, _ = ref_matrix(geo, omega)
    monkeypatch.setenv("SCUFF_CUDA_FAR_ONLY", "1")
    G = ps.RWGGeometry(geo)
    M = G.AssembleBEMMatrix(omega).array
    mask = np.kron(far_edge_mask(G.surface(0)), np.ones((2, 2), dtype=bool))
    err = np.linalg.norm((M - Mref)[mask]) / np.linalg.norm(Mref[mask])
    print("SiO2 far rel err", err, ps.last_timing())
    assert err <= 1e-12


def test_fp64_peak(ps):
    tf = ps.measure_fp64_peak()
    print("FP64 DFMA peak TFLOP/s", tf)
    assert tf > 5.0
