"""Near-field (desingularised cubature + FIPPI) parity.  The FIPPI table is preloaded from the cache file the
UNMODIFIED reference writes (StoreCache), so this isolates the frequency-dependent near kernel: every entry
whose panel pairs are far or desingularised (no common-triangle pair) must match the reference to <=1e-12."""
import os
import numpy as np
import pytest
from conftest import GEODIR, ref_matrix

pytestmark = pytest.mark.gpu


def no_self_panel_mask(s):
    ET = s["ET"]
    P, M = ET[:, 4], ET[:, 5]
    share = (P[:, None] == P[None, :]) | (P[:, None] == M[None, :]) | (M[:, None] == P[None, :]) | (M[:, None] == M[None, :])
    return ~share


@pytest.mark.parametrize("name,omega", [("PECSphere_501", 1.0), ("PECSphere_501", 0.05 + 0.4j), ("PECSphere_255", 0.7j)])
def test_desing_entries(ps, name, omega, tmp_path):
    geo = os.path.join(GEODIR, name + ".scuffgeo")
    cache = str(tmp_path / "ref.cache")
    Mref, _ = ref_matrix(geo, omega, cache_out=cache)
    ps.PreloadCache(cache)
    G = ps.RWGGeometry(geo)
    M = G.AssembleBEMMatrix(omega).array
    mask = no_self_panel_mask(G.surface(0))
    err = np.linalg.norm((M - Mref)[mask]) / np.linalg.norm(Mref[mask])
    worst = np.abs((M - Mref) * mask).max()
    print(name, omega, "entries", mask.sum(), "rel err", err, "max abs", worst, "counts", G.ppi_counts(), "timing", ps.last_timing())
    assert err <= 1e-12
