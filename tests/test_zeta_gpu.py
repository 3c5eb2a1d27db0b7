"""Surface-impedance term of GetSurfaceSurfaceInteractions (AddSurfaceZetaContributionToBEMMatrix,
SurfaceSurfaceInteractions.cc:339-421): a PEC surface with `SURFACE_IMPEDANCE Zeta(w,x,y,z)` adds -Zeta * overlap(f_a, f_b) to
the entries of basis functions that share a panel.  Parity of the full matrix against the reference (<= 1e-10 per block; the
term changes the matrix by ~9 % in these fixtures), in one pass, in column blocks and through the pipelined copy-back."""
import os
import numpy as np
import pytest
from conftest import GEODIR, ref_matrix
from test_full_gpu import block_errors

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,omega,imag", [("ZetaSphere_255", 0.7, False), ("ZetaSphere_255", 0.4j, True),
                                              ("ZetaSpheres_255", 0.25, False)])
def test_surface_impedance(ps, monkeypatch, name, omega, imag, tmp_path):
    geo = os.path.join(GEODIR, name + ".scuffgeo")
    cache = str(tmp_path / "ref.cache")
    Mref, info = ref_matrix(geo, omega, imag=imag, cache_out=cache)
    Mpec, _ = ref_matrix(os.path.join(GEODIR, name.replace("Zeta", "PEC") + ".scuffgeo"), omega, imag=imag)
    assert np.linalg.norm(Mref - Mpec) > 1e-3 * np.linalg.norm(Mpec)          # the fixture really exercises the term
    ps.ClearCache(); ps.PreloadCache(cache)
    G = ps.RWGGeometry(geo)
    for K in ("1", "3"):                                                       # one pass; column blocks + overlapped copy-back
        monkeypatch.setenv("SCUFF_CUDA_PIPELINE_BLOCKS", K)
        Mh = G.AllocateBEMMatrix(imag)
        assert Mh.is_real == bool(info["real"])
        M = G.AssembleBEMMatrix(omega, M=Mh).array
        errs = block_errors(G, M, Mref)
        print(name, omega, "K", K, "block errs", {k: float("%.2e" % v) for k, v in errs.items()})
        assert max(errs.values()) <= 1e-10
        assert np.array_equal(M, M.T)
