"""Transformed geometries (the scuff-cas3D workflow: RWGGeometry::Transform from a .trans file, assemble, UnTransform): the
moved surface is re-sent to the device, pair plans of the old position are dropped, the off-diagonal block is re-assembled;
compared with the UNMODIFIED reference applying the same transformation."""
import os
import numpy as np
import pytest
from conftest import GEODIR, ref_matrix
from test_full_gpu import block_errors

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,omega,imag,indices", [("PECSpheres_327", 0.8j, True, (0, 3, 7)), ("E10Spheres_327", 0.5j, False, (2,))])
def test_transformed_geometry(ps, name, omega, imag, indices, tmp_path):
    geo = os.path.join(GEODIR, name + ".scuffgeo")
    transfile = os.path.join(GEODIR, "Spheres.trans")
    lines = [l.strip() for l in open(transfile) if l.strip() and not l.startswith("#")]
    G = ps.RWGGeometry(geo)
    M = G.AllocateBEMMatrix(imag)
    for n in indices:
        cache = str(tmp_path / ("ref%d.cache" % n))
        Mref, _ = ref_matrix(geo, omega, imag=imag, cache_out=cache, trans=("Spheres.trans", n))
        ps.ClearCache()            # pin the FIPPI table to THIS reference run's records (the reference is not reproducible
        ps.PreloadCache(cache)     # run to run below ~1e-10: racing threads store only one of their records per key)
        G.transform(lines[n])
        A = G.AssembleBEMMatrix(omega, M=M).array
        G.untransform()
        errs = block_errors(G, A, Mref)
        print(name, lines[n], {k: float("%.2e" % v) for k, v in errs.items()})
        assert max(errs.values()) <= 1e-10
    # back at the original position
    cache0 = str(tmp_path / "ref_orig.cache")
    Mref0, _ = ref_matrix(geo, omega, imag=imag, cache_out=cache0)
    ps.ClearCache()
    ps.PreloadCache(cache0)
    A0 = G.AssembleBEMMatrix(omega, M=M).array
    assert max(block_errors(G, A0, Mref0).values()) <= 1e-10
