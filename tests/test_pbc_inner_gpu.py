"""Periodic geometries, step 1 of AssembleBEMMatrixBlock's PBC branch (AssembleBEMMatrix.cc:465-519): the blocks of the
innermost lattice cells (displaced-surface assembly, region omission by RegionIsExtended, straddler edges) stamped with
their Bloch phase factors (StampInNeighborBlock, :72-114).  The checker is the reference's full periodic block minus the
outer-cell contribution it computes from its GBar tables (oracle/ref_driver.cc pbcinner), tolerance 1e-10 relative
Frobenius.  The outer-cell (Ewald) step is checked in tests/test_pbc_outer_gpu.py."""
import os
import numpy as np
import pytest
from conftest import GEODIR, ref_pbcinner, rel_fro

pytestmark = pytest.mark.gpu

CASES = [
    ("PECPlate_40.scuffgeo", 1.0, (0.3, 0.2), 0, 0),            # PEC sheet filling the unit cell: straddlers on both lattice directions
    ("PECPlate_40.scuffgeo", 0.8j, (0.0, 0.0), 0, 0),           # imaginary frequency, kBloch = 0 (all phases 1)
    ("SiSlab_40.scuffgeo", 1.0, (0.3, 0.2), 0, 0),              # dielectric slab: both regions extended
    ("SiSlab_40.scuffgeo", 1.0, (0.3, 0.2), 0, 1),              # off-diagonal block: all nine cells stamped directly
    ("SiSlab_40.scuffgeo", 0.4 + 0.1j, (-0.5, 1.1), 1, 1),
    ("SphereSlabArray.scuffgeo", 0.7, (0.2, -0.4), 0, 0),
    ("SphereSlabArray.scuffgeo", 0.7, (0.2, -0.4), 0, 1),
    ("SphereSlabArray.scuffgeo", 0.7, (0.2, -0.4), 0, 2),       # no common region (the sphere sits in the default exterior medium): zero block
    ("SphereSlabArray.scuffgeo", 0.7, (0.2, -0.4), 2, 2),       # compact sphere inside the cell: its interior region is not extended
]


@pytest.mark.parametrize("geo,omega,kb,nsa,nsb", CASES)
def test_inner_cell_blocks(ps, tmp_path, geo, omega, kb, nsa, nsb):
    path = os.path.join(GEODIR, geo)
    cache = str(tmp_path / "pbc.cache")
    # structured meshes: the multi-threaded reference is not reproducible at the 1e-10 level
    inner, full, info = ref_pbcinner(path, omega, kb, nsa, nsb, cache_out=cache, env={"OMP_NUM_THREADS": "1"})
    ps.ClearCache(); ps.PreloadCache(cache)
    G = ps.RWGGeometry(path)
    assert nsa < G.NumSurfaces and nsb < G.NumSurfaces
    M = ps.HMatrix.new(info["NR"], info["NC"], False)
    G.AssembleInnerCellBlocks(nsa, nsb, omega, kb, M)
    mine = M.array
    if np.linalg.norm(full) == 0:
        assert np.all(mine == 0)
        return
    err = rel_fro(mine, inner)
    print(geo, omega, kb, (nsa, nsb), "inner-cell rel_fro", err, "| outer/full", info["norm_outer"] / info["norm_full"])
    assert err <= 1e-10
