"""FIPPI table built once by all ranks of a sweep (scuffcuda_fippi_set_partition / _prepare_block / _host_records / _commit,
sweep.share_fippi_table): on one GPU, the shares of a 3-rank partition are built one after the other; they are disjoint, their
sum is bit-identical to the table one rank builds alone, and a matrix assembled from the summed table equals the usual one."""
import os
import numpy as np
import pytest
from conftest import GEODIR

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,omega", [("PECSphere_501", 0.9), ("SiSpheres_255", 0.3)])
def test_partitioned_build_equals_full_build(ps, name, omega):
    geo = os.path.join(GEODIR, name + ".scuffgeo")
    ps.ClearCache()
    G = ps.RWGGeometry(geo)
    G.PrepareNearPlans()
    full = ps.fippi_host_records().copy()
    Mfull = G.AssembleBEMMatrix(omega).array.copy()
    assert full.shape[0] > 1000 and np.all(np.abs(full).sum(axis=1) > 0)
    world = 3
    shares = []
    for rank in range(world):
        ps.ClearCache()                                   # drops the plans and the table
        ps.fippi_set_partition(rank, world)
        G.PrepareNearPlans()
        ps.fippi_set_partition(0, 1)
        shares.append(ps.fippi_host_records().copy())
        assert shares[-1].shape == full.shape
        mine = np.arange(full.shape[0]) % world == rank
        assert np.all(shares[-1][~mine] == 0.0) and np.all(np.abs(shares[-1][mine]).sum(axis=1) > 0)
    total = shares[0] + shares[1] + shares[2]
    assert np.array_equal(total, full)                    # same records, bit for bit
    # the last context state holds rank 2's share only: complete it in place and commit, as sweep.share_fippi_table does
    rec = ps.fippi_host_records()
    rec[...] = total
    ps.fippi_commit()
    M = G.AssembleBEMMatrix(omega).array
    err = np.linalg.norm(M - Mfull) / np.linalg.norm(Mfull)
    print(name, "records", full.shape[0], "matrix from the summed table vs full build", err)
    assert err <= 1e-14
    ps.ClearCache()
