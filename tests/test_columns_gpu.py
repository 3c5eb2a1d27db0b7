"""Column-block assembly (one large matrix across GPUs): the union of the column blocks, mirrored once,
equals the ordinary AssembleBEMMatrix result.  Single-GPU check of the device logic (pair pruning, column filter, item
filter); the exchange itself is covered by the gloo test and by bench.py --gpus N."""
import importlib.util
import os
import numpy as np
import pytest
from conftest import GEODIR, ROOT

pytestmark = pytest.mark.gpu


def _sweep():
    spec = importlib.util.spec_from_file_location("sweep", os.path.join(ROOT, "scuff-em_b200", "sweep.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("name,omega,world", [("PECSphere_501", 1.0, 3), ("SiO2Sphere_501", 0.7, 4), ("PECSpheres_255", 0.3, 2)])
def test_column_blocks_reassemble(ps, name, omega, world):
    G = ps.RWGGeometry(os.path.join(GEODIR, name + ".scuffgeo"))
    Mfull = G.AssembleBEMMatrix(omega).array
    N = G.TotalBFs
    U = np.zeros((N, N), dtype=complex)
    for (c0, c1) in _sweep().column_ranges(N, world):
        M = G.AllocateBEMMatrix()
        G.AssembleBEMMatrixColumns(omega, M, c0, c1)
        A = M.array
        assert np.all(A[:, :c0] == 0) and np.all(A[:, c1:] == 0)       # nothing outside the owned columns
        U[:, c0:c1] = A[:, c0:c1]
    assert np.all(np.tril(U, -1) == 0)                                   # upper triangle only, not mirrored
    S = np.triu(U) + np.triu(U, 1).T
    err = np.linalg.norm(S - Mfull) / np.linalg.norm(Mfull)
    print(name, omega, world, "rel err", err)
    assert err <= 1e-14
