"""HMatrix host/device coherence (include/libscuff_b200.h): the reference applications mix AssembleBEMMatrix[Block] with
host-side edits (InsertBlock of T/U blocks, AddEntry, Zero) and then LUFactorize/LUSolve, repeatedly on the SAME matrix
(scuff-cas3D CasimirIntegrand.cc:159-196, scuff-neq).  The device twin must see every host edit and the host array every
device result; nothing may factorize stale data."""
import os
import numpy as np
import pytest
from conftest import GEODIR

pytestmark = pytest.mark.gpu


def _solve(M, rhs):
    X = type(M).new(M.NR, 1, M.is_real)
    for i, v in enumerate(rhs):
        X.SetEntry(i, 0, v)
    M.lu_solve(X)
    return X.array[:, 0]


def test_edit_then_factorize_twice(ps):
    G = ps.RWGGeometry(os.path.join(GEODIR, "PECSphere_255.scuffgeo"))
    N = G.TotalBFs
    rng = np.random.default_rng(0)
    rhs = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    M = G.AllocateBEMMatrix()
    for it in range(2):                        # the second pass must not factorize the first pass's LU factors
        G.AssembleBEMMatrix(0.7 + 0.1 * it, M=M)
        A = M.array.copy()
        M.AddEntry(3, 5, 0.25 - 0.5j)          # host edit after the device assembly
        M.SetEntry(7, 7, M.GetEntry(7, 7) * 2.0)
        A[3, 5] += 0.25 - 0.5j; A[7, 7] *= 2.0
        assert M.state()["DeviceStale"]
        M.lu_factorize()
        x = _solve(M, rhs)
        assert np.linalg.norm(A @ x - rhs) <= 1e-9 * np.linalg.norm(rhs), it


def test_insert_block_on_device_and_host(ps):
    """T/U-block stamping as scuff-cas3D does it: blocks assembled on the device are inserted into the big matrix without
    touching the host; a block edited on the host takes the host path; both give the same matrix"""
    G = ps.RWGGeometry(os.path.join(GEODIR, "PECSpheres_255.scuffgeo"))
    N = G.TotalBFs
    s0, s1 = G.surface(0), G.surface(1)
    T0 = ps.HMatrix.new(s0["NBF"], s0["NBF"]); U = ps.HMatrix.new(s0["NBF"], s1["NBF"])
    G.AssembleBEMMatrixBlock(0, 0, 0.9, T0)
    G.AssembleBEMMatrixBlock(0, 1, 0.9, U)
    full = G.AssembleBEMMatrix(0.9).array
    M = ps.HMatrix.new(N, N)
    M.InsertBlock(T0, 0, 0)
    M.InsertBlock(T0, s1["BFOffset"], s1["BFOffset"])           # the mate's diagonal block
    M.InsertBlock(U, 0, s1["BFOffset"])
    Ut = ps.HMatrix.new(s1["NBF"], s0["NBF"])
    Ut.host_view()[:, :] = U.array                              # transpose written straight into the host array ...
    Ut.MarkHostDirty()                                          # ... and announced
    M.InsertBlock(Ut, s1["BFOffset"], 0)
    assert np.linalg.norm(M.array - full) <= 1e-12 * np.linalg.norm(full)
    # factorize, solve, then overwrite a block again and re-factorize: the factors must not leak into the second system
    rng = np.random.default_rng(1)
    rhs = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    M.lu_factorize()
    x = _solve(M, rhs)
    assert np.linalg.norm(full @ x - rhs) <= 1e-9 * np.linalg.norm(rhs)
    M.InsertBlock(T0, 0, 0); M.InsertBlock(T0, s1["BFOffset"], s1["BFOffset"]); M.InsertBlock(U, 0, s1["BFOffset"]); M.InsertBlock(Ut, s1["BFOffset"], 0)
    M.ZeroBlock(0, s0["NBF"], s1["BFOffset"], s1["NBF"])        # drop the coupling block on the host
    M.ZeroBlock(s1["BFOffset"], s1["NBF"], 0, s0["NBF"])
    A2 = full.copy(); A2[:s0["NBF"], s1["BFOffset"]:] = 0; A2[s1["BFOffset"]:, :s0["NBF"]] = 0
    assert np.linalg.norm(M.array - A2) <= 1e-12 * np.linalg.norm(A2)
    M.lu_factorize()
    x2 = _solve(M, rhs)
    assert np.linalg.norm(A2 @ x2 - rhs) <= 1e-9 * np.linalg.norm(rhs)


def test_block_assembly_keeps_host_edits(ps):
    """AssembleBEMMatrixBlock into a matrix whose other blocks were written on the host must not overwrite them with stale
    device data"""
    G = ps.RWGGeometry(os.path.join(GEODIR, "PECSpheres_255.scuffgeo"))
    N = G.TotalBFs; s1 = G.surface(1)
    M = ps.HMatrix.new(N, N)
    G.AssembleBEMMatrixBlock(0, 0, 0.9, M)                       # creates the device twin
    M.SetEntry(N - 1, N - 1, 42.0)                               # host edit outside the block
    G.AssembleBEMMatrixBlock(0, 1, 0.9, M, ColOffset=s1["BFOffset"])
    assert M.GetEntry(N - 1, N - 1) == 42.0
    M.Zero()
    assert np.all(M.array == 0)
    G.AssembleBEMMatrix(0.9, M=M)
    M.lu_factorize()
    M.SetEntry(0, 0, 1.0)                                        # edit after the factorization: a solve would use stale factors
    with pytest.raises(ps.ScuffError, match="modified on the host"):
        _solve(M, np.ones(N))
