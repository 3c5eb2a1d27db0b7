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


@pytest.mark.parametrize("name,omega,K", [("PECSphere_501", 1.0, 3), ("SiO2Sphere_501", 0.7, 4), ("PECSpheres_255", 0.3, 5),
                                          ("SiSphere_255", 0.25j, 2)])
def test_pipelined_copy_back(ps, monkeypatch, name, omega, K):
    """AssembleBEMMatrix's copy-back overlapped with assembly (K column blocks, trapezoid downloads on the copy stream,
    scuffcuda_matrix_download_region_async) returns the same host matrix as the one-pass assembly + one download."""
    G = ps.RWGGeometry(os.path.join(GEODIR, name + ".scuffgeo"))
    imag = np.real(omega) == 0
    monkeypatch.setenv("SCUFF_CUDA_PIPELINE_BLOCKS", "1")
    M1 = G.AllocateBEMMatrix(imag)
    G.AssembleBEMMatrix(omega, M=M1); A1 = M1.array.copy(); c1 = G.ppi_counts()
    monkeypatch.setenv("SCUFF_CUDA_PIPELINE_BLOCKS", str(K))
    # "upper" (default): one-pass assembly, only the upper triangle crosses PCIe in 2K chunks, host threads write the mirror image;
    # "columns": assembly in K column blocks overlapped with full-matrix downloads
    for mode in ("upper", "columns"):
        monkeypatch.setenv("SCUFF_CUDA_PIPELINE_MODE", mode)
        for threads in ((1, 5, -5) if mode == "upper" else (0,)):
            if threads:
                monkeypatch.setenv("SCUFF_CUDA_MIRROR_THREADS", str(abs(threads)))
                monkeypatch.setenv("SCUFF_CUDA_MIRROR_DMA", "0" if threads < 0 else "1")   # host threads only / host threads + copy engine
            M2 = G.AllocateBEMMatrix(imag)
            d0 = ps.transfer_bytes()[1]
            G.AssembleBEMMatrix(omega, M=M2); A2 = M2.array.copy(); c2 = G.ppi_counts()
            d2h = ps.transfer_bytes()[1] - d0
            assert A1.dtype == A2.dtype and (A1.dtype == np.float64) == bool(imag)
            err = np.linalg.norm(A2 - A1) / np.linalg.norm(A1)
            print(name, omega, K, mode, threads, "pipelined vs one-pass rel err", err, "D2H bytes", d2h, "of", A2.nbytes)
            assert err <= 1e-14
            assert np.array_equal(A2, A2.T)                        # mirrored exactly
            # the HBM copy is the same matrix (LU consumers read it)
            D = M2.device_tensor().cpu().numpy().view(A2.dtype).reshape(A2.shape[1], A2.shape[0]).T
            assert np.array_equal(D, A2)
            assert all(b >= a for a, b in zip(c1, c2))            # pairs straddling a block boundary are visited by both blocks
            if mode == "upper":
                assert list(c1) == list(c2)
                if threads < 0:
                    assert d2h < 0.7 * A2.nbytes                   # host-only mirror: about half the matrix crosses PCIe
