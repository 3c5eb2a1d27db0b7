"""Error behaviour at the C ABI and through the reference-facing API: non-zero status + message, nothing silently ignored."""
import ctypes as C
import os
import numpy as np
import pytest
from conftest import GEODIR

pytestmark = pytest.mark.gpu


def _args(ps, **kw):
    a = ps.SSIArgs()
    a.omega[0] = 1.0
    a.eps_a[0] = 1.0; a.mu_a[0] = 1.0; a.sign_a = 1.0
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def test_c_abi_argument_errors(ps):
    L, ctx = ps.cuda_lib(), ps.context()
    G = ps.RWGGeometry(os.path.join(GEODIR, "PECSpheres_255.scuffgeo"))
    M = G.AssembleBEMMatrix(0.1, to_host=False)          # geometry and a matrix now live on the device
    h = M.device_handle
    err = lambda: L.scuffcuda_last_error(ctx).decode()
    assert L.scuffcuda_assemble_block(ctx, 0, 5, C.byref(_args(ps)), h) != 0 and "surface index" in err()
    assert L.scuffcuda_assemble_block(ctx, 0, 1, C.byref(_args(ps)), 9999) != 0 and "handle" in err()
    assert L.scuffcuda_assemble_block(ctx, 0, 1, C.byref(_args(ps, symmetric=1)), h) != 0 and "symmetric" in err()
    assert L.scuffcuda_assemble_block(ctx, 0, 1, C.byref(_args(ps, row_offset=G.TotalBFs)), h) != 0 and "exceeds" in err()
    assert L.scuffcuda_matrix_copy_block(ctx, h, 0, 0, 10 ** 6, 1, 0, 0) != 0 and "out of range" in err()
    assert L.scuffcuda_matrix_symmetrize(ctx, h, 5, G.TotalBFs) != 0
    hm = C.c_int()
    assert L.scuffcuda_matrix_create(ctx, 0, 4, 0, C.byref(hm)) != 0
    # a frequency so high that the FP64 phase reduction of the far kernel would leave its exact range is refused
    a = _args(ps); a.omega[0] = 1.0e6
    assert L.scuffcuda_assemble_block(ctx, 0, 1, C.byref(a), h) != 0 and "exceeds the supported range" in err()
    # LU of a rectangular matrix / solve before factorisation
    assert L.scuffcuda_matrix_create(ctx, 4, 3, 0, C.byref(hm)) == 0
    assert L.scuffcuda_lu_factor(ctx, hm.value) != 0 and "square" in err()
    rhs = np.zeros(8)
    assert L.scuffcuda_lu_solve(ctx, h, rhs.ctypes.data, 1) != 0 and "factored" in err()
    L.scuffcuda_matrix_destroy(ctx, hm.value)


def test_reference_api_errors(ps):
    G = ps.RWGGeometry(os.path.join(GEODIR, "PECSphere_255.scuffgeo"))
    with pytest.raises(ps.ScuffError, match="Bloch"):
        G.AssembleBEMMatrix(1.0, kBloch=(0.1, 0.2))                    # AssembleBEMMatrix.cc:692-695
    P = ps.RWGGeometry(os.path.join(GEODIR, "SiSlab_40.scuffgeo"))     # periodic: detected and reported, not ignored
    with pytest.raises(ps.ScuffError):
        P.AssembleBEMMatrix(1.0)
    # (a periodic matrix with kBloch given assembles: tests/test_pbc_outer_gpu.py)
    # omega = 0: k = 0 -> GC = 0 (EdgeEdgeInteractions.cc:74-79): the matrix is exactly zero
    Z = G.AssembleBEMMatrix(0.0).array
    assert np.all(Z == 0)


def test_region_linearity_and_accumulate(ps):
    """Linearity the domain offers: the PMCHWT matrix is the sum of its exterior-region and interior-region parts
    (MatProp::Zero() on one region at a time, as scuff-neq / scuff-cas3D use it; SurfaceSurfaceInteractions.cc:177,294), and
    Args->Accumulate adds a block on top of what the matrix holds."""
    G = ps.RWGGeometry(os.path.join(GEODIR, "SiSphere_255.scuffgeo"))
    omega = 0.6
    full = G.AssembleBEMMatrix(omega).array
    G.zero_region(1, True)
    ext = G.AssembleBEMMatrix(omega).array
    G.zero_region(1, False); G.zero_region(0, True)
    inn = G.AssembleBEMMatrix(omega).array
    G.zero_region(0, False)
    assert np.linalg.norm(ext) > 0 and np.linalg.norm(inn) > 0
    assert np.linalg.norm(ext + inn - full) <= 1e-13 * np.linalg.norm(full)
    # accumulate through the C ABI: the same block added twice
    L, ctx = ps.cuda_lib(), ps.context()
    M = G.AssembleBEMMatrix(omega, to_host=False)
    a = _args(ps, symmetric=1, accumulate=1)
    a.omega[0] = omega
    assert L.scuffcuda_assemble_block(ctx, 0, 0, C.byref(a), M.device_handle) == 0, L.scuffcuda_last_error(ctx)
    host = np.zeros((G.TotalBFs, G.TotalBFs), dtype=np.complex128, order="F")
    assert L.scuffcuda_matrix_download(ctx, M.device_handle, host.ctypes.data, G.TotalBFs) == 0
    assert np.linalg.norm(host - (full + ext)) <= 1e-13 * np.linalg.norm(full)
