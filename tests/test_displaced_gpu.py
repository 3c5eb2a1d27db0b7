"""Displaced surface-surface blocks (GetSSIArgStruct::Displacement, SurfaceSurfaceInteractions.cc:426-585): surface b is
shifted by a vector L before the panel pairs are assessed and integrated.  These are the neighbour-cell blocks of the
periodic assembly (AssembleBEMMatrix.cc:72-114); parity against the reference's GetSurfaceSurfaceInteractions through
scuffcuda_assemble_block with use_displacement, tolerance 1e-10 relative Frobenius per block."""
import ctypes as C
import os
import numpy as np
import pytest
from conftest import GEODIR, ref_ssi, rel_fro

pytestmark = pytest.mark.gpu

CASES = [
    # geometry, omega, nsa, nsb, L, (eps_a, mu_a, sign_a), (eps_b, mu_b, sign_b) or None, reference threads
    ("PECSphere_255.scuffgeo", 0.5, 0, 0, (2.05, 0.1, 0.0), (1.0, 1.0, 1.0), None, None),       # near pairs without common vertices
    ("PECSphere_255.scuffgeo", 0.5 + 0.3j, 0, 0, (0.0, 3.5, 1.0), (1.0, 1.0, 1.0), None, None),  # far pairs only, complex frequency
    ("OpenPlate_40.scuffgeo", 1.0, 0, 0, (1.0, 0.0, 0.0), (1.0, 1.0, 1.0), None, "1"),           # panels share edges/vertices across the cell boundary
    ("OpenPlate_40.scuffgeo", 0.7j, 0, 0, (1.0, 1.0, 0.0), (1.0, 1.0, 1.0), None, "1"),          # corner contact (one common vertex), imaginary frequency
    ("PECSpheres_255.scuffgeo", 0.8, 0, 1, (0.5, 0.0, -0.4), (1.0, 1.0, 1.0), None, None),       # two different surfaces, displaced
]


@pytest.mark.parametrize("geo,omega,nsa,nsb,Lvec,ra,rb,threads", CASES)
def test_displaced_block(ps, tmp_path, geo, omega, nsa, nsb, Lvec, ra, rb, threads):
    path = os.path.join(GEODIR, geo)
    cache = str(tmp_path / "ssi.cache")
    env = {"OMP_NUM_THREADS": threads} if threads else None
    ref, info = ref_ssi(path, omega, nsa, nsb, Lvec, cache_out=cache, env=env)
    ps.ClearCache(); ps.PreloadCache(cache)

    L, ctx = ps.cuda_lib(), ps.context()
    G = ps.RWGGeometry(path)
    G.AssembleBEMMatrix(0.1, to_host=False)                     # puts the geometry on the device
    a = ps.SSIArgs()
    a.omega[0], a.omega[1] = float(np.real(omega)), float(np.imag(omega))
    a.eps_a[0], a.mu_a[0], a.sign_a = ra
    if rb:
        a.eps_b[0], a.mu_b[0], a.sign_b = rb
    a.symmetric = 0
    a.use_displacement = 1
    for i in range(3):
        a.displacement[i] = Lvec[i]
    h = C.c_int()
    assert L.scuffcuda_matrix_create(ctx, info["NR"], info["NC"], 0, C.byref(h)) == 0
    assert L.scuffcuda_assemble_block(ctx, nsa, nsb, C.byref(a), h.value) == 0, L.scuffcuda_last_error(ctx)
    mine = np.zeros((info["NR"], info["NC"]), dtype=np.complex128, order="F")
    assert L.scuffcuda_matrix_download(ctx, h.value, mine.ctypes.data, info["NR"]) == 0
    L.scuffcuda_matrix_destroy(ctx, h.value)
    assert np.linalg.norm(ref) > 0
    err = rel_fro(mine, ref)
    print(geo, omega, Lvec, "rel_fro", err)
    assert err <= 1e-10


def test_displaced_block_builds_its_own_table(ps):
    """Same as above without a preloaded FIPPI table: displaced near pairs get their records from the device builder."""
    path = os.path.join(GEODIR, "PECSphere_255.scuffgeo")
    Lvec = (2.05, 0.1, 0.0)
    ref, info = ref_ssi(path, 0.5, 0, 0, Lvec)
    ps.ClearCache()
    L, ctx = ps.cuda_lib(), ps.context()
    G = ps.RWGGeometry(path)
    G.AssembleBEMMatrix(0.1, to_host=False)
    a = ps.SSIArgs()
    a.omega[0] = 0.5
    a.eps_a[0], a.mu_a[0], a.sign_a = 1.0, 1.0, 1.0
    a.use_displacement = 1
    for i in range(3):
        a.displacement[i] = Lvec[i]
    h = C.c_int()
    assert L.scuffcuda_matrix_create(ctx, info["NR"], info["NC"], 0, C.byref(h)) == 0
    assert L.scuffcuda_assemble_block(ctx, 0, 0, C.byref(a), h.value) == 0, L.scuffcuda_last_error(ctx)
    mine = np.zeros((info["NR"], info["NC"]), dtype=np.complex128, order="F")
    assert L.scuffcuda_matrix_download(ctx, h.value, mine.ctypes.data, info["NR"]) == 0
    L.scuffcuda_matrix_destroy(ctx, h.value)
    assert rel_fro(mine, ref) <= 1e-10
