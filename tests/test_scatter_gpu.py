"""Right-hand side, solve and power/force/torque on the device against the unmodified reference's
scuff-scatter flow (oracle/_ref/ref_driver scatter: AssembleBEMMatrix -> AssembleRHSVector -> LUFactorize/LUSolve ->
GetPFTMatrix with the energy/momentum-transfer method, cubature integrals SCUFF_EMTPFTI_EHDERIVATIVES1).

Tolerances: RHS entries 1e-12 relative to the largest entry (78-point rule, same points, same order of vertices);
solution vector 1e-8 (two LU factorisations of a matrix with condition number ~1e4); power/force/torque 1e-9 relative to
the largest entry of its group when both sides start from the SAME solution vector — the reference's same-surface kernels
Im Psi = (kr cos kr - sin kr)/(4 pi r^3), Im Zeta cancel catastrophically for close points (EMTPFT.cc:113-117), which
amplifies last-bit differences of sin/cos."""
import os

import numpy as np
import pytest

from conftest import GEODIR, ref_scatter

pytestmark = pytest.mark.gpu

PW_X = ((1.0, 0.0, 0.0), (0.0, 0.0, 1.0))                 # E0 = x, propagating along +z (the Mie test's field)
PW_OBL = ((0.3 + 0.1j, -0.2j, 0.1), (0.6, 0.0, 0.8))       # complex polarisation, oblique (E0 need not be transverse)


def geo(name):
    return os.path.join(GEODIR, name)


def rel(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300)


def pft_err(P, Pref):
    """worst relative error per group (powers, force, torque), relative to the group's largest reference entry"""
    worst = 0.0
    for cols in ((0, 1), (2, 3, 4), (5, 6, 7)):
        scale = np.max(np.abs(Pref[:, cols]))
        if scale > 0:
            worst = max(worst, np.max(np.abs(P[:, cols] - Pref[:, cols])) / scale)
    return worst


CASES = [("PECSphere_501.scuffgeo", 1.0, PW_X, None),
         ("SiO2Sphere_501.scuffgeo", 1.0, PW_X, None),
         ("SiO2Sphere_501.scuffgeo", 0.1, PW_OBL, None),
         ("PECSpheres_327.scuffgeo", 0.7, PW_OBL, None),
         ("E10Spheres_327.scuffgeo", 1.0, PW_X, ("Spheres.trans", 2)),
         ("GoldSphere_501.scuffgeo", 2.0, PW_X, None),
         ("NestedSpheres.scuffgeo", 0.9, PW_OBL, None),        # a body inside a body: the "A contained in B" branches (EMTPFT.cc:655-660)
         ("OpenPlate_40.scuffgeo", 1.0, PW_OBL, None)]         # open PEC surface: exterior edges carry no basis function


@pytest.mark.parametrize("name,omega,pw,trans", CASES)
def test_rhs_solve_pft_planewave(ps, name, omega, pw, trans):
    RHSr, KNr, PFTr, info = ref_scatter(geo(name), omega, pw, trans=trans)
    G = ps.RWGGeometry(geo(name))
    if trans is not None:
        line = open(os.path.join(GEODIR, trans[0])).read().splitlines()[trans[1]]
        G.transform(line)
    IF = ps.PlaneWave(pw[0], pw[1])
    M = G.AssembleBEMMatrix(omega, to_host=False)
    RHS = G.AssembleRHSVector(omega, IF)
    assert RHS.state()["HostStale"]                      # assembled in HBM
    assert rel(RHS.array, RHSr) <= 1e-12
    KN = G.AllocateRHSVector()
    KN.Copy(RHS)                                         # device-to-device
    M.lu_factorize()
    M.lu_solve_vector(KN)
    assert rel(KN.array, KNr) <= 1e-8
    # same solution vector on both sides: isolates the PFT sums
    KN2 = G.AllocateRHSVector()
    KN2.set(KNr)
    P = G.GetPFTMatrix(KN2, omega, IF, EMTPFTIMethod=0)
    assert pft_err(P, PFTr) <= 1e-9, (P, PFTr)
    # and end to end from the device's own solution
    P2 = G.GetPFTMatrix(KN, omega, IF, EMTPFTIMethod=0)
    assert pft_err(P2, PFTr) <= 1e-6, (P2, PFTr)


@pytest.mark.parametrize("name,omega", [("SiO2Sphere_501.scuffgeo", 1.0), ("NestedSpheres.scuffgeo", 0.9)])
def test_interior_pft(ps, name, omega):
    RHSr, KNr, PFTr, info = ref_scatter(geo(name), omega, PW_X, interior=True)
    G = ps.RWGGeometry(geo(name))
    KN = G.AllocateRHSVector(); KN.set(KNr)
    P = G.GetPFTMatrix(KN, omega, ps.PlaneWave(*PW_X), Interior=True, EMTPFTIMethod=0)
    # the interior absorbed power of nearly lossless SiO2 (6.8e-7) is what remains of sums 1e3 times larger: 1e-8 here
    assert pft_err(P, PFTr) <= 1e-8, (P, PFTr)
    assert P[0, 1] == 0.0                                # EMTPFT.cc:863-865


def test_point_source_and_chain(ps):
    """an IncField the device does not know (PointSource) goes through the tabulated route; chained after a plane wave"""
    name, omega = "SiO2Sphere_501.scuffgeo", 1.0
    src = ((0.0, 0.0, 2.5), (0.0, 1.0 + 0.5j, 0.2))
    RHSr, KNr, PFTr, info = ref_scatter(geo(name), omega, PW_X, ps_src=src)
    G = ps.RWGGeometry(geo(name))
    IF = ps.PlaneWave(*PW_X).chain(ps.PointSource(*src))
    RHS = G.AssembleRHSVector(omega, IF)
    assert rel(RHS.array, RHSr) <= 1e-12
    KN = G.AllocateRHSVector(); KN.set(KNr)
    P = G.GetPFTMatrix(KN, omega, IF, EMTPFTIMethod=0)   # extinction uses the head of the chain only (EMTPFT.cc:499-511)
    assert pft_err(P, PFTr) <= 1e-9, (P, PFTr)


def test_point_source_inside(ps):
    """a source inside the sphere enters the RHS with the opposite sign and the interior medium's eps/mu"""
    name, omega = "SiO2Sphere_501.scuffgeo", 0.5
    src = ((0.1, 0.2, -0.3), (1.0, 0.0, 0.0))
    G = ps.RWGGeometry(geo(name))
    RHS_in = G.AssembleRHSVector(omega, ps.PointSource(*src)).array
    assert np.all(np.isfinite(RHS_in)) and np.max(np.abs(RHS_in)) > 0


def test_refusals(ps):
    G = ps.RWGGeometry(geo("PECSphere_501.scuffgeo"))
    KN = G.AllocateRHSVector()
    with pytest.raises(ps.ScuffError, match="EHDERIVATIVES1"):
        G.GetPFTMatrix(KN, 1.0, ps.PlaneWave(*PW_X), EMTPFTIMethod=1)
    with pytest.raises(ps.ScuffError, match="real positive frequency"):
        G.GetPFTMatrix(KN, 1.0j, ps.PlaneWave(*PW_X), EMTPFTIMethod=0)


def test_origin_follows_transform(ps):
    G = ps.RWGGeometry(geo("E10Spheres_327.scuffgeo"))
    assert np.allclose(G.surface_origin(0), [0, 0, 3]) and np.allclose(G.surface_origin(1), [0, 0, 0])
    G.transform("TRANS 3.4 SURFACE UpperSphere DISP 0 0 0.4")
    assert np.allclose(G.surface_origin(0), [0, 0, 3.4])
    G.untransform()
    assert np.allclose(G.surface_origin(0), [0, 0, 3])


def test_lazy_host_mode(ps, monkeypatch):
    """SCUFF_CUDA_LAZY_HOST=1: AssembleBEMMatrix (the reference-facing call) leaves the matrix in HBM; accessors fetch on demand"""
    name, omega = "PECSphere_501.scuffgeo", 1.0
    RHSr, KNr, PFTr, info = ref_scatter(geo(name), omega, PW_X)
    G = ps.RWGGeometry(geo(name))
    ctx = ps.context(); nbytes = (ps.C.c_ulonglong * 2)()
    ps.cuda_lib().scuffcuda_get_transfer_bytes(ctx, nbytes); d2h0 = nbytes[1]
    monkeypatch.setenv("SCUFF_CUDA_LAZY_HOST", "1")
    M = G.AssembleBEMMatrix(omega)                        # to_host=True path
    monkeypatch.delenv("SCUFF_CUDA_LAZY_HOST")
    assert M.state()["HostStale"]
    IF = ps.PlaneWave(*PW_X)
    KN = G.AssembleRHSVector(omega, IF)
    M.lu_factorize(); M.lu_solve_vector(KN)
    P = G.GetPFTMatrix(KN, omega, IF, EMTPFTIMethod=0)
    ps.cuda_lib().scuffcuda_get_transfer_bytes(ctx, nbytes)
    assert nbytes[1] - d2h0 < 512 * 1024                  # per-CTA partial sums only: the matrix (4 MB) never crossed PCIe
    assert pft_err(P, PFTr) <= 1e-6
    M2 = G.AssembleBEMMatrix(omega)                       # default mode: host copy current
    assert not M2.state()["HostStale"]


@pytest.mark.parametrize("name,omega", [("PECSphere_501", 1.0), ("SiO2Sphere_501", 1.0), ("E10Spheres_327", 0.8)])
def test_against_committed_golden(ps, name, omega):
    """the whole scattering point against the committed reference vectors (tests/golden/scatter_golden.npz, made by
    tests/golden/make_scatter_golden.py from the unmodified reference): no reference binary needed at test time"""
    from conftest import GOLDEN
    Gd = np.load(os.path.join(GOLDEN, "scatter_golden.npz"))
    key = "%s@%g" % (name, omega)
    G = ps.RWGGeometry(geo(name + ".scuffgeo"))
    IF = ps.PlaneWave(*PW_X)
    M = G.AssembleBEMMatrix(omega, to_host=False)
    KN = G.AssembleRHSVector(omega, IF)
    assert rel(KN.array, Gd[key + ":rhs"]) <= 1e-12
    M.lu_factorize(); M.lu_solve_vector(KN)
    assert rel(KN.array, Gd[key + ":kn"]) <= 1e-8
    KNr = G.AllocateRHSVector(); KNr.set(Gd[key + ":kn"])
    assert pft_err(G.GetPFTMatrix(KNr, omega, IF, EMTPFTIMethod=0), Gd[key + ":emtpft0"]) <= 1e-9
    assert pft_err(G.GetDSIPFTMatrix(KNr, omega, IF), Gd[key + ":dsipft"]) <= 1e-10
    F = G.GetFields(IF, KNr, omega, Gd[key + ":x"])
    assert rel(F, Gd[key + ":fields"]) <= 1e-11


@pytest.mark.parametrize("name,omega,kbloch,region", [("SiSlab_40.scuffgeo", 1.1, (0.0, 0.0), "UpperHalfSpace"), ("SiSlab_40.scuffgeo", 1.1, (0.7, 0.9), "UpperHalfSpace"),
                                                      ("PECPlate_40.scuffgeo", 0.8, (0.3, 0.0), None)])
def test_periodic_geometry_planewave_rhs_and_solve(ps, name, omega, kbloch, region):
    """periodic (LATTICE) geometries: a plane wave is its own Bloch-periodic field, so AssembleRHSVector is the same set of inner
    products over the unit cell's basis functions (straddlers reach into the neighbour cell); RHS and the solution of the periodic
    BEM system against the reference (scuff-transmission's flow)"""
    pw = ((1.0, 0.0, 0.0), (0.0, 0.0, -1.0))
    RHSr, KNr, PFTr, info = ref_scatter(geo(name), omega, pw, nopft=True, kbloch=kbloch, pw_region=region, env={"OMP_NUM_THREADS": "1"})
    assert np.max(np.abs(RHSr)) > 0
    G = ps.RWGGeometry(geo(name))
    IF = ps.PlaneWave(pw[0], pw[1], region)
    RHS = G.AssembleRHSVector(omega, IF, kBloch=kbloch)
    assert rel(RHS.array, RHSr) <= 1e-12
    M = G.AssembleBEMMatrix(omega, kBloch=kbloch, to_host=False)
    M.lu_factorize(); M.lu_solve_vector(RHS)
    assert rel(RHS.array, KNr) <= 1e-7
    with pytest.raises(ps.ScuffError, match="plane-wave"):
        G.AssembleRHSVector(omega, ps.PointSource((0, 0, 2.0), (1, 0, 0)), kBloch=kbloch)
