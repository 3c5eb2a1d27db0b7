"""Scattered + incident fields at evaluation points and the reduced-field matrix on the device against the
reference's GetFields / GetRFMatrix (oracle/_ref/ref_driver scatter --fields [--rf]; GetFields.cc:141-395).

Both sides start from the SAME solution vector (the reference's).  Tolerance 1e-11 relative to the largest field component:
the same cubature rules at the same points (16-point rule beyond 4 basis-function radii, 78-point beyond 1)."""
import os

import numpy as np
import pytest

from conftest import GEODIR, ref_scatter, ref_dgf

pytestmark = pytest.mark.gpu

PW_X = ((1.0, 0.0, 0.0), (0.0, 0.0, 1.0))


def geo(name):
    return os.path.join(GEODIR, name)


def points_outside(n, seed, rmin, rmax, centre=(0, 0, 0)):
    rng = np.random.default_rng(seed)
    d = rng.standard_normal((n, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
    return np.asarray(centre) + d * rng.uniform(rmin, rmax, n)[:, None]


def rel(a, b):
    return np.max(np.abs(a - b)) / np.max(np.abs(b))


@pytest.mark.parametrize("name,omega", [("PECSphere_501.scuffgeo", 1.0), ("SiO2Sphere_501.scuffgeo", 1.0), ("GoldSphere_501.scuffgeo", 2.0)])
def test_fields_outside_and_inside(ps, name, omega):
    # outside: from 1.5 radii (78-point rule for the closest basis functions) to 6 radii; inside (dielectrics): |x| < 0.55
    X = np.vstack([points_outside(24, 1, 1.5, 6.0), [[0.1, 0.2, 0.3], [1.3, -1.9, 0.7]]])       # + the two points of tests/Mie/EPFile
    if "PEC" not in name:
        X = np.vstack([X, points_outside(10, 2, 0.0, 0.55)])
    RHSr, KNr, PFTr, info = ref_scatter(geo(name), omega, PW_X, nopft=True, points=X)
    G = ps.RWGGeometry(geo(name))
    KN = G.AllocateRHSVector(); KN.set(KNr)
    IF = ps.PlaneWave(*PW_X)
    G.AssembleRHSVector(omega, IF)                       # UpdateIncFields, as the reference's flow does before GetFields
    F = G.GetFields(IF, KN, omega, X)
    assert rel(F, info["fields"]) <= 1e-11
    # scattered part alone
    Fs = G.GetFields(None, KN, omega, X)
    assert np.max(np.abs(Fs)) > 0


def test_nested_bodies_fields_in_every_region(ps):
    """a sphere inside a sphere: evaluation points in the exterior, in the shell (bounded by both surfaces, with opposite signs) and in the core"""
    name, omega = "NestedSpheres.scuffgeo", 0.9
    X = np.vstack([points_outside(8, 21, 1.2, 4.0), points_outside(10, 22, 0.4, 0.85), points_outside(5, 23, 0.0, 0.15), points_outside(4, 24, 0.26, 0.3)])
    RHSr, KNr, PFTr, info = ref_scatter(geo(name), omega, PW_X, nopft=True, points=X)
    G = ps.RWGGeometry(geo(name))
    KN = G.AllocateRHSVector(); KN.set(KNr)
    IF = ps.PlaneWave(*PW_X)
    G.AssembleRHSVector(omega, IF)
    F = G.GetFields(IF, KN, omega, X)
    assert rel(F, info["fields"]) <= 1e-9


def test_two_bodies_and_rf_matrix(ps):
    name, omega = "E10Spheres_327.scuffgeo", 1.0
    X = np.vstack([points_outside(12, 3, 1.5, 3.0, (0, 0, 1.5)) + [[3.0, 0, 0]], points_outside(4, 4, 0.0, 0.5, (0, 0, 3.0)), points_outside(4, 5, 0.0, 0.5)])
    RHSr, KNr, PFTr, info = ref_scatter(geo(name), omega, PW_X, nopft=True, points=X, rf=True)
    G = ps.RWGGeometry(geo(name))
    KN = G.AllocateRHSVector(); KN.set(KNr)
    IF = ps.PlaneWave(*PW_X)
    G.AssembleRHSVector(omega, IF)
    F = G.GetFields(IF, KN, omega, X)
    assert rel(F, info["fields"]) <= 1e-11
    RF = G.GetRFMatrix(omega, X).array
    assert RF.shape == info["rf"].shape
    assert rel(RF, info["rf"]) <= 1e-11
    # the matrix contracts with the solution to the scattered fields (GetFields.cc:366-376)
    Fs = G.GetFields(None, KN, omega, X)
    assert rel((KNr @ RF).reshape(-1, 6), Fs) <= 1e-12


@pytest.mark.parametrize("name,omega", [("PECSphere_501.scuffgeo", 1.0), ("SiO2Sphere_501.scuffgeo", 1.0), ("PECSphere_255.scuffgeo", 0.3)])
def test_near_fields(ps, name, omega):
    """evaluation points closer than one basis-function radius to the surface (down to 1 % of the sphere radius, outside and — for the
    dielectric — inside): the reference's GetReducedFields_Nearby branch (closed-form static panel integrals + desingularised cubature),
    restated in csrc/nearfield.cuh.  Tolerance 1e-9 relative to the largest field component: the closed forms cancel for points close to
    a panel's plane; both sides evaluate the same expressions, libm vs CUDA log/atan/pow differ in the last bit"""
    X = np.vstack([points_outside(30, 11, 1.01, 1.3), [[0.0, 0.0, 1.02], [0.0, 0.6, 0.85]]])
    if "PEC" not in name:
        X = np.vstack([X, points_outside(20, 12, 0.7, 0.97)])
    RHSr, KNr, PFTr, info = ref_scatter(geo(name), omega, PW_X, nopft=True, points=X, rf=True)
    G = ps.RWGGeometry(geo(name))
    KN = G.AllocateRHSVector(); KN.set(KNr)
    IF = ps.PlaneWave(*PW_X)
    G.AssembleRHSVector(omega, IF)
    F = G.GetFields(IF, KN, omega, X)
    err = np.max(np.abs(F - info["fields"]), axis=1) / np.max(np.abs(info["fields"]))
    print(name, "near-field parity, worst point:", err.max(), "at", X[np.argmax(err)])
    assert err.max() <= 1e-9
    RF = G.GetRFMatrix(omega, X).array
    assert rel(RF, info["rf"]) <= 1e-9


def test_dsipft_default_radius_on_a_coarse_mesh(ps):
    """PECSphere_255: the automatic DSI radius (1.5 x extent) puts cubature points within one basis-function radius of the surface"""
    name, omega = "PECSphere_255.scuffgeo", 1.0
    RHSr, KNr, PFTr, info = ref_scatter(geo(name), omega, PW_X, nopft=True, dsi=(0.0, 302))
    G = ps.RWGGeometry(geo(name))
    KN = G.AllocateRHSVector(); KN.set(KNr)
    IF = ps.PlaneWave(*PW_X)
    G.AssembleRHSVector(omega, IF)
    P = G.GetDSIPFTMatrix(KN, omega, IF)
    ref = info["dsipft"]
    # (the torque on a sphere vanishes by symmetry: its 1e-10 noise is compared on the scale of force x radius)
    for cols in ((0, 1), (2, 3, 4, 5, 6, 7)):
        assert np.max(np.abs(P[:, cols] - ref[:, cols])) <= 1e-9 * np.max(np.abs(ref[:, cols])), (P, ref)


@pytest.mark.parametrize("name,omega,trans,dsi", [("SiO2Sphere_501.scuffgeo", 1.0, None, (0.0, 302)), ("PECSphere_501.scuffgeo", 1.0, None, (2.5, 110)),
                                                  ("E10Spheres_327.scuffgeo", 0.8, ("Spheres.trans", 4), (0.0, 302))])
def test_dsipft(ps, name, omega, trans, dsi):
    """power, force, torque by the displaced-surface-integral method (GetDSIPFT, DSIPFT.cc:449-590; the reference's tests/Mie checks it):
    Lebedev rule on a sphere around each body (automatic radius and an explicit one), fields from the device; the second body of the
    two-sphere case is displaced by a transformation, so its sphere and torque centre move with it"""
    RHSr, KNr, PFTr, info = ref_scatter(geo(name), omega, PW_X, nopft=True, trans=trans, dsi=dsi)
    G = ps.RWGGeometry(geo(name))
    if trans is not None:
        G.transform(open(os.path.join(GEODIR, trans[0])).read().splitlines()[trans[1]])
    KN = G.AllocateRHSVector(); KN.set(KNr)
    IF = ps.PlaneWave(*PW_X)
    G.AssembleRHSVector(omega, IF)
    P = G.GetDSIPFTMatrix(KN, omega, IF, DSIRadius=dsi[0], DSIPoints=dsi[1])
    ref = info["dsipft"]
    for cols in ((0, 1), (2, 3, 4), (5, 6, 7)):
        scale = np.max(np.abs(ref[:, cols]))
        assert np.max(np.abs(P[:, cols] - ref[:, cols])) <= 1e-10 * scale, (P, ref)


@pytest.mark.parametrize("name,omega", [("PECSphere_501.scuffgeo", 1.0), ("SiO2Sphere_501.scuffgeo", 0.7)])
def test_dyadic_green_functions(ps, name, omega):
    """RWGGeometry::GetDyadicGFs (scuff-ldos): one-point (source = destination) and two-point scattering dyadics, and the two-point
    total dyadic (direct contribution of a point source added); reduced-field matrices, the 6 NX solves and the contractions on the device"""
    G = ps.RWGGeometry(geo(name))
    M = G.AssembleBEMMatrix(omega, to_host=False)
    M.lu_factorize()
    X1 = points_outside(6, 31, 1.05, 3.0)                    # incl. near-field points
    GEr, GMr = ref_dgf(geo(name), omega, X1)
    GE, GM = G.GetDyadicGFs(omega, X1, M)
    assert rel(GE, GEr) <= 1e-8 and rel(GM, GMr) <= 1e-8
    X2 = np.hstack([points_outside(5, 32, 1.5, 3.0), points_outside(5, 33, 1.5, 3.0)])
    for total in (False, True):
        GEr, GMr = ref_dgf(geo(name), omega, X2, total=total)
        GE, GM = G.GetDyadicGFs(omega, X2, M, ScatteringOnly=not total)
        assert rel(GE, GEr) <= 1e-8 and rel(GM, GMr) <= 1e-8
