"""CPU-side checks of the scattering-run rows: cubature tables extracted from the reference, the incident-field
classes of the C++ mirror against independent numpy formulas, the committed golden vectors, and "no CPU fallback"."""
import os
import re

import numpy as np
import pytest

from conftest import GEODIR, GOLDEN, ROOT

ZVAC = 376.73031346177


def table(path, name):
    src = open(os.path.join(ROOT, path)).read()
    m = re.search(r"%s\[\d+\] = \{(.*?)\};" % re.escape(name), src, flags=re.S)
    return np.array([float(x) for x in m.group(1).replace("\n", " ").split(",") if x.strip()])


@pytest.mark.parametrize("hdr,name,npts", [("scuff-em_b200/csrc/generated_tcr7.h", "SC_TCR7_H", 16), ("scuff-em_b200/csrc/generated_tcr9.h", "SC_TCR9_H", 21),
                                           ("scuff-em_b200/csrc/generated_tcr20.h", "SC_TCR20_H", 78)])
def test_triangle_rules(hdr, name, npts):
    """(u, v, w) rules of GetTCR: weights sum to 1/2 (the reference multiplies by 2 x area), points inside the triangle, and the rule
    integrates monomials of its order exactly: int u^a v^b = a! b! / (a + b + 2)!"""
    T = table(hdr, name).reshape(npts, 3)
    u, v, w = T.T
    assert abs(w.sum() - 0.5) < 1e-11
    assert np.all(u > 0) and np.all(v > 0) and np.all(u + v < 1)
    from math import factorial
    order = {16: 7, 21: 9, 78: 20}[npts]
    worst = 0.0
    for a in range(order + 1):
        for b in range(order + 1 - a):
            exact = factorial(a) * factorial(b) / factorial(a + b + 2)
            worst = max(worst, abs((w * u ** a * v ** b).sum() - exact) / exact)
    assert worst < 1e-9, worst          # the reference's tables carry 13-16 digits


@pytest.mark.parametrize("n", [6, 110, 302, 590])
def test_lebedev_rules(n):
    L = table("scuff-em_b200/host/generated_lebedev.h", "SC_LEBEDEV_%d" % n).reshape(n, 4)
    assert np.allclose(np.linalg.norm(L[:, :3], axis=1), 1.0, atol=1e-12)
    assert abs(L[:, 3].sum() - 4 * np.pi) < 1e-10                      # weights include the 4 pi of the sphere (DSIPFT.cc:301-310 multiplies by R^2)
    # exact for low-order spherical polynomials: int x^2 = 4 pi / 3, int x^2 y^2 = 4 pi / 15
    if n > 6:
        assert abs((L[:, 3] * L[:, 0] ** 2).sum() - 4 * np.pi / 3) < 1e-10
        assert abs((L[:, 3] * L[:, 0] ** 2 * L[:, 1] ** 2).sum() - 4 * np.pi / 15) < 1e-10


def test_planewave_fields(ps):
    E0 = np.array([0.3 + 0.1j, -0.2j, 0.1]); n = np.array([0.6, 0.0, 0.8]); omega, eps, mu = 1.3, 2.5 + 0.1j, 1.0
    X = np.random.default_rng(0).uniform(-2, 2, (5, 3))
    EH, dEH = ps.PlaneWave(E0, n).eval(omega, X, eps, mu, gradients=True)
    k = np.sqrt(eps * mu) * omega; Z = ZVAC * np.sqrt(mu / eps)
    E = E0[None, :] * np.exp(1j * k * X @ n)[:, None]
    assert np.allclose(EH[:, :3], E, rtol=1e-14, atol=0)
    assert np.allclose(EH[:, 3:], np.cross(n, E) / Z, rtol=1e-13, atol=1e-18)
    assert np.allclose(dEH, 1j * k * n[None, :, None] * EH[:, None, :], rtol=1e-13, atol=1e-18)


def test_pointsource_fields(ps):
    """electric dipole P (moment / eps0) at X0 in a medium: the textbook field E = k^2 e^{ikr}/(4 pi eps r) [ (n x p) x n + (3 n (n.p) - p)(1/(kr)^2 - i/(kr)) ],
    H = k^2 e^{ikr}/(4 pi eps r) (n x p)(1 - 1/(ikr)) / Z; gradients by the reference's central differences (IncField.cc:178-198)"""
    X0 = np.array([0.1, -0.2, 0.3]); P = np.array([0.5, 1.0 + 0.5j, -0.2]); omega, eps, mu = 0.9, 3.0, 1.0
    X = np.random.default_rng(1).uniform(1, 3, (6, 3))
    EH, dEH = ps.PointSource(X0, P).eval(omega, X, eps, mu, gradients=True)
    k = np.sqrt(eps * mu) * omega; Z = ZVAC * np.sqrt(mu / eps)
    R = X - X0; r = np.linalg.norm(R, axis=1); n = R / r[:, None]
    pref = (k * k * np.exp(1j * k * r) / (4 * np.pi * eps * r))[:, None]
    ndp = (n @ P)[:, None]
    kr = (k * r)[:, None]
    E = pref * (np.cross(np.cross(n, P), n) + (3 * n * ndp - P) * (1 / kr ** 2 - 1j / kr))
    H = pref * np.cross(n, P) * (1 - 1 / (1j * kr)) / Z
    assert np.allclose(EH[:, :3], E, rtol=1e-12, atol=0)
    assert np.allclose(EH[:, 3:], H, rtol=1e-12, atol=0)
    # finite-difference gradients approximate the analytic ones to O(Delta^2)
    h = 1e-6
    for i in range(3):
        Xp = X.copy(); Xp[:, i] += h; Xm = X.copy(); Xm[:, i] -= h
        g = (ps.PointSource(X0, P).eval(omega, Xp, eps, mu) - ps.PointSource(X0, P).eval(omega, Xm, eps, mu)) / (2 * h)
        assert np.allclose(dEH[:, i, :], g, rtol=1e-5, atol=1e-9)


def test_scatter_golden_is_consistent():
    """the committed reference vectors (tests/golden/make_scatter_golden.py): shapes, and the physics they must satisfy —
    a PEC body absorbs nothing, the two PFT methods of the reference agree on the powers, total field inside a PEC body is zero"""
    Gd = np.load(os.path.join(GOLDEN, "scatter_golden.npz"))
    for key, nbf, ns in (("PECSphere_501@1", 501, 1), ("SiO2Sphere_501@1", 1002, 1), ("E10Spheres_327@0.8", 1308, 2)):
        assert Gd[key + ":rhs"].shape == (nbf,) and Gd[key + ":kn"].shape == (nbf,)
        emt, dsi = Gd[key + ":emtpft0"], Gd[key + ":dsipft"]
        assert emt.shape == (ns, 8) and dsi.shape == (ns, 8) and Gd[key + ":fields"].shape == (12, 6)
        assert np.allclose(emt[:, 1], dsi[:, 1], rtol=2e-6)                 # scattered power: EMT (cubature integrals) vs DSI
        assert np.allclose(emt[:, 0], dsi[:, 0], rtol=1e-4, atol=1e-9)      # absorbed power
    assert abs(Gd["PECSphere_501@1:emtpft0"][0, 0]) < 1e-9


def test_no_cpu_fallback_for_the_scattering_run(ps):
    if ps.cuda_lib().scuffcuda_device_count() > 0:
        pytest.skip("a CUDA device is present")
    G = ps.RWGGeometry(os.path.join(GEODIR, "PECSphere_255.scuffgeo"))
    with pytest.raises(ps.ScuffError):
        G.AssembleRHSVector(1.0, ps.PlaneWave((1, 0, 0), (0, 0, 1)))


def test_oracle_rhs_and_emtpft_power_pinned_to_the_reference(oracle):
    """numpy restatement (oracle/scuff_oracle.py) of the reference's per-basis-function formulation of the plane-wave right-hand side
    and of the energy/momentum-transfer powers, against the committed reference vectors: pins the oracle for SURVEY 8 row f2"""
    Gd = np.load(os.path.join(GOLDEN, "scatter_golden.npz"))
    for name, pec in (("PECSphere_501", True), ("SiO2Sphere_501", False)):
        S = oracle.parse_geom_dump(os.path.join(GOLDEN, "geom", name + ".geom.txt"))[2][0]
        key = name + "@1"
        rhs = oracle.rhs_planewave(S, 1.0, (1.0, 0.0, 0.0), (0.0, 0.0, 1.0))
        assert np.max(np.abs(rhs - Gd[key + ":rhs"])) <= 1e-13 * np.max(np.abs(Gd[key + ":rhs"]))
    # powers (501^2 pairs of basis functions x 4 panel pairs x 36 / 441 point pairs, vectorised: ~15 s each): PEC (electric currents
    # only) and dielectric (electric and magnetic currents, all four bilinears)
    for name in ("PECSphere_501", "SiO2Sphere_501"):
        S = oracle.parse_geom_dump(os.path.join(GOLDEN, "geom", name + ".geom.txt"))[2][0]
        pabs, pscat = oracle.emtpft_power_single_surface(S, Gd[name + "@1:kn"], 1.0, (1.0, 0.0, 0.0), (0.0, 0.0, 1.0))
        ref = Gd[name + "@1:emtpft0"][0]
        assert abs(pscat - ref[1]) <= 1e-10 * abs(ref[1]), (name, pscat, ref[1])
        assert abs(pabs - ref[0]) <= 1e-10 * max(abs(ref[0]), abs(ref[1])), (name, pabs, ref[0])


def test_torque_centre_follows_the_surface(ps):
    """RWGSurface::Origin (the torque centre of GetPFT) starts at the object's DISPLACED/ROTATED origin and moves with every
    transformation, as in the reference (RWGSurface.cc:423-429, :609, :657); host logic, no device"""
    G = ps.RWGGeometry(os.path.join(GEODIR, "E10Spheres_327.scuffgeo"))
    assert np.allclose(G.surface_origin(0), [0, 0, 3]) and np.allclose(G.surface_origin(1), [0, 0, 0])
    G.transform("TRANS 3.4 SURFACE UpperSphere DISP 0 0 0.4")
    assert np.allclose(G.surface_origin(0), [0, 0, 3.4]) and np.allclose(G.surface_origin(1), [0, 0, 0])
    G.untransform()
    assert np.allclose(G.surface_origin(0), [0, 0, 3])


def test_nested_geometry_regions(ps):
    """a body inside a body: the inner surface's exterior region is the outer body's interior (the branches of the PFT sums and of
    GetFields that the nested fixture exercises on the GPU)"""
    G = ps.RWGGeometry(os.path.join(GEODIR, "NestedSpheres.scuffgeo"))
    assert [G.surface(i)["Regions"] for i in range(2)] == [(0, 1), (1, 2)]
    assert [r[0] for r in G.regions()] == ["EXTERIOR", "Shell", "Core"]
