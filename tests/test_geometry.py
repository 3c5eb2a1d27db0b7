"""Host-side geometry front-end vs the reference's own RWGGeometry (dumps made by oracle/_ref/ref_driver geom,
committed under tests/golden/geom by tests/golden/make_golden.py).  Bit-exact: vertices, panel order, centroids,
radii, edge ordering (= BEM matrix row/column order), region indices, BF offsets, Mate[]."""
import os
import numpy as np
import pytest
from conftest import GEODIR, GOLDEN

COMPACT = ["PECSphere_501", "PECSphere_255", "PECSphere_R0P75_414", "PECSpheres_255", "PECSpheres_327",
           "SiO2Sphere_501", "GoldSphere_501", "SiSphere_255", "SiSpheres_255", "LossySphere_327", "E10Spheres_327", "OpenPlate_40"]


def parse_dump(fn):
    S, hdr, regions = [], {}, []
    for l in open(fn):
        t = l.split()
        if t[0] in ("NumSurfaces", "NumRegions", "TotalBFs", "LDim"):
            hdr[t[0]] = int(t[1])
        elif t[0] == "Region":
            regions.append((t[2], t[3]))
        elif t[0] == "Surface":
            S.append(dict(Label=t[2], IsPEC=int(t[4]), Regions=(int(t[6]), int(t[7])), BFOffset=int(t[9]), Mate=int(t[11]),
                          NV=int(t[13]), NP=int(t[15]), NE=int(t[17]), NBF=int(t[19]), V=[], P=[], E=[]))
        elif t[0] in "VPE":
            S[-1][t[0]].append([float(x) for x in t[1:]])
    return hdr, regions, S


# periodic (LATTICE) geometries: straddler vertices / panels / edges appended as RWGSurface::AddStraddlers does
# (PBCSetup.cc:142-302) -- they define the matrix row/column order of the PBC configurations
PERIODIC = ["SiSlab_40", "SphereSlabArray", "PECPlate_40"]


@pytest.mark.parametrize("name", COMPACT + PERIODIC)
def test_geometry_matches_reference(ps, name):
    hdr, regions, S = parse_dump(os.path.join(GOLDEN, "geom", name + ".geom.txt"))
    G = ps.RWGGeometry(os.path.join(GEODIR, name + ".scuffgeo"))
    assert (G.NumSurfaces, G.TotalBFs, G.NumRegions, G.LDim) == (hdr["NumSurfaces"], hdr["TotalBFs"], hdr["NumRegions"], hdr["LDim"])
    assert [(a.upper(), b.upper()) for a, b in G.regions()] == [(a.upper(), b.upper()) for a, b in regions]
    for ns, s in enumerate(S):
        m = G.surface(ns)
        V, P, E = np.array(s["V"]), np.array(s["P"]), np.array(s["E"])
        assert m["Label"] == s["Label"] and m["Regions"] == s["Regions"]
        assert m["BFOffset"] == s["BFOffset"] and m["Mate"] == s["Mate"] and m["IsPEC"] == bool(s["IsPEC"])
        assert (m["NV"], m["NP"], m["NE"], m["NBF"]) == (s["NV"], s["NP"], s["NE"], s["NBF"])
        assert np.array_equal(m["V"], V)
        assert np.array_equal(m["PVI"], P[:, :3].astype(int))
        assert np.array_equal(m["PC"], P[:, 3:6]) and np.array_equal(m["PR"], P[:, 6])
        np.testing.assert_allclose(m["PA"], P[:, 7], rtol=1e-14)
        assert np.array_equal(m["ET"], E[:, :8].astype(int)) and np.array_equal(m["EL"], E[:, 8])


def test_material_evaluation(ps):
    """eps(w) of parsed materials vs values printed by the reference (ref_driver matrix ... 'eps' field)."""
    G = ps.RWGGeometry(os.path.join(GEODIR, "SiO2Sphere_501.scuffgeo"))
    eps, mu = G.eps_mu(1.0)
    # reference: oracle/_ref/ref_driver matrix SiO2Sphere_501.scuffgeo 1 0  -> eps[1]
    ref = complex(*GOLDEN_EPS["SiO2Sphere_501@1"])
    assert abs(eps[1] - ref) <= 1e-14 * abs(ref)
    assert eps[0] == 1.0 and mu[1] == 1.0
    G = ps.RWGGeometry(os.path.join(GEODIR, "LossySphere_327.scuffgeo"))
    eps, mu = G.eps_mu(0.7)
    ref = complex(*GOLDEN_EPS["LossySphere_327@0.7"])
    assert abs(eps[1] - ref) <= 1e-14 * abs(ref)
    G = ps.RWGGeometry(os.path.join(GEODIR, "GoldSphere_501.scuffgeo"))
    eps, mu = G.eps_mu(0.1j)
    ref = complex(*GOLDEN_EPS["GoldSphere_501@0.1i"])
    assert abs(eps[1] - ref) <= 1e-14 * abs(ref)


GOLDEN_EPS = {}


def setup_module(mod):
    import json
    GOLDEN_EPS.update(json.load(open(os.path.join(GOLDEN, "eps_golden.json"))))


def test_bad_inputs(ps, tmp_path):
    with pytest.raises(ps.ScuffError):
        ps.RWGGeometry(str(tmp_path / "nonexistent.scuffgeo"))
    bad = tmp_path / "bad.scuffgeo"
    bad.write_text("OBJECT A\n\tMESHFILE nothere.msh\nENDOBJECT\n")
    with pytest.raises(ps.ScuffError):
        ps.RWGGeometry(str(bad))
    bad.write_text("FOO\n")
    with pytest.raises(ps.ScuffError):
        ps.RWGGeometry(str(bad))


@pytest.mark.parametrize("name", ["ZetaSphere_255", "OpenPlate_40"])
def test_rwg_overlaps_match_reference(ps, name):
    """RWGSurface::GetOverlap (OPFT.cc:81-226, the basis-function overlap of the surface-impedance term) against the
    reference's values for every overlapping edge pair (tests/golden/overlaps_*.txt, written by ref_driver overlaps)."""
    G = ps.RWGGeometry(os.path.join(GEODIR, name + ".scuffgeo"))
    rows = [l.split() for l in open(os.path.join(GOLDEN, "overlaps_%s.txt" % name))]
    NE = G.surface(0)["NE"]
    ref = {(int(a), int(b)): float(v) for a, b, v in rows}
    assert len(ref) >= NE                                      # at least the diagonal
    for (a, b), v in ref.items():
        mine = G.GetOverlap(0, a, b)
        assert abs(mine - v) <= 1e-14 * max(1.0, abs(v)) and abs(G.GetOverlap(0, b, a) - v) <= 1e-14 * max(1.0, abs(v))
    # edge pairs without a common panel do not overlap
    rng = np.random.default_rng(3)
    for _ in range(300):
        a, b = sorted(int(x) for x in rng.integers(0, NE, 2))
        if (a, b) not in ref:
            assert G.GetOverlap(0, a, b) == 0.0


def test_surface_impedance_expression(ps):
    """SURFACE_IMPEDANCE Zeta(w,x,y,z) (RWGSurface.cc:188-210): parsed from the rest of the line, evaluated with
    w = Omega * MatProp::FreqUnit as AddSurfaceZetaContributionToBEMMatrix does (SurfaceSurfaceInteractions.cc:400-408)."""
    G = ps.RWGGeometry(os.path.join(GEODIR, "ZetaSphere_255.scuffgeo"))
    X, omega = (0.3, -0.2, 0.5), 0.7
    w = omega * 2.99792458e14
    expect = 0.02 + 0.3j + 0.05 * X[0] * X[1] + (w / 2.99792458e14) * (0.01 - 0.02j * X[2])
    assert abs(G.surface_zeta(0, omega, X) - expect) <= 1e-15
    assert ps.RWGGeometry(os.path.join(GEODIR, "PECSphere_255.scuffgeo")).surface_zeta(0, omega, X) is None
