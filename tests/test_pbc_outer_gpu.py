"""Periodic geometries, step 2 of AssembleBEMMatrixBlock's PBC branch (AssembleBEMMatrix.cc:526-560) and everything under it:
the Ewald-summed Bloch-periodic Green's function (GBarVDEwald.cc:931-996), the interpolation-table builder
(CreateGBarAccelerator GBarAccelerator.cc:339-547 with the grid decisions of GetOptimalGridSpacing2D), GetGBar (:823-1005),
the table-evaluating cubature kernel, the kBloch-independent block cache (CreateABMBAccelerator) and the complete periodic
BEM matrix RWGGeometry::AssembleBEMMatrix(Omega, kBloch, M) — BASELINE config 4.  Checker: the unmodified reference
(oracle/_ref) and the committed goldens it wrote (tests/golden/gbar_golden.json).  Tolerance: 1e-10 relative Frobenius per
block for matrices (north_star); Green's function values 1e-10 relative to the largest magnitude of each component."""
import json
import os
import numpy as np
import pytest
from conftest import GEODIR, GOLDEN, ref_matrix, rel_fro
from conftest_pbc import ref_gbar_points, ref_gba

pytestmark = pytest.mark.gpu


def _cmp(a, b, floor=1e-6):
    sc = np.abs(b).max(axis=0)
    return (np.abs(a - b) / (np.abs(b) + floor * sc)).max()


def test_gbar_ewald_golden(ps):
    """device GBarVDEwald against the goldens the reference wrote (value, gradient and mixed derivatives)"""
    cases = json.load(open(os.path.join(GOLDEN, "gbar_golden.json")))
    n = 0
    for c in cases:
        if c["ldim"] != 2 or c["lbv"][0][1] != 0 or c["lbv"][1][0] != 0:
            continue
        G = ps.gbar_ewald(complex(*c["k"]), c["kb"], [c["lbv"][0][:2], c["lbv"][1][:2]], c["R"], exclude_inner=bool(c["excl"]))
        ref = np.array(c["GBarVD"]); ref = ref[:, 0::2] + 1j * ref[:, 1::2]
        err = _cmp(G[:, :ref.shape[1]], ref)
        print(c["k"], c["kb"], c["excl"], "err", err)
        assert err <= 1e-10
        n += 1
    assert n >= 4


@pytest.mark.parametrize("k", [0.1, 1.1, 1.1 * np.sqrt(11.87), 0.5j, 2.0 + 0.3j])
@pytest.mark.parametrize("kb", [(0.0, 0.0), (0.7, 0.9)])
def test_gbar_ewald_vs_reference(ps, tmp_path, k, kb):
    """device GBarVDEwald against the live reference at table-like points (cell corners, z = 0 planes) and random points"""
    rng = np.random.default_rng(7)
    for lbv in ([[1.0, 0.0], [0.0, 1.0]], [[0.7, 0.0], [0.0, 1.3]]):
        g = np.linspace(-0.5, 0.5, 4)
        X, Y, Z = np.meshgrid(g * lbv[0][0], g * lbv[1][1], [0.0, 0.3186, 1.0, -0.4], indexing="ij")
        R = np.concatenate([np.c_[X.ravel(), Y.ravel(), Z.ravel()], rng.uniform(-1.5, 1.5, (60, 3))])
        for excl in (True, False):
            Rr = R if excl else R[np.linalg.norm(R, axis=1) > 1e-3]
            mine = ps.gbar_ewald(k, kb, lbv, Rr, exclude_inner=excl)
            ref = ref_gbar_points(k, kb, lbv, excl, Rr, str(tmp_path))
            err = _cmp(mine[:, :4], ref[:, :4])          # value + gradient: what the matrix uses directly
            errd = _cmp(mine[:, 4:], ref[:, 4:], 1e-3)    # mixed derivatives: enter through the tricubic fit only
            print(k, kb, lbv, excl, "value/gradient", err, "mixed derivatives", errd)
            assert err <= 1e-10 and errd <= 1e-9


TABLE_CASES = [("SiSlab_40.scuffgeo", 1, 0, 0), ("SiSlab_40.scuffgeo", 2, 0, 1), ("SiSlab_40.scuffgeo", 2, 0, 0),
               ("PECPlate_40.scuffgeo", 0, 0, 0), ("SphereSlabArray.scuffgeo", 1, 0, 2), ("SphereSlabArray.scuffgeo", 1, 2, 2),
               ("SphereSlabArray.scuffgeo", 2, 0, 1), ("SphereSlabArray.scuffgeo", 0, 2, 2)]


@pytest.mark.parametrize("geo,nr,nsa,nsb", TABLE_CASES)
@pytest.mark.parametrize("omega,kb", [(0.1, (0.0, 0.0)), (1.1, (0.7, 0.9))])
def test_gbar_table_vs_reference(ps, tmp_path, geo, nr, nsa, nsb, omega, kb):
    """interpolation table: same grid decisions as CreateRegionGBA, same coefficients, same GetGBar values (table lookups
    with Wigner-Seitz wrap and add-back, full Ewald outside the |z| range)"""
    rng = np.random.default_rng(3)
    pts = np.concatenate([rng.uniform(-1, 1, (150, 3)), np.c_[rng.uniform(-1, 1, (80, 2)), np.zeros(80)],
                          np.c_[rng.uniform(-1, 1, (80, 2)), np.ones(80)], np.c_[rng.uniform(-1, 1, (80, 2)), -np.ones(80)]])
    J, ct, ref = ref_gba(os.path.join(GEODIR, geo), omega, kb, nr, nsa, nsb, pts, str(tmp_path))
    if J is None:
        pytest.skip("the reference builds no table for this region (not extended)")
    T = ps.GBarAccelerator(complex(J["kRe"], J["kIm"]), kb, [[1.0, 0.0], [0.0, 1.0]], J["RhoMin"], J["RhoMax"])
    assert T.full == bool(J["full"])
    if "N" in J:
        assert T.N == J["N"], (T.N, J["N"])
        assert np.allclose(T.XMin, J["XMin"], rtol=0, atol=1e-15)
        assert np.allclose(T.DX[:2], J["DX"][:2], rtol=1e-14)
        # the Rho spacing of a zero-width range is the probed optimal spacing itself (a ratio of interpolation errors):
        # it inherits their conditioning
        assert np.allclose(T.DX[2], J["DX"][2], rtol=1e-6)
        mine_ct = T.table().reshape(-1, 4, 4, 4)
        ref_ct = ct.reshape(-1, 4, 4, 4)
        # coefficients of t3^0 do not depend on the Rho spacing
        e0 = np.abs(mine_ct[..., 0] - ref_ct[..., 0]).max() / np.abs(ref_ct[..., 0]).max()
        print(geo, nr, (nsa, nsb), omega, kb, "N", T.N, "coefficient err (r=0)", e0)
        assert e0 <= 1e-9
    mine = T.eval(pts)
    err = _cmp(mine, ref)
    print(geo, nr, (nsa, nsb), omega, kb, "GetGBar err", err)
    assert err <= 1e-9
    T.close()


MATRIX_CASES = [   # unit-test-BEMMatrix.cc:155-203 (tests 8-12) + a lossy / imaginary-frequency point
    ("PECPlate_40.scuffgeo", 0.1, (0.0, 0.0)),
    ("PECPlate_40.scuffgeo", 1.1, (0.7, 0.9)),
    ("SiSlab_40.scuffgeo", 0.1, (0.0, 0.0)),
    ("SiSlab_40.scuffgeo", 1.1, (0.7, 0.9)),
    ("SphereSlabArray.scuffgeo", 1.1, (0.7, 0.9)),
    ("SiSlab_40.scuffgeo", 0.4 + 0.1j, (-0.5, 1.1)),
    ("PECPlate_40.scuffgeo", 0.8j, (0.3, 0.2)),
]


@pytest.mark.parametrize("geo,omega,kb", MATRIX_CASES)
def test_periodic_matrix(ps, tmp_path, geo, omega, kb):
    """RWGGeometry::AssembleBEMMatrix(Omega, kBloch, M) on a LATTICE geometry, every block <= 1e-10 against the
    single-threaded reference (FIPPI records shared through the reference's own cache file)"""
    path = os.path.join(GEODIR, geo)
    cache = str(tmp_path / "pbc.cache")
    ref, info = ref_matrix(path, omega, kbloch=kb, cache_out=cache, env={"OMP_NUM_THREADS": "1"})
    ps.ClearCache(); ps.PreloadCache(cache)
    G = ps.RWGGeometry(path)
    M = G.AssembleBEMMatrix(omega, kBloch=kb).array
    offs = [G.surface(ns)["BFOffset"] for ns in range(G.NumSurfaces)] + [G.TotalBFs]
    worst = 0.0
    for a in range(G.NumSurfaces):
        for b in range(G.NumSurfaces):
            A = M[offs[a]:offs[a + 1], offs[b]:offs[b + 1]]; B = ref[offs[a]:offs[a + 1], offs[b]:offs[b + 1]]
            if np.linalg.norm(B) == 0:
                assert np.all(A == 0)
                continue
            worst = max(worst, rel_fro(A, B))
    print(geo, omega, kb, "worst block rel_fro", worst, "whole", rel_fro(M, ref))
    assert worst <= 1e-10


def test_abmb_accelerator(ps):
    """CreateABMBAccelerator / KBIMBCache (AssembleBEMMatrix.cc:250-334,405-408,477-503): with the cache the inner-cell blocks
    are assembled once per Omega and only re-stamped per kBloch; results equal the uncached blocks"""
    path = os.path.join(GEODIR, "SiSlab_40.scuffgeo")
    G = ps.RWGGeometry(path)
    for nsa, nsb in ((0, 0), (0, 1)):
        sa, sb = G.surface(nsa), G.surface(nsb)
        acc = G.CreateABMBAccelerator(nsa, nsb)
        assert acc
        launches = []
        for omega in (0.9, 1.3):
            for kb in ((0.1, 0.2), (0.5, -0.3), (0.0, 0.0)):
                M0 = ps.HMatrix.new(sa["NBF"], sb["NBF"]); M1 = ps.HMatrix.new(sa["NBF"], sb["NBF"])
                G.AssembleBEMMatrixBlock(nsa, nsb, omega, M0, kBloch=kb)
                n0 = ps.launch_count()
                G.AssembleBEMMatrixBlock(nsa, nsb, omega, M1, kBloch=kb, ABMBCache=acc)
                launches.append(ps.launch_count() - n0)
                err = rel_fro(M1.array, M0.array)
                assert err <= 1e-13, (nsa, nsb, omega, kb, err)
        # the first kBloch of each Omega assembles the blocks, the next two only stamp them
        assert launches[1] < launches[0] and launches[2] < launches[0] and launches[4] < launches[3], launches
        G.DestroyABMBAccelerator(acc)
    C = ps.RWGGeometry(os.path.join(GEODIR, "PECSphere_255.scuffgeo"))
    assert not C.CreateABMBAccelerator(0, 0)          # compact geometry: no accelerator (AssembleBEMMatrix.cc:264-265)
