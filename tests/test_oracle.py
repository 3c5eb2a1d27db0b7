"""Pins the CPU oracle (oracle/scuff_oracle.py, numpy restatement of the far-field path) against
 (a) the reference's OWN known answers, unitTests/unit-test-PPIs.cc:70-79 (cases NCLW/NCMW/NCSW: panel pair (0,223)
     of PECSphere_R0P75_414 — the no-common-vertex cases the 6x6 cubature serves), compared as the reference's
     test compares them (EqualFloat: equal after rounding to float; we allow 2 float ulps), and
 (b) golden vectors written by the UNMODIFIED reference (oracle/_ref) through tests/golden/make_golden.py."""
import json
import os
import numpy as np
import pytest
from conftest import GOLDEN

# unitTests/unit-test-PPIs.cc:70-79
REF_KNOWN = {0.1: (-5.380869460e+00 - 7.929329541e-01j, -1.203784538e-04 - 1.244908468e-07j),
             5.0: (-1.097563488e-03 - 1.833260738e-03j, -8.115226086e-04 + 3.380858352e-04j),
             5.0 + 20.0j: (+4.102927698e-18 + 3.052119470e-17j, -5.053807314e-16 - 5.115425623e-16j)}


def surf(oracle, name):
    return oracle.parse_geom_dump(os.path.join(GOLDEN, "geom", name + ".geom.txt"))[2]


def panel(S, np_):
    return S["V"][S["PVI"][np_]]


def test_reference_known_answers(oracle):
    S = surf(oracle, "PECSphere_R0P75_414")[0]
    assert oracle.rrel_matrix(S, S)[0, 223] > 4.0          # a far pair: PPIALG_LOCUBATURE
    Va, Vb = panel(S, 0), panel(S, 223)
    for k, exact in REF_KNOWN.items():
        H = oracle.ppi_cubature(Va, Va[0], Vb, Vb[0], k)
        for n in range(2):
            for got, want in ((H[n].real, exact[n].real), (H[n].imag, exact[n].imag)):
                assert abs(got - want) <= 3e-7 * abs(want) + 1e-30, (k, n, got, want)


def test_ppis_vs_reference_golden(oracle):
    g = json.load(open(os.path.join(GOLDEN, "ppi_golden.json")))
    S = surf(oracle, g["geometry"])[0]
    nfar = 0
    for kk, rows in zip(g["k"], g["H"]):
        k = complex(kk[0], kk[1])
        for (nsa, npa, iQa, nsb, npb, iQb), r in zip(g["pairs"], rows):
            if int(r[4]) != 0:      # only PPIALG_LOCUBATURE is restated in numpy; the rest is checked with oracle/_ref
                continue
            Va, Vb = panel(S, npa), panel(S, npb)
            H = oracle.ppi_cubature(Va, Va[iQa], Vb, Vb[iQb], k)
            ref = np.array([complex(r[0], r[1]), complex(r[2], r[3])])
            assert np.all(np.abs(H - ref) <= 1e-13 * np.abs(ref) + 1e-300), (k, npa, npb, H, ref)
            nfar += 1
    assert nfar > 50


def test_desingularised_integrand_is_consistent(oracle):
    """ExpRel(x,4) = exp(x) - 1 - x - x^2/2 - x^3/6 to the reference's 1e-8 truncation"""
    for x in (0.3j, 0.05 + 0.2j, -1.5, 2.0j):
        want = np.exp(x) - 1 - x - x * x / 2 - x ** 3 / 6
        assert abs(oracle.exprel(x, 4) - want) <= 1e-7 * abs(want)


@pytest.mark.parametrize("name,omega,eps_in", [("PECSphere_255", 1.0, None), ("PECSphere_255", 0.1j, None),
                                               ("PECSphere_501", 0.7 + 0.3j, None), ("SiSphere_255", 0.01, "golden")])
def test_far_matrix_entries_vs_reference(oracle, name, omega, eps_in):
    """all-far entries of the restated matrix == the reference's AssembleBEMMatrix entries (<= 1e-12)"""
    Z = np.load(os.path.join(GOLDEN, "matrix_samples.npz"))
    key = "%s@%r@0" % (name, complex(omega))
    idx, val = Z[key + "/idx"], Z[key + "/val"]
    S = surf(oracle, name)[0]
    if eps_in == "golden":
        eps_in = complex(*json.load(open(os.path.join(GOLDEN, "eps_golden.json")))["SiSphere_255@0.01"])
    M, mask = oracle.far_matrix_single_surface(S, omega, eps_in=eps_in)
    # reference fills nea<=neb and mirrors: compare on the upper triangle position of each sample
    r, c = np.minimum(idx[:, 0], idx[:, 1]), np.maximum(idx[:, 0], idx[:, 1])
    if not S["IsPEC"]:
        r, c = idx[:, 0], idx[:, 1]
    sel = mask[r, c]
    assert sel.sum() > 1000
    got, want = M[r, c][sel], val[sel]
    assert np.linalg.norm(got - want) <= 1e-12 * np.linalg.norm(want)


@pytest.mark.parametrize("geom,golden", [("PECSphere_255", "ZetaSphere_255"), ("OpenPlate_40", "OpenPlate_40")])
def test_rwg_overlap_vs_reference_golden(oracle, geom, golden):
    """the overlap integrals of the surface-impedance term against the reference's RWGSurface::GetOverlap values"""
    S = surf(oracle, geom)[0]
    n = 0
    for l in open(os.path.join(GOLDEN, "overlaps_%s.txt" % golden)):
        a, b, v = l.split()
        got = oracle.rwg_overlap(S, int(a), int(b))
        assert abs(got - float(v)) <= 1e-14 * max(1.0, abs(float(v))), (a, b, got, v)
        n += 1
    assert n >= S["NE"]


def test_stamp_in_neighbor_block_properties(oracle):
    """Bloch-phase stamping: kBloch = 0 adds B + B^T; the +L and -L cells of a diagonal block are each other's
    conjugate-phase transposes, so the stamped block is Hermitian-symmetric in the phase (B real-symmetric => result real)."""
    rng = np.random.default_rng(5)
    B = rng.standard_normal((6, 6)) + 1j * rng.standard_normal((6, 6))
    M = oracle.stamp_in_neighbor_block(np.zeros((8, 8), complex), B, 1, 2, (1.0, 0.0), (0.0, 0.0), True)
    assert np.allclose(M[1:7, 2:8], B + B.T) and np.all(M[0] == 0) and np.all(M[:, :2] == 0)
    Br = rng.standard_normal((6, 6)); Br = Br + Br.T
    M = oracle.stamp_in_neighbor_block(np.zeros((6, 6), complex), Br, 0, 0, (1.0, 0.5), (0.7, -0.3), True)
    assert np.allclose(M.imag, 0.0) and np.allclose(M.real, 2 * np.cos(0.7 - 0.15) * Br)
    M1 = oracle.stamp_in_neighbor_block(np.zeros((6, 6), complex), B, 0, 0, (1.0, 0.0), (0.4, 0.0), False)
    assert np.allclose(M1, np.exp(0.4j) * B)


def test_gbar_direct_sum_vs_reference_golden(oracle):
    """Periodic Green's function: the brute-force Bloch lattice sum (absolutely convergent for Im k > 0) against the
    reference's Ewald-summed GBarVDEwald values and first derivatives (tests/golden/gbar_golden.json), full sum and with the
    inner 3x3 (3) cells excluded.  Pins the definition (R - L, phase exp(+i kBloch.L)) the outer-cell step must reproduce."""
    cases = json.load(open(os.path.join(GOLDEN, "gbar_golden.json")))
    checked = 0
    for c in cases:
        if c["k"][1] <= 0 or c.get("reference_not_converged"):
            continue
        lbv = [c["lbv"][0]] + ([c["lbv"][1]] if c["ldim"] == 2 else [])
        tol = 1e-12 if c["ldim"] == 2 else 5e-9           # the 1-D Ewald sums of the reference are converged to ~1e-9
        for R, g in list(zip(c["R"], c["GBarVD"]))[:3]:
            G, dG = oracle.gbar_direct_sum(R, complex(*c["k"]), c["kb"], lbv, bool(c["excl"]), ncells=30)
            ref = complex(g[0], g[1]); dref = np.array([complex(g[2 + 2 * m], g[3 + 2 * m]) for m in range(3)])
            assert abs(G - ref) <= tol * abs(ref), (c, R, G, ref)
            assert np.abs(dG - dref).max() <= tol * 10 * np.abs(dref).max(), (c, R, dG, dref)
            checked += 1
    assert checked >= 12


def test_gbar_device_source_host_build_vs_golden(tmp_path):
    """The Ewald restatement the CUDA kernels compile (scuff-em_b200/csrc/gbar.cuh, SCUFF_HD inline source) built for the host by
    g++ (tests/host_harness.cc) reproduces the reference's GBarVDEwald goldens, value and all derivatives, on rectangular
    lattices — the check of the formulas that can run without a GPU.  (The product path only runs the device build.)"""
    import ctypes
    import shutil
    import subprocess
    if not shutil.which("g++"):
        pytest.skip("no g++")
    so = str(tmp_path / "libhh.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(os.path.dirname(__file__), "host_harness.cc")], check=True)
    hh = ctypes.CDLL(so)
    dp = ctypes.POINTER(ctypes.c_double)
    cases = json.load(open(os.path.join(GOLDEN, "gbar_golden.json")))
    n = 0
    for c in cases:
        if c["ldim"] != 2 or c["lbv"][0][1] != 0 or c["lbv"][1][0] != 0:
            continue
        R = np.ascontiguousarray(c["R"], dtype=float); out = np.zeros((len(R), 16))
        kb = np.array(c["kb"], dtype=float); lbv = np.array([c["lbv"][0][:2], c["lbv"][1][:2]], dtype=float).ravel()
        hh.hh_gbar(ctypes.c_double(c["k"][0]), ctypes.c_double(c["k"][1]), kb.ctypes.data_as(dp), lbv.ctypes.data_as(dp), int(c["excl"]),
                   len(R), R.ctypes.data_as(dp), out.ctypes.data_as(dp))
        G = out[:, 0::2] + 1j * out[:, 1::2]
        ref = np.array(c["GBarVD"]); ref = ref[:, 0::2] + 1j * ref[:, 1::2]
        sc = np.abs(ref).max(axis=0)
        err = (np.abs(G[:, :ref.shape[1]] - ref) / (np.abs(ref) + 1e-6 * sc)).max()
        assert err <= 1e-11, (c["k"], c["kb"], err)
        n += 1
    assert n >= 4
