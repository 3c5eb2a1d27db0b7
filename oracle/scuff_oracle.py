"""TEST INFRASTRUCTURE ONLY — CPU (numpy) restatement of the reference's far-field assembly path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may import
this module; the product (scuff-em_b200/) never does and fails loudly without its CUDA library.

What is restated here (HomerReid/scuff-em, paths relative to the reference root), in the REFERENCE's
own formulation (one PPI per (panel pair, Q, Q') — NOT the Q-independent moment formulation the CUDA
kernels use, so that the two are independent derivations of the same numbers):

  * AssessPanelPair's rRel / far test                libs/libscuff/AssessPanelPair.cc:147-172
  * GetTCR(4) 6-point rule                           libs/libTriInt/TCR.cc:83-91
  * GetPPIs_Cubature + AssembleInnerPPIIntegrand     libs/libscuff/PanelPanelInteractions.cc:108-201,211-350
  * ExpRel                                           libs/libscuff/PanelPanelInteractions.cc:84-103
  * GetEdgeEdgeInteractions                          libs/libscuff/EdgeEdgeInteractions.cc:154-188
  * GSSIThread prefactors and stamping               libs/libscuff/SurfaceSurfaceInteractions.cc:165-176,203-327
  * symmetric fill                                   libs/libscuff/AssembleBEMMatrix.cc:753-758

Near / singular panel pairs (rRel <= 4: desingularisation + FIPPI, Taylor-Duffy) are NOT restated:
for those the checker is the UNMODIFIED reference itself, compiled by oracle/Makefile into
oracle/_ref (tests/conftest.py: ref_matrix).

Pinning: tests/test_oracle.py checks this restatement against (a) the reference's own known-answer
values for far panel pairs, unitTests/unit-test-PPIs.cc:70-79 (float precision, as the reference's
EqualFloat does), and (b) golden PPI / matrix-entry vectors produced by oracle/_ref and committed
under tests/golden/ by tests/golden/make_golden.py (<= 1e-13).
"""
import numpy as np

# GetTCR(4): libs/libTriInt/TCR.cc:83-91 (13-digit literals; the numbers are the contract)
TCR4 = np.array([[0.0915762135098, 0.0915762135098, 0.0549758718276],
                 [0.8168475729805, 0.0915762135098, 0.0549758718276],
                 [0.0915762135098, 0.8168475729805, 0.0549758718276],
                 [0.1081030181681, 0.4459484909160, 0.1116907948390],
                 [0.4459484909160, 0.1081030181681, 0.1116907948390],
                 [0.4459484909160, 0.4459484909160, 0.1116907948390]])
DESING_RADIUS = 4.0   # PanelPanelInteractions.cc:57


def parse_geom_dump(fn):
    """tests/golden/geom/*.geom.txt: the reference's own RWGGeometry dumped by oracle/_ref/ref_driver geom"""
    S, hdr, regions = [], {}, []
    for l in open(fn):
        t = l.split()
        if t[0] in ("NumSurfaces", "NumRegions", "TotalBFs", "LDim"):
            hdr[t[0]] = int(t[1])
        elif t[0] == "Region":
            regions.append((t[2], t[3]))
        elif t[0] == "Surface":
            S.append(dict(Label=t[2], IsPEC=bool(int(t[4])), Regions=(int(t[6]), int(t[7])), BFOffset=int(t[9]),
                          Mate=int(t[11]), NV=int(t[13]), NP=int(t[15]), NE=int(t[17]), NBF=int(t[19]), V=[], P=[], E=[]))
        elif t[0] in "VPE":
            S[-1][t[0]].append([float(x) for x in t[1:]])
    for s in S:
        V, P, E = np.array(s["V"]), np.array(s["P"]), np.array(s["E"])
        s.update(V=V, PVI=P[:, :3].astype(int), PC=P[:, 3:6], PR=P[:, 6], PA=P[:, 7],
                 ET=E[:, :8].astype(int), EL=E[:, 8])
    return hdr, regions, S


def exprel(x, n=4):
    """ExpRel(x,n) with the reference's truncation (PanelPanelInteractions.cc:84-103); scalar complex x"""
    term = 1.0 + 0j
    m = 1
    while m < n:
        term *= x / m
        m += 1
    s = 0j
    while m < 100:
        term *= x / m
        s += term
        if abs(term) ** 2 < 1.0e-16 * abs(s) ** 2:   # reference compares norms squared (abs(Term) < 1e-8 abs(Sum))
            break
        m += 1
    return s


def ppi_cubature(Va, Qa, Vb, Qb, k, desingularize=False):
    """GetPPIs_Cubature(order 4): H[0], H[1] of one (panel, Q ; panel', Q') pair.  Va, Vb: (3,3) vertices."""
    ik = 1j * k
    A, B = Va[1] - Va[0], Va[2] - Va[0]
    AP, BP = Vb[1] - Vb[0], Vb[2] - Vb[0]
    H = np.zeros(2, dtype=complex)
    for (u, v, w) in TCR4:
        X = Va[0] + u * A + v * B
        F = X - Qa
        HI = np.zeros(2, dtype=complex)
        for (up, vp, wp) in TCR4:
            XP = Vb[0] + up * AP + vp * BP
            FP = XP - Qb
            R = X - XP
            hPlus = np.dot(F, FP) + 4.0 / (ik * ik)
            FxFP = np.cross(F, FP)
            r = np.sqrt(np.dot(R, R))
            Phi = (exprel(ik * r, 4) if desingularize else np.exp(ik * r)) / (4.0 * np.pi * r)
            if not np.isfinite(Phi.real):
                Phi = 0.0
            Phi *= wp
            Psi = Phi * (ik - 1.0 / r) / r
            HI[0] += hPlus * Phi
            HI[1] += np.dot(FxFP, R) * Psi
        H += w * HI
    return H


def rrel_matrix(Sa, Sb, displacement=None):
    """rRel = |Ca - Cb (- L)| / max(Ra, Rb) for all panel pairs (AssessPanelPair.cc:161-166)"""
    D = Sa["PC"][:, None, :] - Sb["PC"][None, :, :]
    if displacement is not None:
        D = D - np.asarray(displacement)[None, None, :]
    r2 = (D[..., 0] * D[..., 0] + D[..., 1] * D[..., 1]) + D[..., 2] * D[..., 2]
    return np.sqrt(r2) / np.maximum(Sa["PR"][:, None], Sb["PR"][None, :])


def _panel_points(S):
    """cubature points X[p, n, :] and weights of every panel, formed as GetPPIs_Cubature does"""
    V0 = S["V"][S["PVI"][:, 0]]
    A = S["V"][S["PVI"][:, 1]] - V0
    B = S["V"][S["PVI"][:, 2]] - V0
    X = V0[:, None, :] + TCR4[None, :, 0, None] * A[:, None, :] + TCR4[None, :, 1, None] * B[:, None, :]
    return X, TCR4[:, 2]


def far_ppis(Sa, Sb, k, displacement=None):
    """H0[pa,pb,iQa,iQb], H1[...] for ALL panel pairs by 6x6 cubature (vectorised GetPPIs_Cubature),
    plus the far mask rRel > 4.  Memory O(NPa*NPb*36): meant for meshes of a few hundred panels."""
    ik = 1j * k
    Xa, w = _panel_points(Sa)
    Xb, _ = _panel_points(Sb)
    if displacement is not None:
        Xb = Xb + np.asarray(displacement)[None, None, :]
    R = Xa[:, None, :, None, :] - Xb[None, :, None, :, :]            # (Pa,Pb,6,6,3)
    r = np.sqrt(np.einsum("abpqm,abpqm->abpq", R, R))
    with np.errstate(divide="ignore", invalid="ignore"):
        Phi = np.exp(ik * r) / (4.0 * np.pi * r)
        Phi = np.where(np.isfinite(Phi.real), Phi, 0.0)
        ww = w[:, None] * w[None, :]
        Phi = Phi * ww[None, None]
        Psi = np.where(r > 0, Phi * (ik - 1.0 / r) / r, 0.0)
    NPa, NPb = Xa.shape[0], Xb.shape[0]
    H0 = np.zeros((NPa, NPb, 3, 3), dtype=complex)
    H1 = np.zeros((NPa, NPb, 3, 3), dtype=complex)
    Qa = Sa["V"][Sa["PVI"]]                                          # (Pa,3,3)
    Qb = Sb["V"][Sb["PVI"]]
    if displacement is not None:
        Qb = Qb + np.asarray(displacement)[None, None, :]
    for i in range(3):
        F = Xa - Qa[:, i, None, :]                                   # (Pa,6,3)
        for j in range(3):
            FP = Xb - Qb[:, j, None, :]                              # (Pb,6,3)
            hPlus = np.einsum("apm,bqm->abpq", F, FP) + 4.0 / (ik * ik)
            FxFP = np.cross(F[:, None, :, None, :], FP[None, :, None, :, :])
            hTimes = np.einsum("abpqm,abpqm->abpq", FxFP, R)
            H0[:, :, i, j] = np.einsum("abpq,abpq->ab", hPlus, Phi)
            H1[:, :, i, j] = np.einsum("abpq,abpq->ab", hTimes, Psi)
    return H0, H1, rrel_matrix(Sa, Sb, displacement) > DESING_RADIUS


def edge_edge_far(Sa, Sb, k, displacement=None):
    """GC0[nea,neb], GC1[nea,neb] (EdgeEdgeInteractions.cc:154-188) using ONLY far panel pairs
    (non-far PPIs contribute 0), and allfar[nea,neb] = all (up to 4) panel pairs of the edge pair are far."""
    H0, H1, far = far_ppis(Sa, Sb, k, displacement)
    H0 = np.where(far[:, :, None, None], H0, 0.0)
    H1 = np.where(far[:, :, None, None], H1, 0.0)
    Ea, Eb = Sa["ET"], Sb["ET"]
    GC0 = np.zeros((len(Ea), len(Eb)), dtype=complex)
    GC1 = np.zeros_like(GC0)
    allfar = np.ones(GC0.shape, dtype=bool)
    for (ca, ia, sa) in ((4, 6, +1.0), (5, 7, -1.0)):       # P panel (+), M panel (-)
        pa, qa = Ea[:, ca], Ea[:, ia]
        va = pa >= 0
        for (cb, ib, sb) in ((4, 6, +1.0), (5, 7, -1.0)):
            pb, qb = Eb[:, cb], Eb[:, ib]
            vb = pb >= 0
            m = va[:, None] & vb[None, :]
            h0 = H0[np.where(va, pa, 0)[:, None], np.where(vb, pb, 0)[None, :], np.where(va, qa, 0)[:, None], np.where(vb, qb, 0)[None, :]]
            h1 = H1[np.where(va, pa, 0)[:, None], np.where(vb, pb, 0)[None, :], np.where(va, qa, 0)[:, None], np.where(vb, qb, 0)[None, :]]
            GC0 += np.where(m, sa * sb * h0, 0.0)
            GC1 += np.where(m, sa * sb * h1, 0.0)
            allfar &= np.where(m, far[np.where(va, pa, 0)[:, None], np.where(vb, pb, 0)[None, :]], True)
    ll = Sa["EL"][:, None] * Sb["EL"][None, :]
    return ll * GC0, ll * GC1 / (1j * k), allfar


def csqrt2(z):
    """libs/libhrutil/libhrutil.cc:692-699: square root with non-negative imaginary part"""
    r = np.sqrt(complex(z))
    return -r if r.imag < 0 else r


def stamp_block(Sa, Sb, omega, eps, mu, sign, symmetric, displacement=None):
    """One region's contribution to the (Sa,Sb) block using far panel pairs only (GSSIThread
    SurfaceSurfaceInteractions.cc:165-176,203-266).  Returns (block, allfar_mask_on_BF_level)."""
    k = csqrt2(eps * mu) * omega
    GC0, GC1, allfar = edge_edge_far(Sa, Sb, k, displacement)
    PreFac1 = sign * 1j * mu * omega
    PreFac2 = -sign * 1j * k
    PreFac3 = -sign * 1j * eps * omega
    aP, bP = Sa["IsPEC"], Sb["IsPEC"]
    NA, NB = len(Sa["ET"]), len(Sb["ET"])
    if aP and bP:
        B = PreFac1 * GC0
        mask = allfar
    elif aP and not bP:
        B = np.zeros((NA, 2 * NB), dtype=complex)
        B[:, 0::2] = PreFac1 * GC0
        B[:, 1::2] = PreFac2 * GC1
        mask = np.repeat(allfar, 2, axis=1)
    elif not aP and bP:
        B = np.zeros((2 * NA, NB), dtype=complex)
        B[0::2, :] = PreFac1 * GC0
        B[1::2, :] = PreFac2 * GC1
        mask = np.repeat(allfar, 2, axis=0)
    else:
        B = np.zeros((2 * NA, 2 * NB), dtype=complex)
        B[0::2, 0::2] = PreFac1 * GC0
        B[0::2, 1::2] = PreFac2 * GC1
        B[1::2, 0::2] = PreFac2 * GC1
        B[1::2, 1::2] = PreFac3 * GC0
        mask = np.kron(allfar, np.ones((2, 2), dtype=bool))
    return B, mask


def far_matrix_single_surface(S, omega, eps_out=1.0, mu_out=1.0, eps_in=None, mu_in=1.0):
    """Far-pair-only BEM matrix of ONE closed surface: PEC (eps_in None) or PMCHWT dielectric
    (exterior region + interior region, both sign +1 on the diagonal block,
    CountCommonRegions SurfaceSurfaceInteractions.cc:62-90).  Equivalent to the CUDA path run with
    scuffcuda_ssi_args.far_only = 1.  Returns (M, mask of entries whose panel pairs are ALL far)."""
    M, mask = stamp_block(S, S, omega, eps_out, mu_out, +1.0, True)
    if eps_in is not None and not S["IsPEC"]:
        M2, _ = stamp_block(S, S, omega, eps_in, mu_in, +1.0, True)
        M = M + M2
    return M, mask


# ------------------------------------------------------------------------------------------------------------------
# Bookkeeping steps around the integrals (restated for the checker; the product implements them in
# scuff-em_b200/host/Assemble.cc and csrc/matrix_kernels.cu)
# ------------------------------------------------------------------------------------------------------------------
def rwg_overlap(S, nea, neb):
    """RWGSurface::GetOverlap, Overlaps[0] = int f_a . f_b (OPFT.cc:81-226): sum over the panels the two basis functions
    share of  sign * l_a l_b / (2A) * [ (L1.L1 + L1.L2)/4 + L1.DQ/3 + L2.L2/12 + L2.DQ/6 ]  with L1 = Qa' - Qa,
    L2 = Qa'' - Qa', DQ = Qa - Qb (Qa, Qb: the source/sink vertices of the two functions on that panel)."""
    Ea, Eb = S["ET"][nea], S["ET"][neb]
    LL = S["EL"][nea] * S["EL"][neb]

    def panel_term(npanel, iqa, iqb, sign):
        vi = S["PVI"][npanel]
        Qa, Qa1, Qa2, Qb = S["V"][vi[iqa]], S["V"][vi[(iqa + 1) % 3]], S["V"][vi[(iqa + 2) % 3]], S["V"][vi[iqb]]
        L1, L2, DQ = Qa1 - Qa, Qa2 - Qa1, Qa - Qb
        bullet = (L1 @ L1 + L1 @ L2) / 4.0 + (L1 @ DQ) / 3.0 + (L2 @ L2) / 12.0 + (L2 @ DQ) / 6.0
        return sign * LL / (2.0 * S["PA"][npanel]) * bullet

    # ET columns: iV1 iV2 iQP iQM iPPanel iMPanel PIndex MIndex
    O = 0.0
    if Ea[4] == Eb[4]:
        O += panel_term(Ea[4], Ea[6], Eb[6], +1.0)
    if Ea[4] == Eb[5]:
        O += panel_term(Ea[4], Ea[6], Eb[7], -1.0)
    if Ea[5] != -1 and Ea[5] == Eb[4]:
        O += panel_term(Ea[5], Ea[7], Eb[6], -1.0)
    if Ea[5] != -1 and Ea[5] == Eb[5]:
        O += panel_term(Ea[5], Ea[7], Eb[7], +1.0)
    return O


def stamp_in_neighbor_block(M, B, row_offset, col_offset, L, kbloch, use_symmetry):
    """StampInNeighborBlock (AssembleBEMMatrix.cc:72-114): M[block] += BPF * B (+ conj(BPF) * B^T), BPF = exp(i kBloch.L)"""
    bpf = np.exp(1j * (kbloch[0] * L[0] + kbloch[1] * L[1]))
    nr, nc = B.shape
    M[row_offset:row_offset + nr, col_offset:col_offset + nc] += bpf * B + (np.conj(bpf) * B.T if use_symmetry else 0.0)
    return M


def gbar_direct_sum(R, k, kbloch, lbv, exclude_inner=False, ncells=40):
    """Bloch-periodic scalar Green's function by brute-force lattice summation,
         GBar(R) = sum_L exp(i kBloch.L) exp(ik|R-L|) / (4 pi |R-L|),   L = n1 L1 (+ n2 L2),
    the quantity GBarVDEwald (GBarVDEwald.cc:931-996) evaluates by Ewald's method (sign convention R-L: the one that matches the
    reference's values to 1e-15 on rectangular lattices; on oblique lattices the reference's truncated Ewald sums are themselves
    ~1 % off this sum, tests/golden/gbar_golden.json keeps such a case flagged).  Only usable as a checker where the sum converges absolutely, i.e.
    Im k > 0 (exp(-kappa |L|) decay).  exclude_inner drops the cells |n1|,|n2| <= 1 (`ExcludeInnerCells`, the part the
    periodic assembly integrates directly, AssembleBEMMatrix.cc:465-519).  Returns (GBar, grad GBar[3])."""
    R = np.asarray(R, float); k = complex(k)
    lbv = [np.asarray(l, float) for l in lbv]
    rng = range(-ncells, ncells + 1)
    G = 0.0j; dG = np.zeros(3, complex)
    for n1 in rng:
        for n2 in (rng if len(lbv) == 2 else (0,)):
            if exclude_inner and abs(n1) <= 1 and abs(n2) <= 1:
                continue
            L = n1 * lbv[0] + (n2 * lbv[1] if len(lbv) == 2 else 0.0)
            X = R - np.array([L[0], L[1], 0.0])
            r = np.sqrt(X @ X)
            phi = np.exp(1j * (kbloch[0] * L[0] + kbloch[1] * L[1])) * np.exp(1j * k * r) / (4.0 * np.pi * r)
            G += phi
            dG += phi * (1j * k - 1.0 / r) * X / r
    return G, dG


def far_edge_pairs_gc0_extended(S, nea, neb, k):
    """GC[0] = l_a l_b (H++ - H+- - H-+ + H--)[0] (EdgeEdgeInteractions.cc:154-188) of a batch of FAR edge pairs of one surface in
    80-bit extended precision (np.longdouble), in the reference's formulation (6x6 rule, hPlus = F.F' + 4/(ik)^2): the
    yardstick for how much of a double-precision difference is rounding noise of the ill-conditioned sum (for k h << 1 the
    four panel-pair terms cancel to ~ (k h)^2 of their size).  Returns (GC0 [complex longdouble], far mask): entries whose
    four panel pairs are not all far (rRel > 4) are flagged False and must be ignored.
    S: dict with V, PVI, PC, PR, ET, EL as returned by pyscuff RWGGeometry.surface() / parse_geom_dump."""
    LD, CLD = np.longdouble, np.clongdouble
    nea = np.asarray(nea); neb = np.asarray(neb)
    V = S["V"].astype(LD); T = TCR4.astype(LD)
    ik = CLD(1j) * CLD(k)
    c4 = CLD(4.0) / (ik * ik)
    ET = S["ET"]
    out = np.zeros(len(nea), dtype=CLD)
    far = np.ones(len(nea), dtype=bool)
    fourpi = LD(4.0) * np.arccos(LD(-1.0))
    for sa, (ipa, iqa) in enumerate(((4, 2), (5, 3))):          # P panel / Q = iQP ; M panel / Q = iQM
        for sb, (ipb, iqb) in enumerate(((4, 2), (5, 3))):
            pa, pb = ET[nea, ipa], ET[neb, ipb]
            ok = (pa >= 0) & (pb >= 0)
            pa_, pb_ = np.where(ok, pa, 0), np.where(ok, pb, 0)
            D = S["PC"][pa_] - S["PC"][pb_]
            rrel = np.sqrt((D * D).sum(1)) / np.maximum(S["PR"][pa_], S["PR"][pb_])
            far &= (~ok) | (rrel > DESING_RADIUS)
            Qa, Qb = V[ET[nea, iqa]], V[ET[neb, iqb]]
            Va, Vb = V[S["PVI"][pa_]], V[S["PVI"][pb_]]                      # (m,3,3)
            A, B = Va[:, 1] - Va[:, 0], Va[:, 2] - Va[:, 0]
            AP, BP = Vb[:, 1] - Vb[:, 0], Vb[:, 2] - Vb[:, 0]
            X = Va[:, None, 0, :] + T[None, :, 0, None] * A[:, None, :] + T[None, :, 1, None] * B[:, None, :]      # (m,6,3)
            XP = Vb[:, None, 0, :] + T[None, :, 0, None] * AP[:, None, :] + T[None, :, 1, None] * BP[:, None, :]
            F, FP = X - Qa[:, None, :], XP - Qb[:, None, :]
            R = X[:, :, None, :] - XP[:, None, :, :]
            r = np.sqrt((R * R).sum(-1))
            with np.errstate(divide="ignore", invalid="ignore"):
                Phi = np.exp(ik * r.astype(CLD)) / (fourpi * r)
            hPlus = (F[:, :, None, :] * FP[:, None, :, :]).sum(-1).astype(CLD) + c4
            ww = (T[:, 2][:, None] * T[:, 2][None, :])[None]
            H0 = (hPlus * Phi * ww).sum((1, 2))
            sign = (1 if sa == 0 else -1) * (1 if sb == 0 else -1)
            out += np.where(ok, sign * H0, 0)
    return S["EL"][nea].astype(LD) * S["EL"][neb].astype(LD) * out, far


# ------------------------------------------------------------------------------------------------------------------
# Scattering-run callers: the reference's formulation per basis function / per pair of basis
# functions, vectorised — independent of the surface-current formulation of the CUDA kernels (csrc/rhs_pft.cu).
#   AssembleRHSVector / GetInnerProducts      libs/libscuff/AssembleRHSVector.cc:172-229,282-345
#   PlaneWave::GetFields                      libs/libIncField/PlaneWave.cc:75-88
#   GetScatteredPFTIntegrals (method 0) +     libs/libscuff/EMTPFT.cc:82-176,308-334,640-725 (power only)
#   GetExtinctionPFTT                         libs/libscuff/EMTPFT.cc:389-531 (power only)
#   GetBFBFCubature2 / GetPanelCubature       libs/libscuff/PanelCubature.cc:155-236,791-893
# Pinned by tests/test_scatter_cpu.py against the committed reference vectors tests/golden/scatter_golden.npz.
# ------------------------------------------------------------------------------------------------------------------
ZVAC = 376.73031346177


def load_rule(header, name, npts):
    """(u, v, w) table of a triangle rule as extracted from the compiled reference (scuff-em_b200/csrc/generated_tcr*.h)"""
    import os
    import re
    src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "scuff-em_b200", "csrc", header)).read()
    body = re.search(r"%s\[\d+\] = \{(.*?)\};" % re.escape(name), src, flags=re.S).group(1)
    return np.array([float(x) for x in body.replace("\n", " ").split(",") if x.strip()]).reshape(npts, 3)


def planewave_fields(X, omega, E0, nhat, eps=1.0, mu=1.0):
    k = np.sqrt(eps * mu + 0j) * omega
    Z = ZVAC * np.sqrt((mu + 0j) / eps)
    E = np.asarray(E0, dtype=complex)[None, :] * np.exp(1j * k * (X @ np.asarray(nhat, dtype=float)))[:, None]
    return E, np.cross(np.asarray(nhat, dtype=float)[None, :], E) / Z


def rhs_planewave(S, omega, E0, nhat, eps=1.0, mu=1.0, sign=-1.0):
    """RHS entries of one surface for a plane wave whose sources lie in the surface's exterior region (sign -1): 78-point rule over
    the positive panel (QP, V1, V2) minus the negative panel (V1, V2, QM); E-entries divided by ZVAC, H-entries for dielectric surfaces"""
    T = load_rule("generated_tcr20.h", "SC_TCR20_H", 78)
    V, ET, EL = S["V"], S["ET"], S["EL"]
    out = np.zeros(S["NBF"], dtype=complex)
    for ne in range(S["NE"]):
        iV1, iV2, iQP, iQM = ET[ne, :4]
        acc = np.zeros(2, dtype=complex)
        for pm, (A0, A1, A2, Q) in enumerate(((V[iQP], V[iV1], V[iV2], V[iQP]), (V[iV1], V[iV2], V[iQM] if iQM >= 0 else None, V[iQM] if iQM >= 0 else None))):
            if Q is None:
                continue
            X = A0[None, :] + T[:, 0, None] * (A1 - A0)[None, :] + T[:, 1, None] * (A2 - A0)[None, :]
            E, H = planewave_fields(X, omega, E0, nhat, eps, mu)
            f = X - Q[None, :]
            w = T[:, 2] * EL[ne] * (1.0 if pm == 0 else -1.0)       # PreFac x Jacobian x w = (l/2A)(2A) w
            acc[0] += np.sum(w * np.sum(f * E, axis=1)); acc[1] += np.sum(w * np.sum(f * H, axis=1))
        if S["IsPEC"]:
            out[ne] = sign * acc[0] / ZVAC
        else:
            out[2 * ne] = sign * acc[0] / ZVAC; out[2 * ne + 1] = sign * acc[1]
    return out


def _bf_points(S, ne, rule):
    """cubature points, basis-function values and weights of one RWG function as GetBFBFCubature2 forms them: both panels
    parametrised from their source vertex, X = Q + u (V1 - Q) + v (V2 - Q)"""
    V, ET, EL = S["V"], S["ET"], S["EL"]
    iV1, iV2, iQP, iQM = ET[ne, :4]
    X, b, div, W = [], [], [], []
    for sgn, iQ in ((1.0, iQP), (-1.0, iQM)):
        if iQ < 0:
            continue
        Q, A, B = V[iQ], V[iV1] - V[iQ], V[iV2] - V[iQ]
        area2 = np.linalg.norm(np.cross(A, B))
        pre = sgn * EL[ne] / area2
        d = rule[:, 0, None] * A[None, :] + rule[:, 1, None] * B[None, :]
        X.append(d + Q[None, :]); b.append(pre * d); div.append(np.full(len(rule), 2.0 * pre)); W.append(area2 * rule[:, 2])
    return np.vstack(X), np.vstack(b), np.concatenate(div), np.concatenate(W)


def emtpft_power_single_surface(S, KN, omega, E0, nhat, eps=1.0, mu=1.0):
    """(PAbs, PScat) of a single closed surface in a medium (eps, mu) by the energy/momentum-transfer method with the cubature
    integrals (EMTPFTIMethod 0), exterior expansion: PScat from all ordered pairs of basis functions (same surface: only the
    imaginary parts of the kernels, order 9 if the two functions share a vertex, else 4), PAbs = extinction - PScat."""
    T4, T9 = TCR4, load_rule("generated_tcr9.h", "SC_TCR9_H", 21)
    NE = S["NE"]
    k = float(np.real(np.sqrt(eps * mu + 0j) * omega))
    pec = S["IsPEC"]
    kc = KN if pec else KN[0::2]
    nc = np.zeros(NE, dtype=complex) if pec else -ZVAC * KN[1::2]
    pts = {4: [_bf_points(S, ne, T4) for ne in range(NE)], 9: [_bf_points(S, ne, T9) for ne in range(NE)]}
    verts = [set(int(v) for v in S["ET"][ne, :4] if v >= 0) for ne in range(NE)]

    def kernels(r):
        kr = k * r
        small = np.abs(kr) < 1e-3
        rs = np.where(small, 1.0, r)
        phi = np.where(small, (1.0 - kr * kr / 6.0) * k / (4 * np.pi), np.sin(kr) / (4 * np.pi * rs))
        psi = np.where(small, -(1.0 - kr * kr / 10.0) * k ** 3 / (12 * np.pi), (kr * np.cos(kr) - np.sin(kr)) / (4 * np.pi * rs ** 3))
        return phi, psi

    pscat = 0.0
    for a in range(NE):
        for order in (4, 9):
            bs = [b for b in range(NE) if (len(verts[a] & verts[b]) > 0) == (order == 9)]
            if not bs:
                continue
            Xa, ba, da, Wa = pts[order][a]
            Xb = np.vstack([pts[order][b][0] for b in bs]); bb = np.vstack([pts[order][b][1] for b in bs])
            db = np.concatenate([pts[order][b][2] for b in bs]); Wb = np.concatenate([pts[order][b][3] for b in bs])
            starts = np.concatenate([[0], np.cumsum([len(pts[order][b][3]) for b in bs])[:-1]])       # column ranges of the functions b
            R = Xa[:, None, :] - Xb[None, :, :]
            r = np.linalg.norm(R, axis=2)
            phi, psi = kernels(r)
            W = Wa[:, None] * Wb[None, :]
            pefie = ba @ bb.T - np.outer(da, db) / (k * k)
            pmfie = np.einsum("ijm,ijm->ij", np.cross(ba[:, None, :], bb[None, :, :]), R)                 # (bA x bB).R
            qkk = np.add.reduceat((W * pefie * omega * mu * phi).sum(axis=0), starts)                      # Im QKK (= Im QNN mu/eps)
            qkn = np.add.reduceat((W * pmfie * psi).sum(axis=0), starts)                                   # Im QKNmNK
            kb, nb = kc[bs], nc[bs]
            u0kk = np.conj(kc[a]) * kb * ZVAC
            e0nn = np.conj(nc[a]) * nb / ZVAC
            knmnk = np.conj(kc[a]) * nb - np.conj(nc[a]) * kb
            pscat += 0.5 * np.sum(u0kk.real * qkk + e0nn.real * (qkk * eps / mu) + knmnk.imag * qkn)
    # extinction: 21-point rule of every panel in the panel's own vertex order, per basis function
    pext = 0.0
    V, PVI = S["V"], S["PVI"]
    for ne in range(NE):
        iV1, iV2, iQP, iQM = S["ET"][ne, :4]
        for sgn, npan, iQ in ((1.0, S["ET"][ne, 4], iQP), (-1.0, S["ET"][ne, 5], iQM)):
            if iQ < 0 or npan < 0:
                continue
            V0, V1, V2 = V[PVI[npan, 0]], V[PVI[npan, 1]], V[PVI[npan, 2]]
            area2 = np.linalg.norm(np.cross(V1 - V0, V2 - V0))
            X = V0[None, :] + T9[:, 0, None] * (V1 - V0)[None, :] + T9[:, 1, None] * (V2 - V0)[None, :]
            b = sgn * S["EL"][ne] / area2 * (X - V[iQ][None, :])
            E, H = planewave_fields(X, omega, E0, nhat, eps, mu)
            w = area2 * T9[:, 2]
            pext += 0.5 * np.real(np.conj(kc[ne]) * np.sum(w * np.sum(b * E, axis=1)) + np.conj(nc[ne]) * np.sum(w * np.sum(b * H, axis=1)))
    return pext - pscat, pscat
