// Bloch-periodic Helmholtz Green's function by Ewald summation, values and derivatives, for 2-D lattices:
//   GBar(R) = sum_L exp(i kBloch.L) exp(ik|R-L|) / (4 pi |R-L|)      (optionally without the 3x3 innermost cells)
// Restates the reference's GBarVDEwald (libs/libscuff/GBarVDEwald.cc:931-996): same Ewald parameter (GetRLBasis
// :186-240), same real-space (GetGBarNearby :742-822, AddGShort :636-740) and spectral (GetGBarDistant :465-558,
// AddGLong2D :298-330, GetEEF :268-293) shell order and the same stopping rule (three consecutive shells that change
// no component by more than 1e-8 relative), the singular-kBloch nudge (:969-975) and the real-space removal of the inner
// cells' long-range part (AddGLongRealSpace :824-900).  The truncation error of the reference's sums (1e-8) is part of
// the contract, so the order of summation is kept; individual terms only need to agree to rounding.
//
// exp(a) erfc(b) for complex b goes through the Faddeeva function w(z) = exp(-z^2) erfc(-iz), Im z >= 0, evaluated
// with Weideman's N=40 rational expansion (generated_faddeeva.h; 2.5e-14 relative against the Faddeeva package the
// reference vendors) — branch-free, 40 complex FMAs, which is what a GPU thread wants.
//
// SCUFF_HD: the same inline source is compiled by g++ into tests/host_harness.cc so that the restatement can be
// checked against the reference on a machine without a GPU.  The product path only ever calls the device
// instantiation (gbar.cu, pbc_kernel.cu).
#pragma once
#include "gbar_params.h"
#include "generated_faddeeva.h"

namespace scuffcuda {

#if defined(__CUDACC__)
__constant__ double c_weid[SC_WEID_N] = SC_WEID_COEFFS;
#endif
static const double h_weid[SC_WEID_N] = SC_WEID_COEFFS;

#define SC_GB_NSUM 8
#define SC_GB_NMAX 10000
#define SC_GB_RELTOL 1.0e-8
#define SC_GB_ABSTOL 0.0


// ---- complex helpers (principal branches as in C99 / libstdc++, signed zeros respected) ----
SCUFF_HD cd gb_cexp(cd z)
{ const double e = exp(z.x); double s, c; sincos(z.y, &s, &c); return mk(e * c, e * s); }
SCUFF_HD double gb_cabs(cd z) { return hypot(z.x, z.y); }
SCUFF_HD cd gb_csqrt(cd z)
{ if (z.x == 0.0 && z.y == 0.0) return mk(0.0, z.y);
  const double r = hypot(z.x, z.y);
  if (z.x >= 0.0) { const double t = sqrt(0.5 * (r + z.x)); return mk(t, z.y / (2.0 * t)); }
  const double t = sqrt(0.5 * (r - z.x));
  return mk(fabs(z.y) / (2.0 * t), copysign(t, z.y));
}
SCUFF_HD bool gb_cfinite(cd z) { return isfinite(z.x) && isfinite(z.y); }
SCUFF_HD cd cmulI(cd z) { return mk(-z.y, z.x); }       // i z

// Faddeeva function w(z), Im z >= 0 (Weideman 1994, N = 40)
SCUFF_HD cd faddeeva_w_upper(cd z)
{
#if defined(__CUDA_ARCH__)
  const double *a = c_weid;
#else
  const double *a = h_weid;
#endif
  const double L = SC_WEID_L;
  const cd iz = cmulI(z);
  const cd den = mk(L - iz.x, -iz.y);                     // L - iz
  const cd Z = cdiv(mk(L + iz.x, iz.y), den);
  cd p = mk(a[0], 0.0);
#pragma unroll 8
  for (int n = 1; n < SC_WEID_N; n++) { const cd t = p * Z; p = mk(t.x + a[n], t.y); }
  const cd id = cinv(den);
  const cd id2 = id * id;
  return 2.0 * (p * id2) + 0.5641895835477562869 * id;    // 1/sqrt(pi)
}
// any z: w(z) = 2 exp(-z^2) - w(-z) in the lower half plane
SCUFF_HD cd faddeeva_w(cd z)
{ if (z.y >= 0.0) return faddeeva_w_upper(z);
  const cd mz2 = mk((z.y - z.x) * (z.x + z.y), -2.0 * z.x * z.y);
  return 2.0 * gb_cexp(mz2) - faddeeva_w_upper(-z);
}
// erfcx(b) = exp(b^2) erfc(b) = w(ib)
SCUFF_HD cd gb_cerfcx(cd b) { return faddeeva_w(cmulI(b)); }

// exp(a) erfc(b) without overflow (erfc_s, GBarVDEwald.cc:246-257)
SCUFF_HD cd erfc_s(cd a, cd b)
{ const double x = b.x, y = b.y;
  const cd mb2 = mk((y - x) * (x + y), -2.0 * x * y);     // -b^2
  if (x >= 0.0) return gb_cexp(a + mb2) * gb_cerfcx(b);
  return 2.0 * gb_cexp(a) - gb_cexp(a + mb2) * gb_cerfcx(-b);
}

// EEF = e^{QZ} erfc(Q/2E + EZ) + e^{-QZ} erfc(Q/2E - EZ) and its Z derivative (GetEEF, GBarVDEwald.cc:268-291)
SCUFF_HD void gbar_eef(double z, double E, cd Q, cd &EEF, cd &EEFPrime)
{ const double c = 1.1283791670955125739 * E;             // M_2_SQRTPI * E
  const cd hQ = mk(0.5 * Q.x / E, 0.5 * Q.y / E);
  const cd Qz = Q * z;
  cd Arg = mk(hQ.x + z * E, hQ.y);
  const cd Plus = erfc_s(Qz, Arg);
  const cd dPlus = Q * Plus - c * gb_cexp(Qz - Arg * Arg);
  Arg = mk(hQ.x - z * E, hQ.y);
  const cd Minus = erfc_s(-Qz, Arg);
  const cd dMinus = c * gb_cexp(-Qz - Arg * Arg) - Q * Minus;
  EEF = Plus + Minus;
  EEFPrime = dPlus + dMinus;
  if (!gb_cfinite(EEF)) EEF = mk(0.0, 0.0);
  if (!gb_cfinite(EEFPrime)) EEFPrime = mk(0.0, 0.0);
}

// one reciprocal-lattice vector of the spectral sum (AddGLong2D, GBarVDEwald.cc:298-330)
SCUFF_HD void gbar_add_long2d(const double R[3], cd k, const double P[2], int n1, int n2, const double Gamma[2][2], double E,
                              cd *Sum, bool &Singular)
{ if (Singular) return;
  const double PmG0 = P[0] - n1 * Gamma[0][0] - n2 * Gamma[1][0];
  const double PmG1 = P[1] - n1 * Gamma[0][1] - n2 * Gamma[1][1];
  const cd k2 = k * k;
  const cd Q = gb_csqrt(mk(PmG0 * PmG0 + PmG1 * PmG1 - k2.x, -k2.y));
  if (gb_cabs(Q) < 1.0e-4 * gb_cabs(k)) { Singular = true; return; }
  const double ph = PmG0 * R[0] + PmG1 * R[1];
  double s, c; sincos(ph, &s, &c);
  const cd PreFactor = cdiv(mk(c, s), Q);
  cd EEF, EEFPrime;
  gbar_eef(R[2], E, Q, EEF, EEFPrime);
  const cd T = PreFactor * EEF, TP = PreFactor * EEFPrime;
  Sum[0] += T;
  Sum[1] += cmulI(T) * PmG0;
  Sum[2] += cmulI(T) * PmG1;
  Sum[3] += TP;
  Sum[4] += T * (-PmG0 * PmG1);
  Sum[5] += cmulI(TP) * PmG0;
  Sum[6] += cmulI(TP) * PmG1;
  Sum[7] += TP * (-PmG0 * PmG1);
}

// one real-space cell of the FULL periodic Green's function, values and mixed derivatives (AddGFull, :566-603)
SCUFF_HD void gbar_add_full(const double R[3], cd k, const double kBloch[2], double Lx, double Ly, cd *Sum)
{ double s, c; sincos(Lx * kBloch[0] + Ly * kBloch[1], &s, &c);
  const cd Phase = mk(c, s);
  const double X = R[0] - Lx, Y = R[1] - Ly, Z = R[2];
  const double r2 = X * X + Y * Y + Z * Z, r = sqrt(r2);
  if (r < 1.0e-8) return;
  const cd IKR = mk(-k.y * r, k.x * r);
  const cd Phi = gb_cexp(IKR) * (1.0 / (4.0 * M_PI * r));
  Sum[0] += Phase * Phi;
  const cd Psi = (mk(IKR.x - 1.0, IKR.y) * Phi) * (1.0 / r2);
  cd t = mk(IKR.x - 3.0, IKR.y);                     // -3 + IKR
  t = IKR * t; t.x += 3.0;                           // 3 + IKR(-3 + IKR)
  const cd Zeta = (t * Phi) * (1.0 / (r2 * r2));
  t = mk(IKR.x - 6.0, IKR.y); t = IKR * t; t.x += 15.0; t = IKR * t; t.x -= 15.0;
  const cd Ups = (t * Phi) * (1.0 / (r2 * r2 * r2));
  const cd PPsi = Phase * Psi, PZeta = Phase * Zeta;
  Sum[1] += PPsi * X; Sum[2] += PPsi * Y; Sum[3] += PPsi * Z;
  Sum[4] += PZeta * (X * Y); Sum[5] += PZeta * (X * Z); Sum[6] += PZeta * (Y * Z);
  Sum[7] += (Phase * Ups) * (X * Y * Z);
}

// one real-space cell of the short-range (erfc-screened) sum (AddGShort, GBarVDEwald.cc:636-737)
SCUFF_HD void gbar_add_short(const double R[3], cd k, const double kBloch[2], int n1, int n2, const double LBV[2][2], double E, cd *Sum)
{ const double L0 = n1 * LBV[0][0] + n2 * LBV[1][0], L1 = n1 * LBV[0][1] + n2 * LBV[1][1];
  if (E == 0.0) { gbar_add_full(R, k, kBloch, L0, L1, Sum); return; }
  double s, c; sincos(kBloch[0] * L0 + kBloch[1] * L1, &s, &c);
  const cd Phase = mk(c / (8.0 * M_PI), s / (8.0 * M_PI));
  const double X = R[0] - L0, Y = R[1] - L1, Z = R[2];
  const double rml2 = X * X + Y * Y + Z * Z, rml = sqrt(rml2);
  if (rml < 1.0e-6) return;
  const double rml3 = rml2 * rml, rml4 = rml3 * rml, rml5 = rml4 * rml, rml6 = rml5 * rml, rml7 = rml6 * rml;
  const double E2 = E * E, E4 = E2 * E2;
  const cd ik = cmulI(k), ikr = ik * rml;
  const cd hk = mk(ik.x / (2.0 * E), ik.y / (2.0 * E));               // i k / (2E)
  const cd gp = erfc_s(ikr, mk(E * rml + hk.x, hk.y));
  const cd gm = erfc_s(-ikr, mk(E * rml - hk.x, -hk.y));
  const cd ggP = gp + gm, ggM = gp - gm;
  const cd k2 = k * k;
  const cd g4 = (-2.0 * 1.1283791670955125739 * E) * gb_cexp(mk(-E2 * rml2 + k2.x / (4.0 * E2), k2.y / (4.0 * E2)));
  const cd g4ik = g4 + ik * ggM;
  Sum[0] += Phase * (ggP * (1.0 / rml));
  cd Term = g4ik * (1.0 / rml2) - ggP * (1.0 / rml3);
  cd PT = Phase * Term;
  Sum[1] += PT * X; Sum[2] += PT * Y; Sum[3] += PT * Z;
  Term = ggP * (3.0 / rml5) - g4ik * (3.0 / rml4) - (k2 * ggP) * (1.0 / rml3) - g4 * (2.0 * E2 / rml2);
  PT = Phase * Term;
  Sum[4] += PT * (X * Y); Sum[5] += PT * (X * Z); Sum[6] += PT * (Y * Z);
  Term = ggP * (-15.0 / rml7) + g4ik * (15.0 / rml6) + (k2 * ggP) * (6.0 / rml5) + g4 * (10.0 * E2 / rml4)
         - (k2 * g4ik) * (1.0 / rml4) + g4 * (4.0 * E4 / rml2);
  Sum[7] += (Phase * Term) * (X * Y * Z);
}

// maximal relative / absolute change of the eight sums in one shell (the reference's stopping test)
SCUFF_HD bool gbar_shell_converged(const cd *Sum, cd *Last)
{ double MaxRel = 0.0, MaxAbs = 0.0;
  for (int ns = 0; ns < SC_GB_NSUM; ns++)
   { const double Delta = gb_cabs(Sum[ns] - Last[ns]);
     if (Delta > MaxAbs) MaxAbs = Delta;
     const double AbsSum = gb_cabs(Sum[ns]);
     if (AbsSum > 0.0 && Delta > MaxRel * AbsSum) MaxRel = Delta / AbsSum;
     Last[ns] = Sum[ns]; }
  return MaxAbs < SC_GB_ABSTOL || MaxRel < SC_GB_RELTOL;
}

// real-space (short-range) lattice sum (GetGBarNearby, GBarVDEwald.cc:742-818)
SCUFF_HD void gbar_nearby(const double R[3], const GBarParams &P, const double kBloch[2], double E, cd *Sum)
{ for (int ns = 0; ns < SC_GB_NSUM; ns++) Sum[ns] = mk(0.0, 0.0);
  for (int n1 = -1; n1 <= 1; n1++)
   for (int n2 = -1; n2 <= 1; n2++)
    if (!P.excludeInner) gbar_add_short(R, P.k, kBloch, n1, n2, P.LBV, E, Sum);      // |n1|,|n2| <= 1 are all inner cells
  cd Last[SC_GB_NSUM];
  for (int ns = 0; ns < SC_GB_NSUM; ns++) Last[ns] = Sum[ns];
  int Conv = 0;
  for (int NN = 2; Conv < 3 && NN <= SC_GB_NMAX; NN++)
   { for (int n = -NN; n < NN; n++)
      { gbar_add_short(R, P.k, kBloch, n, NN, P.LBV, E, Sum);
        gbar_add_short(R, P.k, kBloch, NN, -n, P.LBV, E, Sum);
        gbar_add_short(R, P.k, kBloch, -n, -NN, P.LBV, E, Sum);
        gbar_add_short(R, P.k, kBloch, -NN, n, P.LBV, E, Sum); }
     if (gbar_shell_converged(Sum, Last)) Conv++; else Conv = 0; }
}

// spectral (long-range) lattice sum (GetGBarDistant, GBarVDEwald.cc:465-558); returns true if a lattice vector hits
// the light cone (Q ~ 0)
SCUFF_HD bool gbar_distant(const double R[3], const GBarParams &P, const double kBloch[2], const double Gamma[2][2], double E, cd *Sum)
{ for (int ns = 0; ns < SC_GB_NSUM; ns++) Sum[ns] = mk(0.0, 0.0);
  if (E == 0.0) return false;
  bool Singular = false;
  for (int m1 = -1; m1 <= 1; m1++)
   for (int m2 = -1; m2 <= 1; m2++)
    gbar_add_long2d(R, P.k, kBloch, m1, m2, Gamma, E, Sum, Singular);
  if (Singular) return true;
  cd Last[SC_GB_NSUM];
  for (int ns = 0; ns < SC_GB_NSUM; ns++) Last[ns] = Sum[ns];
  int Conv = 0;
  for (int NN = 2; Conv < 3 && NN <= SC_GB_NMAX; NN++)
   { for (int m = -NN; m < NN; m++)
      { gbar_add_long2d(R, P.k, kBloch, m, NN, Gamma, E, Sum, Singular);
        gbar_add_long2d(R, P.k, kBloch, NN, -m, Gamma, E, Sum, Singular);
        gbar_add_long2d(R, P.k, kBloch, -m, -NN, Gamma, E, Sum, Singular);
        gbar_add_long2d(R, P.k, kBloch, -NN, m, Gamma, E, Sum, Singular); }
     if (Singular) return true;
     if (gbar_shell_converged(Sum, Last)) Conv++; else Conv = 0; }
  const double PreFactor = (Gamma[0][0] * Gamma[1][1] - Gamma[0][1] * Gamma[1][0]) / (16.0 * M_PI * M_PI);
  for (int ns = 0; ns < SC_GB_NSUM; ns++) Sum[ns] = Sum[ns] * PreFactor;
  return false;
}

// long-range part of one real-space cell, = full - short, or its small-r expansion (AddGLongRealSpace, :824-900)
SCUFF_HD void gbar_add_long_realspace(const double R[3], const GBarParams &P, const double kBloch[2], int n1, int n2, double E, cd *Sum)
{ if (E == 0.0) return;
  const cd k = P.k;
  const double L0 = n1 * P.LBV[0][0] + n2 * P.LBV[1][0], L1 = n1 * P.LBV[0][1] + n2 * P.LBV[1][1];
  const double X = R[0] - L0, Y = R[1] - L1, Z = R[2];
  const double rml2 = X * X + Y * Y + Z * Z, rml = sqrt(rml2);
  const bool Smallr = (rml * E < 0.1) && (rml * gb_cabs(k) < 1.0);
  if (Smallr)
   { const double PI32 = 5.5683279968317078453;
     const cd k2 = k * k, k3 = k2 * k;
     const double E2 = E * E;
     const cd ExpFac = gb_cexp(mk(0.25 * k2.x / E2, 0.25 * k2.y / E2));
     // 1 + erf(ik/2E) = erfc(-ik/2E) = exp(k^2/4E^2) w(k/2E)
     const cd ErfFac = ExpFac * faddeeva_w(mk(k.x / (2.0 * E), k.y / (2.0 * E)));
     const cd ikE = cmulI(k) * ErfFac, ik3E = cmulI(k3) * ErfFac;
     const cd C0 = ExpFac * (E / (2.0 * PI32)) + ikE * (1.0 / (4.0 * M_PI));
     const cd C2 = -((ExpFac * E) * mk(2.0 * E2 + k2.x, k2.y)) * (1.0 / (12.0 * PI32)) - ik3E * (1.0 / (24.0 * M_PI));
     const cd k4 = k2 * k2;
     const cd C4 = ((ExpFac * E) * mk(12.0 * E2 * E2 + 2.0 * E2 * k2.x + k4.x, 2.0 * E2 * k2.y + k4.y)) * (1.0 / (240.0 * PI32))
                   + (ik3E * k2) * (1.0 / (480.0 * M_PI));
     double s, c; sincos(kBloch[0] * L0 + kBloch[1] * L1, &s, &c);
     const cd Phase = mk(c, s);
     Sum[0] += Phase * (C0 + C2 * rml2 + C4 * (rml2 * rml2));
     const cd g = C2 * 2.0 + C4 * (4.0 * rml2);
     Sum[1] += Phase * (g * X); Sum[2] += Phase * (g * Y); Sum[3] += Phase * (g * Z);
     Sum[4] += Phase * (C4 * (8.0 * X * Y)); Sum[5] += Phase * (C4 * (8.0 * X * Z)); Sum[6] += Phase * (C4 * (8.0 * Y * Z));
     return; }
  cd GFull[SC_GB_NSUM], GShort[SC_GB_NSUM];
  for (int ns = 0; ns < SC_GB_NSUM; ns++) { GFull[ns] = mk(0.0, 0.0); GShort[ns] = mk(0.0, 0.0); }
  gbar_add_full(R, k, kBloch, L0, L1, GFull);
  gbar_add_short(R, k, kBloch, n1, n2, P.LBV, E, GShort);
  for (int ns = 0; ns < SC_GB_NSUM; ns++) Sum[ns] += GFull[ns] - GShort[ns];
}

// GBar and its derivatives {1, dx, dy, dz, dxdy, dxdz, dydz, dxdydz} (GBarVDEwald, GBarVDEwald.cc:931-996), E = EOpt
SCUFF_HD void gbar_vd_ewald(const double R[3], const GBarParams &P, cd *GBarVD)
{ if (P.k.x == 0.0 && P.k.y == 0.0) { for (int ns = 0; ns < SC_GB_NSUM; ns++) GBarVD[ns] = mk(0.0, 0.0); return; }
  // reciprocal basis and Ewald parameter (GetRLBasis :221-238)
  const double Area = P.LBV[0][0] * P.LBV[1][1] - P.LBV[0][1] * P.LBV[1][0];
  double Gamma[2][2];
  Gamma[0][0] = 2.0 * M_PI * P.LBV[1][1] / Area;  Gamma[0][1] = -2.0 * M_PI * P.LBV[0][1] / Area;
  Gamma[1][0] = -2.0 * M_PI * P.LBV[1][0] / Area; Gamma[1][1] = 2.0 * M_PI * P.LBV[0][0] / Area;
  const double EOpt1 = sqrt(M_PI / Area);
  const double G12 = Gamma[0][0] * Gamma[0][0] + Gamma[0][1] * Gamma[0][1];
  const double G22 = Gamma[1][0] * Gamma[1][0] + Gamma[1][1] * Gamma[1][1];
  const double EOpt2 = sqrt(norm2(P.k) + G12 + G22) / 10.0;
  const double E = fmax(EOpt1, EOpt2);
  double kBloch[2] = { P.kBloch[0], P.kBloch[1] };
  cd Near[SC_GB_NSUM], Dist[SC_GB_NSUM];
  gbar_nearby(R, P, kBloch, E, Near);
  if (gbar_distant(R, P, kBloch, Gamma, E, Dist))
   { kBloch[0] += 1.0e-2 * gb_cabs(P.k);
     gbar_nearby(R, P, kBloch, E, Near);
     gbar_distant(R, P, kBloch, Gamma, E, Dist); }
  for (int ns = 0; ns < SC_GB_NSUM; ns++) GBarVD[ns] = Near[ns] + Dist[ns];
  if (P.excludeInner)
   { cd Inner[SC_GB_NSUM];
     for (int ns = 0; ns < SC_GB_NSUM; ns++) Inner[ns] = mk(0.0, 0.0);
     for (int n1 = -1; n1 <= 1; n1++)
      for (int n2 = -1; n2 <= 1; n2++)
       gbar_add_long_realspace(R, P, kBloch, n1, n2, E, Inner);
     for (int ns = 0; ns < SC_GB_NSUM; ns++) GBarVD[ns] -= Inner[ns]; }
}

}  // namespace scuffcuda
