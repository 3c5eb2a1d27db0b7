// Integrands of the 33 frequency-independent panel-panel integrals (FIPPIs) of one panel pair:
//   xMxpRM3[3] = int int (x-x')/r^3, xXxpRM3[3] = int int (x cross x')/r^3,
//   uvupvpRM1/R1/R2[9] = int int {1,u',v',u,uu',uv',v,vu',vv'} r^{-1, 1, 2}
// Reference: no common vertices — CFDIntegrand3D QIFIPPICubature.cc:187-325 (innermost v' integral in
// closed form, remaining (u, v/u, u') in [0,1]^3 by hcubature, ABSTOL 1e-10 RELTOL 1e-6 MAXEVALS 1e6);
// common edge / vertex — TaylorDuffySum_FIPPI QIFIPPITaylorDuffyV2P0.cc:96-221 with
// GetAlphaBetaGamma_FIPPI :345-428, GetQFPIntegrals_FIPPI :430-480 and the ScriptP polynomial tables
// (generated_tables.h), hcubature in 1-D / 2-D with ABSTOL 0 RELTOL 1e-5 MAXFEVALS 10000.
// Decision-critical arithmetic is kept free of FMA contraction (hmul/hadd/hsub).
#pragma once
#include "hcubature.cuh"
#include "fippi_math.cuh"

namespace scuffcuda {

struct PXEntry { unsigned char np, d, n, jl, pow; double coef; };
struct PCEntry { unsigned char slot, d, n, jl, vec, pow; double coef; };
struct FIPPITables
{ const PXEntry *cePX; int cePXn; const PCEntry *cePC; int cePCn;
  const PXEntry *cvPX; int cvPXn; const PCEntry *cvPC; int cvPCn;
};

// ---------------------------------------------------------------------------------------
// ncv == 0
// ---------------------------------------------------------------------------------------
struct CFD3D
{ double V0[3], A[3], B[3], V0P[3], AP[3], BP[3];
  SCUFF_HD void setup(const double *const Va[3], const double *const Vb[3])
  { for (int m = 0; m < 3; m++)
     { V0[m] = Va[0][m]; A[m] = hsub(Va[1][m], Va[0][m]); B[m] = hsub(Va[2][m], Va[1][m]);
       V0P[m] = Vb[0][m]; AP[m] = hsub(Vb[1][m], Vb[0][m]); BP[m] = hsub(Vb[2][m], Vb[1][m]); }
  }
  SCUFF_HD void operator()(const double *x, double *fval) const
  { const double u = x[0], v = hmul(u, x[1]), up = x[2];
    double X[3], Y[3], Z[3], a2 = 0, BPdY = 0, Y2 = 0;
    for (int m = 0; m < 3; m++)
     { X[m] = hadd(hadd(V0[m], hmul(u, A[m])), hmul(v, B[m]));
       Z[m] = hadd(V0P[m], hmul(up, AP[m]));
       Y[m] = hsub(X[m], Z[m]);
       a2 = hadd(a2, hmul(BP[m], BP[m])); BPdY = hadd(BPdY, hmul(BP[m], Y[m])); Y2 = hadd(Y2, hmul(Y[m], Y[m]));
     }
    const double a = sqrt(a2), a3 = hmul(a, a2), vp0 = -BPdY / a2, vp02 = hmul(vp0, vp0), vp03 = hmul(vp02, vp0), vp04 = hmul(vp03, vp0);
    const double b2 = hsub(Y2 / a2, vp02);
    double OneRM3, vpRM3, OneRM1, vpRM1, OneR1, vpR1, OneR2, vpR2;
    if (b2 > 1.0e-10)
     { const double S1 = sqrt(hadd(b2, vp02));
       const double Sum = hadd(vp0, up);
       const double S2 = sqrt(hadd(b2, hmul(Sum, Sum)));
       const double LogFac = log(hadd(S2, hadd(up, vp0)) / hadd(S1, vp0));
       const double Sum3 = hmul(hmul(Sum, Sum), Sum), Sum4 = hmul(Sum3, Sum);
       OneRM3 = hmul(u, hsub(hadd(up, vp0) / S2, vp0 / S1)) / hmul(a3, b2);
       vpRM3 = hadd(hmul(-vp0, OneRM3), hmul(u, hsub(1.0 / S1, 1.0 / S2)) / a3);
       OneRM1 = hmul(u, LogFac) / a;
       vpRM1 = hadd(hmul(-vp0, OneRM1), hmul(u, hsub(S2, S1)) / a);
       OneR1 = hmul(hmul(hmul(u, a), 0.5), hsub(hadd(hmul(b2, LogFac), hmul(hadd(up, vp0), S2)), hmul(vp0, S1)));
       vpR1 = hadd(hmul(-vp0, OneR1), hmul(hmul(u, a), hsub(hmul(hmul(S2, S2), S2), hmul(hmul(S1, S1), S1))) / 3.0);
       OneR2 = hmul(hmul(u, a2), hadd(hsub(Sum3, vp03) / 3.0, hmul(up, b2)));
       vpR2 = hadd(hmul(-vp0, OneR2), hmul(hmul(u, a2), hadd(hsub(Sum4, vp04) / 4.0, hmul(hmul(up, hadd(vp0, hmul(0.5, up))), b2))));
     }
    else
     { const double q = hadd(vp0, up), q2 = hmul(q, q), q3 = hmul(q2, q), q4 = hmul(q3, q);
       OneRM3 = hmul(u, fabs(hsub(1.0 / vp02, 1.0 / q2))) / hmul(2.0, a3);
       vpRM3 = hadd(hmul(-vp0, OneRM3), hmul(u, fabs(hsub(1.0 / vp0, 1.0 / q))) / a3);
       OneRM1 = hmul(u, fabs(log(q / vp0))) / a;
       vpRM1 = hadd(hmul(-vp0, OneRM1), hmul(u, up) / a);
       OneR1 = hmul(hmul(u, a), fabs(hsub(q2, vp02))) / 2.0;
       OneR2 = hmul(hmul(u, a2), fabs(hsub(q3, vp03))) / 3.0;
       vpR1 = hadd(hmul(-vp0, OneR1), OneR2 / a);
       vpR2 = hadd(hmul(-vp0, OneR2), hmul(hmul(u, a2), fabs(hsub(q4, vp04))) / 4.0);
     }
    double CT[3], LT[3];
    CT[0] = hsub(hmul(X[1], Z[2]), hmul(X[2], Z[1])); CT[1] = hsub(hmul(X[2], Z[0]), hmul(X[0], Z[2])); CT[2] = hsub(hmul(X[0], Z[1]), hmul(X[1], Z[0]));
    LT[0] = hsub(hmul(X[1], BP[2]), hmul(X[2], BP[1])); LT[1] = hsub(hmul(X[2], BP[0]), hmul(X[0], BP[2])); LT[2] = hsub(hmul(X[0], BP[1]), hmul(X[1], BP[0]));
    int nf = 0;
    for (int m = 0; m < 3; m++) fval[nf++] = hsub(hmul(Y[m], OneRM3), hmul(BP[m], vpRM3));
    for (int m = 0; m < 3; m++) fval[nf++] = hadd(hmul(CT[m], OneRM3), hmul(LT[m], vpRM3));
    const double One[3] = { OneRM1, OneR1, OneR2 }, Vp[3] = { vpRM1, vpR1, vpR2 };
    for (int k = 0; k < 3; k++)
     { const double o = One[k], w = Vp[k];
       fval[nf++] = o;            fval[nf++] = hmul(up, o);           fval[nf++] = w;
       fval[nf++] = hmul(u, o);   fval[nf++] = hmul(hmul(u, up), o);  fval[nf++] = hmul(u, w);
       fval[nf++] = hmul(v, o);   fval[nf++] = hmul(hmul(v, up), o);  fval[nf++] = hmul(v, w);
     }
  }
};

// ---------------------------------------------------------------------------------------
// ncv == 1, 2: twice-integrable Taylor-Duffy reduction
// ---------------------------------------------------------------------------------------
SCUFF_HD void pqr_to_abg2(double P, double Q, double R, double &A, double &B, double &G2)
{ A = sqrt(P); B = Q / P;
  const double RoP = R / P;
  G2 = hsub(RoP, hmul(B, B));
  if (fabs(G2) < hmul(1.0e-6, RoP)) { B = 0.0; G2 = RoP; }
}

// int_0^1 [(y+P)^2+Q2]^{p/2} dy and int_0^1 y [...] dy for p = -3, -1, 1, 2
SCUFF_HD void qfp_integrals(double P, double Q2, double I[4], double Iy[4])
{ const double P2 = hmul(P, P), PP1 = hadd(P, 1.0);
  if (fabs(Q2) < 1.0e-8)
   { const double Sg = P > 0.0 ? 1.0 : -1.0;
     I[0] = hmul(Sg, hadd(P, 0.5)) / hmul(hmul(P2, PP1), PP1);
     Iy[0] = hadd(hmul(-P, I[0]), Sg / hmul(P, PP1));
     I[1] = hmul(Sg, log(hadd(1.0, 1.0 / P)));
     Iy[1] = hadd(hmul(-P, I[1]), hmul(Sg, 1.0));
     I[2] = hmul(Sg, hadd(P, 0.5));
     Iy[2] = hadd(hmul(-P, I[2]), hmul(Sg, hadd(hmul(P, PP1), 1.0 / 3.0)));
     I[3] = hadd(hmul(P, PP1), 1.0 / 3.0);
     Iy[3] = hadd(hadd(hadd(hadd(hmul(-P, I[3]), hmul(P, P2)), hmul(3.0 / 2.0, P2)), P), 0.25);
   }
  else
   { const double S2 = hadd(hmul(P, P), Q2), S = sqrt(S2), S3 = hmul(S, S2), S4 = hmul(S2, S2);
     const double T2 = hadd(hmul(PP1, PP1), Q2), T = sqrt(T2), T3 = hmul(T, T2), T4 = hmul(T2, T2);
     const double LogFac = log(hadd(PP1, T) / hadd(P, S));
     I[0] = hsub(PP1 / T, P / S) / Q2;
     Iy[0] = hsub(hadd(hmul(-P, I[0]), 1.0 / S), 1.0 / T);
     I[1] = LogFac;
     Iy[1] = hadd(hmul(-P, I[1]), hsub(T, S));
     I[2] = hadd(hadd(T, hmul(P, hsub(T, S))), hmul(Q2, LogFac)) / 2.0;
     Iy[2] = hadd(hmul(-P, I[2]), hsub(T3, S3) / 3.0);
     I[3] = hsub(hadd(T2, S2) / 2.0, 1.0 / 6.0);
     Iy[3] = hadd(hmul(-P, I[3]), hsub(T4, S4) / 4.0);
   }
}

struct TDFIPPI
{ const FIPPITables *T;
  int commonEdge;
  double vec[12][3];    // A,B,AP,BP,V0xAP,V0xBP,AxV0,BxV0,AxAP,AxBP,BxAP,BxBP
  double A2, B2, AP2, BP2, L2, AdB, AdAP, AdBP, AdL, BdAP, BdBP, APdBP, BPdL;

  // V1,V2,V3,V2P,V3P as ComputeQIFIPPIData_TaylorDuffyV2P0 receives them (OVa[0..2], OVb[1], OVb[2])
  SCUFF_HD void setup(const FIPPITables *tab, const double *V1, const double *V2, const double *V3, const double *V2P, const double *V3P)
  { T = tab;
    double *A = vec[0], *B = vec[1], *AP = vec[2], *BP = vec[3], L[3];
    for (int m = 0; m < 3; m++) { A[m] = hsub(V2[m], V1[m]); B[m] = hsub(V3[m], V2[m]); AP[m] = hsub(V2P[m], V1[m]); BP[m] = hsub(V3P[m], V2P[m]); }
    for (int m = 0; m < 3; m++) L[m] = hsub(BP[m], B[m]);
    A2 = dot3(A, A); B2 = dot3(B, B); AP2 = dot3(AP, AP); BP2 = dot3(BP, BP); L2 = dot3(L, L);
    AdB = dot3(A, B); AdAP = dot3(A, AP); AdBP = dot3(A, BP); AdL = dot3(A, L);
    BdAP = dot3(B, AP); BdBP = dot3(B, BP); APdBP = dot3(AP, BP); BPdL = dot3(BP, L);
    cross3(V1, AP, vec[4]); cross3(V1, BP, vec[5]); cross3(A, V1, vec[6]); cross3(B, V1, vec[7]);
    cross3(A, AP, vec[8]); cross3(A, BP, vec[9]); cross3(B, AP, vec[10]); cross3(B, BP, vec[11]);
    // VecEqualFloat(V2, V2P): common edge iff the second vertices agree to float precision
    commonEdge = ((float)V2[0] == (float)V2P[0]) && ((float)V2[1] == (float)V2P[1]) && ((float)V2[2] == (float)V2P[2]);
  }

  SCUFF_HD void operator()(const double *y, double *f) const
  { const int NumRegions = commonEdge ? 6 : 2, nOff = commonEdge ? 2 : 3;
    const int nMinC = 1, nMaxC = commonEdge ? 3 : 2, nMinX = 0, nMaxX = commonEdge ? 3 : 2;
    const double Jac = commonEdge ? y[0] : y[1];
    double Av[6], Bv[6], G2v[6];
    if (commonEdge)
     { const double y1 = y[0], y12 = y1 * y1;
       pqr_to_abg2((BP2 - 2.0 * BPdL + L2) * y12, y1 * (L2 + BPdL * (y1 - 1.0) - (AdL + L2 - AdBP) * y1),
                   L2 - 2 * AdL * y1 - 2 * L2 * y1 + A2 * y12 + 2 * AdL * y12 + L2 * y12, Av[0], Bv[0], G2v[0]);
       pqr_to_abg2(BP2 * y12, (BPdL + AdBP * y1 - BPdL * y1) * y1,
                   L2 + 2 * AdL * y1 - 2 * L2 * y1 + (A2 - 2 * AdL + L2) * y12, Av[1], Bv[1], G2v[1]);
       pqr_to_abg2((A2 + 2 * AdBP + BP2) * y12, y1 * (AdL * (-1.0 + y1) + BPdL * (-1.0 + y1) - (AdBP + BP2) * y1),
                   L2 + 2 * BPdL * y1 - 2 * L2 * y1 + BP2 * y12 - 2 * BPdL * y12 + L2 * y12, Av[2], Bv[2], G2v[2]);
       pqr_to_abg2((A2 + 2 * AdBP - 2 * AdL + BP2 - 2 * BPdL + L2) * y12, y1 * (AdL - L2 - (AdBP + BP2) * y1 + BPdL * (1 + y1)),
                   L2 - 2 * BPdL * y1 + BP2 * y12, Av[3], Bv[3], G2v[3]);
       pqr_to_abg2(A2 * y12, (AdBP * y1 - AdL) * y1, L2 - 2 * BPdL * y1 + BP2 * y12, Av[4], Bv[4], G2v[4]);
       pqr_to_abg2(A2 * y12, y1 * (AdL + AdBP * y1 - AdL * y1),
                   L2 + 2 * BPdL * y1 - 2 * L2 * y1 + BP2 * y12 - 2 * BPdL * y12 + L2 * y12, Av[5], Bv[5], G2v[5]);
     }
    else
     { const double y1 = y[0], y2 = y[1], y22 = y2 * y2;
       pqr_to_abg2(BP2 * y22, -AdBP * y2 - BdBP * y1 * y2 + APdBP * y22,
                   A2 + B2 * y1 * y1 + AP2 * y2 * y2 + 2.0 * AdB * y1 - 2.0 * AdAP * y2 - 2.0 * BdAP * y1 * y2, Av[0], Bv[0], G2v[0]);
       pqr_to_abg2(B2 * y22, AdB * y22 - BdAP * y2 - BdBP * y1 * y2,
                   A2 * y2 * y2 + AP2 + BP2 * y1 * y1 - 2 * AdAP * y2 - 2 * AdBP * y2 * y1 + 2 * APdBP * y1, Av[1], Bv[1], G2v[1]);
     }
    // J / L functions for the four kernels r^-3, r^-1, r, r^2
    double JM3[6][7], LM3[6][7], Jp[3][6][7], Lp[3][6][7];
    for (int d = 0; d < NumRegions; d++)
     { double I[4], Iy[4];
       qfp_integrals(Bv[d], G2v[d], I, Iy);
       const double a = Av[d], a2 = a * a, a3 = a2 * a;
       for (int n = nMinC + nOff; n <= nMaxC + nOff; n++)
        { JM3[d][n] = I[0] / (a3 * (1.0 + n - 3.0)); LM3[d][n] = Iy[0] / (a3 * (1.0 + n - 3.0)); }
       for (int n = nMinX + nOff; n <= nMaxX + nOff; n++)
        { Jp[0][d][n] = I[1] / (a * (1.0 + n - 1.0)); Lp[0][d][n] = Iy[1] / (a * (1.0 + n - 1.0));
          Jp[1][d][n] = a * I[2] / (1.0 + n + 1.0);   Lp[1][d][n] = a * Iy[2] / (1.0 + n + 1.0);
          Jp[2][d][n] = a2 * I[3] / (1.0 + n + 2.0);  Lp[2][d][n] = a2 * Iy[3] / (1.0 + n + 2.0); }
     }
    // ScriptP polynomials from the coefficient tables
    double PC[6][6][4][2], PX[9][6][4][2];
    for (int a = 0; a < 6; a++) for (int d = 0; d < NumRegions; d++) for (int n = 0; n < 4; n++) { PC[a][d][n][0] = 0; PC[a][d][n][1] = 0; }
    for (int a = 0; a < 9; a++) for (int d = 0; d < NumRegions; d++) for (int n = 0; n < 4; n++) { PX[a][d][n][0] = 0; PX[a][d][n][1] = 0; }
    double mon[4];
    if (commonEdge) { mon[0] = 1.0; mon[1] = y[0]; mon[2] = y[0] * y[0]; mon[3] = 0.0; }
    else { mon[0] = 1.0; mon[1] = y[0]; mon[2] = y[1]; mon[3] = y[0] * y[1]; }
    const PXEntry *px = commonEdge ? T->cePX : T->cvPX; const int npx = commonEdge ? T->cePXn : T->cvPXn;
    const PCEntry *pc = commonEdge ? T->cePC : T->cvPC; const int npc = commonEdge ? T->cePCn : T->cvPCn;
    for (int e = 0; e < npx; e++) { const PXEntry &E = px[e]; PX[E.np][E.d][E.n][E.jl] += E.coef * mon[E.pow]; }
    for (int e = 0; e < npc; e++)
     { const PCEntry &E = pc[e]; const double c = E.coef * mon[E.pow];
       for (int m = 0; m < 3; m++) PC[3 * E.slot + m][E.d][E.n][E.jl] += c * vec[E.vec][m]; }
    int ns = 0;
    for (int np = 0; np < 6; np++, ns++)
     { double s = 0;
       for (int d = 0; d < NumRegions; d++) for (int n = nMinC; n <= nMaxC; n++)
        s += Jac * (PC[np][d][n][0] * JM3[d][n + nOff] + PC[np][d][n][1] * LM3[d][n + nOff]);
       f[ns] = s; }
    for (int nk = 0; nk < 3; nk++)
     for (int np = 0; np < 9; np++, ns++)
      { double s = 0;
        for (int d = 0; d < NumRegions; d++) for (int n = nMinX; n <= nMaxX; n++)
         s += Jac * (PX[np][d][n][0] * Jp[nk][d][n + nOff] + PX[np][d][n][1] * Lp[nk][d][n + nOff]);
        f[ns] = s; }
  }
};

}  // namespace scuffcuda
