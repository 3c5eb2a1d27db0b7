// Taylor-Duffy for panel pairs with ONE or TWO common vertices (TD_COMMONVERTEX / TD_COMMONEDGE), used by the reference
// in the short-wavelength regime |k| R > pi and whenever the FIPPI cache is disabled
// (PanelPanelInteractions.cc:449-491).  Reference semantics restated here:
//   TaylorDuffy()               libs/libscuff/TaylorDuffy.cc:277-368   (driver, once-integrable Helmholtz kernels)
//   TaylorDuffySum()            :152-258   integrand  sum_d sum_n P_dn(y) K_n(X_d(y)) * Jacobian/(4 pi)
//   GetX / GetScriptP / GetScriptK  :491-547, :552-622, :896-935
//   ComputeGeometricParameters  :1265-1526 (C, M, V, S inputs of the three P polynomials TD_UNITY, TD_PMCHWG1, TD_PMCHWC)
//   pcubature                   libs/libSGJC/pcubature.c:300-400: p-adaptive tensor Clenshaw-Curtis, refine the dimension
//                               with the largest error estimate, stop when converged (RelTol 1e-10, per component) or
//                               numEval > MaxEval = 1000 (the 3-D common-vertex case normally stops on the budget, so the
//                               level sequence is part of the contract).
// The Upsilon coefficient maps come from generated_tables.h as data.  One thread per (panel pair, Q_a, Q_b) item; the
// function-value cache of the cubature lives in a strided HBM scratch block.
#include <vector>
#include <algorithm>
#include "kernels.h"
#include "sw_kernels.h"
#include "classify.cuh"
#include "stamp.cuh"
#include "td_math.cuh"

namespace scuffcuda {

#define TDG_NUMMONO 22
#define TDG_MAXPTS 2304            /* >= 27 + 1000 + last increment (see run_td_general) */
#define TDG_FDIM 6
#define TDG_MAXENT 48
#define SC_TD_MAXLEVEL 10          /* device Clenshaw-Curtis tables hold levels <= 10 (generated_tables.h SC_CC_MAXM) */

// monomial ordering of TaylorDuffy.cc:53-79
enum { M_1 = 0, M_Y1 = 1, M_Y12 = 2, M_Y2 = 3, M_Y1Y2 = 4, M_Y22 = 5, M_Y12Y2 = 7, M_Y1Y22 = 8, M_Y12Y22 = 10,
       M_Y2Y3 = 15, M_Y1Y2Y3 = 17, M_Y22Y3 = 18, M_Y12Y2Y3 = 19, M_Y1Y22Y32 = 20, M_Y22Y32 = 21 };

struct TDGen
{ int cas, numRegions, nOffset, hk;   // hk: high-k kernels (twice-integrable path, TaylorDuffy.cc:336-342)
  double A2, B2, AP2, BP2, L2, AdB, AdAP, AdBP, AdL, BdAP, BdBP, APdBP, BPdL;
  double U[3][6][TD_NUMWPOWERS][TDG_NUMMONO];     // [P slot][region][n][monomial]
  int nMin[3], nMax[3];
  cd k;
};

__device__ void tdg_setup(const TDTables &T, int cas, const double *V1, const double *V2, const double *V3,
                          const double *V2Pin, const double *V3P, const double *Q, const double *QP, cd k, TDGen &W, int hk = 0)
{ const double *V2P = (cas >= 2) ? V2 : V2Pin;              // TaylorDuffy.cc:1286-1289
  if (cas == 3) V3P = V3;
  double A[3], B[3], AP[3], BP[3], L[3], D[3], DP[3], DeltaD[3];
  sub3(V2, V1, A); sub3(V3, V2, B); sub3(V2P, V1, AP); sub3(V3P, V2P, BP);
  W.cas = cas; W.k = k; W.hk = hk;
  W.A2 = dot3(A, A); W.AdB = dot3(A, B); W.B2 = dot3(B, B);
  W.AP2 = W.BP2 = W.L2 = W.AdAP = W.AdBP = W.AdL = W.BdAP = W.BdBP = W.APdBP = W.BPdL = 0.0;
  if (cas == 2)
   { sub3(BP, B, L);
     W.BP2 = dot3(BP, BP); W.AdBP = dot3(A, BP); W.L2 = dot3(L, L); W.AdL = dot3(A, L); W.BPdL = dot3(BP, L);
     W.numRegions = 6; W.nOffset = 2; }
  else if (cas == 3) { W.numRegions = 3; W.nOffset = 1; }
  else
   { W.AP2 = dot3(AP, AP); W.BP2 = dot3(BP, BP); W.AdAP = dot3(A, AP); W.AdBP = dot3(A, BP);
     W.BdAP = dot3(B, AP); W.BdBP = dot3(B, BP); W.APdBP = dot3(AP, BP);
     W.numRegions = 2; W.nOffset = 3; }
  sub3(V1, Q, D); sub3(V1, QP, DP); sub3(QP, Q, DeltaD);
  double DPxD[3]; cross3(DP, D, DPxD);
  for (int np = 0; np < 3; np++)
   { // inputs: C[4], M11 M12 M13 M14 M22 M23 M24 M33 M34 M44, V[4], S
     double in[19];
     for (int i = 0; i < 19; i++) in[i] = 0.0;
     if (np == 0) in[18] = 1.0;                                          // TD_UNITY
     else if (np == 1)                                                   // TD_PMCHWG1
      { in[4 + 2] = dot3(A, AP); in[4 + 3] = dot3(A, BP); in[4 + 5] = dot3(B, AP); in[4 + 6] = dot3(B, BP);
        in[14] = dot3(A, DP); in[15] = dot3(B, DP); in[16] = dot3(D, AP); in[17] = dot3(D, BP);
        in[18] = dot3(D, DP); }
     else                                                                // TD_PMCHWC
      { double S[3];
        cross3(A, AP, S); in[4 + 2] = -dot3(S, DeltaD);
        cross3(A, BP, S); in[4 + 3] = -dot3(S, DeltaD);
        cross3(B, AP, S); in[4 + 5] = -dot3(S, DeltaD);
        cross3(B, BP, S); in[4 + 6] = -dot3(S, DeltaD);
        in[14] = -dot3(A, DPxD); in[15] = -dot3(B, DPxD); in[16] = dot3(AP, DPxD); in[17] = dot3(BP, DPxD); }
     for (int d = 0; d < 6; d++) for (int n = 0; n < TD_NUMWPOWERS; n++) for (int m = 0; m < TDG_NUMMONO; m++) W.U[np][d][n][m] = 0.0;
     for (int e = T.upsStart[cas]; e < T.upsStart[cas + 1]; e++)
      { const UpsEntry &E = T.ups[e];
        W.U[np][E.d][E.n][E.mono] += E.coef * in[E.input]; }
     int nMin = TD_NUMWPOWERS - 1, nMax = -1;
     for (int n = 0; n < TD_NUMWPOWERS; n++) for (int d = 0; d < W.numRegions; d++) for (int m = 0; m < TDG_NUMMONO; m++)
      if (fabs(W.U[np][d][n][m]) > 1.0e-8) { if (n < nMin) nMin = n; if (n > nMax) nMax = n; }
     W.nMin[np] = nMin; W.nMax[np] = nMax;
   }
}

// ---- high-k ("very short wavelength", |k| R > 20) kernels: twice-integrable Taylor-Duffy path ----------------------
// GetQFPIntegral (TaylorDuffy.cc:627-770), negative powers only:  Int = int_0^1 [(y+P)^2 + Q2]^{p/2} dy,
// Inty = int_0^1 y [(y+P)^2 + Q2]^{p/2} dy, with the reference's small-Q2 switches (they change the numbers)
__device__ void qfp_integral(double P, double Q2, int p, double &I, double &Iy)
{ const double P2 = P * P, PP1 = P + 1.0, PP12 = PP1 * PP1, Q4 = Q2 * Q2;
  const double S2 = P * P + Q2, S = sqrt(S2), T2 = PP1 * PP1 + Q2, T = sqrt(T2);
  const double SignP = (P > 0.0) ? 1.0 : -1.0;
  switch (p)
   { case -1:
       I = (Q2 < 1.0e-8) ? SignP * log(1.0 + 1.0 / P) : log((PP1 + T) / (P + S));
       Iy = -P * I + (T - S); return;
     case -2:
       if (Q2 < 1.0e-8) I = 1.0 / (P * PP1);
       else { const double Q = sqrt(Q2); I = (atan(PP1 / Q) - atan(P / Q)) / Q; }
       Iy = -P * I + log(T / S); return;
     case -3:
       I = (Q2 < 1.0e-8) ? SignP * (1.0 / (2.0 * S2) - 1.0 / (2.0 * T2)) : (PP1 / T - P / S) / Q2;
       Iy = -P * I + 1.0 / S - 1.0 / T; return;
     case -4:
       if (Q2 < 1.0e-6) { const double S3 = S * S2, T3 = T * T2; I = SignP * (1.0 / (3.0 * S3) - 1.0 / (3.0 * T3)); }
       else { const double Q = sqrt(Q2), at = (atan(PP1 / Q) - atan(P / Q)) / Q; I = ((Q2 - P * PP1) / (S2 * T2) + at) / (2.0 * Q2); }
       Iy = -P * I + 1.0 / (2.0 * S2) - 1.0 / (2.0 * T2); return;
     case -5:
      { const double S3 = S2 * S, T3 = T2 * T;
        if (Q2 < 1.0e-4) { const double S4 = S3 * S, T4 = T3 * T; I = SignP * (1.0 / (4.0 * S4) - 1.0 / (4.0 * T4)); }
        else I = (PP1 * (2.0 * PP12 + 3.0 * Q2) * S3 - P * (2.0 * P2 + 3.0 * Q2) * T3) / (3.0 * Q4 * S3 * T3);
        Iy = -P * I + 1.0 / (3.0 * S3) - 1.0 / (3.0 * T3); return; }
     case -6:
      { const double S4 = S2 * S2, T4 = T2 * T2;
        if (Q2 < 1.0e-4) { const double S5 = S4 * S, T5 = T4 * T; I = SignP * (1.0 / (5.0 * S5) - 1.0 / (5.0 * T5)); }
        else
         { const double Q = sqrt(Q2), at = (atan(PP1 / Q) - atan(P / Q)) / Q;
           I = (3.0 * PP1 / T2 + 2.0 * Q2 * PP1 / T4 - 3.0 * P / S2 - 2.0 * Q2 * P / S4 + 3 * at) / (8.0 * Q4); }
        Iy = -P * I + 1.0 / (4.0 * S4) - 1.0 / (4.0 * T4); return; }
     default: I = Iy = nan(""); return;       // the reference ErrExit()s for other powers
   }
}

// Px^2 + 2Qx + R -> A^2 [(x+B)^2 + G2]  (PQRtoABG2, TaylorDuffy.cc:375-389, including the B = 0 snap)
__device__ __forceinline__ void pqr_to_abg2(double P, double Q, double R, double &A, double &B, double &G2)
{ A = sqrt(P); B = Q / P;
  const double ROverP = R / P;
  G2 = ROverP - B * B;
  if (fabs(G2) < 1.0e-6 * ROverP) { B = 0.0; G2 = ROverP; }
}

// z^p for complex z, real p, as std::pow(complex, double): exp(p log z)
__device__ __forceinline__ cd cpow_(cd z, double p)
{ if (z.y == 0.0 && z.x > 0.0) return mk(pow(z.x, p), 0.0);
  const double lr = log(cabs_(z)), th = atan2(z.y, z.x);
  const double e = exp(p * lr); double s, c; sincos(p * th, &s, &c);
  return mk(e * c, e * s);
}

__device__ void tdg_integrand_hk(const TDGen &W, const double *y, double f[TDG_FDIM])
{ const int NR = W.numRegions, nOff = W.nOffset;
  double Jac = 1.0;
  if (W.cas == 2) Jac = y[0]; else if (W.cas == 1) Jac = y[1];
  for (int i = 0; i < TDG_FDIM; i++) f[i] = 0.0;
  if (Jac == 0.0) return;
  double Av[6], Bv[6], G2v[6];
  if (W.cas == 3)        // GetAlphaBetaGamma2, TaylorDuffy.cc:398-418
   { const double A2 = W.A2, AdB = W.AdB, B2 = W.B2, C2 = A2 + B2 + 2.0 * AdB, CdB = AdB + B2;
     Av[0] = sqrt(B2); Av[1] = sqrt(C2); Av[2] = sqrt(A2);
     Bv[0] = AdB / B2; Bv[1] = -CdB / C2; Bv[2] = AdB / A2;
     G2v[0] = A2 / B2 - (AdB * AdB) / (B2 * B2); G2v[1] = B2 / C2 - (CdB * CdB) / (C2 * C2); G2v[2] = B2 / A2 - (AdB * AdB) / (A2 * A2); }
  else if (W.cas == 2)   // :419-457
   { const double y1 = y[0], y12 = y1 * y1, A2 = W.A2, BP2 = W.BP2, L2 = W.L2, AdBP = W.AdBP, AdL = W.AdL, BPdL = W.BPdL;
     pqr_to_abg2((BP2 - 2.0 * BPdL + L2) * y12, y1 * (L2 + BPdL * (y1 - 1.0) - (AdL + L2 - AdBP) * y1),
                 L2 - 2 * AdL * y1 - 2 * L2 * y1 + A2 * y12 + 2 * AdL * y12 + L2 * y12, Av[0], Bv[0], G2v[0]);
     pqr_to_abg2(BP2 * y12, (BPdL + AdBP * y1 - BPdL * y1) * y1, L2 + 2 * AdL * y1 - 2 * L2 * y1 + (A2 - 2 * AdL + L2) * y12, Av[1], Bv[1], G2v[1]);
     pqr_to_abg2((A2 + 2 * AdBP + BP2) * y12, y1 * (AdL * (-1 + y1) + BPdL * (-1 + y1) - (AdBP + BP2) * y1),
                 L2 + 2 * BPdL * y1 - 2 * L2 * y1 + BP2 * y12 - 2 * BPdL * y12 + L2 * y12, Av[2], Bv[2], G2v[2]);
     pqr_to_abg2((A2 + 2 * AdBP - 2 * AdL + BP2 - 2 * BPdL + L2) * y12, y1 * (AdL - L2 - (AdBP + BP2) * y1 + BPdL * (1 + y1)),
                 L2 - 2 * BPdL * y1 + BP2 * y12, Av[3], Bv[3], G2v[3]);
     pqr_to_abg2(A2 * y12, (AdBP * y1 - AdL) * y1, L2 - 2 * BPdL * y1 + BP2 * y12, Av[4], Bv[4], G2v[4]);
     pqr_to_abg2(A2 * y12, y1 * (AdL + AdBP * y1 - AdL * y1), L2 + 2 * BPdL * y1 - 2 * L2 * y1 + BP2 * y12 - 2 * BPdL * y12 + L2 * y12,
                 Av[5], Bv[5], G2v[5]); }
  else                   // :458-484
   { const double y1 = y[0], y2 = y[1], y22 = y2 * y2;
     pqr_to_abg2(W.BP2 * y22, -W.AdBP * y2 - W.BdBP * y1 * y2 + W.APdBP * y22,
                 W.A2 + W.B2 * y1 * y1 + W.AP2 * y2 * y2 + 2.0 * W.AdB * y1 - 2.0 * W.AdAP * y2 - 2.0 * W.BdAP * y1 * y2, Av[0], Bv[0], G2v[0]);
     pqr_to_abg2(W.B2 * y22, W.AdB * y22 - W.BdAP * y2 - W.BdBP * y1 * y2,
                 W.A2 * y2 * y2 + W.AP2 + W.BP2 * y1 * y1 - 2 * W.AdAP * y2 - 2 * W.AdBP * y2 * y1 + 2 * W.APdBP * y1, Av[1], Bv[1], G2v[1]); }
  const double Fact[7] = { 1.0, 1.0, 2.0, 6.0, 24.0, 120.0, 620.0 };     // TaylorDuffy.cc:777 (620, as written)
  const cd mIK = mk(W.k.y, -W.k.x);                                       // -i k
  const int numPK = (W.cas == 3) ? 2 : 3;
  for (int npk = 0; npk < numPK; npk++)
   { const int nMin = W.nMin[npk], nMax = W.nMax[npk];
     cd Sum = mk(0.0, 0.0);
     for (int n = nMin; n <= nMax; n++)
      for (int d = 0; d < NR; d++)
       { const double *Up = W.U[npk][d][n];
         double P0, P1;
         if (W.cas == 3) { P0 = Up[M_1]; P1 = Up[M_Y1]; }
         else if (W.cas == 2)
          { const double y1 = y[0], y12 = y1 * y1;
            P0 = Up[M_1] + Up[M_Y1] * y1 + Up[M_Y12] * y12;
            P1 = Up[M_Y2] + Up[M_Y1Y2] * y1 + Up[M_Y12Y2] * y12; }
         else
          { const double y1 = y[0], y12 = y1 * y1, y2 = y[1], y22 = y2 * y2;
            P0 = Up[M_1] + Up[M_Y1] * y1 + Up[M_Y2] * y2 + Up[M_Y12] * y12 + Up[M_Y1Y2] * y1 * y2 + Up[M_Y22] * y22
                 + Up[M_Y12Y2] * y12 * y2 + Up[M_Y1Y22] * y1 * y22;
            P1 = Up[M_Y2Y3] * y2 + Up[M_Y1Y2Y3] * y1 * y2 + Up[M_Y22Y3] * y22 + Up[M_Y12Y2Y3] * y12 * y2; }
         const int nn = n + nOff;
         double I, Iy; qfp_integral(Bv[d], G2v[d], -(nn + 1), I, Iy);
         const cd z = mk(mIK.x * Av[d], mIK.y * Av[d]);
         cd J, L;
         if (npk < 2)      // TD_HIGHK_HELMHOLTZ
          { const cd pw = cpow_(z, -(double)nn); const double c = Fact[nn - 1];
            const cd t = mk(c * pw.x, c * pw.y);
            J = mk(t.x * I / Av[d], t.y * I / Av[d]); L = mk(t.x * Iy / Av[d], t.y * Iy / Av[d]); }
         else              // TD_HIGHK_GRADHELMHOLTZ
          { const double A3 = Av[d] * Av[d] * Av[d];
            const cd pw = cpow_(z, -(double)(nn - 2)); const double c = -(nn - 1) * Fact[nn - 3];
            const cd t = mk(c * pw.x, c * pw.y);
            J = mk(t.x * I / A3, t.y * I / A3); L = mk(t.x * Iy / A3, t.y * Iy / A3); }
         Sum += P0 * J + P1 * L;
       }
     const double s = Jac / (4.0 * M_PI);
     f[2 * npk] = Sum.x * s; f[2 * npk + 1] = Sum.y * s;
   }
}

// integrand (TaylorDuffySum, once-integrable branch): f[0..5] = Re/Im of the three P x K integrals
__device__ void tdg_integrand(const TDGen &W, const double *y, double f[TDG_FDIM])
{ if (W.hk) { tdg_integrand_hk(W, y, f); return; }
  const int NR = W.numRegions, nOff = W.nOffset;
  double yInner, Jac;
  if (W.cas == 2) { yInner = y[1]; Jac = y[0]; } else { yInner = y[2]; Jac = y[1]; }
  if (Jac == 0.0) { for (int i = 0; i < TDG_FDIM; i++) f[i] = 0.0; return; }
  double X[6];
  if (W.cas == 2)
   { const double y1 = y[0], y2 = y[1]; double u1, u2, xi2;
#define TDG_XE(d) X[d] = sqrt(u1 * u1 * W.A2 + u2 * u2 * W.BP2 + xi2 * xi2 * W.L2 + 2.0 * (u1 * (u2 * W.AdBP + xi2 * W.AdL) + u2 * xi2 * W.BPdL))
     u1 = -y1;      u2 = -y1 * y2;         xi2 = (1.0 - y1 + y1 * y2); TDG_XE(0);
     u1 = y1;       u2 = y1 * y2;          xi2 = (1.0 - y1);           TDG_XE(1);
     u1 = -y1 * y2; u2 = y1 * (1.0 - y2);  xi2 = (1.0 - y1);           TDG_XE(2);
     u1 = y1 * y2;  u2 = -y1 * (1.0 - y2); xi2 = (1.0 - y1 * y2);      TDG_XE(3);
     u1 = -y1 * y2; u2 = -y1;              xi2 = 1.0;                  TDG_XE(4);
     u1 = y1 * y2;  u2 = y1;               xi2 = (1.0 - y1);           TDG_XE(5);
#undef TDG_XE
   }
  else
   { const double y1 = y[0], y2 = y[1], y3 = y[2]; double xi1, xi2, eta1, eta2;
#define TDG_XV(d) X[d] = sqrt(xi1 * xi1 * W.A2 + xi2 * xi2 * W.B2 + eta1 * eta1 * W.AP2 + eta2 * eta2 * W.BP2 \
                              + 2.0 * xi1 * (xi2 * W.AdB - eta1 * W.AdAP - eta2 * W.AdBP) - 2.0 * xi2 * (eta1 * W.BdAP + eta2 * W.BdBP) \
                              + 2.0 * eta1 * eta2 * W.APdBP)
     xi1 = 1.0; xi2 = y1;      eta1 = y2;  eta2 = y2 * y3; TDG_XV(0);
     xi1 = y2;  xi2 = y2 * y3; eta1 = 1.0; eta2 = y1;      TDG_XV(1);
#undef TDG_XV
   }
  // ExpRel table per region for the union of the n ranges: HELMHOLTZ needs n in [nMin+nOff, nMax+nOff] of P slots 0,1;
  // GRADHELMHOLTZ needs n-2 .. n-1 of slot 2
  int lo = 99, hi = -1;
  for (int np = 0; np < 2; np++) if (W.nMax[np] >= W.nMin[np]) { lo = min(lo, W.nMin[np] + nOff); hi = max(hi, W.nMax[np] + nOff); }
  if (W.nMax[2] >= W.nMin[2]) { lo = min(lo, W.nMin[2] + nOff - 2); hi = max(hi, W.nMax[2] + nOff - 1); }
  if (lo < 0) lo = 0;
  const cd IK = mk(-W.k.y, W.k.x);
  cd E[6][9], eIKXv[6];
  for (int d = 0; d < NR; d++)
   { cd IKX = mk(IK.x * X[d], IK.y * X[d]), mIKX = mk(-IKX.x, -IKX.y);
     eIKXv[d] = cexp_(IKX);
     const bool small = (mIKX.y == 0.0) ? (fabs(mIKX.x) < 0.1) : (cabs_(mIKX) < 0.1);
     cd emIKX = mk(0.0, 0.0);
     if (!small) emIKX = (mIKX.y == 0.0) ? mk(exp(mIKX.x), 0.0) : cexp_(mIKX);
     for (int n = lo; n <= hi && n < 9; n++) E[d][n] = exprel_v3p0_pre(n, mIKX, emIKX, small);
   }
  for (int npk = 0; npk < 3; npk++)
   { const int nMin = W.nMin[npk], nMax = W.nMax[npk];
     cd Sum = mk(0.0, 0.0);
     for (int n = nMin; n <= nMax; n++)
      for (int d = 0; d < NR; d++)
       { const double *Up = W.U[npk][d][n];
         double P0, P1;
         if (W.cas == 2)
          { const double y1 = y[0], y12 = y1 * y1;
            P0 = Up[M_1] + Up[M_Y1] * y1 + Up[M_Y12] * y12;
            P1 = Up[M_Y2] + Up[M_Y1Y2] * y1 + Up[M_Y12Y2] * y12; }
         else
          { const double y1 = y[0], y12 = y1 * y1, y2 = y[1], y22 = y2 * y2;
            P0 = Up[M_1] + Up[M_Y1] * y1 + Up[M_Y2] * y2 + Up[M_Y12] * y12 + Up[M_Y1Y2] * y1 * y2 + Up[M_Y22] * y22
                 + Up[M_Y12Y2] * y12 * y2 + Up[M_Y1Y22] * y1 * y22;
            P1 = Up[M_Y2Y3] * y2 + Up[M_Y1Y2Y3] * y1 * y2 + Up[M_Y22Y3] * y22 + Up[M_Y12Y2Y3] * y12 * y2; }
         const int nn = n + nOff;
         cd Kv;
         if (npk < 2)      // TD_HELMHOLTZ: e^{ikX} ExpRel(n,-ikX) / (n X)
          { cd t = eIKXv[d] * E[d][nn]; const double dn = nn * X[d], rdn = 1.0 / dn; Kv = mk(td_div_rcp(t.x, dn, rdn), td_div_rcp(t.y, dn, rdn)); }
         else              // TD_GRADHELMHOLTZ
          { cd t1 = IK * E[d][nn - 1]; const double d1 = (double)nn - 1.0; t1 = mk(t1.x / d1, t1.y / d1);
            const double d2 = ((double)nn - 2.0) * X[d]; cd t2 = mk(E[d][nn - 2].x / d2, E[d][nn - 2].y / d2);
            cd t = eIKXv[d] * (t1 - t2); const double xx = X[d] * X[d]; Kv = mk(t.x / xx, t.y / xx); }
         Sum += (P0 + yInner * P1) * Kv;
       }
     const double s = Jac / (4.0 * M_PI);
     f[2 * npk] = Sum.x * s; f[2 * npk + 1] = Sum.y * s;
   }
}

// ---- p-adaptive tensor Clenshaw-Curtis cubature over [0,1]^DIM (pcubature.c) ----
struct PCEnt { unsigned char m[3]; signed char mi; int off; int cnt[3]; };

template <int DIM>
__device__ __forceinline__ double pc_weight(const TDTables &T, const PCEnt &e, int d, int li, int Md)
{ // weight of local point li of entry e along dimension d for a rule of level Md (Md < 0: trivial midpoint rule)
  int pairIdx;
  if (d != e.mi)
   { if (li == 0) return (Md < 0) ? 2.0 : T.ccw[Md + (1 << Md) - 1];
     pairIdx = (li - 1) >> 1; }
  else
   { const int lo = e.m[d] ? (1 << (e.m[d] - 1)) : 0;
     pairIdx = lo + (li >> 1); }
  if (Md < 0 || pairIdx >= (1 << Md)) return 0.0;
  return T.ccw[Md + (1 << Md) - 1 + 1 + pairIdx];
}

template <int DIM>
__device__ __forceinline__ double pc_coord(const TDTables &T, const PCEnt &e, int d, int li)
{ if (d != e.mi)
   { if (li == 0) return 0.5;
     const int i = (li - 1) >> 1; return ((li - 1) & 1) ? 0.5 - 0.5 * T.ccx[i] : 0.5 + 0.5 * T.ccx[i]; }
  const int lo = e.m[d] ? (1 << (e.m[d] - 1)) : 0; const int i = lo + (li >> 1);
  return (li & 1) ? 0.5 - 0.5 * T.ccx[i] : 0.5 + 0.5 * T.ccx[i];
}

// ---- group-cooperative evaluation: TDG_GROUP lanes share one item.  The points of every cache entry are enumerated
// linearly (q = (i0*n1 + i1)*n2 + i2, the storage order); lanes stride q, both when evaluating the integrand at new points
// and when forming the quadrature sums, which are then reduced with warp shuffles (identical totals on all lanes, so every
// decision is group-uniform).  Decisions are those of pcubature.c; only the summation order of the quadrature sums differs.
#define TDG_GROUP 8

__device__ __forceinline__ double grp_sum(double v, unsigned gmask)
{
#pragma unroll
  for (int o = TDG_GROUP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
  return v; }

// integral of the rule with levels M[] (M[md] decremented when md < DIM) from the cached values; cache[(pt*FDIM + j)]
template <int DIM>
__device__ void pc_evals(const TDTables &T, const PCEnt *ent, int nEnt, const int *m, int md, const double *cache,
                         int gl, unsigned gmask, double *val)
{ for (int j = 0; j < TDG_FDIM; j++) val[j] = 0.0;
  int M[3];
  for (int d = 0; d < DIM; d++) M[d] = m[d] - (d == md ? 1 : 0);
  double V = 1.0; for (int d = 0; d < DIM; d++) V *= 0.5;
  for (int ie = 0; ie < nEnt; ie++)
   { const PCEnt &e = ent[ie];
     if (!(e.mi >= DIM || (int)e.m[e.mi] + (e.mi == md ? 1 : 0) <= m[e.mi])) continue;
     const int n1 = DIM > 1 ? e.cnt[1] : 1, n2 = DIM > 2 ? e.cnt[2] : 1;
     const int total = e.cnt[0] * n1 * n2;
     for (int q = gl; q < total; q += TDG_GROUP)
      { const int i2 = q % n2, i1 = (q / n2) % n1, i0 = q / (n1 * n2);
        double w = V * pc_weight<DIM>(T, e, 0, i0, M[0]);
        if (DIM > 1 && w != 0.0) w *= pc_weight<DIM>(T, e, 1, i1, M[1]);
        if (DIM > 2 && w != 0.0) w *= pc_weight<DIM>(T, e, 2, i2, M[2]);
        if (w == 0.0) continue;
        const double *cv = cache + (size_t)(e.off + q) * TDG_FDIM;
        for (int j = 0; j < TDG_FDIM; j++) val[j] += cv[j] * w;
      }
   }
  for (int j = 0; j < TDG_FDIM; j++) val[j] = grp_sum(val[j], gmask);
}

template <int DIM>
__device__ int pc_add_entry(const TDTables &T, const TDGen &W, PCEnt *ent, int &nEnt, int &nPts, const int *m, int mi,
                            double *cache, int gl, unsigned gmask)
{ if (nEnt >= TDG_MAXENT) return 2;
  PCEnt &e = ent[nEnt];
  for (int d = 0; d < 3; d++) { e.m[d] = d < DIM ? (unsigned char)m[d] : 0; e.cnt[d] = 1; }
  e.mi = (signed char)mi; e.off = nPts;
  int total = 1;
  for (int d = 0; d < DIM; d++)
   { e.cnt[d] = (d == mi) ? (m[d] == 0 ? 2 : (1 << m[d])) : ((1 << (m[d] + 1)) + 1);
     total *= e.cnt[d]; }
  if (nPts + total > TDG_MAXPTS) return 2;
  const int n1 = DIM > 1 ? e.cnt[1] : 1, n2 = DIM > 2 ? e.cnt[2] : 1;
  for (int q = gl; q < total; q += TDG_GROUP)
   { const int i2 = q % n2, i1 = (q / n2) % n1, i0 = q / (n1 * n2);
     double y[3] = { 0.0, 0.0, 0.0 }, f[TDG_FDIM];
     y[0] = pc_coord<DIM>(T, e, 0, i0);
     if (DIM > 1) y[1] = pc_coord<DIM>(T, e, 1, i1);
     if (DIM > 2) y[2] = pc_coord<DIM>(T, e, 2, i2);
     tdg_integrand(W, y, f);
     double *cv = cache + (size_t)(nPts + q) * TDG_FDIM;
     for (int j = 0; j < TDG_FDIM; j++) cv[j] = f[j];
   }
  __syncwarp(gmask);
  nPts += total; nEnt++;
  return 0;
}

template <int DIM>
__device__ int pcubature_nd(const TDTables &T, const TDGen &W, double relTol, double absTol, int maxEval,
                            double *cache, int gl, unsigned gmask, double val[TDG_FDIM], double err[TDG_FDIM])
{ PCEnt ent[TDG_MAXENT]; int nEnt = 0, nPts = 0, numEval = 0;
  int m[3] = { 0, 0, 0 };
  int st = pc_add_entry<DIM>(T, W, ent, nEnt, nPts, m, DIM, cache, gl, gmask);
  if (st) return st;
  for (;;)
   { pc_evals<DIM>(T, ent, nEnt, m, DIM, cache, gl, gmask, val);
     for (int j = 0; j < TDG_FDIM; j++) err[j] = 0.0;
     int mi = 0; double maxerr = 0.0;
     for (int i = 0; i < DIM; i++)
      { double val1[TDG_FDIM], emax = 0.0;
        pc_evals<DIM>(T, ent, nEnt, m, i, cache, gl, gmask, val1);
        for (int j = 0; j < TDG_FDIM; j++)
         { const double e = fabs(val[j] - val1[j]);
           if (e > emax) emax = e;
           if (e > err[j]) err[j] = e; }
        if (emax > maxerr) { maxerr = emax; mi = i; }
      }
     bool conv = true;
     for (int j = 0; j < TDG_FDIM; j++) if (err[j] > absTol && err[j] > fabs(val[j]) * relTol) { conv = false; break; }
     if (conv || (numEval > maxEval && maxEval)) return 0;
     m[mi] += 1;
     if (m[mi] > SC_TD_MAXLEVEL) return 1;
     int newPts = 1;
     for (int d = 0; d < DIM; d++) newPts *= (d == mi) ? (1 << m[d]) : ((1 << (m[d] + 1)) + 1);
     st = pc_add_entry<DIM>(T, W, ent, nEnt, nPts, m, mi, cache, gl, gmask);
     if (st) return st;
     numEval += newPts;
   }
}

// items[t] = 36*nearIndex + 18*highk + 9*swapped + 3*i + j
__global__ void __launch_bounds__(64)
td_general_kernel(const NearArgs A, const TDTables T, const int *items, int nItems, double *scratch, int *itemStatus)
{ const int t = (blockIdx.x * blockDim.x + threadIdx.x) / TDG_GROUP;          // item handled by this group of lanes
  const int gl = threadIdx.x & (TDG_GROUP - 1);
  if (t >= nItems) return;                                                  // whole groups leave together
  const unsigned gmask = ((1u << TDG_GROUP) - 1u) << ((threadIdx.x & 31) & ~(TDG_GROUP - 1));
  double *cache = scratch + (size_t)t * TDG_MAXPTS * TDG_FDIM;
  const int item = items[t];
  const NearPair N = A.near[item / 36];
  const int hk = (item % 36) / 18, swapped = (item % 18) / 9, i = (item % 9) / 3, j = item % 3;
  const PanelRec &Pa = A.pa[N.pa], &Pb = A.pb[N.pb];
  const int ea = Pa.ei[i], eb = Pb.ei[j];
  double Va[9], Vb[9];
  for (int q = 0; q < 9; q++) { Va[q] = Pa.V[q]; Vb[q] = A.useDisp ? (Pb.V[q] + 1.0 * A.L[q % 3]) : Pb.V[q]; }
  const double *V1, *V2, *V3, *V2P, *V3P, *Q, *QP; int ncv = N.ncv;
  if (!swapped)
   { V1 = Va + 3 * N.ia[0]; V2 = Va + 3 * N.ia[1]; V3 = Va + 3 * N.ia[2]; V2P = Vb + 3 * N.ib[1]; V3P = Vb + 3 * N.ib[2];
     Q = Va + 3 * i; QP = Vb + 3 * j; }
  else
   { // the reference evaluates PPI(pb,pa) for this entry: AssessPanelPair with the panels exchanged
     unsigned char ja[3], jb[3];
     ncv = assess_vertices(Vb, Va, fmax(Pa.R, Pb.R), ja, jb);
     V1 = Vb + 3 * ja[0]; V2 = Vb + 3 * ja[1]; V3 = Vb + 3 * ja[2]; V2P = Va + 3 * jb[1]; V3P = Va + 3 * jb[2];
     Q = Vb + 3 * j; QP = Va + 3 * i; }
  TDGen W;
  const cd k = mk(A.K.kr, A.K.ki);
  tdg_setup(T, ncv, V1, V2, V3, V2P, V3P, Q, QP, k, W, hk);
  double val[TDG_FDIM], err[TDG_FDIM];
  for (int q = 0; q < TDG_FDIM; q++) val[q] = 0.0;
  // integral dimension = 4 - WhichCase - TwiceIntegrable (TaylorDuffy.cc:355)
  const int dim = 4 - ncv - hk;
  int st = 0;
  if (dim == 0) tdg_integrand_hk(W, nullptr, val);
  else if (dim == 1) st = pcubature_nd<1>(T, W, 1.0e-10, 0.0, 1000, cache, gl, gmask, val, err);
  else if (dim == 2) st = pcubature_nd<2>(T, W, 1.0e-10, 0.0, 1000, cache, gl, gmask, val, err);
  else st = pcubature_nd<3>(T, W, 1.0e-10, 0.0, 1000, cache, gl, gmask, val, err);
  if (gl != 0) return;
  itemStatus[t] = st;
  if (st) return;
  const cd R0 = mk(val[0], val[1]), R1 = mk(val[2], val[3]), R2 = mk(val[4], val[5]);
  const cd k2 = k * k;
  cd H0 = R1 - cdiv(mk(4.0 * R0.x, 4.0 * R0.y), k2), H1 = (ncv == 3) ? mk(0.0, 0.0) : R2;
  // a diagonal edge pair (ea == eb, pa != pb) of a symmetric block receives PPI(pa,pb) + PPI(pb,pa): both orders are
  // separate items, scatter_ppi would double each
  if (A.S.symmetric && N.pa != N.pb && ea == eb) { H0 = 0.5 * H0; H1 = 0.5 * H1; }
  scatter_ppi(A.S, A.K, N.pa, N.pb, ea, eb, Pa.sl[i], Pb.sl[j], H0, H1);
  atomicAdd(A.counters + (hk ? 3 : 2), 1ull);
}

int run_td_general(scuffcuda_context *ctx, const NearArgs &A, const TDTables &T, const int *items, int nItems)
{ if (nItems <= 0) return 0;
  // capacity: 27 initial points + at most MaxEval (1000) before the last refinement + the last increment (< what is
  // already there) -> TDG_MAXPTS points x 6 doubles per item; items are processed in batches
  const size_t perItem = (size_t)TDG_MAXPTS * TDG_FDIM * sizeof(double);
  const int batch = std::min(nItems, 8192);
  const long threads = (((long)batch * TDG_GROUP + 63) / 64) * 64;
  const size_t hdr = (((size_t)batch * 2 * sizeof(int) + 255) / 256) * 256;
  const size_t need = hdr + (size_t)batch * perItem;
  if (need > ctx->scratchBytes)
   { if (ctx->dscratch) cudaFree(ctx->dscratch); ctx->dscratch = nullptr; ctx->scratchBytes = 0;
     if (cudaMalloc(&ctx->dscratch, need) != cudaSuccess) { cudaGetLastError(); ctx->err = "Taylor-Duffy scratch allocation failed"; return 1; }
     ctx->scratchBytes = need; }
  int *dItems = (int *)ctx->dscratch;
  int *hStatus = (int *)mapped_host_scratch(ctx, (size_t)batch * sizeof(int), 1);
  if (!hStatus) { ctx->err = "Taylor-Duffy: cannot allocate mapped host memory"; return 1; }
  double *scr = (double *)((char *)ctx->dscratch + hdr);
  std::vector<int> status(batch);
  for (int first = 0; first < nItems; first += batch)
   { const int n = std::min(batch, nItems - first);
     cudaMemcpyAsync(dItems, items + first, n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
     ctx->launches++;
     td_general_kernel<<<(unsigned)(threads / 64), 64, 0, ctx->stream>>>(A, T, dItems, n, scr, hStatus);
     cudaError_t e = cudaStreamSynchronize(ctx->stream);
     std::copy(hStatus, hStatus + n, status.begin());
     if (e != cudaSuccess) { ctx->err = std::string("td_general_kernel: ") + cudaGetErrorString(e); return 1; }
     for (int q = 0; q < n; q++)
      if (status[q]) { ctx->err = status[q] == 1 ? "Taylor-Duffy cubature exceeded the tabulated Clenshaw-Curtis levels"
                                                 : "Taylor-Duffy cubature exceeded its function-value cache"; return 1; }
   }
  return 0;
}

}  // namespace scuffcuda
