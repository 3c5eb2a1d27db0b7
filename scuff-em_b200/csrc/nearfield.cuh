// Reduced fields of one RWG basis function at an evaluation point closer than one basis-function radius: the branch of
// GetRFMatrix that calls GetReducedFields_Nearby (libs/libscuff/GetFields.cc:296-299, GetNearFields.cc:60-905).
//
// What the reference does, restated (only what e and h need — no second derivatives):
//   G(r) = 1/(4 pi r) + ik/(4 pi) + (ik)^2 r/(8 pi) + G_DS(r): the first three terms of the small-k expansion are integrated over
//   the panel in closed form, the remainder G_DS = [exp(ikr) - 1 - ikr - (ikr)^2/2]/(4 pi r) by the 21-point rule.
//   * closed forms (GetEdgeIntegrals :67-157, GetScriptIJ :190-299, GetStaticPanelIntegrals :307-535): in the panel's (u,v,w) frame
//     (u along V0->V1, w the right-handed normal, origin = projection of the evaluation point) each panel edge contributes line
//     integrals I_p, J_p, dJ_p/dl for p = -5,-3,-1,1; they assemble into the potentials int r^p and int rho r^p over the triangle and
//     their gradients, with the "interior point correction" when the projection falls inside the triangle (or on its boundary);
//   * reduced vector potential a^(n) = int (x - Q) r^n, scalar potential p^(n) = int r^n (n = -1, 1; the k^0 terms are the area
//     and area x (centroid - Q)), gradients d_mu p and d_mu a_nu; e = a + grad p / k^2, h = curl a; RWG prefactors l/(2A), l/A;
//     positive panel minus negative panel.
// Same formulas, thresholds (1e-4, 1e-8 x length scale, EqualFloat on the angle sum) and order of operations as the reference:
// where the closed forms lose digits (points almost in the panel's plane or on an edge line) both sides lose the same ones.
#pragma once
#include "ppi_math.cuh"
#include "generated_tcr9.h"

namespace scuffcuda {

#define NF_NUMPS 4          /* p = -5, -3, -1, 1 */

// line integrals along one panel edge (GetEdgeIntegrals, GetNearFields.cc:67-157)
__device__ inline void nf_edge_integrals(double sM, double sP, double t, double w, double I[NF_NUMPS], double J[NF_NUMPS], double dJdl[NF_NUMPS])
{ const double PV[NF_NUMPS] = { -5.0, -3.0, -1.0, 1.0 };
  const double LengthScale = fabs(sP - sM);
  const double sP2 = sP * sP, sP3 = sP2 * sP, sM2 = sM * sM, sM3 = sM2 * sM;
  const double t2 = t * t, w2 = w * w, w3 = w2 * w, X2 = t2 + w2, X4 = X2 * X2, X = sqrt(X2);
  const double RP = sqrt(sP2 + X2), RM = sqrt(sM2 + X2);
  double LogFac;
  if (X > 1.0e-4 * LengthScale) LogFac = log((RP + sP) / (RM + sM));
  else if ((fabs(sM) > 1.0e-4 * LengthScale) && (fabs(sP) > 1.0e-4 * LengthScale)) LogFac = log(fabs(sP) / fabs(sM));
  else LogFac = 0.0;
  double atanFac;
  if (fabs(w) < 1.0e-8 * LengthScale || fabs(t) < 1.0e-8 * LengthScale) atanFac = 0.0;
  else atanFac = atan((sP * w) / (RP * t)) - atan((sM * w) / (RM * t));
  if (fabs(t) < 1.0e-8 * LengthScale) { for (int n = 0; n < NF_NUMPS; n++) I[n] = 0.0; }
  else
   { if (fabs(w) < 1.0e-8 * LengthScale)
      { I[0] = sP * (2.0 * sP * sP + 3.0 * t * t) / (3.0 * t * t * t * RP * RP * RP) - sM * (2.0 * sM * sM + 3.0 * t * t) / (3.0 * t * t * t * RM * RM * RM);
        I[1] = (sP / RP - sM / RM) / t; }
     else
      { I[0] = t * (sM / RM - sP / RP) / (w2 * X2) + atanFac / w3;
        I[1] = atanFac / w; }
     I[2] = w * atanFac + t * LogFac;
     I[3] = t * (RP * sP - RM * sM + LogFac * (t2 + 3.0 * w2)) / 2.0 + w3 * atanFac; }
  if (X < 1.0e-4 * LengthScale) J[0] = 0.0; else J[0] = (sP / RP - sM / RM) / X2;
  J[1] = LogFac;
  J[2] = (RP * sP - RM * sM + LogFac * X2) / 2.0;
  J[3] = (2.0 * RP * sP3 - 2.0 * RM * sM3 + 5.0 * X2 * (RP * sP - RM * sM) + 3.0 * LogFac * X4) / 8.0;
  for (int n = 0; n < NF_NUMPS; n++) dJdl[n] = 0.0;
  if (fabs(RM) > 1.0e-8 * LengthScale) for (int n = 1; n < NF_NUMPS; n++) dJdl[n] += pow(RM, PV[n] + 2.0);
  if (fabs(RP) > 1.0e-8 * LengthScale) for (int n = 1; n < NF_NUMPS; n++) dJdl[n] -= pow(RP, PV[n] + 2.0);
}

// int r^p / 4 pi ... over the triangle with vertices (u_i, v_i, 0), observation point (0, 0, w0) (GetScriptIJ :190-299);
// only the first derivatives that e and h need
__device__ inline void nf_script_ij(const double u[3], const double v[3], double w0, double LengthScale,
                                    double ScriptI[NF_NUMPS], double dScriptIMu[NF_NUMPS][3], double ScriptJMu[NF_NUMPS][3], double dScriptJMuNu[NF_NUMPS][3][3])
{ const double PV[NF_NUMPS] = { -5.0, -3.0, -1.0, 1.0 };
  for (int n = 0; n < NF_NUMPS; n++)
   { ScriptI[n] = 0.0;
     for (int m = 0; m < 3; m++) { dScriptIMu[n][m] = ScriptJMu[n][m] = 0.0; for (int q = 0; q < 3; q++) dScriptJMuNu[n][m][q] = 0.0; } }
  double Beta = 0.0;
  for (int i = 0; i < 3; i++)
   { const double ui = u[i], uip1 = u[(i + 1) % 3], vi = v[i], vip1 = v[(i + 1) % 3];
     double lHat[2] = { uip1 - ui, vip1 - vi };
     const double Length = sqrt(lHat[0] * lHat[0] + lHat[1] * lHat[1]);
     lHat[0] /= Length; lHat[1] /= Length;
     const double mHat[2] = { lHat[1], -lHat[0] };
     const double t = -(ui * mHat[0] + vi * mHat[1]);
     if (!(fabs(t) < 1.0e-8 * LengthScale))
      Beta += acos((ui * uip1 + vi * vip1) / (sqrt(ui * ui + vi * vi) * sqrt(uip1 * uip1 + vip1 * vip1)));
     const double sM = ui * lHat[0] + vi * lHat[1], sP = uip1 * lHat[0] + vip1 * lHat[1];
     double I[NF_NUMPS], J[NF_NUMPS], dJdl[NF_NUMPS];
     nf_edge_integrals(sM, sP, t, w0, I, J, dJdl);
     for (int n = 0; n < NF_NUMPS; n++)
      { const double P = PV[n], Denom = P + 2.0;
        ScriptI[n] -= I[n] / Denom;
        ScriptJMu[n][0] += mHat[0] * J[n] / Denom;
        ScriptJMu[n][1] += mHat[1] * J[n] / Denom;
        if (n > 0)
         { const double dJdm = (P + 2.0) * t * J[n - 1];
           for (int Mu = 0; Mu < 2; Mu++)
            for (int Nu = 0; Nu < 2; Nu++)
             dScriptJMuNu[n][Mu][Nu] += (dJdl[n] * lHat[Mu] + dJdm * mHat[Mu]) * mHat[Nu] / Denom; } } }
  for (int n = 1; n < NF_NUMPS; n++)
   { const double P = PV[n];
     dScriptIMu[n][0] = -P * ScriptJMu[n - 1][0];
     dScriptIMu[n][1] = -P * ScriptJMu[n - 1][1];
     dScriptIMu[n][2] = P * w0 * ScriptI[n - 1];
     dScriptJMuNu[n][2][0] = P * w0 * ScriptJMu[n - 1][0];
     dScriptJMuNu[n][2][1] = P * w0 * ScriptJMu[n - 1][1];
     dScriptJMuNu[n][2][2] = P * w0 * ScriptJMu[n - 1][2]; }
  // interior point correction: the projection of the observation point lies inside the triangle (angle sum 2 pi) or on its
  // boundary (pi); EqualFloat = equal as floats (libhrutil/Vector.cc:270-271)
  if ((float)Beta == (float)M_PI || (float)Beta == (float)(2.0 * M_PI))
   { const double Sign = w0 > 0.0 ? 1.0 : -1.0, w0Abs = fabs(w0);
     if (w0Abs > 1.0e-8 * LengthScale)
      { const double w02 = w0 * w0; double w0Powers[NF_NUMPS];
        w0Powers[0] = pow(fabs(w0), PV[0] + 2.0);
        for (int n = 1; n < NF_NUMPS; n++) w0Powers[n] = w0Powers[n - 1] * w02;
        for (int n = 0; n < NF_NUMPS; n++)
         { const double PP2 = PV[n] + 2.0;
           ScriptI[n] -= Beta * w0Powers[n] / PP2;
           dScriptIMu[n][2] -= Sign * Beta * w0Powers[n] / w0Abs; } } }
}

// static potentials a^(n), grad p^(n), grad a^(n) (n = -1, +1) of the half basis function with source vertex iQ of one panel
// (GetStaticPanelIntegrals :307-535, first-derivative part); vectors in the global frame
__device__ inline void nf_static_panel(const PanelRec &Pan, int iQ, const double X0[3], double an[2][3], double dpn[2][3], double dan[2][3][3])
{ const double PV[NF_NUMPS] = { -5.0, -3.0, -1.0, 1.0 };
  const double *V0 = Pan.V, *V1 = Pan.V + 3, *V2 = Pan.V + 6, *Q = Pan.V + 3 * iQ, *XC = Pan.C;
  double uHat[3], vHat[3], wHat[3];
  { double n = 0.0; for (int m = 0; m < 3; m++) { uHat[m] = V1[m] - V0[m]; n += uHat[m] * uHat[m]; } n = sqrt(n); for (int m = 0; m < 3; m++) uHat[m] /= n; }
  for (int m = 0; m < 3; m++) vHat[m] = V2[m] - V0[m];
  wHat[0] = uHat[1] * vHat[2] - uHat[2] * vHat[1]; wHat[1] = uHat[2] * vHat[0] - uHat[0] * vHat[2]; wHat[2] = uHat[0] * vHat[1] - uHat[1] * vHat[0];
  { double n = sqrt(wHat[0] * wHat[0] + wHat[1] * wHat[1] + wHat[2] * wHat[2]); for (int m = 0; m < 3; m++) wHat[m] /= n; }
  vHat[0] = wHat[1] * uHat[2] - wHat[2] * uHat[1]; vHat[1] = wHat[2] * uHat[0] - wHat[0] * uHat[2]; vHat[2] = wHat[0] * uHat[1] - wHat[1] * uHat[0];
  double Rho0[3];
  for (int m = 0; m < 3; m++) Rho0[m] = X0[m] - XC[m];
  const double w0 = Rho0[0] * wHat[0] + Rho0[1] * wHat[1] + Rho0[2] * wHat[2];
  for (int m = 0; m < 3; m++) Rho0[m] = X0[m] + (-w0) * wHat[m];
  double QBarMu[3] = { 0.0, 0.0, 0.0 };
  for (int m = 0; m < 3; m++) { const double qb = Q[m] + (-1.0) * Rho0[m]; QBarMu[0] += qb * uHat[m]; QBarMu[1] += qb * vHat[m]; }
  double u[3], v[3];
  for (int i = 0; i < 3; i++)
   { u[i] = v[i] = 0.0;
     for (int m = 0; m < 3; m++) { const double d = Pan.V[3 * i + m] - Rho0[m]; u[i] += d * uHat[m]; v[i] += d * vHat[m]; } }
  double ScriptI[NF_NUMPS], dScriptIMu[NF_NUMPS][3], ScriptJMu[NF_NUMPS][3], dScriptJMuNu[NF_NUMPS][3][3];
  nf_script_ij(u, v, w0, Pan.R, ScriptI, dScriptIMu, ScriptJMu, dScriptJMuNu);
  double anMu[2][3], dpnMu[2][3], danMuNu[2][3][3];
  for (int NI = 0; NI < 2; NI++)
   { const int PI = NI == 0 ? 2 : 3; const double P = PV[PI];
     for (int Mu = 0; Mu < 3; Mu++)
      { dpnMu[NI][Mu] = dScriptIMu[PI][Mu];
        anMu[NI][Mu] = ScriptJMu[PI][Mu] - QBarMu[Mu] * ScriptI[PI];
        danMuNu[NI][Mu][2] = 0.0;
        danMuNu[NI][2][Mu] = P * w0 * (ScriptJMu[PI - 1][Mu] - QBarMu[Mu] * ScriptI[PI - 1]); }
     for (int Mu = 0; Mu < 2; Mu++)
      for (int Nu = 0; Nu < 2; Nu++)
       danMuNu[NI][Mu][Nu] = dScriptJMuNu[PI][Mu][Nu] + ((Mu == Nu) ? ScriptI[PI] : 0.0) + P * QBarMu[Nu] * ScriptJMu[PI - 1][Mu]; }
  const double *M[3] = { uHat, vHat, wHat };
  for (int nn = 0; nn < 2; nn++)
   { for (int i = 0; i < 3; i++)
      { an[nn][i] = dpn[nn][i] = 0.0;
        for (int Mu = 0; Mu < 3; Mu++) { an[nn][i] += M[Mu][i] * anMu[nn][Mu]; dpn[nn][i] += M[Mu][i] * dpnMu[nn][Mu]; } }
     for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
       { dan[nn][i][j] = 0.0;
         for (int Mu = 0; Mu < 3; Mu++) for (int Nu = 0; Nu < 3; Nu++) dan[nn][i][j] += M[Mu][i] * M[Nu][j] * danMuNu[nn][Mu][Nu]; } }
}

__device__ __forceinline__ cd nf_cexp(cd z) { const double e = exp(z.x); double s, c; sincos(z.y, &s, &c); return mk(e * c, e * s); }

// desingularised kernel and its gradient (GetGDS :544-611): G_DS = ERB3/(4 pi r), ERB_n = exp(x) - sum_{m<n} x^m/m!, x = ikr
__device__ inline cd nf_gds(const double R[3], cd k, cd dG[3])
{ const double r = sqrt(R[0] * R[0] + R[1] * R[1] + R[2] * R[2]);
  if (r == 0.0) { dG[0] = dG[1] = dG[2] = mk(0.0, 0.0); return mk(0.0, 0.0); }
  const cd IK = mk(-k.y, k.x), x = IK * r;
  const cd Term1 = x, xT1 = x * Term1, Term2 = mk(xT1.x / 2.0, xT1.y / 2.0), xT2 = x * Term2;
  cd Term = mk(xT2.x / 3.0, xT2.y / 3.0), ERB3;
  if (sqrt(norm2(x)) > 0.1) ERB3 = nf_cexp(x) - mk(1.0, 0.0) - Term1 - Term2;
  else
   { cd Sum = Term;
     for (int m = 4; m < 100; m++)
      { Term = Term * mk(x.x / (double)m, x.y / (double)m);
        Sum += Term;
        if (norm2(Term) < 1.0e-8 * 1.0e-8 * norm2(Sum)) break; }
     ERB3 = Sum; }
  const cd ERB2 = ERB3 + Term2;
  const double Denom = 4.0 * M_PI * r, Denom2 = Denom * r, Denom3 = Denom2 * r;
  const cd ie2 = IK * ERB2;
  const cd Factor1 = mk(ie2.x / Denom2 - ERB3.x / Denom3, ie2.y / Denom2 - ERB3.y / Denom3);
  for (int m = 0; m < 3; m++) dG[m] = Factor1 * R[m];
  return mk(ERB3.x / Denom, ERB3.y / Denom);
}

// e, h of the half basis function (panel Pan, source vertex iQ) at X0 (GetReducedPotentials_Nearby :702-797 + GetReducedFields_Nearby :837-866)
__device__ inline void nf_reduced_fields_panel(const PanelRec &Pan, int iQ, const double X0[3], cd k, const double *tcr9, cd e[3], cd h[3])
{ double an[2][3], dpn[2][3], dan[2][3][3];
  nf_static_panel(Pan, iQ, X0, an, dpn, dan);
  const double *Q = Pan.V + 3 * iQ, *V1 = Pan.V + 3 * ((iQ + 1) % 3), *V2 = Pan.V + 3 * ((iQ + 2) % 3);
  double A[3], B[3], AxB[3];
  for (int m = 0; m < 3; m++) { A[m] = V1[m] - Q[m]; B[m] = V2[m] - V1[m]; }
  // panel area as InitRWGPanel computes it: half the norm of (V1-V0) x (V2-V0)
  { double a0[3], b0[3]; for (int m = 0; m < 3; m++) { a0[m] = Pan.V[3 + m] - Pan.V[m]; b0[m] = Pan.V[6 + m] - Pan.V[m]; }
    AxB[0] = a0[1] * b0[2] - a0[2] * b0[1]; AxB[1] = a0[2] * b0[0] - a0[0] * b0[2]; AxB[2] = a0[0] * b0[1] - a0[1] * b0[0]; }
  const double Area = 0.5 * sqrt(AxB[0] * AxB[0] + AxB[1] * AxB[1] + AxB[2] * AxB[2]);
  double a0v[3]; for (int m = 0; m < 3; m++) a0v[m] = Area * (Pan.C[m] - Q[m]);
  // desingularised part: 21-point rule, X = Q + (u+v) A + v B (GetDSReducedPotentials :617-697)
  cd aDS[3], dpDS[3], daDS[3][3];
  for (int m = 0; m < 3; m++) { aDS[m] = dpDS[m] = mk(0.0, 0.0); for (int n = 0; n < 3; n++) daDS[m][n] = mk(0.0, 0.0); }
  const double PreFactor = 2.0 * Area;
  for (int p = 0; p < SC_TCR9_NPTS; p++)
   { double u = tcr9[3 * p]; const double v = tcr9[3 * p + 1], w = tcr9[3 * p + 2] * PreFactor;
     u += v;
     double F[3], R[3];
     for (int m = 0; m < 3; m++) { F[m] = u * A[m] + v * B[m]; const double X = F[m] + Q[m]; R[m] = X0[m] - X; }
     cd dG[3]; const cd GG = nf_gds(R, k, dG);
     for (int m = 0; m < 3; m++)
      { dpDS[m] += dG[m] * w;
        aDS[m] += GG * (w * F[m]);
        for (int n = 0; n < 3; n++) daDS[m][n] += dG[m] * (w * F[n]); } }
  const cd IK = mk(-k.y, k.x);
  const double CM1 = 1.0 / (4.0 * M_PI);
  const cd C0 = IK * (1.0 / (4.0 * M_PI)), CP1 = (IK * IK) * (1.0 / (8.0 * M_PI));
  cd a[3], dp[3], da[3][3];
  for (int m = 0; m < 3; m++)
   { a[m] = mk(CM1 * an[0][m], 0.0) + C0 * a0v[m] + CP1 * an[1][m] + aDS[m];
     dp[m] = mk(CM1 * dpn[0][m], 0.0) + CP1 * dpn[1][m] + dpDS[m];
     for (int n = 0; n < 3; n++) da[m][n] = mk(CM1 * dan[0][m][n], 0.0) + CP1 * dan[1][m][n] + daDS[m][n]; }
  double Length = 0.0; for (int m = 0; m < 3; m++) Length += (V1[m] - V2[m]) * (V1[m] - V2[m]); Length = sqrt(Length);
  const double aPrefac = Length / (2.0 * Area), pPrefac = 2.0 * aPrefac;
  const cd k2 = k * k;
  for (int m = 0; m < 3; m++)
   { const int m1 = (m + 1) % 3, m2 = (m + 2) % 3;
     e[m] = a[m] * aPrefac + cdiv(dp[m] * pPrefac, k2);
     h[m] = (da[m1][m2] - da[m2][m1]) * aPrefac; }
}

}  // namespace scuffcuda
