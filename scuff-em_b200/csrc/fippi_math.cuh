// Q-dependent assembly of the frequency-independent panel-panel integrals (FIPPIs) and their
// contribution to the desingularised panel-pair integrals.
// Reference semantics: GetQDFIPPIData libs/libscuff/QDFIPPI.cc:67-260 (Q-dependent combination of the
// 33 cached integrals, r^0 closed forms :50-58, translated-record fix-up :143-158, sign flip :253-258),
// QIFIPPIData layout libscuffInternals.h:282-293, and the series coefficients of
// GetPanelPanelInteractions PanelPanelInteractions.cc:68-75,524-553.
#pragma once
#include "ppi_math.cuh"

namespace scuffcuda {

// record layout (36 doubles): xMxpRM3[3], xXxpRM3[3], uvupvpRM1[9], uvupvpR1[9], uvupvpR2[9], x0[3]
#define FIPPI_REC_DOUBLES 36

struct QDFIPPI
{ double hTimesRM3;
  double hDotRM1, hNablaRM1, hTimesRM1;
  double hDotR0, hNablaR0, hTimesR0;
  double hDotR1, hNablaR1, hTimesR1;
  double hDotR2, hNablaR2;
};

SCUFF_HD double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
SCUFF_HD void cross3(const double *a, const double *b, double *c)
{ c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0]; }
SCUFF_HD void sub3(const double *a, const double *b, double *c) { c[0] = a[0] - b[0]; c[1] = a[1] - b[1]; c[2] = a[2] - b[2]; }

// index order of the nine (u,v,u',v') monomials: 1, u', v', u, uu', uv', v, vu', vv'
// OVa/OVb: canonically ordered vertices; OQa/OQb: the RWG source/sink vertices in the same
// (possibly swapped) panel order; Qa/Qb: the un-swapped ones; CentroidForL: centroid of the panel
// that became OVa (Vb if flipped), used only for the translated-record correction.
SCUFF_HD void qdfippi(const double *const OVa[3], const double *OQa, const double *const OVb[3], const double *OQb,
                      const double *Qa, const double *Qb, const double CentroidForL[3], int flipped,
                      const double *rec, QDFIPPI &o)
{ const double *xMxpRM3 = rec, *xXxpRM3 = rec + 3, *RM1 = rec + 6, *R1 = rec + 15, *R2 = rec + 24, *x0 = rec + 33;
  double A[3], B[3], V0mQ[3], AP[3], BP[3], V0PmQP[3];
  sub3(OVa[1], OVa[0], A); sub3(OVa[2], OVa[1], B); sub3(OVa[0], OQa, V0mQ);
  sub3(OVb[1], OVb[0], AP); sub3(OVb[2], OVb[1], BP); sub3(OVb[0], OQb, V0PmQP);
  // dot-product coefficients in monomial order
  double cd[9] = { dot3(V0mQ, V0PmQP), dot3(V0mQ, AP), dot3(V0mQ, BP),
                   dot3(A, V0PmQP), dot3(A, AP), dot3(A, BP),
                   dot3(B, V0PmQP), dot3(B, AP), dot3(B, BP) };
  const double *V0 = OVa[0], *V0P = OVb[0];
  double QamQb[3], QaxQb[3], V0mV0P[3], s[3];
  sub3(V0, V0P, V0mV0P); sub3(Qa, Qb, QamQb); cross3(Qa, Qb, QaxQb);
  // cross-product ("times") coefficients in monomial order
  double ct[9];
  cross3(V0, V0P, s); ct[0] = dot3(QaxQb, V0mV0P) + dot3(QamQb, s);
  cross3(V0, AP, s);  ct[1] = -dot3(QaxQb, AP) + dot3(QamQb, s);
  cross3(V0, BP, s);  ct[2] = -dot3(QaxQb, BP) + dot3(QamQb, s);
  cross3(A, V0P, s);  ct[3] = dot3(QaxQb, A) + dot3(QamQb, s);
  cross3(A, AP, s);   ct[4] = dot3(QamQb, s);
  cross3(A, BP, s);   ct[5] = dot3(QamQb, s);
  cross3(B, V0P, s);  ct[6] = dot3(QaxQb, B) + dot3(QamQb, s);
  cross3(B, AP, s);   ct[7] = dot3(QamQb, s);
  cross3(B, BP, s);   ct[8] = dot3(QamQb, s);

  o.hTimesRM3 = dot3(QaxQb, xMxpRM3) + dot3(QamQb, xXxpRM3);
  // record computed for a translated copy of this pair? (FIPPICache keys are translation invariant)
  double L[3] = { x0[0] - CentroidForL[0], x0[1] - CentroidForL[1], x0[2] - CentroidForL[2] };
  if (sqrt(dot3(L, L)) > 1.0e-8 * sqrt(dot3(A, A)))
   { cross3(L, QamQb, s); o.hTimesRM3 -= dot3(s, xMxpRM3); }

  // the reference sums the nine terms in the order 1,u',v',u,uu',uv',v,vu',vv' for hDot and
  // 1,u,v,u',v',uu',uv',vu',vv' for hTimes
  const int td[9] = { 0, 3, 6, 1, 2, 4, 5, 7, 8 };
  // r^0 integrals of the monomials over the two unit triangles (closed form)
  const double R0[9] = { 1.0 / 4.0, 1.0 / 6.0, 1.0 / 12.0, 1.0 / 6.0, 1.0 / 9.0, 1.0 / 18.0, 1.0 / 12.0, 1.0 / 18.0, 1.0 / 36.0 };
  double dM1 = 0, d0 = 0, d1 = 0, d2 = 0, tM1 = 0, t0 = 0, t1 = 0;
  for (int n = 0; n < 9; n++) { dM1 += RM1[n] * cd[n]; d0 += R0[n] * cd[n]; d1 += R1[n] * cd[n]; d2 += R2[n] * cd[n]; }
  for (int q = 0; q < 9; q++) { int n = td[q]; tM1 += RM1[n] * ct[n]; t0 += R0[n] * ct[n]; t1 += R1[n] * ct[n]; }
  o.hDotRM1 = dM1; o.hNablaRM1 = 4.0 * RM1[0]; o.hTimesRM1 = tM1;
  o.hDotR0 = d0;   o.hNablaR0 = 1.0;           o.hTimesR0 = t0;
  o.hDotR1 = d1;   o.hNablaR1 = 4.0 * R1[0];   o.hTimesR1 = t1;
  o.hDotR2 = d2;   o.hNablaR2 = 4.0 * R2[0];
  if (flipped) { o.hTimesRM3 *= -1.0; o.hTimesRM1 *= -1.0; o.hTimesR0 *= -1.0; o.hTimesR1 *= -1.0; }
}

// singular-term contributions: H0 += sum_n (ik)^n/(4 pi) A_n (hDot_n + hNabla_n/(ik)^2),
// H1 += sum_n (ik)^n/(4 pi) B_n hTimes_n, A = {1,1,1/2,1/6}, B = {-1,-,1/2,1/3,1/6}
SCUFF_HD void add_fippi_terms(const QDFIPPI &q, const KParams &K, cd &H0, cd &H1)
{ cd ik = K.ik;
  cd ooik2 = cinv(ik * ik);
  cd PF0 = mk(1.0 / (4.0 * M_PI), 0.0), PF1 = ik * PF0, PF2 = ik * PF1, PF3 = ik * PF2, PF4 = ik * PF3;
  cd t;
  t = mk(q.hDotRM1, 0.0) + ooik2 * q.hNablaRM1; H0 += PF0 * t;
  t = mk(q.hDotR0, 0.0) + ooik2 * q.hNablaR0;   H0 += PF1 * t;
  t = mk(q.hDotR1, 0.0) + ooik2 * q.hNablaR1;   H0 += (0.5 * PF2) * t;
  t = mk(q.hDotR2, 0.0) + ooik2 * q.hNablaR2;   H0 += ((1.0 / 6.0) * PF3) * t;
  H1 += (-1.0 * q.hTimesRM3) * PF0;
  H1 += (0.5 * q.hTimesRM1) * PF2;
  H1 += ((1.0 / 3.0) * q.hTimesR0) * PF3;
  H1 += ((1.0 / 6.0) * q.hTimesR1) * PF4;
}

}  // namespace scuffcuda
