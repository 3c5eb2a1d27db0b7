// Panel-pair classification with the reference's exact floating-point decisions.
// Reference: AssessPanelPair (AssessPanelPair.cc:52-103,147-172), the displaced variant in
// GetPanelPanelInteractions (PanelPanelInteractions.cc:389-414), VecClose (libhrutil/Vector.cc:294-297),
// CanonicallyOrderVertices / VLT (AssessPanelPair.cc:257-455).
// Every threshold comparison uses explicitly-rounded mul/add (no FMA contraction) so that it falls
// on the same side as the reference's SSE2 arithmetic.
#pragma once
#include "ppi_math.cuh"

namespace scuffcuda {

#if defined(__CUDA_ARCH__)
#define SC_MUL(a, b) __dmul_rn(a, b)
#define SC_ADD(a, b) __dadd_rn(a, b)
#define SC_SUB(a, b) __dsub_rn(a, b)
#else
#define SC_MUL(a, b) ((a) * (b))
#define SC_ADD(a, b) ((a) + (b))
#define SC_SUB(a, b) ((a) - (b))
#endif

#define SCUFF_DESING_RADIUS 4.0

// D = Ca - Cb (- L).  Returns rRel = |D| / max(Ra,Rb).
SCUFF_HD double pair_rrel(const double Ca[3], double Ra, const double Cb[3], double Rb,
                          bool useDisp, const double L[3], double D[3])
{ if (useDisp)
   { D[0] = SC_SUB(SC_SUB(Ca[0], Cb[0]), L[0]);
     D[1] = SC_SUB(SC_SUB(Ca[1], Cb[1]), L[1]);
     D[2] = SC_SUB(SC_SUB(Ca[2], Cb[2]), L[2]); }
  else
   { D[0] = SC_SUB(Ca[0], Cb[0]); D[1] = SC_SUB(Ca[1], Cb[1]); D[2] = SC_SUB(Ca[2], Cb[2]); }
  double r2 = SC_MUL(D[0], D[0]);
  r2 = SC_ADD(r2, SC_MUL(D[1], D[1]));
  r2 = SC_ADD(r2, SC_MUL(D[2], D[2]));
  return sqrt(r2) / fmax(Ra, Rb);
}

SCUFF_HD bool vec_close(const double *a, const double *b, double abstol)
{ double s = SC_ADD(SC_ADD(fabs(SC_SUB(a[0], b[0])), fabs(SC_SUB(a[1], b[1]))), fabs(SC_SUB(a[2], b[2])));
  return s <= SC_MUL(3.0, abstol);
}

// AssessPanelPair(Va,Vb,Radius): count common vertices and produce the common-first order.
// Va/Vb are 9-double arrays; ia/ib receive vertex permutations.
SCUFF_HD int assess_vertices(const double *Va, const double *Vb, double rMax, unsigned char ia[3], unsigned char ib[3])
{ int ncv = 0; int CVIa[3], CVIb[3];
  double thr = fabs(SC_MUL(1.0e-6, rMax));
  for (int a = 0; a < 3; a++)
   for (int b = 0; b < 3; b++)
    if (vec_close(Va + 3 * a, Vb + 3 * b, thr))
     { if (ncv < 3) { CVIa[ncv] = a; CVIb[ncv] = b; } ncv++; }
  if (ncv > 3) ncv = 3;   // degenerate input; the reference would overflow its arrays
  ia[0] = 0; ia[1] = 1; ia[2] = 2; ib[0] = 0; ib[1] = 1; ib[2] = 2;
  if (ncv == 1)
   { ia[0] = CVIa[0]; ia[1] = (CVIa[0] + 1) % 3; ia[2] = (CVIa[0] + 2) % 3;
     ib[0] = CVIb[0]; ib[1] = (CVIb[0] + 1) % 3; ib[2] = (CVIb[0] + 2) % 3; }
  else if (ncv == 2)
   { ia[0] = CVIa[0]; ia[1] = CVIa[1]; ia[2] = 3 - CVIa[0] - CVIa[1];
     ib[0] = CVIb[0]; ib[1] = CVIb[1]; ib[2] = 3 - CVIb[0] - CVIb[1]; }
  else if (ncv == 3)
   { ia[0] = CVIa[0]; ia[1] = CVIa[1]; ia[2] = CVIa[2];
     ib[0] = CVIb[0]; ib[1] = CVIb[1]; ib[2] = CVIb[2]; }
  return ncv;
}

// VLT: 'vertex less than' (AssessPanelPair.cc:257-277)
SCUFF_HD int vlt(const double *V1, const double *V2)
{ double DV;
  DV = SC_SUB(V1[0], V2[0]);
  if (fabs(DV) > SC_MUL(1.0e-6, fabs(V1[0]))) return DV < 0.0 ? 1 : 0;
  DV = SC_SUB(V1[1], V2[1]);
  if (fabs(DV) > SC_MUL(1.0e-6, fabs(V1[1]))) return DV < 0.0 ? 1 : 0;
  DV = SC_SUB(V1[2], V2[2]);
  if (fabs(DV) > SC_MUL(1.0e-6, fabs(V1[2]))) return DV < 0.0 ? 1 : 0;
  return 1;
}

// CanonicallyOrderVertices.  Va[i], Vb[i] are the vertices in AssessPanelPair order.
// Outputs index arrays into Va / Vb: OVa[i] = (flipped ? Vb : Va)[oa[i]], OVb[i] = (flipped ? Va : Vb)[ob[i]].
// NOTE the ncv==0 branch reproduces the reference as written (AssessPanelPair.cc:413-451): panel b's
// vertices are permuted with panel a's permutation and the swap test compares Va[iMina] with Va[iMinb].
SCUFF_HD int canonical_order(const double *const Va[3], const double *const Vb[3], int ncv,
                             unsigned char oa[3], unsigned char ob[3])
{ int iMina, iMeda = 0, iMaxa, iMinb, iMaxb;
  if (ncv == 3)
   { if (vlt(Va[0], Va[1])) { iMina = 0; iMaxa = 1; } else { iMina = 1; iMaxa = 0; }
     if (vlt(Va[2], Va[iMina])) { iMeda = iMina; iMina = 2; }
     else if (vlt(Va[2], Va[iMaxa])) { iMeda = 2; }
     else { iMeda = iMaxa; iMaxa = 2; }
     oa[0] = iMina; oa[1] = iMeda; oa[2] = iMaxa; ob[0] = iMina; ob[1] = iMeda; ob[2] = iMaxa;
     return 0;
   }
  else if (ncv == 2)
   { int a01Min = vlt(Va[0], Va[1]) ? 0 : 1;
     oa[0] = a01Min; oa[1] = 1 - a01Min; oa[2] = 2; ob[0] = a01Min; ob[1] = 1 - a01Min; ob[2] = 2;
     return vlt(Va[2], Vb[2]) ? 0 : 1;
   }
  else if (ncv == 1)
   { int a12Min = vlt(Va[1], Va[2]) ? 1 : 2;
     int b12Min = vlt(Vb[1], Vb[2]) ? 1 : 2;
     if (vlt(Va[a12Min], Vb[b12Min]))
      { oa[0] = 0; oa[1] = a12Min; oa[2] = 3 - a12Min; ob[0] = 0; ob[1] = b12Min; ob[2] = 3 - b12Min; return 0; }
     else
      { // OVa = Vb[0,b12Min,3-b12Min], OVb = Va[0,a12Min,3-a12Min]
        oa[0] = 0; oa[1] = b12Min; oa[2] = 3 - b12Min; ob[0] = 0; ob[1] = a12Min; ob[2] = 3 - a12Min; return 1; }
   }
  else
   { if (vlt(Va[0], Va[1])) { iMina = 0; iMaxa = 1; } else { iMina = 1; iMaxa = 0; }
     if (vlt(Va[2], Va[iMina])) { iMeda = iMina; iMina = 2; }
     else if (vlt(Va[2], Va[iMaxa])) { iMeda = 2; }
     else { iMeda = iMaxa; iMaxa = 2; }
     if (vlt(Vb[0], Vb[1])) { iMinb = 0; iMaxb = 1; } else { iMinb = 1; iMaxb = 0; }
     if (vlt(Vb[2], Vb[iMinb])) { iMinb = 2; }
     else if (vlt(Vb[2], Vb[iMaxb])) { }
     else { iMaxb = 2; }
     (void)iMaxb;
     // both outcomes use panel a's permutation for both panels (as the reference does)
     oa[0] = iMina; oa[1] = iMeda; oa[2] = iMaxa; ob[0] = iMina; ob[1] = iMeda; ob[2] = iMaxa;
     return vlt(Va[iMina], Va[iMinb]) ? 0 : 1;
   }
}

// FIPPI cache key (FIPPICache.cc:120-156): 15 floats, vertex differences taken in FLOAT arithmetic
// relative to OVa[0] — translation invariant to float precision.
SCUFF_HD void fippi_key(const double *const OVa[3], const double *const OVb[3], float k[15])
{ for (int m = 0; m < 3; m++)
   { float v0 = (float)OVa[0][m];
     k[0 + m] = ((float)OVa[1][m]) - v0;
     k[3 + m] = ((float)OVa[2][m]) - v0;
     k[6 + m] = ((float)OVb[0][m]) - v0;
     k[9 + m] = ((float)OVb[1][m]) - v0;
     k[12 + m] = ((float)OVb[2][m]) - v0;
   }
}

}  // namespace scuffcuda
