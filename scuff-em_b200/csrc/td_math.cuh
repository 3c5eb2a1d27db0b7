// Taylor-Duffy evaluation of singular panel-pair integrals (frequency dependent) and the p-adaptive
// Clenshaw-Curtis cubature that integrates the reduced integrand.
//
// Reference semantics: TaylorDuffy() libs/libscuff/TaylorDuffy.cc:277-368 (driver), TaylorDuffySum
// :152-258 (integrand: sum over sub-regions d and powers n of P_dn(y) K_n(X_d(y))), GetX :491-547,
// GetScriptP :552-622, GetScriptK :896-935, ExpRelV3P0 :829-891, ComputeGeometricParameters :1265-1526,
// and pcubature libs/libSGJC/pcubature.c:300-400 (nested Clenshaw-Curtis levels, error = difference
// to the previous level, per-component convergence test converged.h, stop when numEval > MaxEval).
// The polynomial coefficient maps (CMVStoUpsilon) come from generated_tables.h as data.
//
// Here: the shared integrand machinery and the common-triangle case (ncv == 3, 1-D integral), which the reference
// uses at every frequency (td_kernel.cu); the common-edge / common-vertex cases (2-D / 3-D, reached in the
// short-wavelength regime |k| R > pi or with SCUFF_DISABLE_CACHE) and the high-k kernels are in td_general.cu.
#pragma once
#include "ppi_math.cuh"
#include "fippi_math.cuh"

namespace scuffcuda {

struct UpsEntry { unsigned char cas, d, n, mono, input; double coef; };

struct TDTables
{ const double *ccx, *ccw;          // Clenshaw-Curtis nodes / weights (clencurt.h layout)
  const UpsEntry *ups; int upsStart[5];
};

#define TD_NUMWPOWERS 5
#define TD_Y1TERM 1

// complex helpers with std::complex-like semantics
SCUFF_HD cd cexp_(cd z) { double e = exp(z.x), s, c; sincos(z.y, &s, &c); return mk(e * c, e * s); }
SCUFF_HD double cabs_(cd z) { return hypot(z.x, z.y); }


// x / d for a small positive integer d (the divisors m + n of the ExpRel series).  The reference divides; an IEEE division costs
// ~25 instructions on the device (MUFU.RCP64H seed, Newton steps, range checks, slow-path call) and these divisions were half of
// the Taylor-Duffy kernels' issue slots.  With y = RN(1/d) tabulated, q0 = RN(x y), r = x - q0 d (exact in one FMA) and
// q = RN(q0 + r y) is the correctly rounded quotient (Markstein's division step: q0 is a faithful estimate) — three FP64
// instructions, same bits.  The host build (formula checks without a GPU) keeps the plain division.
#if defined(__CUDACC__)
static __constant__ double c_td_rcp[128] = { 0.0, 1.0, 0.5, 0.3333333333333333, 0.25, 0.2, 0.16666666666666666, 0.14285714285714285, 0.125, 0.1111111111111111, 0.1, 0.09090909090909091, 0.08333333333333333, 0.07692307692307693, 0.07142857142857142, 0.06666666666666667, 0.0625, 0.058823529411764705, 0.05555555555555555, 0.05263157894736842, 0.05, 0.047619047619047616, 0.045454545454545456, 0.043478260869565216, 0.041666666666666664, 0.04, 0.038461538461538464, 0.037037037037037035, 0.03571428571428571, 0.034482758620689655, 0.03333333333333333, 0.03225806451612903, 0.03125, 0.030303030303030304, 0.029411764705882353, 0.02857142857142857, 0.027777777777777776, 0.02702702702702703, 0.02631578947368421, 0.02564102564102564, 0.025, 0.024390243902439025, 0.023809523809523808, 0.023255813953488372, 0.022727272727272728, 0.022222222222222223, 0.021739130434782608, 0.02127659574468085, 0.020833333333333332, 0.02040816326530612, 0.02, 0.0196078431372549, 0.019230769230769232, 0.018867924528301886, 0.018518518518518517, 0.01818181818181818, 0.017857142857142856, 0.017543859649122806, 0.017241379310344827, 0.01694915254237288, 0.016666666666666666, 0.01639344262295082, 0.016129032258064516, 0.015873015873015872, 0.015625, 0.015384615384615385, 0.015151515151515152, 0.014925373134328358, 0.014705882352941176, 0.014492753623188406, 0.014285714285714285, 0.014084507042253521, 0.013888888888888888, 0.0136986301369863, 0.013513513513513514, 0.013333333333333334, 0.013157894736842105, 0.012987012987012988, 0.01282051282051282, 0.012658227848101266, 0.0125, 0.012345679012345678, 0.012195121951219513, 0.012048192771084338, 0.011904761904761904, 0.011764705882352941, 0.011627906976744186, 0.011494252873563218, 0.011363636363636364, 0.011235955056179775, 0.011111111111111112, 0.01098901098901099, 0.010869565217391304, 0.010752688172043012, 0.010638297872340425, 0.010526315789473684, 0.010416666666666666, 0.010309278350515464, 0.01020408163265306, 0.010101010101010102, 0.01, 0.009900990099009901, 0.00980392156862745, 0.009708737864077669, 0.009615384615384616, 0.009523809523809525, 0.009433962264150943, 0.009345794392523364, 0.009259259259259259, 0.009174311926605505, 0.00909090909090909, 0.009009009009009009, 0.008928571428571428, 0.008849557522123894, 0.008771929824561403, 0.008695652173913044, 0.008620689655172414, 0.008547008547008548, 0.00847457627118644, 0.008403361344537815, 0.008333333333333333, 0.008264462809917356, 0.00819672131147541, 0.008130081300813009, 0.008064516129032258, 0.008, 0.007936507936507936, 0.007874015748031496 };
#endif
SCUFF_HD double td_div_int(double x, int d)
{
#if defined(__CUDA_ARCH__)
  const double y = c_td_rcp[d], dv = (double)d;
  const double q0 = x * y, r = fma(-q0, dv, x);
  return fma(r, y, q0);
#else
  return x / ((double)d);
#endif
}
// x / d given y = RN(1/d) (one IEEE division shared by several quotients)
SCUFF_HD double td_div_rcp(double x, double d, double y)
{
#if defined(__CUDA_ARCH__)
  const double q0 = x * y, r = fma(-q0, d, x);
  return fma(r, y, q0);
#else
  (void)y; return x / d;
#endif
}

// ExpRel(n,Z) = (e^Z - first n Taylor terms) / (Z^n/n!)   (TaylorDuffy.cc:829-891)
SCUFF_HD cd exprel_v3p0(int n, cd Z)
{ if (Z.y == 0.0)
   { double X = Z.x;
     if (fabs(X) < 0.1)
      { double Sum = 1.0, Term = 1.0;
        for (int m = 1; m < 100; m++)
         { Term *= td_div_int(X, m + n); Sum += Term;
           if (fabs(Term) < 1.0e-8 * fabs(Sum)) break; }
        return mk(Sum, 0.0);
      }
     double Sum = exp(X), Term = 1.0;
     for (int m = 0; m < n; m++) { Sum -= Term; Term *= td_div_int(X, m + 1); }
     return mk(Sum / Term, 0.0);
   }
  if (cabs_(Z) < 0.1)
   { cd Sum = mk(1.0, 0.0), Term = mk(1.0, 0.0);
     for (int m = 1; m < 100; m++)
      { Term = Term * mk(td_div_int(Z.x, m + n), td_div_int(Z.y, m + n)); Sum += Term;
        if (norm2(Term) < 1.0e-16 * norm2(Sum)) break; }
     return Sum;
   }
  cd Sum = cexp_(Z), Term = mk(1.0, 0.0);
  for (int m = 0; m < n; m++)
   { Sum -= Term; Term = Term * mk(td_div_int(Z.x, m + 1), td_div_int(Z.y, m + 1)); }
  return cdiv(Sum, Term);
}

// per-call workspace for the common-triangle case with the two P polynomials the assembly needs:
// P0 = 1 (TD_UNITY), P1 = (x-Q).(x'-Q') (TD_PMCHWG1); kernel e^{ikr}/(4 pi r) (TD_HELMHOLTZ)
struct TDCommonTriangle
{ double A2, AdB, B2;
  double U[2][3][TD_NUMWPOWERS][2];   // [P][region][n][ {1, y} ]
  int nMin[2], nMax[2];
  cd k;
};

SCUFF_HD void td_ct_setup(const TDTables &T, const double *V1, const double *V2, const double *V3,
                          const double *Q, const double *QP, cd k, TDCommonTriangle &W)
{ double A[3], B[3], D[3], DP[3];
  sub3(V2, V1, A); sub3(V3, V2, B); sub3(V1, Q, D); sub3(V1, QP, DP);
  W.A2 = dot3(A, A); W.AdB = dot3(A, B); W.B2 = dot3(B, B); W.k = k;
  for (int np = 0; np < 2; np++)
   { double in[19];
     for (int i = 0; i < 19; i++) in[i] = 0.0;
     if (np == 0) in[18] = 1.0;                                   // S = 1
     else
      { // common triangle: A' = A, B' = B  (inputs: 4 C, then M11 M12 M13 M14 M22 M23 M24 M33 M34 M44, V1..V4, S)
        in[4 + 2] = dot3(A, A); in[4 + 3] = dot3(A, B); in[4 + 5] = dot3(B, A); in[4 + 6] = dot3(B, B);
        in[14] = dot3(A, DP); in[15] = dot3(B, DP); in[16] = dot3(D, A); in[17] = dot3(D, B);
        in[18] = dot3(D, DP);
      }
     for (int d = 0; d < 3; d++) for (int n = 0; n < TD_NUMWPOWERS; n++) { W.U[np][d][n][0] = 0.0; W.U[np][d][n][1] = 0.0; }
     int nMin = TD_NUMWPOWERS - 1, nMax = -1;
     double U2[3][TD_NUMWPOWERS];   // y^2 monomial: contributes to the nMin/nMax scan only
     for (int d = 0; d < 3; d++) for (int n = 0; n < TD_NUMWPOWERS; n++) U2[d][n] = 0.0;
     for (int e = T.upsStart[3]; e < T.upsStart[4]; e++)
      { const UpsEntry &E = T.ups[e];
        double v = E.coef * in[E.input];
        if (E.mono == 0) W.U[np][E.d][E.n][0] += v;
        else if (E.mono == TD_Y1TERM) W.U[np][E.d][E.n][1] += v;
        else U2[E.d][E.n] += v;
      }
     for (int n = 0; n < TD_NUMWPOWERS; n++) for (int d = 0; d < 3; d++)
      if (fabs(W.U[np][d][n][0]) > 1.0e-8 || fabs(W.U[np][d][n][1]) > 1.0e-8 || fabs(U2[d][n]) > 1.0e-8)
       { if (n < nMin) nMin = n; if (n > nMax) nMax = n; }
     W.nMin[np] = nMin; W.nMax[np] = nMax;
   }
}

// ExpRel(n,Z) with exp(Z) supplied by the caller (same value cexp_(Z) would return: hoisted out of the loop over n)
SCUFF_HD cd exprel_v3p0_pre(int n, cd Z, cd expZ, bool small)
{ if (small) return exprel_v3p0(n, Z);
  if (Z.y == 0.0)
   { double X = Z.x, Sum = expZ.x, Term = 1.0;
     for (int m = 0; m < n; m++) { Sum -= Term; Term *= td_div_int(X, m + 1); }
     return mk(Sum / Term, 0.0);
   }
  cd Sum = expZ, Term = mk(1.0, 0.0);
  for (int m = 0; m < n; m++)
   { Sum -= Term; Term = Term * mk(td_div_int(Z.x, m + 1), td_div_int(Z.y, m + 1)); }
  return cdiv(Sum, Term);
}

// integrand at y in [0,1]: f[0..3] = Re/Im of the two P x Helmholtz integrals (TaylorDuffySum).
// K_n(X_d) depends on neither P polynomial, so it is evaluated once for the union of their n ranges, and the two
// complex exponentials per sub-region are evaluated once (the reference recomputes them for every n and P; same values).
SCUFF_HD void td_ct_integrand(const TDCommonTriangle &W, double y, double f[4])
{ double X[3], u1, u2;
  u1 = 1.0; u2 = y;       X[0] = sqrt(W.A2 * u1 * u1 + 2.0 * W.AdB * u1 * u2 + W.B2 * u2 * u2);
  u1 = y;   u2 = (y - 1); X[1] = sqrt(W.A2 * u1 * u1 + 2.0 * W.AdB * u1 * u2 + W.B2 * u2 * u2);
  u1 = y;   u2 = 1.0;     X[2] = sqrt(W.A2 * u1 * u1 + 2.0 * W.AdB * u1 * u2 + W.B2 * u2 * u2);
  const cd IK = mk(-W.k.y, W.k.x);
  const int nLo = (W.nMin[0] < W.nMin[1] ? W.nMin[0] : W.nMin[1]), nHi = (W.nMax[0] > W.nMax[1] ? W.nMax[0] : W.nMax[1]);
  cd K[3][7];
  for (int d = 0; d < 3; d++)
   { cd IKX = mk(IK.x * X[d], IK.y * X[d]), eIKX = cexp_(IKX), mIKX = mk(-IKX.x, -IKX.y);
     // exprel_v3p0's own branch conditions (TaylorDuffy.cc:829-891)
     const bool small = (mIKX.y == 0.0) ? (fabs(mIKX.x) < 0.1) : (cabs_(mIKX) < 0.1);
     cd emIKX = mk(0.0, 0.0);
     if (!small) emIKX = (mIKX.y == 0.0) ? mk(exp(mIKX.x), 0.0) : cexp_(mIKX);
     for (int n = nLo + 1; n <= nHi + 1; n++)
      { cd t = eIKX * exprel_v3p0_pre(n, mIKX, emIKX, small); const double dn = n * X[d], rdn = 1.0 / dn; K[d][n] = mk(td_div_rcp(t.x, dn, rdn), td_div_rcp(t.y, dn, rdn)); }
   }
  for (int np = 0; np < 2; np++)
   { cd Sum = mk(0.0, 0.0);
     for (int n = W.nMin[np]; n <= W.nMax[np]; n++)
      for (int d = 0; d < 3; d++)
       { double p = W.U[np][d][n][0] + y * W.U[np][d][n][1];
         Sum += p * K[d][n + 1]; }
     const double s = 1.0 / (4.0 * M_PI);     // Jacobian = 1
     f[2 * np] = Sum.x * s; f[2 * np + 1] = Sum.y * s;
   }
}

// 1-D p-adaptive Clenshaw-Curtis over [0,1] for an FDIM-component integrand (pcubature.c semantics).
// cache: caller-provided scratch of at least FDIM*(2^(maxM+1)+1) doubles, element e at cache[e*stride]
// (stride = number of threads sharing the scratch block on the device, so accesses coalesce).
template <int FDIM, typename F>
SCUFF_HD int pcubature_1d(const TDTables &T, F &func, double relTol, double absTol, int maxEval,
                          double *cache, size_t stride, double val[FDIM], double err[FDIM], int maxM)
{ const double c = 0.5, r = 0.5, V = 0.5;
  int m = 0, numEval = 0;
  { double fv[FDIM];
    func(c, fv);                  for (int j = 0; j < FDIM; j++) cache[(size_t)(j) * stride] = fv[j];
    func(c + r * T.ccx[0], fv);   for (int j = 0; j < FDIM; j++) cache[(size_t)(FDIM + j) * stride] = fv[j];
    func(c - r * T.ccx[0], fv);   for (int j = 0; j < FDIM; j++) cache[(size_t)(2 * FDIM + j) * stride] = fv[j];
  }
  double prev[FDIM];
  for (int j = 0; j < FDIM; j++) prev[j] = 0.0;
  for (;;)
   { // rule of order m from the cached values; the error estimate is the difference to the rule of order m-1, which
     // is exactly the previous iteration's value (same points, weights and summation order), or the midpoint rule at m = 0
     { double acc[FDIM];
       for (int j = 0; j < FDIM; j++) acc[j] = 0.0;
       const int mid = m;
       const double *wb = T.ccw + mid + (1 << mid) - 1;
       double w = V * wb[0];
       for (int j = 0; j < FDIM; j++) acc[j] += cache[(size_t)(j) * stride] * w;
       w = V * wb[1];
       for (int j = 0; j < FDIM; j++) acc[j] += cache[(size_t)(FDIM + j) * stride] * w;
       for (int j = 0; j < FDIM; j++) acc[j] += cache[(size_t)(2 * FDIM + j) * stride] * w;
       for (int kk = 1; kk <= mid; kk++)
        { const double *wk = wb + 1 + (1 << (kk - 1));
          const int nx = 1 << (kk - 1);
          const double *cv = cache + (size_t)FDIM * ((1 << kk) + 1) * stride;   // level kk starts after 2^kk+1 points
#pragma unroll 2
          for (int i = 0; i < nx; i++)
           { double ww = V * wk[i];
             for (int j = 0; j < FDIM; j++) acc[j] += cv[(size_t)((2 * i) * FDIM + j) * stride] * ww;
             for (int j = 0; j < FDIM; j++) acc[j] += cv[(size_t)((2 * i + 1) * FDIM + j) * stride] * ww;
           }
        }
       if (m == 0) { const double w0 = V * 2; for (int j = 0; j < FDIM; j++) prev[j] = 0.0 + cache[(size_t)(j) * stride] * w0; }
       for (int j = 0; j < FDIM; j++) { val[j] = acc[j]; err[j] = fabs(val[j] - prev[j]); prev[j] = acc[j]; }
     }
     bool conv = true;
     for (int j = 0; j < FDIM; j++) if (err[j] > absTol && err[j] > fabs(val[j]) * relTol) { conv = false; break; }
     if (conv || (numEval > maxEval && maxEval)) return 0;
     m += 1;
     if (m > maxM) return 1;
     const int nx = 1 << (m - 1);
     const double *x = T.ccx + nx;
     double *cv = cache + (size_t)FDIM * ((1 << m) + 1) * stride;
     for (int i = 0; i < nx; i++)
      { double fv[FDIM];
        func(c + r * x[i], fv); for (int j = 0; j < FDIM; j++) cv[(size_t)((2 * i) * FDIM + j) * stride] = fv[j];
        func(c - r * x[i], fv); for (int j = 0; j < FDIM; j++) cv[(size_t)((2 * i + 1) * FDIM + j) * stride] = fv[j];
      }
     numEval += 2 * nx;
   }
}

struct TDCTFunctor
{ const TDCommonTriangle *W;
  SCUFF_HD void operator()(double y, double f[4]) const { td_ct_integrand(*W, y, f); }
};

// H[0], H[1] of a common-triangle panel pair (PanelPanelInteractions.cc:449-491 with ncv == 3):
// H0 = R1 - 4 R0 / k^2, H1 = 0.  Vertices in AssessPanelPair order (V1,V2,V3).  Returns pcubature status.
SCUFF_HD int td_common_triangle(const TDTables &T, const double *V1, const double *V2, const double *V3,
                                const double *Q, const double *QP, cd k, double *cache, size_t stride, cd &H0, cd &H1, int maxM = 9)
{ TDCommonTriangle W;
  td_ct_setup(T, V1, V2, V3, Q, QP, k, W);
  TDCTFunctor F; F.W = &W;
  double val[4], err[4];
  int st = pcubature_1d<4>(T, F, 1.0e-10, 0.0, 1000, cache, stride, val, err, maxM);
  cd R0 = mk(val[0], val[1]), R1 = mk(val[2], val[3]);
  cd k2 = k * k;
  cd t = cdiv(mk(4.0 * R0.x, 4.0 * R0.y), k2);
  H0 = R1 - t; H1 = mk(0.0, 0.0);
  return st;
}

#define TD_CACHE_DOUBLES (4 * ((1 << 10) + 1))   /* MaxEval = 1000 stops at level 9: 1025 points */

}  // namespace scuffcuda
