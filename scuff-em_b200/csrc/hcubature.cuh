// h-adaptive cubature with the exact decision sequence of the reference's libSGJC (S. G. Johnson's
// hcubature, non-"parallel" branch): embedded Genz-Malik degree-7/5 rule in 2-D/3-D, 15-point
// Gauss-Kronrod in 1-D, max-error binary heap, bisection along the dimension of largest fourth
// difference, per-component convergence test (ERROR_INDIVIDUAL), evaluation budget maxEval.
// Reference: libs/libSGJC/hcubature.c:455-576 (Genz-Malik), :619-756 (GK15), :778-880 (heap),
// :897-1013 (rulecubature), converged.h, libSGJC.c:46-61 (adapt_integrate).
//
// Why exact: the reference's FIPPI integrals are only converged to 1e-5..1e-6 (QIFIPPICubature.cc:41-43,
// QIFIPPITaylorDuffyV2P0.cc:47-49), so matching its matrix to 1e-10 requires taking the same
// subdivision decisions; sums are accumulated in the same order and without FMA contraction.
//
// One integration = one thread.  Regions live in a caller-provided pool (HBM scratch).
#pragma once
#include "ppi_math.cuh"
#include "classify.cuh"   // SC_MUL / SC_ADD / SC_SUB (no FMA contraction)

namespace scuffcuda {

#define HC_FDIM 33

template <int DIM> struct HCRegion
{ double val[HC_FDIM], err[HC_FDIM];
  double c[DIM], hw[DIM];
  double vol, errmax;
  int splitDim; int pad;
};

// strictly ordered arithmetic helpers
SCUFF_HD double hmul(double a, double b) { return SC_MUL(a, b); }
SCUFF_HD double hadd(double a, double b) { return SC_ADD(a, b); }
SCUFF_HD double hsub(double a, double b) { return SC_SUB(a, b); }

// ---- Genz-Malik 7/5 rule on one region (DIM = 2 or 3) ----
template <int DIM, typename F>
SCUFF_HD void gm_eval(F &f, HCRegion<DIM> &R)
{ const double lambda2 = 0.3585685828003180919906451539079374954541;
  const double lambda4 = 0.9486832980505137995996680633298155601160;
  const double lambda5 = 0.6882472016116852977216287342936235251269;
  const double weight2 = 980. / 6561., weight4 = 200. / 19683.;
  const double weightE2 = 245. / 486., weightE4 = 25. / 729.;
  const double ratio = (lambda2 * lambda2) / (lambda4 * lambda4);
  const double weight1 = ((double)(12824 - 9120 * DIM + 400 * DIM * DIM)) / 19683.0;
  const double weight3 = ((double)(1820 - 400 * DIM)) / 19683.0;
  const double weight5 = 6859.0 / 19683.0 / ((double)(1U << DIM));
  const double weightE1 = ((double)(729 - 950 * DIM + 50 * DIM * DIM)) / 729.0;
  const double weightE3 = ((double)(265 - 100 * DIM)) / 1458.0;

  double p[DIM], r2[DIM], r4[DIM], r5[DIM];
  for (int i = 0; i < DIM; i++) { p[i] = R.c[i]; r2[i] = hmul(R.hw[i], lambda2); r4[i] = hmul(R.hw[i], lambda4); r5[i] = hmul(R.hw[i], lambda5); }

  double val0[HC_FDIM], sum2[HC_FDIM], sum3[HC_FDIM], sum4[HC_FDIM], sum5[HC_FDIM];
  double v0[HC_FDIM], v1[HC_FDIM], v2[HC_FDIM], v3[HC_FDIM];
  double diff[DIM];
  f(p, val0);
  for (int j = 0; j < HC_FDIM; j++) { sum2[j] = 0; sum3[j] = 0; sum4[j] = 0; sum5[j] = 0; }
  for (int k = 0; k < DIM; k++)
   { p[k] = hsub(R.c[k], r2[k]); f(p, v0);
     p[k] = hadd(R.c[k], r2[k]); f(p, v1);
     p[k] = hsub(R.c[k], r4[k]); f(p, v2);
     p[k] = hadd(R.c[k], r4[k]); f(p, v3);
     p[k] = R.c[k];
     double dk = 0;
     for (int j = 0; j < HC_FDIM; j++)
      { double s01 = hadd(v0[j], v1[j]), s23 = hadd(v2[j], v3[j]);
        sum2[j] = hadd(sum2[j], s01); sum3[j] = hadd(sum3[j], s23);
        double tv = hmul(2, val0[j]);
        double t = hsub(hsub(s01, tv), hmul(ratio, hsub(s23, tv)));
        dk = hadd(dk, fabs(t));
      }
     diff[k] = dk;
   }
  // (lambda4, lambda4, 0, ...) points
  for (int i = 0; i < DIM - 1; i++)
   { p[i] = hsub(R.c[i], r4[i]);
     for (int jj = i + 1; jj < DIM; jj++)
      { p[jj] = hsub(R.c[jj], r4[jj]); f(p, v0); for (int j = 0; j < HC_FDIM; j++) sum4[j] = hadd(sum4[j], v0[j]);
        p[i] = hadd(R.c[i], r4[i]);   f(p, v0); for (int j = 0; j < HC_FDIM; j++) sum4[j] = hadd(sum4[j], v0[j]);
        p[jj] = hadd(R.c[jj], r4[jj]); f(p, v0); for (int j = 0; j < HC_FDIM; j++) sum4[j] = hadd(sum4[j], v0[j]);
        p[i] = hsub(R.c[i], r4[i]);   f(p, v0); for (int j = 0; j < HC_FDIM; j++) sum4[j] = hadd(sum4[j], v0[j]);
        p[jj] = R.c[jj];
      }
     p[i] = R.c[i];
   }
  // (lambda5, ..., lambda5) points in Gray-code order
  { unsigned signs = 0;
    for (int i = 0; i < DIM; i++) p[i] = hadd(R.c[i], r5[i]);
    for (unsigned i = 0;; ++i)
     { f(p, v0); for (int j = 0; j < HC_FDIM; j++) sum5[j] = hadd(sum5[j], v0[j]);
       unsigned d = 0, n = i; while (n & 1u) { n >>= 1; d++; }   // least-significant zero bit of i
       if (d >= (unsigned)DIM) break;
       unsigned mask = 1u << d; signs ^= mask;
       p[d] = (signs & mask) ? hsub(R.c[d], r5[d]) : hadd(R.c[d], r5[d]);
     }
  }
  for (int j = 0; j < HC_FDIM; j++)
   { double a = hmul(weight1, val0[j]);
     a = hadd(a, hmul(weight2, sum2[j])); a = hadd(a, hmul(weight3, sum3[j]));
     a = hadd(a, hmul(weight4, sum4[j])); a = hadd(a, hmul(weight5, sum5[j]));
     double result = hmul(R.vol, a);
     double b = hmul(weightE1, val0[j]);
     b = hadd(b, hmul(weightE2, sum2[j])); b = hadd(b, hmul(weightE3, sum3[j])); b = hadd(b, hmul(weightE4, sum4[j]));
     double res5th = hmul(R.vol, b);
     R.val[j] = result; R.err[j] = fabs(hsub(res5th, result));
   }
  double maxdiff = 0; int dimDiffMax = 0;
  for (int i = 0; i < DIM; i++) if (diff[i] > maxdiff) { maxdiff = diff[i]; dimDiffMax = i; }
  R.splitDim = dimDiffMax;
}

// ---- 15-point Gauss-Kronrod rule on one 1-D region ----
template <typename F>
SCUFF_HD void gk15_eval(F &f, HCRegion<1> &R)
{ const double xgk[8] = { 0.991455371120812639206854697526329, 0.949107912342758524526189684047851,
                          0.864864423359769072789712788640926, 0.741531185599394439863864773280788,
                          0.586087235467691130294144838258730, 0.405845151377397166906606412076961,
                          0.207784955007898467600689403773245, 0.000000000000000000000000000000000 };
  const double wg[4] = { 0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
                         0.381830050505118944950369775488975, 0.417959183673469387755102040816327 };
  const double wgk[8] = { 0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
                          0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
                          0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
                          0.204432940075298892414161999234649, 0.209482141084727828012999174891714 };
  const double center = R.c[0], halfwidth = R.hw[0];
  double vals[15][HC_FDIM];
  double pt[1];
  int np = 0;
  pt[0] = center; f(pt, vals[np++]);
  for (int j = 0; j < 3; j++)
   { double w = hmul(halfwidth, xgk[2 * j + 1]);
     pt[0] = hsub(center, w); f(pt, vals[np++]); pt[0] = hadd(center, w); f(pt, vals[np++]); }
  for (int j = 0; j < 4; j++)
   { double w = hmul(halfwidth, xgk[2 * j]);
     pt[0] = hsub(center, w); f(pt, vals[np++]); pt[0] = hadd(center, w); f(pt, vals[np++]); }
  R.splitDim = 0;
  for (int k = 0; k < HC_FDIM; k++)
   { double result_gauss = hmul(vals[0][k], wg[3]);
     double result_kronrod = hmul(vals[0][k], wgk[7]);
     double result_abs = fabs(result_kronrod);
     int n = 1;
     for (int j = 0; j < 3; j++)
      { int j2 = 2 * j + 1;
        double v = hadd(vals[n][k], vals[n + 1][k]);
        result_gauss = hadd(result_gauss, hmul(wg[j], v));
        result_kronrod = hadd(result_kronrod, hmul(wgk[j2], v));
        result_abs = hadd(result_abs, hmul(wgk[j2], hadd(fabs(vals[n][k]), fabs(vals[n + 1][k]))));
        n += 2; }
     for (int j = 0; j < 4; j++)
      { int j2 = 2 * j;
        result_kronrod = hadd(result_kronrod, hmul(wgk[j2], hadd(vals[n][k], vals[n + 1][k])));
        result_abs = hadd(result_abs, hmul(wgk[j2], hadd(fabs(vals[n][k]), fabs(vals[n + 1][k]))));
        n += 2; }
     R.val[k] = hmul(result_kronrod, halfwidth);
     double mean = hmul(result_kronrod, 0.5);
     double result_asc = hmul(wgk[7], fabs(hsub(vals[0][k], mean)));
     n = 1;
     for (int j = 0; j < 3; j++)
      { int j2 = 2 * j + 1;
        result_asc = hadd(result_asc, hmul(wgk[j2], hadd(fabs(hsub(vals[n][k], mean)), fabs(hsub(vals[n + 1][k], mean)))));
        n += 2; }
     for (int j = 0; j < 4; j++)
      { int j2 = 2 * j;
        result_asc = hadd(result_asc, hmul(wgk[j2], hadd(fabs(hsub(vals[n][k], mean)), fabs(hsub(vals[n + 1][k], mean)))));
        n += 2; }
     double err = hmul(fabs(hsub(result_kronrod, result_gauss)), halfwidth);
     result_abs = hmul(result_abs, halfwidth);
     result_asc = hmul(result_asc, halfwidth);
     if (result_asc != 0 && err != 0)
      { double scale = pow(hmul(200, err) / result_asc, 1.5);
        err = (scale < 1) ? hmul(result_asc, scale) : result_asc; }
     if (result_abs > 2.2250738585072014e-308 / (50 * 2.2204460492503131e-16))
      { double min_err = hmul(50 * 2.2204460492503131e-16, result_abs);
        if (min_err > err) err = min_err; }
     R.err[k] = err;
   }
}

template <int DIM, typename F> SCUFF_HD void rule_eval(F &f, HCRegion<DIM> &R);
template <int DIM> struct RuleDispatch
{ template <typename F> static SCUFF_HD void eval(F &f, HCRegion<DIM> &R) { gm_eval<DIM>(f, R); } };
template <> struct RuleDispatch<1>
{ template <typename F> static SCUFF_HD void eval(F &f, HCRegion<1> &R) { gk15_eval(f, R); } };

template <int DIM> SCUFF_HD int rule_points() { return DIM == 1 ? 15 : (1 + 4 * DIM + 2 * DIM * (DIM - 1) + (1 << DIM)); }

// Adaptive integration over [0,1]^DIM.  pool: capacity regions; heap: capacity ints.
// Returns 0 on success, 2 if the pool capacity was exceeded (caller retries with a larger pool).
template <int DIM, typename F>
SCUFF_HD int hcubature_run(F &f, size_t maxEval, double reqAbsError, double reqRelError,
                           HCRegion<DIM> *pool, int *heap, int capacity, double *val, double *err)
{ const int npts = rule_points<DIM>();
  double totV[HC_FDIM], totE[HC_FDIM];
  for (int j = 0; j < HC_FDIM; j++) { totV[j] = 0; totE[j] = 0; }
  int nheap = 0, nregions = 0;
  size_t numEval = 0;

  auto push = [&](int ri)
  { HCRegion<DIM> &R = pool[ri];
    double em = 0; for (int j = 0; j < HC_FDIM; j++) if (R.err[j] > em) em = R.err[j];
    R.errmax = em;
    for (int j = 0; j < HC_FDIM; j++) { totV[j] = hadd(totV[j], R.val[j]); totE[j] = hadd(totE[j], R.err[j]); }
    int insert = nheap++;
    while (insert)
     { int parent = (insert - 1) / 2;
       if (em <= pool[heap[parent]].errmax) break;
       heap[insert] = heap[parent]; insert = parent; }
    heap[insert] = ri;
  };
  auto pop = [&]() -> int
  { int ret = heap[0];
    int n = --nheap; heap[0] = heap[n];
    int i = 0, child;
    while ((child = i * 2 + 1) < n)
     { int largest;
       if (pool[heap[child]].errmax <= pool[heap[i]].errmax) largest = i; else largest = child;
       if (++child < n && pool[heap[largest]].errmax < pool[heap[child]].errmax) largest = child;
       if (largest == i) break;
       int sw = heap[i]; heap[i] = heap[largest]; heap[i = largest] = sw; }
    HCRegion<DIM> &R = pool[ret];
    for (int j = 0; j < HC_FDIM; j++) { totV[j] = hsub(totV[j], R.val[j]); totE[j] = hsub(totE[j], R.err[j]); }
    return ret;
  };

  { HCRegion<DIM> &R = pool[nregions++];
    double vol = 1;
    for (int i = 0; i < DIM; i++) { R.c[i] = 0.5 * (0.0 + 1.0); R.hw[i] = 0.5 * (1.0 - 0.0); vol = hmul(vol, hmul(2, R.hw[i])); }
    R.vol = vol; R.splitDim = 0;
    RuleDispatch<DIM>::eval(f, R);
    push(0);
    numEval += npts;
  }
  while (numEval < maxEval || !maxEval)
   { bool conv = true;
     for (int j = 0; j < HC_FDIM; j++) if (totE[j] > reqAbsError && totE[j] > fabs(totV[j]) * reqRelError) { conv = false; break; }
     if (conv) break;
     if (nregions >= capacity) return 2;
     int r0 = pop();
     int r1 = nregions++;
     HCRegion<DIM> &R = pool[r0], &R2 = pool[r1];
     const int d = R.splitDim;
     R.hw[d] = hmul(R.hw[d], 0.5); R.vol = hmul(R.vol, 0.5);
     for (int i = 0; i < DIM; i++) { R2.c[i] = R.c[i]; R2.hw[i] = R.hw[i]; }
     R2.vol = R.vol; R2.splitDim = R.splitDim;
     R.c[d] = hsub(R.c[d], R.hw[d]); R2.c[d] = hadd(R2.c[d], R.hw[d]);
     RuleDispatch<DIM>::eval(f, R);
     RuleDispatch<DIM>::eval(f, R2);
     push(r0); push(r1);
     numEval += 2 * npts;
   }
  for (int j = 0; j < HC_FDIM; j++) { val[j] = 0; err[j] = 0; }
  for (int i = 0; i < nheap; i++)
   { const HCRegion<DIM> &R = pool[heap[i]];
     for (int j = 0; j < HC_FDIM; j++) { val[j] = hadd(val[j], R.val[j]); err[j] = hadd(err[j], R.err[j]); } }
  return 0;
}

}  // namespace scuffcuda
