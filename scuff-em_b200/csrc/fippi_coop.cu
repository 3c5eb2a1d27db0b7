// Warp-cooperative FIPPI builder: 3-D hcubature of CFDIntegrand3D (no common vertex) and the 1-D / 2-D hcubature of the
// Taylor-Duffy reduced integrand (common edge / common vertex).
//
// Same decisions and the same bits as the one-thread-per-record builder (fippi_build.cuh / hcubature.cuh), organised for
// latency instead: one WARP per record.  The 33 points of the Genz-Malik rule are evaluated by the lanes in parallel (the
// coordinates of point q are formed with the very operations the serial code uses), the function values go through shared
// memory, and lane j then forms every sum of component j in the serial order (Genz-Malik in 2-D/3-D, 15-point Gauss-Kronrod in 1-D) — so the per-component results, the error
// estimates, the fourth differences (summed over the components in order by one lane), the split dimension, the heap order
// and the convergence test are bit-identical to the serial replica of the reference's hcubature.  The max-error heap lives in
// shared memory.  Records that outgrow this pass's region pool are redone by the serial passes with larger pools.
// Compiled with -fmad=false like the other FIPPI units.
#include "fippi_build.cuh"

namespace scuffcuda {

#define FC_F_DBL (33 * HC_FDIM)          /* function values of one region */
#define FC_FIXED_DBL (FC_F_DBL + 3 * HC_FDIM + 4)   /* + |t| table + diff[3] */
// shared memory per warp for a region pool of `cap` regions: fixed part + heap errmax (double) + heap region index (short)
__host__ __device__ inline size_t fc_warp_bytes(int cap) { return (((size_t)(FC_FIXED_DBL + cap) * 8 + (size_t)cap * 2) + 15) / 16 * 16; }

template <int DIM> struct FCRegionT { double val[HC_FDIM], err[HC_FDIM], c[DIM], hw[DIM], vol; int splitDim, pad; };
static_assert(sizeof(FCRegionT<3>) <= FIPPI_REGION_BYTES, "region size");
template <int DIM> __host__ __device__ constexpr int fc_npts() { return DIM == 1 ? 15 : (1 + 4 * DIM + 2 * DIM * (DIM - 1) + (1 << DIM)); }

// q-th point of the rule in the order the serial code (hcubature.cuh gm_eval / gk15_eval) visits them
template <int DIM>
__device__ __forceinline__ void fc_point(int q, const double *c, const double *hw, double *p)
{ if (DIM == 1)
   { const double xgk[8] = { 0.991455371120812639206854697526329, 0.949107912342758524526189684047851,
                             0.864864423359769072789712788640926, 0.741531185599394439863864773280788,
                             0.586087235467691130294144838258730, 0.405845151377397166906606412076961,
                             0.207784955007898467600689403773245, 0.000000000000000000000000000000000 };
     if (q == 0) { p[0] = c[0]; return; }
     const bool first = q <= 6;
     const int j = first ? ((q - 1) >> 1) : ((q - 7) >> 1), plus = first ? ((q - 1) & 1) : ((q - 7) & 1);
     const double w = hmul(hw[0], xgk[first ? 2 * j + 1 : 2 * j]);
     p[0] = plus ? hadd(c[0], w) : hsub(c[0], w);
     return; }
  const double lambda2 = 0.3585685828003180919906451539079374954541;
  const double lambda4 = 0.9486832980505137995996680633298155601160;
  const double lambda5 = 0.6882472016116852977216287342936235251269;
  for (int d = 0; d < DIM; d++) p[d] = c[d];
  if (q == 0) return;
  if (q <= 4 * DIM)
   { const int k = (q - 1) >> 2, t = (q - 1) & 3;
     const double r = hmul(hw[k], t < 2 ? lambda2 : lambda4);
     p[k] = (t & 1) ? hadd(c[k], r) : hsub(c[k], r);
     return; }
  const int n4 = 2 * DIM * (DIM - 1);
  if (q <= 4 * DIM + n4)
   { const int pr = (q - 4 * DIM - 1) >> 2, t = (q - 4 * DIM - 1) & 3;   // pairs (0,1),(0,2),(1,2); signs (-,-),(+,-),(+,+),(-,+)
     const int i = pr == 2 ? 1 : 0, jj = pr == 0 ? 1 : 2;
     const double ri = hmul(hw[i], lambda4), rj = hmul(hw[jj], lambda4);
     p[i] = (t == 1 || t == 2) ? hadd(c[i], ri) : hsub(c[i], ri);
     p[jj] = (t >= 2) ? hadd(c[jj], rj) : hsub(c[jj], rj);
     return; }
  { const int gray[8] = { 0, 1, 3, 2, 6, 7, 5, 4 };             // bit set = minus, Gray-code order of hcubature.c
    const int s = gray[q - 4 * DIM - n4 - 1];
    for (int d = 0; d < DIM; d++)
     { const double r = hmul(hw[d], lambda5); p[d] = (s >> d) & 1 ? hsub(c[d], r) : hadd(c[d], r); } }
}

// component j of the Genz-Malik rule from the shared function values (the serial order of gm_eval<DIM>)
template <int DIM>
__device__ __forceinline__ void fc_component(const double *F, double *T, int j, double vol, const double *hwUnused, double &val, double &err)
{ const double lambda2 = 0.3585685828003180919906451539079374954541, lambda4 = 0.9486832980505137995996680633298155601160;
  const double weight2 = 980. / 6561., weight4 = 200. / 19683., weightE2 = 245. / 486., weightE4 = 25. / 729.;
  const double ratio = (lambda2 * lambda2) / (lambda4 * lambda4);
  const double weight1 = ((double)(12824 - 9120 * DIM + 400 * DIM * DIM)) / 19683.0;
  const double weight3 = ((double)(1820 - 400 * DIM)) / 19683.0;
  const double weight5 = 6859.0 / 19683.0 / ((double)(1U << DIM));
  const double weightE1 = ((double)(729 - 950 * DIM + 50 * DIM * DIM)) / 729.0;
  const double weightE3 = ((double)(265 - 100 * DIM)) / 1458.0;
  const int n4 = 2 * DIM * (DIM - 1), n5 = 1 << DIM;
  const double val0 = F[j];
  double sum2 = 0, sum3 = 0, sum4 = 0, sum5 = 0;
  for (int k = 0; k < DIM; k++)
   { const double v0 = F[(1 + 4 * k) * HC_FDIM + j], v1 = F[(2 + 4 * k) * HC_FDIM + j], v2 = F[(3 + 4 * k) * HC_FDIM + j], v3 = F[(4 + 4 * k) * HC_FDIM + j];
     const double s01 = hadd(v0, v1), s23 = hadd(v2, v3);
     sum2 = hadd(sum2, s01); sum3 = hadd(sum3, s23);
     const double tv = hmul(2, val0);
     const double t = hsub(hsub(s01, tv), hmul(ratio, hsub(s23, tv)));
     T[k * HC_FDIM + j] = fabs(t); }
  for (int q = 4 * DIM + 1; q <= 4 * DIM + n4; q++) sum4 = hadd(sum4, F[q * HC_FDIM + j]);
  for (int q = 4 * DIM + n4 + 1; q <= 4 * DIM + n4 + n5; q++) sum5 = hadd(sum5, F[q * HC_FDIM + j]);
  double a = hmul(weight1, val0);
  a = hadd(a, hmul(weight2, sum2)); a = hadd(a, hmul(weight3, sum3));
  a = hadd(a, hmul(weight4, sum4)); a = hadd(a, hmul(weight5, sum5));
  const double result = hmul(vol, a);
  double b = hmul(weightE1, val0);
  b = hadd(b, hmul(weightE2, sum2)); b = hadd(b, hmul(weightE3, sum3)); b = hadd(b, hmul(weightE4, sum4));
  const double res5th = hmul(vol, b);
  val = result; err = fabs(hsub(res5th, result));
}

// component k of the 15-point Gauss-Kronrod rule (the serial order of gk15_eval)
template <>
__device__ __forceinline__ void fc_component<1>(const double *F, double *T, int k, double vol, const double *hw, double &val, double &errOut)
{ const double wg[4] = { 0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
                         0.381830050505118944950369775488975, 0.417959183673469387755102040816327 };
  const double wgk[8] = { 0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
                          0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
                          0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
                          0.204432940075298892414161999234649, 0.209482141084727828012999174891714 };
  const double halfwidth = hw[0];
#define FV(n) F[(n) * HC_FDIM + k]
  double result_gauss = hmul(FV(0), wg[3]);
  double result_kronrod = hmul(FV(0), wgk[7]);
  double result_abs = fabs(result_kronrod);
  int n = 1;
  for (int j = 0; j < 3; j++)
   { const int j2 = 2 * j + 1;
     const double v = hadd(FV(n), FV(n + 1));
     result_gauss = hadd(result_gauss, hmul(wg[j], v));
     result_kronrod = hadd(result_kronrod, hmul(wgk[j2], v));
     result_abs = hadd(result_abs, hmul(wgk[j2], hadd(fabs(FV(n)), fabs(FV(n + 1)))));
     n += 2; }
  for (int j = 0; j < 4; j++)
   { const int j2 = 2 * j;
     result_kronrod = hadd(result_kronrod, hmul(wgk[j2], hadd(FV(n), FV(n + 1))));
     result_abs = hadd(result_abs, hmul(wgk[j2], hadd(fabs(FV(n)), fabs(FV(n + 1)))));
     n += 2; }
  val = hmul(result_kronrod, halfwidth);
  const double mean = hmul(result_kronrod, 0.5);
  double result_asc = hmul(wgk[7], fabs(hsub(FV(0), mean)));
  n = 1;
  for (int j = 0; j < 3; j++)
   { const int j2 = 2 * j + 1;
     result_asc = hadd(result_asc, hmul(wgk[j2], hadd(fabs(hsub(FV(n), mean)), fabs(hsub(FV(n + 1), mean)))));
     n += 2; }
  for (int j = 0; j < 4; j++)
   { const int j2 = 2 * j;
     result_asc = hadd(result_asc, hmul(wgk[j2], hadd(fabs(hsub(FV(n), mean)), fabs(hsub(FV(n + 1), mean)))));
     n += 2; }
#undef FV
  double err = hmul(fabs(hsub(result_kronrod, result_gauss)), halfwidth);
  result_abs = hmul(result_abs, halfwidth);
  result_asc = hmul(result_asc, halfwidth);
  if (result_asc != 0 && err != 0)
   { const double scale = pow(hmul(200, err) / result_asc, 1.5);
     err = (scale < 1) ? hmul(result_asc, scale) : result_asc; }
  if (result_abs > 2.2250738585072014e-308 / (50 * 2.2204460492503131e-16))
   { const double min_err = hmul(50 * 2.2204460492503131e-16, result_abs);
     if (min_err > err) err = min_err; }
  errOut = err;
  (void)T; (void)vol;
}

// integrand adapters: setup from the 18 vertex doubles of a record
template <int DIM> struct FCFunctor;
template <> struct FCFunctor<3>
{ CFD3D f;
  __device__ bool setup(const FIPPITables &, const double *V) { const double *a[3] = { V, V + 3, V + 6 }, *b[3] = { V + 9, V + 12, V + 15 }; f.setup(a, b); return true; }
};
template <> struct FCFunctor<1>
{ TDFIPPI f;
  __device__ bool setup(const FIPPITables &T, const double *V) { f.setup(&T, V, V + 3, V + 6, V + 12, V + 15); return f.commonEdge != 0; }
};
template <> struct FCFunctor<2>
{ TDFIPPI f;
  __device__ bool setup(const FIPPITables &T, const double *V) { f.setup(&T, V, V + 3, V + 6, V + 12, V + 15); return f.commonEdge == 0; }
};

template <int FC_WARPS, int DIM>
__global__ void __launch_bounds__(32 * FC_WARPS)
fippi_coop_kernel(const FIPPITables TB, const double *verts, const int *items, int nItems, char *pool, int FC_CAP, double *out, int *status)
{ extern __shared__ __align__(16) unsigned char fsmRaw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * FC_WARPS + warp;
  if (t >= nItems) return;                              // whole warps leave together
  double *F = (double *)(fsmRaw + (size_t)warp * fc_warp_bytes(FC_CAP)), *T = F + FC_F_DBL, *diffS = T + 3 * HC_FDIM, *hErr = diffS + 4;
  short *hIdx = (short *)(hErr + FC_CAP);
  const int it = items[t];
  const double *V = verts + 18 * (size_t)it;
  typedef FCRegionT<DIM> FCRegion;
  constexpr int NPTS = fc_npts<DIM>();
  FCFunctor<DIM> fw;
  if (!fw.setup(TB, V)) { if (lane == 0) status[t] = 3; return; }          // integration case mismatch
  const auto &f = fw.f;
  FCRegion *regs = (FCRegion *)(pool + (size_t)t * ((size_t)FC_CAP * FIPPI_REGION_BYTES));
  double totV[2] = { 0, 0 }, totE[2] = { 0, 0 };        // component `lane` (slot 0) and, on lane 0, component 32 (slot 1)
  int nheap = 0, nregions = 0; size_t numEval = 0;
  const size_t maxEval = (DIM == 3) ? 1000000 : 10000; const double reqAbs = (DIM == 3) ? 1.0e-10 : 0.0, reqRel = (DIM == 3) ? 1.0e-6 : 1.0e-5;

  // evaluates the rule on region ri (geometry already stored), writes val/err/splitDim, returns errmax
  auto eval_region = [&](int ri) -> double
  { FCRegion &R = regs[ri];
    double c[DIM], hw[DIM]; for (int i = 0; i < DIM; i++) { c[i] = R.c[i]; hw[i] = R.hw[i]; } const double vol = R.vol;
    for (int q = lane; q < NPTS; q += 32)
     { double p[DIM], fv[HC_FDIM]; fc_point<DIM>(q, c, hw, p); f(p, fv);
       for (int j = 0; j < HC_FDIM; j++) F[q * HC_FDIM + j] = fv[j]; }
    __syncwarp();
    double v, e, em;
    fc_component<DIM>(F, T, lane, vol, hw, v, e);
    R.val[lane] = v; R.err[lane] = e; em = e;
    if (lane == 0) { fc_component<DIM>(F, T, 32, vol, hw, v, e); R.val[32] = v; R.err[32] = e; if (e > em) em = e; }
    __syncwarp();
    if (DIM > 1 && lane < DIM) { double dk = 0; for (int j = 0; j < HC_FDIM; j++) dk = hadd(dk, T[lane * HC_FDIM + j]); diffS[lane] = dk; }
    for (int o = 16; o > 0; o >>= 1) { const double x = __shfl_xor_sync(0xffffffffu, em, o); if (x > em) em = x; }
    __syncwarp();
    if (lane == 0)
     { double maxdiff = 0; int dm = 0;
       if (DIM > 1) for (int i = 0; i < DIM; i++) if (diffS[i] > maxdiff) { maxdiff = diffS[i]; dm = i; }
       R.splitDim = dm; }
    __syncwarp();
    return em;
  };
  auto push = [&](int ri, double em)
  { const FCRegion &R = regs[ri];
    totV[0] = hadd(totV[0], R.val[lane]); totE[0] = hadd(totE[0], R.err[lane]);
    if (lane == 0)
     { totV[1] = hadd(totV[1], R.val[32]); totE[1] = hadd(totE[1], R.err[32]);
       int insert = nheap;
       while (insert)
        { const int parent = (insert - 1) / 2;
          if (em <= hErr[parent]) break;
          hErr[insert] = hErr[parent]; hIdx[insert] = hIdx[parent]; insert = parent; }
       hErr[insert] = em; hIdx[insert] = (short)ri; }
    nheap++;
    __syncwarp();
  };
  auto pop = [&]() -> int
  { int ret = 0;
    if (lane == 0)
     { ret = hIdx[0];
       const int n = nheap - 1; hErr[0] = hErr[n]; hIdx[0] = hIdx[n];
       int i = 0, child;
       while ((child = i * 2 + 1) < n)
        { int largest;
          if (hErr[child] <= hErr[i]) largest = i; else largest = child;
          if (++child < n && hErr[largest] < hErr[child]) largest = child;
          if (largest == i) break;
          const double se = hErr[i]; hErr[i] = hErr[largest]; hErr[largest] = se;
          const short si = hIdx[i]; hIdx[i] = hIdx[largest]; hIdx[largest] = si; i = largest; } }
    nheap--;
    ret = __shfl_sync(0xffffffffu, ret, 0);
    const FCRegion &R = regs[ret];
    totV[0] = hsub(totV[0], R.val[lane]); totE[0] = hsub(totE[0], R.err[lane]);
    if (lane == 0) { totV[1] = hsub(totV[1], R.val[32]); totE[1] = hsub(totE[1], R.err[32]); }
    return ret;
  };

  if (lane == 0)
   { FCRegion &R = regs[0]; double vol = 1;
     for (int i = 0; i < DIM; i++) { R.c[i] = 0.5 * (0.0 + 1.0); R.hw[i] = 0.5 * (1.0 - 0.0); vol = hmul(vol, hmul(2, R.hw[i])); }
     R.vol = vol; R.splitDim = 0; }
  __syncwarp();
  nregions = 1;
  { const double em = eval_region(0); push(0, em); numEval += NPTS; }
  int st = 0;
  while (numEval < maxEval)
   { bool notconv = (totE[0] > reqAbs && totE[0] > fabs(totV[0]) * reqRel);
     if (lane == 0 && totE[1] > reqAbs && totE[1] > fabs(totV[1]) * reqRel) notconv = true;
     if (!__any_sync(0xffffffffu, notconv)) break;
     if (nregions >= FC_CAP) { st = 2; break; }
     const int r0 = pop();
     const int r1 = nregions++;
     if (lane == 0)
      { FCRegion &R = regs[r0], &R2 = regs[r1];
        const int d = R.splitDim;
        R.hw[d] = hmul(R.hw[d], 0.5); R.vol = hmul(R.vol, 0.5);
        for (int i = 0; i < DIM; i++) { R2.c[i] = R.c[i]; R2.hw[i] = R.hw[i]; }
        R2.vol = R.vol; R2.splitDim = R.splitDim;
        R.c[d] = hsub(R.c[d], R.hw[d]); R2.c[d] = hadd(R2.c[d], R.hw[d]); }
     __syncwarp();
     const double em0 = eval_region(r0);
     const double em1 = eval_region(r1);
     push(r0, em0); push(r1, em1);
     numEval += 2 * NPTS;
   }
  if (lane == 0) status[t] = st;
  if (st) return;
  // sum over the regions in heap order, component by component
  double *o = out + FIPPI_REC_DOUBLES * (size_t)it;
  { double v = 0; for (int i = 0; i < nheap; i++) v = hadd(v, regs[hIdx[i]].val[lane]); o[lane] = v;
    if (lane == 0) { double w = 0; for (int i = 0; i < nheap; i++) w = hadd(w, regs[hIdx[i]].val[32]); o[32] = w; } }
  if (lane < 3) o[33 + lane] = (V[lane] + V[3 + lane] + V[6 + lane]) / 3.0;
}

// launcher used by fippi_kernel.cu; pool must hold n * cap * FIPPI_REGION_BYTES bytes.  Small pools: 4 records per CTA;
// large pools (heap of up to 15 200 entries = 150 KB of shared memory): one record per CTA.  Returns false if the heap
// does not fit into shared memory (the caller then uses the one-thread-per-record kernel).
// c: integration case (0: 3-D no common vertex, 1: 1-D common edge, 2: 2-D common vertex)
template <int DIM>
static bool fc_launch(cudaStream_t s, const FIPPITables &T, const double *verts, const int *items, int n, char *pool, int cap, double *out, int *status)
{ if (cap > 32000) return false;
  if (fc_warp_bytes(cap) * 4 <= 100 * 1024)
   { const size_t smem = fc_warp_bytes(cap) * 4;
     cudaFuncSetAttribute(fippi_coop_kernel<4, DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
     fippi_coop_kernel<4, DIM><<<(n + 3) / 4, 128, smem, s>>>(T, verts, items, n, pool, cap, out, status); }
  else
   { const size_t smem = fc_warp_bytes(cap);
     if (smem > 220 * 1024) return false;
     cudaFuncSetAttribute(fippi_coop_kernel<1, DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
     fippi_coop_kernel<1, DIM><<<n, 32, smem, s>>>(T, verts, items, n, pool, cap, out, status); }
  return true;
}
bool fippi_launch_coop(cudaStream_t s, int c, const FIPPITables &T, const double *verts, const int *items, int n, char *pool, int cap, double *out, int *status)
{ if (c == 0) return fc_launch<3>(s, T, verts, items, n, pool, cap, out, status);
  if (c == 1) return fc_launch<1>(s, T, verts, items, n, pool, cap, out, status);
  return fc_launch<2>(s, T, verts, items, n, pool, cap, out, status);
}

}  // namespace scuffcuda
