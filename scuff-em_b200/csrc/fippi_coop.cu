// Warp-cooperative FIPPI builder for the no-common-vertex case (3-D hcubature of CFDIntegrand3D).
//
// Same decisions and the same bits as the one-thread-per-record builder (fippi_build.cuh / hcubature.cuh), organised for
// latency instead: one WARP per record.  The 33 points of the Genz-Malik rule are evaluated by the lanes in parallel (the
// coordinates of point q are formed with the very operations the serial code uses), the function values go through shared
// memory, and lane j then forms every sum of component j in the serial order — so the per-component results, the error
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

struct FCRegion { double val[HC_FDIM], err[HC_FDIM], c[3], hw[3], vol; int splitDim, pad; };
static_assert(sizeof(FCRegion) <= FIPPI_REGION_BYTES, "region size");

// q-th point of the 3-D Genz-Malik rule in the order gm_eval<3> visits them
__device__ __forceinline__ void fc_point(int q, const double *c, const double *hw, double *p)
{ const double lambda2 = 0.3585685828003180919906451539079374954541;
  const double lambda4 = 0.9486832980505137995996680633298155601160;
  const double lambda5 = 0.6882472016116852977216287342936235251269;
  p[0] = c[0]; p[1] = c[1]; p[2] = c[2];
  if (q == 0) return;
  if (q <= 12)
   { const int k = (q - 1) >> 2, t = (q - 1) & 3;
     const double r = hmul(hw[k], t < 2 ? lambda2 : lambda4);
     p[k] = (t & 1) ? hadd(c[k], r) : hsub(c[k], r);
     return; }
  if (q <= 24)
   { const int pr = (q - 13) >> 2, t = (q - 13) & 3;          // pairs (0,1),(0,2),(1,2); signs (-,-),(+,-),(+,+),(-,+)
     const int i = pr == 2 ? 1 : 0, jj = pr == 0 ? 1 : 2;
     const double ri = hmul(hw[i], lambda4), rj = hmul(hw[jj], lambda4);
     p[i] = (t == 1 || t == 2) ? hadd(c[i], ri) : hsub(c[i], ri);
     p[jj] = (t >= 2) ? hadd(c[jj], rj) : hsub(c[jj], rj);
     return; }
  { const int gray[8] = { 0, 1, 3, 2, 6, 7, 5, 4 };             // bit set = minus, Gray-code order of hcubature.c
    const int s = gray[q - 25];
    for (int d = 0; d < 3; d++)
     { const double r = hmul(hw[d], lambda5); p[d] = (s >> d) & 1 ? hsub(c[d], r) : hadd(c[d], r); } }
}

// component j of the rule from the shared function values (the serial order of gm_eval<3>)
__device__ __forceinline__ void fc_component(const double *F, double *T, int j, double vol, double &val, double &err)
{ const int DIM = 3;
  const double lambda2 = 0.3585685828003180919906451539079374954541, lambda4 = 0.9486832980505137995996680633298155601160;
  const double weight2 = 980. / 6561., weight4 = 200. / 19683., weightE2 = 245. / 486., weightE4 = 25. / 729.;
  const double ratio = (lambda2 * lambda2) / (lambda4 * lambda4);
  const double weight1 = ((double)(12824 - 9120 * DIM + 400 * DIM * DIM)) / 19683.0;
  const double weight3 = ((double)(1820 - 400 * DIM)) / 19683.0;
  const double weight5 = 6859.0 / 19683.0 / ((double)(1U << DIM));
  const double weightE1 = ((double)(729 - 950 * DIM + 50 * DIM * DIM)) / 729.0;
  const double weightE3 = ((double)(265 - 100 * DIM)) / 1458.0;
  const double val0 = F[j];
  double sum2 = 0, sum3 = 0, sum4 = 0, sum5 = 0;
  for (int k = 0; k < 3; k++)
   { const double v0 = F[(1 + 4 * k) * HC_FDIM + j], v1 = F[(2 + 4 * k) * HC_FDIM + j], v2 = F[(3 + 4 * k) * HC_FDIM + j], v3 = F[(4 + 4 * k) * HC_FDIM + j];
     const double s01 = hadd(v0, v1), s23 = hadd(v2, v3);
     sum2 = hadd(sum2, s01); sum3 = hadd(sum3, s23);
     const double tv = hmul(2, val0);
     const double t = hsub(hsub(s01, tv), hmul(ratio, hsub(s23, tv)));
     T[k * HC_FDIM + j] = fabs(t); }
  for (int q = 13; q <= 24; q++) sum4 = hadd(sum4, F[q * HC_FDIM + j]);
  for (int q = 25; q <= 32; q++) sum5 = hadd(sum5, F[q * HC_FDIM + j]);
  double a = hmul(weight1, val0);
  a = hadd(a, hmul(weight2, sum2)); a = hadd(a, hmul(weight3, sum3));
  a = hadd(a, hmul(weight4, sum4)); a = hadd(a, hmul(weight5, sum5));
  const double result = hmul(vol, a);
  double b = hmul(weightE1, val0);
  b = hadd(b, hmul(weightE2, sum2)); b = hadd(b, hmul(weightE3, sum3)); b = hadd(b, hmul(weightE4, sum4));
  const double res5th = hmul(vol, b);
  val = result; err = fabs(hsub(res5th, result));
}

template <int FC_WARPS>
__global__ void __launch_bounds__(32 * FC_WARPS)
fippi_coop_kernel(const double *verts, const int *items, int nItems, char *pool, int FC_CAP, double *out, int *status)
{ extern __shared__ __align__(16) unsigned char fsmRaw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * FC_WARPS + warp;
  if (t >= nItems) return;                              // whole warps leave together
  double *F = (double *)(fsmRaw + (size_t)warp * fc_warp_bytes(FC_CAP)), *T = F + FC_F_DBL, *diffS = T + 3 * HC_FDIM, *hErr = diffS + 4;
  short *hIdx = (short *)(hErr + FC_CAP);
  const int it = items[t];
  const double *V = verts + 18 * (size_t)it;
  const double *OVa[3] = { V, V + 3, V + 6 }, *OVb[3] = { V + 9, V + 12, V + 15 };
  CFD3D f; f.setup(OVa, OVb);
  FCRegion *regs = (FCRegion *)(pool + (size_t)t * ((size_t)FC_CAP * FIPPI_REGION_BYTES));
  double totV[2] = { 0, 0 }, totE[2] = { 0, 0 };        // component `lane` (slot 0) and, on lane 0, component 32 (slot 1)
  int nheap = 0, nregions = 0; size_t numEval = 0;
  const size_t maxEval = 1000000; const double reqAbs = 1.0e-10, reqRel = 1.0e-6;

  // evaluates the rule on region ri (geometry already stored), writes val/err/splitDim, returns errmax
  auto eval_region = [&](int ri) -> double
  { FCRegion &R = regs[ri];
    const double c[3] = { R.c[0], R.c[1], R.c[2] }, hw[3] = { R.hw[0], R.hw[1], R.hw[2] }; const double vol = R.vol;
    for (int q = lane; q < 33; q += 32)
     { double p[3], fv[HC_FDIM]; fc_point(q, c, hw, p); f(p, fv);
       for (int j = 0; j < HC_FDIM; j++) F[q * HC_FDIM + j] = fv[j]; }
    __syncwarp();
    double v, e, em;
    fc_component(F, T, lane, vol, v, e);
    R.val[lane] = v; R.err[lane] = e; em = e;
    if (lane == 0) { fc_component(F, T, 32, vol, v, e); R.val[32] = v; R.err[32] = e; if (e > em) em = e; }
    __syncwarp();
    if (lane < 3) { double dk = 0; for (int j = 0; j < HC_FDIM; j++) dk = hadd(dk, T[lane * HC_FDIM + j]); diffS[lane] = dk; }
    for (int o = 16; o > 0; o >>= 1) { const double x = __shfl_xor_sync(0xffffffffu, em, o); if (x > em) em = x; }
    __syncwarp();
    if (lane == 0)
     { double maxdiff = 0; int dm = 0; for (int i = 0; i < 3; i++) if (diffS[i] > maxdiff) { maxdiff = diffS[i]; dm = i; }
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
     for (int i = 0; i < 3; i++) { R.c[i] = 0.5 * (0.0 + 1.0); R.hw[i] = 0.5 * (1.0 - 0.0); vol = hmul(vol, hmul(2, R.hw[i])); }
     R.vol = vol; R.splitDim = 0; }
  __syncwarp();
  nregions = 1;
  { const double em = eval_region(0); push(0, em); numEval += 33; }
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
        for (int i = 0; i < 3; i++) { R2.c[i] = R.c[i]; R2.hw[i] = R.hw[i]; }
        R2.vol = R.vol; R2.splitDim = R.splitDim;
        R.c[d] = hsub(R.c[d], R.hw[d]); R2.c[d] = hadd(R2.c[d], R.hw[d]); }
     __syncwarp();
     const double em0 = eval_region(r0);
     const double em1 = eval_region(r1);
     push(r0, em0); push(r1, em1);
     numEval += 66;
   }
  if (lane == 0) status[t] = st;
  if (st) return;
  // sum over the regions in heap order, component by component
  double *o = out + FIPPI_REC_DOUBLES * (size_t)it;
  { double v = 0; for (int i = 0; i < nheap; i++) v = hadd(v, regs[hIdx[i]].val[lane]); o[lane] = v;
    if (lane == 0) { double w = 0; for (int i = 0; i < nheap; i++) w = hadd(w, regs[hIdx[i]].val[32]); o[32] = w; } }
  if (lane < 3) o[33 + lane] = (OVa[0][lane] + OVa[1][lane] + OVa[2][lane]) / 3.0;
}

// launcher used by fippi_kernel.cu; pool must hold n * cap * FIPPI_REGION_BYTES bytes.  Small pools: 4 records per CTA;
// large pools (heap of up to 15 200 entries = 150 KB of shared memory): one record per CTA.  Returns false if the heap
// does not fit into shared memory (the caller then uses the one-thread-per-record kernel).
bool fippi_launch_coop(cudaStream_t s, const double *verts, const int *items, int n, char *pool, int cap, double *out, int *status)
{ if (cap > 32000) return false;
  if (fc_warp_bytes(cap) * 4 <= 100 * 1024)
   { const size_t smem = fc_warp_bytes(cap) * 4;
     cudaFuncSetAttribute(fippi_coop_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
     fippi_coop_kernel<4><<<(n + 3) / 4, 128, smem, s>>>(verts, items, n, pool, cap, out, status); }
  else
   { const size_t smem = fc_warp_bytes(cap);
     if (smem > 220 * 1024) return false;
     cudaFuncSetAttribute(fippi_coop_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
     fippi_coop_kernel<1><<<n, 32, smem, s>>>(verts, items, n, pool, cap, out, status); }
  return true;
}

}  // namespace scuffcuda
