// Taylor-Duffy kernel: panel pairs that share all three vertices (ncv == 3; the reference uses
// Taylor-Duffy for them at every frequency, PanelPanelInteractions.cc:449-491).  One thread per
// (pair, Q_a, Q_b) combination runs the 1-D p-adaptive Clenshaw-Curtis cubature of td_math.cuh with the
// reference's tolerances (RelTol 1e-10, MaxEval 1000), its function-value cache living in a strided
// HBM scratch block.
#include <vector>
#include <map>
#include <mutex>
#include <cstdlib>
#include "kernels.h"
#include "stamp.cuh"
#include "td_math.cuh"
#include "sw_kernels.h"
#include "generated_tables.h"

namespace scuffcuda {

static_assert(sizeof(SCUpsEntry) == sizeof(UpsEntry), "generated table layout");

struct TDDeviceTables { double *ccx = nullptr, *ccw = nullptr; UpsEntry *ups = nullptr; int device = -1; };
static TDDeviceTables g_tables[64];
static std::mutex g_tablesLock;      // several contexts (one host thread each) may reach the first use together

static int td_tables(scuffcuda_context *ctx, TDTables &T)
{ std::lock_guard<std::mutex> lock(g_tablesLock);
  TDDeviceTables &G = g_tables[ctx->device & 63];
  if (!G.ccx)
   { if (cudaMalloc(&G.ccx, sizeof SC_CC_X_H) != cudaSuccess || cudaMalloc(&G.ccw, sizeof SC_CC_W_H) != cudaSuccess
         || cudaMalloc(&G.ups, sizeof SC_UPS_H) != cudaSuccess) { ctx->err = "Taylor-Duffy table allocation failed"; return 1; }
     cudaMemcpy(G.ccx, SC_CC_X_H, sizeof SC_CC_X_H, cudaMemcpyHostToDevice);
     cudaMemcpy(G.ccw, SC_CC_W_H, sizeof SC_CC_W_H, cudaMemcpyHostToDevice);
     cudaMemcpy(G.ups, SC_UPS_H, sizeof SC_UPS_H, cudaMemcpyHostToDevice);
     cudaDeviceSynchronize();   // pageable H2D copies may return before the DMA lands; the kernels run on non-blocking streams
   }
  T.ccx = G.ccx; T.ccw = G.ccw; T.ups = G.ups;
  for (int i = 0; i < 5; i++) T.upsStart[i] = SC_UPS_START_H[i];
  return 0;
}

// items[t] = 9*nearIndex + 3*i + j : one (pair, Q_a, Q_b) combination that feeds a matrix entry
__global__ void __launch_bounds__(128, 5)
td_ct_kernel(const NearArgs A, const TDTables T, const int *items, int nItems, double *scratch, int maxM, int *itemStatus)
{ const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nItems) return;
  const int item = items[t];
  const NearPair N = A.near[item / 9];
  const int i = (item % 9) / 3, j = item % 3;
  const PanelRec &Pa = A.pa[N.pa], &Pb = A.pb[N.pb];
  const int ea = Pa.ei[i], eb = Pb.ei[j];
  double Va[9], Vb[9];
  for (int q = 0; q < 9; q++) { Va[q] = Pa.V[q]; Vb[q] = A.useDisp ? (Pb.V[q] + 1.0 * A.L[q % 3]) : Pb.V[q]; }
  cd H0, H1;
  int st = td_common_triangle(T, Va + 3 * N.ia[0], Va + 3 * N.ia[1], Va + 3 * N.ia[2], Va + 3 * i, Vb + 3 * j,
                              mk(A.K.kr, A.K.ki), scratch + t, (size_t)gridDim.x * blockDim.x, H0, H1, maxM);
  itemStatus[t] = st;
  if (st) return;     // needs more Clenshaw-Curtis levels than this pass holds: redone by the next pass
  scatter_ppi(A.S, A.K, N.pa, N.pb, ea, eb, Pa.sl[i], Pb.sl[j], H0, H1);
  atomicAdd(A.counters + 2, (A.S.symmetric && N.pa != N.pb && ea == eb) ? 2ull : 1ull);
}

// Cooperative pair-level kernel.  All (Q_a,Q_b) combinations of a common-triangle panel pair share the expensive part of
// the reduced integrand — K_n(X_d(y)) depends only on the triangle and k — so TD_GROUP lanes take one PAIR: the 2^m new
// points of Clenshaw-Curtis level m are evaluated in parallel by the lanes (K once per point, then the <= 9 polynomial
// contractions), the quadrature sums are reduced with warp shuffles, and every item keeps its own convergence state, so
// each one stops exactly where the reference's independent pcubature call stops (pcubature.c: RelTol 1e-10 per
// component, numEval > MaxEval = 1000); only the floating-point summation order of the quadrature sum differs.
#ifndef TD_CT_MINB
#define TD_CT_MINB 8      /* register cap asked of ptxas: 65536 / (128 * MINB) = 64: the kernel is latency-bound and shares the SMs with the
                             HBM-bound gather, so resident warps count for more than spills.  Measured in the step (gather and Taylor-Duffy side by side,
                             N = 20 280): 168 regs 16.7 ms, 96 regs 16.2 ms, 64 regs 15.95 ms with 2.4 KB of local memory per thread */
#endif
#define TD_GROUP_DEFAULT 8
#define TD_COOP_MAXM 9
#define TD_COOP_NPTS ((1 << (TD_COOP_MAXM + 1)) + 1)
#define TD_COOP_UDBL 64                                          /* doubles of coefficients per item (60 used) */
#define TD_COOP_PAIR_DBL (9 * TD_COOP_UDBL + 9 * 4 * TD_COOP_NPTS)  /* scratch doubles per pair */

// f[it][0..3] for all items of the pair at one point y
__device__ void td_ct_pair_eval(double A2, double AdB, double B2, cd k, int nLo, int nHi, int nIt, const double *Uc, const int *nmm,
                                double y, double *fout /* [nIt][4] at stride 4 */)
{ double X[3], u1, u2;
  u1 = 1.0; u2 = y;       X[0] = sqrt(A2 * u1 * u1 + 2.0 * AdB * u1 * u2 + B2 * u2 * u2);
  u1 = y;   u2 = (y - 1); X[1] = sqrt(A2 * u1 * u1 + 2.0 * AdB * u1 * u2 + B2 * u2 * u2);
  u1 = y;   u2 = 1.0;     X[2] = sqrt(A2 * u1 * u1 + 2.0 * AdB * u1 * u2 + B2 * u2 * u2);
  const cd IK = mk(-k.y, k.x);
  cd K[3][7];
  for (int d = 0; d < 3; d++)
   { cd IKX = mk(IK.x * X[d], IK.y * X[d]), eIKX = cexp_(IKX), mIKX = mk(-IKX.x, -IKX.y);
     const bool small = (mIKX.y == 0.0) ? (fabs(mIKX.x) < 0.1) : (cabs_(mIKX) < 0.1);
     cd emIKX = mk(0.0, 0.0);
     if (!small) emIKX = (mIKX.y == 0.0) ? mk(exp(mIKX.x), 0.0) : cexp_(mIKX);
     for (int n = nLo + 1; n <= nHi + 1; n++)
      { cd t = eIKX * exprel_v3p0_pre(n, mIKX, emIKX, small); const double dn = n * X[d], rdn = 1.0 / dn; K[d][n] = mk(td_div_rcp(t.x, dn, rdn), td_div_rcp(t.y, dn, rdn)); }
   }
  const double s = 1.0 / (4.0 * M_PI);
  for (int it = 0; it < nIt; it++)
   { const double *U = Uc + it * TD_COOP_UDBL;        // [np][d][n][2]
     for (int np = 0; np < 2; np++)
      { cd Sum = mk(0.0, 0.0);
        for (int n = nmm[4 * it + 2 * np]; n <= nmm[4 * it + 2 * np + 1]; n++)
         for (int d = 0; d < 3; d++)
          { const double *u = U + ((np * 3 + d) * TD_NUMWPOWERS + n) * 2;
            Sum += (u[0] + y * u[1]) * K[d][n + 1]; }
        fout[4 * it + 2 * np] = Sum.x * s; fout[4 * it + 2 * np + 1] = Sum.y * s;
      }
   }
}

// pairs[g] = 512*nearIndex + 9-bit mask of the (i,j) combinations (bit 3i+j) that feed a matrix entry
template <int TD_GROUP>
__global__ void __launch_bounds__(128, TD_CT_MINB)
td_ct_pair_kernel(const NearArgs A, const TDTables T, const int *pairs, int nPairs, double *scratch, int *pairStatus)
{ const int g = (blockIdx.x * blockDim.x + threadIdx.x) / TD_GROUP;
  const int gl = threadIdx.x & (TD_GROUP - 1);
  if (g >= nPairs) return;                                                  // whole groups leave together
  const unsigned gmask = (TD_GROUP == 32 ? 0xffffffffu : ((1u << (TD_GROUP & 31)) - 1u)) << ((threadIdx.x & 31) & ~(TD_GROUP - 1));
  const int code = pairs[g];
  const NearPair N = A.near[code >> 9];
  const PanelRec &Pa = A.pa[N.pa], &Pb = A.pb[N.pb];
  int itIJ[9], nIt = 0;
  for (int b = 0; b < 9; b++) if (code & (1 << b)) itIJ[nIt++] = b;
  double Va[9], Vb[9];
  for (int q = 0; q < 9; q++) { Va[q] = Pa.V[q]; Vb[q] = A.useDisp ? (Pb.V[q] + 1.0 * A.L[q % 3]) : Pb.V[q]; }
  const cd k = mk(A.K.kr, A.K.ki);
  // (the per-item coefficients were tried in shared memory, 72 KB per CTA: the occupancy loss made the kernel slower)
  double *Uc = scratch + (size_t)g * TD_COOP_PAIR_DBL;                      // [it][64]
  double *cache = Uc + 9 * TD_COOP_UDBL;                                    // [point][nIt][4]
  __shared__ int s_nmm[128 / TD_GROUP][36];                                 // nMin/nMax per item and polynomial
  int *nmm = s_nmm[threadIdx.x / TD_GROUP];
  double A2 = 0, AdB = 0, B2 = 0;
  { // per-item polynomial coefficients (lanes split the items); the geometric parameters are item independent
    TDCommonTriangle W;
    for (int it = gl; it < nIt + TD_GROUP; it += TD_GROUP)
     { const int itc = it < nIt ? it : 0;
       td_ct_setup(T, Va + 3 * N.ia[0], Va + 3 * N.ia[1], Va + 3 * N.ia[2], Va + 3 * (itIJ[itc] / 3), Vb + 3 * (itIJ[itc] % 3), k, W);
       A2 = W.A2; AdB = W.AdB; B2 = W.B2;
       if (it < nIt)
        { double *U = Uc + it * TD_COOP_UDBL;
          for (int np = 0; np < 2; np++) for (int d = 0; d < 3; d++) for (int n = 0; n < TD_NUMWPOWERS; n++)
           { U[((np * 3 + d) * TD_NUMWPOWERS + n) * 2] = W.U[np][d][n][0]; U[((np * 3 + d) * TD_NUMWPOWERS + n) * 2 + 1] = W.U[np][d][n][1]; }
          nmm[4 * it] = W.nMin[0]; nmm[4 * it + 1] = W.nMax[0]; nmm[4 * it + 2] = W.nMin[1]; nmm[4 * it + 3] = W.nMax[1]; }
       if (it >= nIt) break;
     }
  }
  __syncwarp(gmask);
  int nLo = TD_NUMWPOWERS, nHi = -1;
  for (int q = 0; q < 2 * nIt; q++) { nLo = min(nLo, nmm[2 * q]); nHi = max(nHi, nmm[2 * q + 1]); }
  const double c = 0.5, r = 0.5, V = 0.5;
  const int pstride = 4 * nIt;
  if (gl < 3)
   { const double y = (gl == 0) ? c : (gl == 1 ? c + r * T.ccx[0] : c - r * T.ccx[0]);
     td_ct_pair_eval(A2, AdB, B2, k, nLo, nHi, nIt, Uc, nmm, y, cache + (size_t)gl * pstride); }
  __syncwarp(gmask);
  int m = 0, numEval = 0, st = 0;
  unsigned active = (1u << nIt) - 1u;
  double prev[9][4], fin[9][4];
  for (;;)
   { const double *wb = T.ccw + m + (1 << m) - 1;
     const int npts = (1 << (m + 1)) + 1;
     const bool budget = numEval > 1000;
     for (int it = 0; it < nIt; it++)
      { if (!(active & (1u << it))) continue;
        double acc[4] = { 0, 0, 0, 0 };
        for (int e = gl; e < npts; e += TD_GROUP)
         { const double w = V * (e == 0 ? wb[0] : wb[1 + ((e - 1) >> 1)]);
           const double *cv = cache + (size_t)e * pstride + 4 * it;
           for (int q = 0; q < 4; q++) acc[q] += cv[q] * w; }
#pragma unroll
        for (int o = TD_GROUP / 2; o > 0; o >>= 1)
         for (int q = 0; q < 4; q++) acc[q] += __shfl_xor_sync(gmask, acc[q], o);
        if (m == 0) for (int q = 0; q < 4; q++) prev[it][q] = cache[4 * it + q] * (V * 2);   // trivial (midpoint) rule
        bool conv = true;
        for (int q = 0; q < 4; q++)
         { const double err = fabs(acc[q] - prev[it][q]); prev[it][q] = acc[q];
           if (err > 0.0 && err > fabs(acc[q]) * 1.0e-10) conv = false; }
        if (conv || budget) { for (int q = 0; q < 4; q++) fin[it][q] = acc[q]; active &= ~(1u << it); }
      }
     if (!active) break;
     m += 1;
     if (m > TD_COOP_MAXM) { st = 1; break; }
     const int nx = 1 << (m - 1);
     const double *x = T.ccx + nx;
     double *cvl = cache + (size_t)((1 << m) + 1) * pstride;
     for (int q2 = gl; q2 < 2 * nx; q2 += TD_GROUP)
      { const double y = (q2 & 1) ? c - r * x[q2 >> 1] : c + r * x[q2 >> 1];
        td_ct_pair_eval(A2, AdB, B2, k, nLo, nHi, nIt, Uc, nmm, y, cvl + (size_t)q2 * pstride); }
     __syncwarp(gmask);
     numEval += 2 * nx;
   }
  if (gl == 0) pairStatus[g] = st;
  if (st) return;
  const cd k2 = k * k;
  for (int it = gl; it < nIt; it += TD_GROUP)
   { const int i = itIJ[it] / 3, j = itIJ[it] % 3;
     const int ea = Pa.ei[i], eb = Pb.ei[j];
     const cd R0 = mk(fin[it][0], fin[it][1]), R1 = mk(fin[it][2], fin[it][3]);
     const cd H0 = R1 - cdiv(mk(4.0 * R0.x, 4.0 * R0.y), k2), H1 = mk(0.0, 0.0);
     scatter_ppi(A.S, A.K, N.pa, N.pb, ea, eb, Pa.sl[i], Pb.sl[j], H0, H1);
     atomicAdd(A.counters + 2, (A.S.symmetric && N.pa != N.pb && ea == eb) ? 2ull : 1ull);
   }
}

// ---- self-test of td_math.cuh td_div_int (tests/test_fastmath_gpu.py): out[6 i + 5] = x[i] / ((i % 127) + 1) ----
__global__ void td_div_selftest_kernel(int n, const double *x, double *out)
{ const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[6 * i + 5] = td_div_int(x[i], (i % 127) + 1);
}
int run_td_div_selftest(scuffcuda_context *ctx, int n, const double *xh, double *outh)
{ double *dx = nullptr, *dout = nullptr;
  if (cudaMalloc(&dx, n * sizeof(double)) != cudaSuccess || cudaMalloc(&dout, 6 * (size_t)n * sizeof(double)) != cudaSuccess)
   { cudaFree(dx); ctx->err = "selftest allocation failed"; return 1; }
  cudaMemcpyAsync(dx, xh, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(dout, outh, 6 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  ctx->launches++;
  td_div_selftest_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(n, dx, dout);
  cudaMemcpyAsync(outh, dout, 6 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  cudaFree(dx); cudaFree(dout);
  if (e != cudaSuccess) { ctx->err = std::string("td_div selftest: ") + cudaGetErrorString(e); return 1; }
  return 0;
}

static int td_grow_scratch(scuffcuda_context *ctx, size_t need)
{ if (need <= ctx->scratchBytes) return 0;
  if (ctx->dscratch) cudaFree(ctx->dscratch); ctx->dscratch = nullptr; ctx->scratchBytes = 0;
  if (cudaMalloc(&ctx->dscratch, need) != cudaSuccess) { cudaGetLastError(); ctx->err = "Taylor-Duffy scratch allocation failed"; return 1; }
  ctx->scratchBytes = need; return 0;
}

int run_td(scuffcuda_context *ctx, const Plan &P, const NearArgs &A, int kmode)
{ (void)kmode;
  // classify on the host (cheap: O(#near pairs)): which pairs need Taylor-Duffy / high-order cubature
  const Surface &Sa = ctx->surfaces[P.nsa], &Sb = ctx->surfaces[P.nsb];
  // The algorithm of a near pair depends on the frequency only through the short-wavelength tests |k| R > pi, 20
  // (near_algorithm).  While |k| * (largest panel radius) <= pi no pair can switch, so the item lists of the previous
  // frequency point of the sweep are reused (they are O(#near pairs) host work otherwise).
  struct ListCache { bool valid = false; uint64_t genA = 0, genB = 0; int nsa = -1, nsb = -1, forceTD = 0, nNear = 0; bool symmetric = false, useDisp = false;
                     double L[3] = { 0, 0, 0 }; std::vector<int> items, hoIdx, genItems; int colLo = 0, colHi = 0; };
  static std::map<scuffcuda_context *, ListCache> g_cache;     // per context (a device may carry several contexts)
  static std::mutex g_cacheLock;
  g_cacheLock.lock(); ListCache &C = g_cache[ctx]; g_cacheLock.unlock();
  double Rmax = 0.0;
  for (const PanelRec &Q : Sa.hpanels) Rmax = std::max(Rmax, Q.R);
  for (const PanelRec &Q : Sb.hpanels) Rmax = std::max(Rmax, Q.R);
  const bool noSW = !(A.kabs * Rmax > SW_THRESHOLD);
  const bool reuse = C.valid && noSW && C.colLo == A.S.colLo && C.colHi == A.S.colHi && C.genA == P.genA && C.genB == P.genB && C.nsa == P.nsa && C.nsb == P.nsb && C.forceTD == A.forceTD
                     && C.nNear == P.nNear && C.symmetric == P.symmetric && C.useDisp == P.useDisp
                     && C.L[0] == P.L[0] && C.L[1] == P.L[1] && C.L[2] == P.L[2];
  std::vector<int> &items = C.items, &hoIdx = C.hoIdx, &genItems = C.genItems;
  if (!reuse)
   { C.valid = false; items.clear(); hoIdx.clear(); genItems.clear();
  for (int n = 0; n < P.nNear; n++)
   { const NearPair &N = P.hnear[n]; bool verySW;
     int alg = near_algorithm(N.ncv, A.kabs, Sa.hpanels[N.pa].R, Sb.hpanels[N.pb].R, A.forceTD, verySW);
     if (alg == 4) continue;
     if (alg == 1) { hoIdx.push_back(n); continue; }                  // high-order cubature
     const int hk = (alg == 2 && verySW) ? 18 : 0;                    // high-k Taylor-Duffy kernels (PPIALG_HKTD)
     // the (Q_a,Q_b) combinations that feed a matrix entry
     const PanelRec &Pa = Sa.hpanels[N.pa], &Pb = Sb.hpanels[N.pb];
     for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
      { const int ea = Pa.ei[i], eb = Pb.ei[j];
        if (ea < 0 || eb < 0) continue;
        { // column-block assembly: only items whose matrix column is owned
          const int w = A.S.bPEC ? 1 : 2, ec = A.S.symmetric ? std::max(ea, eb) : eb;
          if (A.S.colOff + w * ec + w - 1 < A.S.colLo || A.S.colOff + w * ec >= A.S.colHi) continue; }
        if (N.ncv == 3)
         { if (A.S.symmetric && N.pa == N.pb && ea > eb) continue;
           if (hk) genItems.push_back(36 * n + hk + 3 * i + j); else items.push_back(9 * n + 3 * i + j); }
        else if (!A.S.symmetric) genItems.push_back(36 * n + hk + 3 * i + j);
        else
         { // symmetric block, pa < pb visited once: entry (min,max)(ea,eb) is EEI(nea<=neb), whose panel order decides
           // which of PPI(pa,pb) / PPI(pb,pa) the reference evaluates; a diagonal edge pair gets both
           if (ea <= eb) genItems.push_back(36 * n + hk + 3 * i + j);
           if (ea >= eb) genItems.push_back(36 * n + hk + 9 + 3 * i + j); }
      }
   }
     C.colLo = A.S.colLo; C.colHi = A.S.colHi;
     C.valid = noSW; C.genA = P.genA; C.genB = P.genB; C.nsa = P.nsa; C.nsb = P.nsb; C.forceTD = A.forceTD; C.nNear = P.nNear;
     C.symmetric = P.symmetric; C.useDisp = P.useDisp; for (int q = 0; q < 3; q++) C.L[q] = P.L[q];
   }
  if (run_ho(ctx, A, hoIdx.data(), (int)hoIdx.size(), kmode)) return 1;
  if (!genItems.empty())
   { TDTables TG; if (td_tables(ctx, TG)) return 1;
     if (run_td_general(ctx, A, TG, genItems.data(), (int)genItems.size())) return 1; }
  if (items.empty()) return 0;
  TDTables T; if (td_tables(ctx, T)) return 1;
  std::vector<int> work(items);
  static int coopEnv = -1;
  if (coopEnv < 0) { const char *e = getenv("SCUFF_CUDA_TD_COOP"); coopEnv = e ? atoi(e) : 1; }
  if (coopEnv)
   { // cooperative pair-level kernel: group the items by panel pair (items of a pair are contiguous in `work`)
     std::vector<int> pairs;
     for (size_t q = 0; q < work.size(); q++)
      { const int n = work[q] / 9, b = work[q] % 9;
        if (!pairs.empty() && (pairs.back() >> 9) == n) pairs.back() |= (1 << b); else pairs.push_back(512 * n + (1 << b)); }
     const size_t perPair = (size_t)TD_COOP_PAIR_DBL * sizeof(double);
     const int batch = (int)std::min<size_t>(pairs.size(), ((size_t)12 << 30) / perPair);
     const size_t hdr = (((size_t)batch * 2 * sizeof(int) + 255) / 256) * 256;
     if (td_grow_scratch(ctx, hdr + (size_t)batch * perPair)) return 1;
     int *dPairs = (int *)ctx->dscratch;
     int *hStatus = (int *)mapped_host_scratch(ctx, (size_t)batch * sizeof(int), 1);
     if (!hStatus) { ctx->err = "Taylor-Duffy: cannot allocate mapped host memory"; return 1; }
     double *scr = (double *)((char *)ctx->dscratch + hdr);
     std::vector<int> status;
     for (size_t first = 0; first < pairs.size(); first += batch)
      { const int n = (int)std::min<size_t>(batch, pairs.size() - first);
        cudaMemcpyAsync(dPairs, pairs.data() + first, n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
        ctx->launches++;
        static const int group = []{ const char *e = getenv("SCUFF_CUDA_TD_GROUP"); int g = e ? atoi(e) : TD_GROUP_DEFAULT;
                                     return (g == 8 || g == 16 || g == 32) ? g : TD_GROUP_DEFAULT; }();
        const long threads = (long)n * group;
        const unsigned grid = (unsigned)((threads + 127) / 128);
        if (group == 8) td_ct_pair_kernel<8><<<grid, 128, 0, ctx->stream>>>(A, T, dPairs, n, scr, hStatus);
        else if (group == 16) td_ct_pair_kernel<16><<<grid, 128, 0, ctx->stream>>>(A, T, dPairs, n, scr, hStatus);
        else td_ct_pair_kernel<32><<<grid, 128, 0, ctx->stream>>>(A, T, dPairs, n, scr, hStatus);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        status.assign(hStatus, hStatus + n);
        if (e != cudaSuccess) { ctx->err = std::string("td_ct_pair_kernel: ") + cudaGetErrorString(e); return 1; }
        for (int q = 0; q < n; q++)
         if (status[q]) { ctx->err = "Taylor-Duffy cubature exceeded the tabulated Clenshaw-Curtis levels"; return 1; }
      }
     return 0;
   }
  // pass 0: every item at once with room for Clenshaw-Curtis levels <= 6 (129 points; the common case),
  // pass 1: the items that asked for more, with the full MaxEval=1000 budget (level 9, 1025 points), in batches
  // with HBM to spare every item gets room for the full budget at once (level 9: 64 KB per item) and nobody restarts
  const bool roomy = (size_t)items.size() * 4 * 1025 * sizeof(double) <= ((size_t)8 << 30);
  const int passM[2] = { roomy ? 9 : 6, 9 }, passBatch[2] = { 1 << 30, 32768 };
  std::vector<int> status;
  for (int pass = 0; pass < 2 && !work.empty(); pass++)
   { const int maxM = passM[pass];
     const size_t perItem = (size_t)4 * ((1u << (maxM + 1)) + 1) * sizeof(double);   // level m holds 2^(m+1)+1 points
     const int batch = (int)std::min<size_t>(work.size(), (size_t)passBatch[pass]);
     const int threads = ((batch + 127) / 128) * 128;
     const size_t hdr = (((size_t)batch * 2 * sizeof(int) + 255) / 256) * 256;
     if (td_grow_scratch(ctx, hdr + (size_t)threads * perItem)) return 1;
     int *dItems = (int *)ctx->dscratch;
     int *hStatus = (int *)mapped_host_scratch(ctx, (size_t)batch * sizeof(int), 1);
     if (!hStatus) { ctx->err = "Taylor-Duffy: cannot allocate mapped host memory"; return 1; }
     double *scr = (double *)((char *)ctx->dscratch + hdr);
     std::vector<int> next;
     for (size_t first = 0; first < work.size(); first += batch)
      { const int n = (int)std::min<size_t>(batch, work.size() - first);
        cudaMemcpyAsync(dItems, work.data() + first, n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
        ctx->launches++;
        // the grid always spans `threads` threads so that the scratch stride is fixed
        td_ct_kernel<<<threads / 128, 128, 0, ctx->stream>>>(A, T, dItems, n, scr, maxM, hStatus);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        status.assign(hStatus, hStatus + n);
        if (e != cudaSuccess) { ctx->err = std::string("td_ct_kernel: ") + cudaGetErrorString(e); return 1; }
        for (int q = 0; q < n; q++) if (status[q]) next.push_back(work[first + q]);
      }
     work.swap(next);
   }
  if (!work.empty()) { ctx->err = "Taylor-Duffy cubature exceeded the tabulated Clenshaw-Curtis levels"; return 1; }
  return 0;
}

}  // namespace scuffcuda
