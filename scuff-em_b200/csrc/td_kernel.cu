// Taylor-Duffy kernel: panel pairs that share all three vertices (ncv == 3; the reference uses
// Taylor-Duffy for them at every frequency, PanelPanelInteractions.cc:449-491).  One thread per
// (pair, Q_a, Q_b) combination runs the 1-D p-adaptive Clenshaw-Curtis cubature of td_math.cuh with the
// reference's tolerances (RelTol 1e-10, MaxEval 1000), its function-value cache living in a strided
// HBM scratch block.
#include <vector>
#include "kernels.h"
#include "stamp.cuh"
#include "td_math.cuh"
#include "sw_kernels.h"
#include "generated_tables.h"

namespace scuffcuda {

static_assert(sizeof(SCUpsEntry) == sizeof(UpsEntry), "generated table layout");

struct TDDeviceTables { double *ccx = nullptr, *ccw = nullptr; UpsEntry *ups = nullptr; int device = -1; };
static TDDeviceTables g_tables[64];

static int td_tables(scuffcuda_context *ctx, TDTables &T)
{ TDDeviceTables &G = g_tables[ctx->device & 63];
  if (!G.ccx)
   { if (cudaMalloc(&G.ccx, sizeof SC_CC_X_H) != cudaSuccess || cudaMalloc(&G.ccw, sizeof SC_CC_W_H) != cudaSuccess
         || cudaMalloc(&G.ups, sizeof SC_UPS_H) != cudaSuccess) { ctx->err = "Taylor-Duffy table allocation failed"; return 1; }
     cudaMemcpy(G.ccx, SC_CC_X_H, sizeof SC_CC_X_H, cudaMemcpyHostToDevice);
     cudaMemcpy(G.ccw, SC_CC_W_H, sizeof SC_CC_W_H, cudaMemcpyHostToDevice);
     cudaMemcpy(G.ups, SC_UPS_H, sizeof SC_UPS_H, cudaMemcpyHostToDevice);
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
  static ListCache g_cache[64];
  ListCache &C = g_cache[ctx->device & 63];
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
     int *dItems = (int *)ctx->dscratch, *dStatus = dItems + batch;
     double *scr = (double *)((char *)ctx->dscratch + hdr);
     std::vector<int> next;
     for (size_t first = 0; first < work.size(); first += batch)
      { const int n = (int)std::min<size_t>(batch, work.size() - first);
        cudaMemcpyAsync(dItems, work.data() + first, n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
        ctx->launches++;
        // the grid always spans `threads` threads so that the scratch stride is fixed
        td_ct_kernel<<<threads / 128, 128, 0, ctx->stream>>>(A, T, dItems, n, scr, maxM, dStatus);
        status.resize(n);
        cudaMemcpyAsync(status.data(), dStatus, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { ctx->err = std::string("td_ct_kernel: ") + cudaGetErrorString(e); return 1; }
        for (int q = 0; q < n; q++) if (status[q]) next.push_back(work[first + q]);
      }
     work.swap(next);
   }
  if (!work.empty()) { ctx->err = "Taylor-Duffy cubature exceeded the tabulated Clenshaw-Curtis levels"; return 1; }
  return 0;
}

}  // namespace scuffcuda
