// Taylor-Duffy kernel: panel pairs that share all three vertices (ncv == 3; the reference uses
// Taylor-Duffy for them at every frequency, PanelPanelInteractions.cc:449-491).  One thread per
// (pair, Q_a, Q_b) combination runs the 1-D p-adaptive Clenshaw-Curtis cubature of td_math.cuh with the
// reference's tolerances (RelTol 1e-10, MaxEval 1000), its function-value cache living in a strided
// HBM scratch block.
#include <vector>
#include "kernels.h"
#include "stamp.cuh"
#include "td_math.cuh"
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

__global__ void __launch_bounds__(64)
td_ct_kernel(const NearArgs A, const TDTables T, const int *tdIdx, int first, int nItems, double *scratch, int *status)
{ const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nItems) return;
  const int item = first + t;
  const NearPair N = A.near[tdIdx[item / 9]];
  const int i = (item % 9) / 3, j = item % 3;
  const PanelRec &Pa = A.pa[N.pa], &Pb = A.pb[N.pb];
  const int ea = Pa.ei[i], eb = Pb.ei[j];
  if (ea < 0 || eb < 0) return;
  if (A.S.symmetric && N.pa == N.pb && ea > eb) return;
  double Va[9], Vb[9];
  for (int q = 0; q < 9; q++) { Va[q] = Pa.V[q]; Vb[q] = A.useDisp ? (Pb.V[q] + 1.0 * A.L[q % 3]) : Pb.V[q]; }
  cd H0, H1;
  int st = td_common_triangle(T, Va + 3 * N.ia[0], Va + 3 * N.ia[1], Va + 3 * N.ia[2], Va + 3 * i, Vb + 3 * j,
                              mk(A.K.kr, A.K.ki), scratch + t, (size_t)gridDim.x * blockDim.x, H0, H1);
  if (st) atomicExch(status, 1);
  scatter_ppi(A.S, A.K, N.pa, N.pb, ea, eb, Pa.sl[i], Pb.sl[j], H0, H1);
  atomicAdd(A.counters + 2, (A.S.symmetric && N.pa != N.pb && ea == eb) ? 2ull : 1ull);
}

int run_td(scuffcuda_context *ctx, const Plan &P, const NearArgs &A, int kmode)
{ (void)kmode;
  // classify on the host (cheap: O(#near pairs)): which pairs need Taylor-Duffy / high-order cubature
  const Surface &Sa = ctx->surfaces[P.nsa], &Sb = ctx->surfaces[P.nsb];
  std::vector<int> td;
  for (int n = 0; n < P.nNear; n++)
   { const NearPair &N = P.hnear[n]; bool verySW;
     int alg = near_algorithm(N.ncv, A.kabs, Sa.hpanels[N.pa].R, Sb.hpanels[N.pb].R, A.forceTD, verySW);
     if (alg == 4) continue;
     if (alg == 2 && N.ncv == 3 && !verySW) { td.push_back(n); continue; }
     ctx->err = "short-wavelength regime (|k|*panel radius > pi) or SCUFF_DISABLE_CACHE: high-order cubature / "
                "common-edge and common-vertex Taylor-Duffy are not yet available on the device path";
     return 1;
   }
  if (td.empty()) return 0;
  TDTables T; if (td_tables(ctx, T)) return 1;
  int *dIdx = nullptr, *dStatus = nullptr;
  const int batch = 32768;
  size_t need = td.size() * sizeof(int) + 256 + (size_t)batch * TD_CACHE_DOUBLES * sizeof(double);
  if (need > ctx->scratchBytes)
   { if (ctx->dscratch) cudaFree(ctx->dscratch); ctx->dscratch = nullptr; ctx->scratchBytes = 0;
     if (cudaMalloc(&ctx->dscratch, need) != cudaSuccess) { ctx->err = "Taylor-Duffy scratch allocation failed"; return 1; }
     ctx->scratchBytes = need; }
  dStatus = (int *)ctx->dscratch; dIdx = (int *)((char *)ctx->dscratch + 256);
  double *scr = (double *)((char *)ctx->dscratch + 256 + ((td.size() * sizeof(int) + 255) / 256) * 256);
  cudaMemsetAsync(dStatus, 0, sizeof(int), ctx->stream);
  cudaMemcpyAsync(dIdx, td.data(), td.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
  cudaStreamSynchronize(ctx->stream);   // td (host vector) must outlive the copy
  const int total = (int)td.size() * 9;
  for (int first = 0; first < total; first += batch)
   { int n = std::min(batch, total - first);
     // grid covers exactly `batch` threads so that the scratch stride is fixed
     ctx->launches++;
     td_ct_kernel<<<batch / 64, 64, 0, ctx->stream>>>(A, T, dIdx, first, n, scr, dStatus);
   }
  int st = 0;
  cudaMemcpyAsync(&st, dStatus, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { ctx->err = std::string("td_ct_kernel: ") + cudaGetErrorString(e); return 1; }
  if (st) { ctx->err = "Taylor-Duffy cubature exceeded the tabulated Clenshaw-Curtis levels"; return 1; }
  return 0;
}

}  // namespace scuffcuda
