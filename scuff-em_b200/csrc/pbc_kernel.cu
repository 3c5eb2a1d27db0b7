// Outer-cell step of the periodic assembly (AssembleBEMMatrix.cc:526-557 -> GetSurfaceSurfaceInteractions with GBA1/GBA2 ->
// GetPanelPanelInteractions PanelPanelInteractions.cc:420-424 -> GetPPIs_Cubature with AssembleInnerPPIIntegrand's GBA
// branch :124-153): EVERY panel pair of the two surfaces is integrated with the 6x6 rule, the kernel being the
// Bloch-periodic Green's function without its innermost 3x3 cells,
//     H0 = sum w w' (F.F' + 4/(ik)^2) GBar(R) ,   H1 = sum w w' (F x F') . grad GBar(R) ,   R = X - X' ,
// GBar and its gradient coming from the tricubic table (gbar_table.cuh) or, outside the table's |z| range, from a full
// Ewald sum (gbar.cuh) — which for the reference's fixtures is the common case, because RWGSurface::UpdateBoundingBox
// looks at three vertices only.
//
// Mapping: one cubature POINT PAIR per thread (an Ewald sum is ~10^5 flop, so the 36 points of a panel pair are the
// natural unit of parallelism), 36 threads per panel pair, 8 panel pairs per CTA.  The 36 weighted kernel values go
// through shared memory into the 3x3 barycentric moments of GBar and of its three gradient components (36 complex
// sums, one per thread, fixed summation order), and nine threads contract them to the (Q,Q') combinations
//     H0_ij = sum_{X in {A,B}, Y in {A',B'}} (X.Y) <(x - d_i)(y - d_j) GBar> + 4/(ik)^2 <GBar>
//     H1_ij = sum_{X,Y} (X x Y) . <(x - d_i)(y - d_j) grad GBar>
// (same barycentric algebra as far_kernel.cu; grad GBar is not parallel to R, hence the vector moments) and scatter.
#include "pbc.h"
#include "gbar_table.cuh"
#include "classify.cuh"
#include "stamp.cuh"

namespace scuffcuda {

#define PBC_PAIRS 8
#define PBC_THREADS (36 * PBC_PAIRS)

struct PbcArgs
{ const PanelRec *pa; int NPa;
  const PanelRec *pb; int NPb;
  GBarTableView T;
  KParams K; StampParams S;
  unsigned long long *counters;
};

__device__ __forceinline__ double tcr4_u(int p)
{ return p == 1 ? 0.8168475729805 : (p == 0 || p == 2) ? 0.0915762135098 : (p == 3) ? 0.1081030181681 : 0.4459484909160; }
__device__ __forceinline__ double tcr4_v(int p)
{ return p == 2 ? 0.8168475729805 : (p == 0 || p == 1) ? 0.0915762135098 : (p == 4) ? 0.1081030181681 : 0.4459484909160; }

__global__ void __launch_bounds__(PBC_THREADS)
pbc_outer_kernel(const PbcArgs A)
{ __shared__ double sVal[PBC_PAIRS][36][8];   // w w' { GBar, d/dx, d/dy, d/dz } as (re, im)
  __shared__ cd sMom[PBC_PAIRS][4][3][3];     // [component][a-side {1,u,v}][b-side {1,u',v'}]
  const int lp = threadIdx.x / 36, t = threadIdx.x % 36;
  const long pair = (long)blockIdx.x * PBC_PAIRS + lp;
  const bool live = pair < (long)A.NPa * A.NPb;
  const int pa = live ? (int)(pair / A.NPb) : 0, pb = live ? (int)(pair % A.NPb) : 0;
  const PanelRec &Pa = A.pa[pa], &Pb = A.pb[pb];
  if (live)
   { const int p = t / 6, pp = t % 6;
     // X = V0 + u A + v B in the reference's operation order (PanelPanelInteractions.cc:292-296,316-320), so that the
     // table's |z|-range test sees the reference's R_z to the last bit
     double R[3];
#pragma unroll
     for (int m = 0; m < 3; m++)
      { const double Am = SC_SUB(Pa.V[3 + m], Pa.V[m]), Bm = SC_SUB(Pa.V[6 + m], Pa.V[m]);
        const double X = SC_ADD(SC_ADD(Pa.V[m], SC_MUL(tcr4_u(p), Am)), SC_MUL(tcr4_v(p), Bm));
        const double APm = SC_SUB(Pb.V[3 + m], Pb.V[m]), BPm = SC_SUB(Pb.V[6 + m], Pb.V[m]);
        const double XP = SC_ADD(SC_ADD(Pb.V[m], SC_MUL(tcr4_u(pp), APm)), SC_MUL(tcr4_v(pp), BPm));
        R[m] = SC_SUB(X, XP); }
     cd G, dG[3];
     gbar_get(A.T, R, G, dG);
     const double w = tcr4_w(p) * tcr4_w(pp);
     double *v = sVal[lp][t];
     v[0] = w * G.x; v[1] = w * G.y;
#pragma unroll
     for (int m = 0; m < 3; m++) { v[2 + 2 * m] = w * dG[m].x; v[3 + 2 * m] = w * dG[m].y; } }
  __syncthreads();
  if (live)
   { // thread t -> moment (component, m, n)
     const int comp = t / 9, m = (t % 9) / 3, n = t % 3;
     double sr = 0.0, si = 0.0;
     for (int p = 0; p < 6; p++)
      { const double fa = (m == 0) ? 1.0 : (m == 1 ? tcr4_u(p) : tcr4_v(p));
        for (int pp = 0; pp < 6; pp++)
         { const double f = fa * ((n == 0) ? 1.0 : (n == 1 ? tcr4_u(pp) : tcr4_v(pp)));
           sr += f * sVal[lp][6 * p + pp][2 * comp]; si += f * sVal[lp][6 * p + pp][2 * comp + 1]; } }
     sMom[lp][comp][m][n] = mk(sr, si); }
  __syncthreads();
  if (!live || t >= 9) return;
  const int i = t / 3, j = t % 3;
  const int ea = Pa.ei[i], eb = Pb.ei[j];
  if (ea < 0 || eb < 0) return;
  double Xa[2][3], Yb[2][3];
#pragma unroll
  for (int m = 0; m < 3; m++)
   { Xa[0][m] = Pa.V[3 + m] - Pa.V[m]; Xa[1][m] = Pa.V[6 + m] - Pa.V[m];
     Yb[0][m] = Pb.V[3 + m] - Pb.V[m]; Yb[1][m] = Pb.V[6 + m] - Pb.V[m]; }
  cd H0 = A.K.c4 * sMom[lp][0][0][0], H1 = mk(0.0, 0.0);
#pragma unroll
  for (int al = 1; al <= 2; al++)
#pragma unroll
   for (int be = 1; be <= 2; be++)
    { const bool si = (i == al), sj = (j == be);
      const double *X = Xa[al - 1], *Y = Yb[be - 1];
      const double dotXY = X[0] * Y[0] + X[1] * Y[1] + X[2] * Y[2];
      const double cr[3] = { X[1] * Y[2] - X[2] * Y[1], X[2] * Y[0] - X[0] * Y[2], X[0] * Y[1] - X[1] * Y[0] };
#pragma unroll
      for (int comp = 0; comp < 4; comp++)
       { const cd (*M)[3] = sMom[lp][comp];
         cd S = M[al][be];
         if (sj) S -= M[al][0];
         if (si) S -= M[0][be];
         if (si && sj) S += M[0][0];
         if (comp == 0) H0 += S * dotXY; else H1 += S * cr[comp - 1]; } }
  scatter_ppi<true>(A.S, A.K, pa, pb, ea, eb, Pa.sl[i], Pb.sl[j], H0, H1);
  atomicAdd(A.counters, 1ULL);
}

int run_pbc_outer(scuffcuda_context *ctx, const Surface &Sa, const Surface &Sb, const GBarTable &T, const KParams &K, const StampParams &SP,
                  unsigned long long *counters)
{ extern GBarTableView gbar_view(const GBarTable &T);
  PbcArgs A;
  A.pa = Sa.dpanels; A.NPa = Sa.NP; A.pb = Sb.dpanels; A.NPb = Sb.NP;
  A.T = gbar_view(T); A.K = K; A.S = SP; A.counters = counters;
  const long nPairs = (long)Sa.NP * Sb.NP;
  const long blocks = (nPairs + PBC_PAIRS - 1) / PBC_PAIRS;
  if (blocks > 0x7fffffffL) { ctx->err = "periodic outer-cell step: too many panel pairs for one launch"; return 1; }
  ctx->launches++;
  pbc_outer_kernel<<<(unsigned)blocks, PBC_THREADS, 0, ctx->stream>>>(A);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { ctx->err = std::string("pbc_outer_kernel: ") + cudaGetErrorString(e); return 1; }
  return 0;
}

}  // namespace scuffcuda
