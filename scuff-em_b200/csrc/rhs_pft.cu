// Right-hand-side vector and power / force / torque of a scattering problem on the device: with the
// BEM matrix assembled and factored in HBM, the incident-field projections, the solve and the energy/momentum-transfer
// sums run on the same device, so that one frequency point of a scuff-scatter run (scuff-scatter.cc:402-520) ends with
// 8 doubles per surface crossing PCIe instead of the matrix.
//
// What the reference computes, restated:
//   * AssembleRHSVector (AssembleRHSVector.cc:365-428): RHS[edge] = -+ <f_edge, E_inc>/ZVAC (and -+ <f_edge, H_inc> for a
//     dielectric surface), each inner product by the 78-point order-20 rule over the edge's positive panel (vertices QP, V1,
//     V2) minus the same over its negative panel (V1, V2, QM) (GetInnerProducts :172-229, TriIntFixed libTriInt.cc:279-322);
//     fields whose sources lie in the surface's exterior region enter with -, interior region with + (:297-306).
//   * GetEMTPFTMatrix (EMTPFT.cc:533-905) with the cubature integrals of SCUFF_EMTPFTI_EHDERIVATIVES1 (:82-176, :308-334):
//     for every ordered pair of basis functions (a, b) that act on each other, sums over the Order x Order points of the
//     four panel pairs (GetBFBFCubature2, PanelCubature.cc:791-893; Order = 9 if the two functions share a vertex
//     (AssessBFPair), else 4) of kernels Phi, Psi, Zeta times polynomial factors, contracted with the bilinears
//     conj(k_a) k_b, ... of the solution (GetKNBilinears, GetPFT.cc:95-132).  On one surface only the imaginary parts of the
//     kernels enter (:103-121).  Plus the extinction terms (GetExtinctionPFTT :466-531): the 21-point rule of every panel
//     with the incident field and its gradient.
//
// B200 formulation (not the reference's): the sums are bilinear in the expansion coefficients, so the surface currents
//   K(x) = sum_e k_e f_e(x),  N(x) = sum_e n_e f_e(x)  are tabulated once at the cubature points of every panel
// (pft_currents_kernel) and ONE pass over all ordered pairs of cubature points (pft_pair_kernel, an N-body style tile loop:
// a thread owns one point of surface a, a CTA stages tiles of the points of surface b in shared memory) accumulates
// power, force and both torque terms — 36 point pairs per panel pair instead of 9 basis-function pairs x 4 x 36.  The basis
// function pairs that share a vertex (O(N) of them, listed on the host by the reference's VecClose test) get the difference
// "order 9 minus order 4" from pft_near_kernel, one warp per pair.  All reductions are fixed-order (per-CTA partial sums,
// summed on the host), so results are reproducible run to run.
#include <vector>
#include <map>
#include <algorithm>
#include <complex>
#include <string>
#include "kernels.h"
#include "fastmath.cuh"
#include "generated_tcr20.h"
#include "generated_tcr9.h"
#include "generated_tcr7.h"
#include "nearfield.cuh"

#define ZVAC 376.73031346177      /* libscuff.h:78 */
#define TENTHIRDS 3.33333333333333333333333   /* PFTOptions.h:64 */
#define CVTHRESHOLD 1.0e-6        /* AssessPanelPair.cc:39 */
#define PFT_NQ 10                 /* PScat, F[3], T1[3], T2[3] (PAbs comes from the extinction pass) */
#define PFT_MAXPW 8

namespace scuffcuda {

struct EdgeRec { double QP[3], V1[3], V2[3], QM[3]; double len; int hasM; int pad; int pP, iP, pM, iM; };   // 128 B; pP/iP: positive panel and the index of QP in it, pM/iM: negative panel

// one cubature point of a panel with the surface currents there (160 B)
struct __align__(16) CurPt
{ double x[3]; double w;            // position; weight = 2 A w_p
  double Kr[3], Ki[3];              // K(x)
  double Nr[3], Ni[3];              // N(x) = sum -ZVAC KN[2e+1] f_e(x)
  double dKr, dKi, dNr, dNi;        // surface divergences
};
static_assert(sizeof(CurPt) == 160, "CurPt layout");

struct PlaneWaveDev { cd E0[3]; double n[3]; cd K; cd invZ; double sign; };

struct NearBF { int ea, eb; };

struct PftNearList { uint64_t genA = 0, genB = 0; int n = 0; NearBF *d = nullptr; };

struct PftState
{ double *dtcr20 = nullptr, *dtcr9 = nullptr, *dtcr4 = nullptr, *dtcr7 = nullptr;
  double *dfx = nullptr; size_t fxCap = 0;          // evaluation points + field sums of scuffcuda_get_fields
  std::map<std::pair<int, int>, PftNearList> nearLists;
  CurPt *dcur = nullptr; size_t curCap = 0;          // current points of all surfaces (order 4, then order 9)
  double *dpart = nullptr; size_t partCap = 0;       // per-CTA partial sums
  double *dtab = nullptr; size_t tabCap = 0;         // tabulated incident fields
};

#define PCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { ctx->err = std::string("rhs/pft: ") + #call + ": " + cudaGetErrorString(e_); return 1; } } while (0)
static int pfail(scuffcuda_context *ctx, const std::string &msg) { ctx->err = msg; return 1; }

static PftState *pft_state(scuffcuda_context *ctx)
{ if (!ctx->pftState) ctx->pftState = new PftState();
  return (PftState *)ctx->pftState;
}
void pft_release(scuffcuda_context *ctx)
{ PftState *P = (PftState *)ctx->pftState; if (!P) return;
  if (P->dtcr20) cudaFree(P->dtcr20); if (P->dtcr9) cudaFree(P->dtcr9); if (P->dtcr4) cudaFree(P->dtcr4); if (P->dtcr7) cudaFree(P->dtcr7); if (P->dfx) cudaFree(P->dfx);
  for (auto &l : P->nearLists) if (l.second.d) cudaFree(l.second.d);
  if (P->dcur) cudaFree(P->dcur); if (P->dpart) cudaFree(P->dpart); if (P->dtab) cudaFree(P->dtab);
  delete P; ctx->pftState = nullptr;
}
static int pft_tables(scuffcuda_context *ctx)
{ PftState *P = pft_state(ctx);
  if (P->dtcr20) return 0;
  static const double U4[6] = SCUFF_TCR4_U, V4[6] = SCUFF_TCR4_V, W4[6] = SCUFF_TCR4_W;
  double T4[18]; for (int p = 0; p < 6; p++) { T4[3 * p] = U4[p]; T4[3 * p + 1] = V4[p]; T4[3 * p + 2] = W4[p]; }
  PCK(cudaMalloc(&P->dtcr4, sizeof T4)); PCK(cudaMalloc(&P->dtcr9, sizeof SC_TCR9_H)); PCK(cudaMalloc(&P->dtcr20, sizeof SC_TCR20_H));
  PCK(cudaMalloc(&P->dtcr7, sizeof SC_TCR7_H));
  PCK(cudaMemcpyAsync(P->dtcr7, SC_TCR7_H, sizeof SC_TCR7_H, cudaMemcpyHostToDevice, ctx->stream));
  PCK(cudaMemcpyAsync(P->dtcr4, T4, sizeof T4, cudaMemcpyHostToDevice, ctx->stream));
  PCK(cudaMemcpyAsync(P->dtcr9, SC_TCR9_H, sizeof SC_TCR9_H, cudaMemcpyHostToDevice, ctx->stream));
  PCK(cudaMemcpyAsync(P->dtcr20, SC_TCR20_H, sizeof SC_TCR20_H, cudaMemcpyHostToDevice, ctx->stream));
  PCK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
template <class T> static int grow(scuffcuda_context *ctx, T *&p, size_t &cap, size_t need)
{ if (need <= cap) return 0;
  if (p) cudaFree(p); p = nullptr; cap = 0;
  PCK(cudaMalloc(&p, need * sizeof(T))); cap = need; return 0;
}

static void host_edge(const Surface &S, int ne, EdgeRec &E)
{ const int *t = &S.edgeTopo[8 * (size_t)ne];
  for (int m = 0; m < 3; m++)
   { E.V1[m] = S.V[3 * (size_t)t[0] + m]; E.V2[m] = S.V[3 * (size_t)t[1] + m]; E.QP[m] = S.V[3 * (size_t)t[2] + m];
     E.QM[m] = t[3] >= 0 ? S.V[3 * (size_t)t[3] + m] : 0.0; }
  E.len = S.edgeLen[ne]; E.hasM = (t[3] >= 0 && t[5] >= 0); E.pad = 0;
  E.pP = t[4]; E.iP = t[6]; E.pM = t[5]; E.iM = t[7];
}
static int ensure_edges(scuffcuda_context *ctx, Surface &S)
{ if (S.dedges && S.edgesGen == S.generation && S.generation != 0) return 0;
  if ((int)S.edgeLen.size() != S.NE) return pfail(ctx, "rhs/pft: surface has no edge table");
  std::vector<EdgeRec> E(S.NE);
  for (int ne = 0; ne < S.NE; ne++) host_edge(S, ne, E[ne]);
  if (S.dedges) { cudaFree(S.dedges); S.dedges = nullptr; }
  PCK(cudaMalloc(&S.dedges, std::max<size_t>(1, S.NE) * sizeof(EdgeRec)));
  PCK(cudaMemcpyAsync(S.dedges, E.data(), (size_t)S.NE * sizeof(EdgeRec), cudaMemcpyHostToDevice, ctx->stream));
  PCK(cudaStreamSynchronize(ctx->stream));
  ctx->h2dBytes += (unsigned long long)S.NE * sizeof(EdgeRec);
  S.edgesGen = S.generation;
  return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// right-hand side
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ cd cexp_i(cd z)    // exp(i z)
{ double s, c; sincos(z.x, &s, &c); const double e = (z.y == 0.0) ? 1.0 : exp(-z.y); return mk(e * c, e * s); }

// sum of the plane waves' fields at X, each with its sign (PlaneWave::GetFields, libIncField/PlaneWave.cc:75-88)
__device__ __forceinline__ void planewave_fields(const PlaneWaveDev *pw, int npw, const double X[3], cd E[3], cd H[3])
{ for (int m = 0; m < 3; m++) { E[m] = mk(0.0, 0.0); H[m] = mk(0.0, 0.0); }
  for (int i = 0; i < npw; i++)
   { const PlaneWaveDev &W = pw[i];
     const double nx = W.n[0] * X[0] + W.n[1] * X[1] + W.n[2] * X[2];
     const cd ef = cexp_i(W.K * nx) * W.sign;
     cd e[3]; for (int m = 0; m < 3; m++) e[m] = W.E0[m] * ef;
     for (int m = 0; m < 3; m++) E[m] += e[m];
     H[0] += (W.n[1] * e[2] - W.n[2] * e[1]) * W.invZ;
     H[1] += (W.n[2] * e[0] - W.n[0] * e[2]) * W.invZ;
     H[2] += (W.n[0] * e[1] - W.n[1] * e[0]) * W.invZ; }
}

struct RhsArgs
{ const EdgeRec *edges; int NE; const double *tcr20; int npw; PlaneWaveDev pw[PFT_MAXPW];
  const double *tab;              // 12 doubles per point ((ne*2 + pm)*78 + np), already signed and summed; or null
  double *vec; int isReal; int offset; int isPEC;
};

__global__ void __launch_bounds__(128) rhs_points_kernel(const EdgeRec *edges, int NE, const double *tcr20, double *X)
{ const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)NE * 2 * SC_TCR20_NPTS) return;
  const int np = (int)(i % SC_TCR20_NPTS), pm = (int)((i / SC_TCR20_NPTS) & 1), ne = (int)(i / (2 * SC_TCR20_NPTS));
  const EdgeRec &E = edges[ne];
  const double u = tcr20[3 * np], v = tcr20[3 * np + 1];
  const double *A0 = pm ? E.V1 : E.QP, *A1 = pm ? E.V2 : E.V1, *A2 = pm ? E.QM : E.V2;
  for (int m = 0; m < 3; m++) X[3 * i + m] = (pm && !E.hasM) ? E.QP[m] : A0[m] + u * (A1[m] - A0[m]) + v * (A2[m] - A0[m]);
}

// one warp per edge, lanes over the 2 x 78 cubature points
__global__ void __launch_bounds__(128) rhs_kernel(const __grid_constant__ RhsArgs A)
{ const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= A.NE) return;
  const EdgeRec &E = A.edges[warp];
  double acc[4] = { 0.0, 0.0, 0.0, 0.0 };      // Re/Im <f,E>, Re/Im <f,H>
  for (int idx = lane; idx < 2 * SC_TCR20_NPTS; idx += 32)
   { const int pm = idx / SC_TCR20_NPTS, np = idx - pm * SC_TCR20_NPTS;
     if (pm && !E.hasM) continue;
     const double u = A.tcr20[3 * np], v = A.tcr20[3 * np + 1], w = A.tcr20[3 * np + 2];
     // TriIntFixed(.., QP, V1, V2) on the positive panel, TriIntFixed(.., V1, V2, QM) on the negative one
     const double *A0 = pm ? E.V1 : E.QP, *A1 = pm ? E.V2 : E.V1, *A2 = pm ? E.QM : E.V2, *Q = pm ? E.QM : E.QP;
     double X[3], f[3];
     for (int m = 0; m < 3; m++) { X[m] = A0[m] + u * (A1[m] - A0[m]) + v * (A2[m] - A0[m]); f[m] = X[m] - Q[m]; }
     cd Ef[3], Hf[3];
     planewave_fields(A.pw, A.npw, X, Ef, Hf);
     if (A.tab)
      { const double *t = A.tab + 12 * ((size_t)(warp * 2 + pm) * SC_TCR20_NPTS + np);
        for (int m = 0; m < 3; m++) { Ef[m] += mk(t[2 * m], t[2 * m + 1]); Hf[m] += mk(t[6 + 2 * m], t[7 + 2 * m]); } }
     // PreFac * Jacobian * w = (l / 2A)(2A) w = l w; the negative panel enters with a minus sign
     const double s = (pm ? -w : w) * E.len;
     acc[0] += s * (f[0] * Ef[0].x + f[1] * Ef[1].x + f[2] * Ef[2].x);
     acc[1] += s * (f[0] * Ef[0].y + f[1] * Ef[1].y + f[2] * Ef[2].y);
     acc[2] += s * (f[0] * Hf[0].x + f[1] * Hf[1].x + f[2] * Hf[2].x);
     acc[3] += s * (f[0] * Hf[0].y + f[1] * Hf[1].y + f[2] * Hf[2].y); }
  for (int q = 0; q < 4; q++)
   for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
  if (lane) return;
  const int row = A.offset + (A.isPEC ? warp : 2 * warp);
  if (A.isReal)
   { A.vec[row] = acc[0] / ZVAC; if (!A.isPEC) A.vec[row + 1] = acc[2]; }
  else
   { A.vec[2 * row] = acc[0] / ZVAC; A.vec[2 * row + 1] = acc[1] / ZVAC;
     if (!A.isPEC) { A.vec[2 * row + 2] = acc[2]; A.vec[2 * row + 3] = acc[3]; } }
}

static int make_pw(scuffcuda_context *ctx, const double omega[2], int n, const scuffcuda_planewave *pw, PlaneWaveDev *out)
{ if (n < 0 || n > PFT_MAXPW) return pfail(ctx, "rhs/pft: at most 8 plane waves per call");
  typedef std::complex<double> cx;
  for (int i = 0; i < n; i++)
   { const cx Om(omega[0], omega[1]), eps(pw[i].eps[0], pw[i].eps[1]), mu(pw[i].mu[0], pw[i].mu[1]);
     const cx K = std::sqrt(eps * mu) * Om, Z = ZVAC * std::sqrt(mu / eps), iZ = 1.0 / Z;
     for (int m = 0; m < 3; m++) { out[i].E0[m] = mk(pw[i].e0[2 * m], pw[i].e0[2 * m + 1]); out[i].n[m] = pw[i].nhat[m]; }
     out[i].K = mk(K.real(), K.imag()); out[i].invZ = mk(iZ.real(), iZ.imag()); out[i].sign = pw[i].sign; }
  return 0;
}

static Matrix *find_matrix(scuffcuda_context *ctx, int h)
{ auto it = ctx->matrices.find(h); return it == ctx->matrices.end() ? nullptr : &it->second; }

// ------------------------------------------------------------------------------------------------------------------
// power / force / torque
// ------------------------------------------------------------------------------------------------------------------
struct PftParams
{ double k, omega;             // real wavenumber in the region, real frequency
  double zmu, eoz;             // ZVAC * MuRel, EpsRel / ZVAC
  double ook2;                 // 1 / k^2
  double origin[3];            // torque centre of the surface the force acts on (RWGSurface::Origin)
};

// surface currents at the cubature points of every panel of one surface (panel vertex order: GetPanelCubature,
// PanelCubature.cc:155-236; coefficients as GetKNCoefficients reads them, RWGGeometry.cc:957-980)
__global__ void __launch_bounds__(128)
pft_currents_kernel(const PanelRec *panels, int NP, const double *tcr, int npts, const double *kn, int offset, int isPEC, int useN, CurPt *out)
{ const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NP * npts) return;
  const int np = i / npts, p = i - np * npts;
  const PanelRec &P = panels[np];
  const double u = tcr[3 * p], v = tcr[3 * p + 1], w = tcr[3 * p + 2];
  double A[3], B[3], X[3];
  for (int m = 0; m < 3; m++) { A[m] = P.V[3 + m] - P.V[m]; B[m] = P.V[6 + m] - P.V[m]; X[m] = P.V[m] + u * A[m] + v * B[m]; }
  const double cx = A[1] * B[2] - A[2] * B[1], cy = A[2] * B[0] - A[0] * B[2], cz = A[0] * B[1] - A[1] * B[0];
  const double twoA = sqrt(cx * cx + cy * cy + cz * cz);
  CurPt C;
  for (int m = 0; m < 3; m++) { C.x[m] = X[m]; C.Kr[m] = C.Ki[m] = C.Nr[m] = C.Ni[m] = 0.0; }
  C.dKr = C.dKi = C.dNr = C.dNi = 0.0; C.w = twoA * w;
  for (int e = 0; e < 3; e++)
   { const int ne = P.ei[e]; if (ne < 0) continue;
     const double pre = P.sl[e] / twoA;                       // +-l/(2A)
     const int row = offset + (isPEC ? ne : 2 * ne);
     const double kr = kn[2 * row], ki = kn[2 * row + 1];
     double nr = 0.0, ni = 0.0;
     if (!isPEC && useN) { nr = -ZVAC * kn[2 * row + 2]; ni = -ZVAC * kn[2 * row + 3]; }
     for (int m = 0; m < 3; m++)
      { const double f = pre * (X[m] - P.V[3 * e + m]);
        C.Kr[m] += kr * f; C.Ki[m] += ki * f; C.Nr[m] += nr * f; C.Ni[m] += ni * f; }
     C.dKr += 2.0 * pre * kr; C.dKi += 2.0 * pre * ki; C.dNr += 2.0 * pre * nr; C.dNi += 2.0 * pre * ni; }
  out[i] = C;
}

__device__ __forceinline__ cd cmul_helper(cd a, cd b, double s) { const cd p = a * b; return mk(p.x * s, p.y * s); }

// complex Phi, Psi, Zeta of ScatteredPFTIntegrand1 for two different surfaces (EMTPFT.cc:122-128); the same-surface
// (imaginary parts only) variant is inlined in pft_point_pair
__device__ __forceinline__ void pft_kernels_complex(double k, double r2, cd &Phi, cd &Psi, cd &Zeta)
{ const double rinv = rsqrt_fast(r2), r = r2 * rinv, kr = k * r, kr2 = kr * kr;
  double s, c; sincos_fast(kr, s, c);
  const double g = rinv * (1.0 / (4.0 * M_PI)), rinv2 = rinv * rinv;
  Phi = mk(c * g, s * g);
  Psi = cmul_helper(Phi, mk(-1.0, kr), rinv2);
  Zeta = cmul_helper(Phi, mk(3.0 - kr2, -3.0 * kr), rinv2 * rinv2);
}

// one cubature point pair: a* = conjugated currents at the point of surface a, b = currents at the point of surface b;
// acc[0] PScat, acc[1..3] force density, acc[4..6] first torque term — all without the weights of point a.
// HASN: both surfaces carry magnetic currents (if either is PEC the reference drops every N bilinear, GetPFT.cc:103-109).
// With S1p = ZVAC mu (K*.K' - div K* div K'/k^2) + (eps/ZVAC)(N*.N' - ...), V1 = ZVAC mu K* x K' + (eps/ZVAC) N* x N',
// C = K* x N' - N* x K', PM = C.R, D = (K*.R) N' - (N*.R) K' - R (K*.N' - N*.K'), the integrands of EMTPFT.cc:131-175 are
//   PScat  = -Re( i w S1p Phi + PM Psi )                 = w Im(S1p Phi) - Re(PM Psi)
//   F_m    = -Im( i S1p R_m Psi - (C_m Psi + PM R_m Zeta)/w ) = -R_m Re(S1p Psi) + Im(C_m Psi + PM R_m Zeta)/w
//   T1_m   = -Im( i V1_m Phi - D_m Psi/w )               = -Re(V1_m Phi) + Im(D_m Psi)/w
// SAME (one surface): Phi = i phi, Psi = i psi, Zeta = i zeta are purely imaginary (:103-121), so only
//   PScat = w phi Re S1p + psi Im PM,  F_m = R_m psi Im S1p + (psi Re C_m + R_m zeta Re PM)/w,  T1_m = phi Im V1_m + psi Re D_m/w
// are formed (half of the complex arithmetic; no Zeta at all without magnetic currents).
__device__ __forceinline__ double cre(double ar, double ai, double br, double bi) { return ar * br - ai * bi; }
__device__ __forceinline__ double cim(double ar, double ai, double br, double bi) { return ar * bi + ai * br; }

template <bool HASN, bool SAME>
__device__ __forceinline__ void pft_point_pair(const PftParams &Q, const double xa[3],
                                               const double aKr[3], const double aKi[3], const double aNr[3], const double aNi[3],
                                               double adKr, double adKi, double adNr, double adNi,
                                               const CurPt &b, double acc[7])
{ double R[3]; for (int m = 0; m < 3; m++) R[m] = xa[m] - b.x[m];
  const double r2 = R[0] * R[0] + R[1] * R[1] + R[2] * R[2];
  const double w = b.w, ft = TENTHIRDS * w, oow = 1.0 / Q.omega;
  if (SAME)
   { // imaginary parts of the kernels with the reference's small-argument series
     // r = r2/sqrt(r2) and 1/r from one branch-free rsqrt, sin/cos by the far kernel's branch-free FP64 routines (fastmath.cuh);
     // the reference's expressions (kr cos kr - sin kr) etc. are kept as they are (their cancellation is part of its answer)
     const double k = Q.k, rinv = rsqrt_fast(fmax(r2, 1.0e-300)), r = r2 * rinv, kr = k * r, kr2 = kr * kr, k2 = k * k;
     double phi, psi, zeta = 0.0;
     if (fabs(kr) < 1.0e-3)
      { const double k3 = k2 * k;
        phi = (1.0 - kr2 / 6.0) * k / (4.0 * M_PI);
        psi = -(1.0 - kr2 / 10.0) * k3 / (12.0 * M_PI);
        if (HASN) zeta = (1.0 - kr2 / 14.0) * (k3 * k2) / (60.0 * M_PI); }
     else
      { double sn, cs; sincos_fast(kr, sn, cs);
        const double g = rinv * (1.0 / (4.0 * M_PI)), rinv2 = rinv * rinv, g3 = g * rinv2;
        phi = sn * g;
        psi = (kr * cs - sn) * g3;
        if (HASN) zeta = (-3.0 * kr * cs + (3.0 - kr2) * sn) * (g3 * rinv2); }
     double kkr = 0, kki = 0, nnr = 0, nni = 0, knr = 0, nkr = 0, aRr = 0, nRr = 0, aRi = 0, nRi = 0, kni = 0, nki = 0;
     for (int m = 0; m < 3; m++)
      { kkr += cre(aKr[m], aKi[m], b.Kr[m], b.Ki[m]); kki += cim(aKr[m], aKi[m], b.Kr[m], b.Ki[m]);
        if (HASN)
         { nnr += cre(aNr[m], aNi[m], b.Nr[m], b.Ni[m]); nni += cim(aNr[m], aNi[m], b.Nr[m], b.Ni[m]);
           knr += cre(aKr[m], aKi[m], b.Nr[m], b.Ni[m]); kni += cim(aKr[m], aKi[m], b.Nr[m], b.Ni[m]);
           nkr += cre(aNr[m], aNi[m], b.Kr[m], b.Ki[m]); nki += cim(aNr[m], aNi[m], b.Kr[m], b.Ki[m]);
           aRr += aKr[m] * R[m]; aRi += aKi[m] * R[m]; nRr += aNr[m] * R[m]; nRi += aNi[m] * R[m]; } }
     double s1r = (kkr - cre(adKr, adKi, b.dKr, b.dKi) * Q.ook2) * Q.zmu, s1i = (kki - cim(adKr, adKi, b.dKr, b.dKi) * Q.ook2) * Q.zmu;
     if (HASN) { s1r += (nnr - cre(adNr, adNi, b.dNr, b.dNi) * Q.ook2) * Q.eoz; s1i += (nni - cim(adNr, adNi, b.dNr, b.dNi) * Q.ook2) * Q.eoz; }
     double Cr[3], Ci[3], V1i[3], Dr[3];
     for (int m = 0; m < 3; m++)
      { const int m1 = (m + 1) % 3, m2 = (m + 2) % 3;
        V1i[m] = (cim(aKr[m1], aKi[m1], b.Kr[m2], b.Ki[m2]) - cim(aKr[m2], aKi[m2], b.Kr[m1], b.Ki[m1])) * Q.zmu;
        if (HASN)
         { V1i[m] += (cim(aNr[m1], aNi[m1], b.Nr[m2], b.Ni[m2]) - cim(aNr[m2], aNi[m2], b.Nr[m1], b.Ni[m1])) * Q.eoz;
           Cr[m] = (cre(aKr[m1], aKi[m1], b.Nr[m2], b.Ni[m2]) - cre(aKr[m2], aKi[m2], b.Nr[m1], b.Ni[m1]))
                 - (cre(aNr[m1], aNi[m1], b.Kr[m2], b.Ki[m2]) - cre(aNr[m2], aNi[m2], b.Kr[m1], b.Ki[m1]));
           Ci[m] = (cim(aKr[m1], aKi[m1], b.Nr[m2], b.Ni[m2]) - cim(aKr[m2], aKi[m2], b.Nr[m1], b.Ni[m1]))
                 - (cim(aNr[m1], aNi[m1], b.Kr[m2], b.Ki[m2]) - cim(aNr[m2], aNi[m2], b.Kr[m1], b.Ki[m1]));
           Dr[m] = cre(aRr, aRi, b.Nr[m], b.Ni[m]) - cre(nRr, nRi, b.Kr[m], b.Ki[m]) - (knr - nkr) * R[m]; } }
     double p = Q.omega * phi * s1r, pmr = 0.0;
     if (HASN) { pmr = Cr[0] * R[0] + Cr[1] * R[1] + Cr[2] * R[2]; p += psi * (Ci[0] * R[0] + Ci[1] * R[1] + Ci[2] * R[2]); }
     acc[0] += w * p;
     const double psis1 = psi * s1i;
     for (int m = 0; m < 3; m++)
      { double f = R[m] * psis1, t = phi * V1i[m];
        if (HASN) { f += (psi * Cr[m] + R[m] * zeta * pmr) * oow; t += psi * Dr[m] * oow; }
        acc[1 + m] += ft * f; acc[4 + m] += ft * t; }
     (void)kni; (void)nki;
     return; }
  cd Phi, Psi, Zeta;
  pft_kernels_complex(Q.k, r2, Phi, Psi, Zeta);
  // conj(K_a) = (aKr, -aKi) etc. are passed already conjugated
  cd kk = mk(0.0, 0.0), nn = mk(0.0, 0.0), kn = mk(0.0, 0.0), nk = mk(0.0, 0.0), aR = mk(0.0, 0.0), nR = mk(0.0, 0.0);
  for (int m = 0; m < 3; m++)
   { kk += mk(aKr[m], aKi[m]) * mk(b.Kr[m], b.Ki[m]);
     aR += mk(aKr[m] * R[m], aKi[m] * R[m]);
     if (HASN)
      { nn += mk(aNr[m], aNi[m]) * mk(b.Nr[m], b.Ni[m]);
        kn += mk(aKr[m], aKi[m]) * mk(b.Nr[m], b.Ni[m]);
        nk += mk(aNr[m], aNi[m]) * mk(b.Kr[m], b.Ki[m]);
        nR += mk(aNr[m] * R[m], aNi[m] * R[m]); } }
  const cd ddk = mk(adKr, adKi) * mk(b.dKr, b.dKi);
  cd S1p = (kk - ddk * Q.ook2) * Q.zmu;
  if (HASN) { const cd ddn = mk(adNr, adNi) * mk(b.dNr, b.dNi); S1p += (nn - ddn * Q.ook2) * Q.eoz; }
  // V1 = zmu (a x b) + eoz (n x m);  C = a x m - n x b
  cd V1[3], C[3], D[3];
  for (int m = 0; m < 3; m++)
   { const int m1 = (m + 1) % 3, m2 = (m + 2) % 3;
     const cd a1 = mk(aKr[m1], aKi[m1]), a2 = mk(aKr[m2], aKi[m2]), b1 = mk(b.Kr[m1], b.Ki[m1]), b2 = mk(b.Kr[m2], b.Ki[m2]);
     V1[m] = (a1 * b2 - a2 * b1) * Q.zmu;
     if (HASN)
      { const cd n1 = mk(aNr[m1], aNi[m1]), n2 = mk(aNr[m2], aNi[m2]), q1 = mk(b.Nr[m1], b.Ni[m1]), q2 = mk(b.Nr[m2], b.Ni[m2]);
        V1[m] += (n1 * q2 - n2 * q1) * Q.eoz;
        C[m] = (a1 * q2 - a2 * q1) - (n1 * b2 - n2 * b1);
        D[m] = aR * mk(b.Nr[m], b.Ni[m]) - nR * mk(b.Kr[m], b.Ki[m]) - (kn - nk) * R[m]; }
     else { C[m] = mk(0.0, 0.0); D[m] = mk(0.0, 0.0); } }
  const cd sPhi = S1p * Phi, sPsi = S1p * Psi;
  double p = Q.omega * sPhi.y;
  cd PM = mk(0.0, 0.0);
  if (HASN) { PM = C[0] * R[0] + C[1] * R[1] + C[2] * R[2]; p -= (PM * Psi).x; }
  acc[0] += w * p;
  const cd pmZ = PM * Zeta;
  for (int m = 0; m < 3; m++)
   { double f = -R[m] * sPsi.x;
     double t = -(V1[m] * Phi).x;
     if (HASN) { f += ((C[m] * Psi).y + pmZ.y * R[m]) * oow; t += (D[m] * Psi).y * oow; }
     acc[1 + m] += ft * f; acc[4 + m] += ft * t; }
}

#define PFT_TILE 64
#define PFT_THREADS 128

struct PairArgs
{ const CurPt *A; int nA; const CurPt *B; int nB; int bChunk;    // blockIdx.y covers points [y*bChunk, (y+1)*bChunk) of B
  PftParams Q; double *part;                                      // PFT_NQ doubles per CTA
};

__device__ __forceinline__ void block_reduce_store(double v[PFT_NQ], double *out)
{ __shared__ double red[PFT_THREADS / 32][PFT_NQ];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int q = 0; q < PFT_NQ; q++)
   for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
  if (lane == 0) for (int q = 0; q < PFT_NQ; q++) red[warp][q] = v[q];
  __syncthreads();
  if (threadIdx.x < PFT_NQ)
   { double s = 0.0; for (int wq = 0; wq < PFT_THREADS / 32; wq++) s += red[wq][threadIdx.x];
     out[threadIdx.x] = s; }
}

template <bool HASN, bool SAME>
__global__ void __launch_bounds__(PFT_THREADS) pft_pair_kernel(const __grid_constant__ PairArgs X)
{ __shared__ CurPt tile[PFT_TILE];
  const int ia = blockIdx.x * PFT_THREADS + threadIdx.x;
  const bool live = ia < X.nA;
  CurPt a = X.A[live ? ia : 0];
  double aKi[3], aNi[3];
  for (int m = 0; m < 3; m++) { aKi[m] = -a.Ki[m]; aNi[m] = -a.Ni[m]; }
  double acc[7] = { 0, 0, 0, 0, 0, 0, 0 };
  const int b0 = blockIdx.y * X.bChunk, b1 = min(X.nB, b0 + X.bChunk);
  for (int t0 = b0; t0 < b1; t0 += PFT_TILE)
   { const int nt = min(PFT_TILE, b1 - t0);
     __syncthreads();
     // cooperative copy of the tile (20 doubles per point)
     { const double *src = (const double *)(X.B + t0); double *dst = (double *)tile;
       for (int i = threadIdx.x; i < nt * 20; i += PFT_THREADS) dst[i] = src[i]; }
     __syncthreads();
     if (live)
      for (int j = 0; j < nt; j++)
       pft_point_pair<HASN, SAME>(X.Q, a.x, a.Kr, aKi, a.Nr, aNi, a.dKr, -a.dKi, a.dNr, -a.dNi, tile[j], acc);
   }
  double v[PFT_NQ];
  const double wa = live ? a.w : 0.0;
  for (int q = 0; q < 7; q++) v[q] = wa * acc[q];
  // second torque term: (x_a - origin) x force density (EMTPFT.cc:166-168)
  const double xt[3] = { a.x[0] - X.Q.origin[0], a.x[1] - X.Q.origin[1], a.x[2] - X.Q.origin[2] };
  v[7] = xt[1] * v[3] - xt[2] * v[2]; v[8] = xt[2] * v[1] - xt[0] * v[3]; v[9] = xt[0] * v[2] - xt[1] * v[1];
  block_reduce_store(v, X.part + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * PFT_NQ);
}

// basis-function pairs that share a vertex: (order 9) - (order 4) of the pair's four panel combinations; one warp per pair
struct NearPftArgs
{ const NearBF *list; int n; const EdgeRec *ea, *eb; const double *kn; int offA, offB, pecA, pecB;
  const double *tcr9, *tcr4; PftParams Q; double *part; };

__device__ __forceinline__ void bf_point(const EdgeRec &E, int pm, double u, double v, double X[3], double f[3], double &div, double &jac)
{ const double *Q = pm ? E.QM : E.QP;
  double A[3], B[3];
  for (int m = 0; m < 3; m++) { A[m] = E.V1[m] - Q[m]; B[m] = E.V2[m] - Q[m]; }
  const double cx = A[1] * B[2] - A[2] * B[1], cy = A[2] * B[0] - A[0] * B[2], cz = A[0] * B[1] - A[1] * B[0];
  jac = sqrt(cx * cx + cy * cy + cz * cz);                    // 2 A
  const double pre = (pm ? -E.len : E.len) / jac;             // +- l/(2A)
  for (int m = 0; m < 3; m++) { const double d = u * A[m] + v * B[m]; X[m] = d + Q[m]; f[m] = pre * d; }
  div = 2.0 * pre;
}

template <bool HASN, bool SAME>
__global__ void __launch_bounds__(PFT_THREADS) pft_near_kernel(const __grid_constant__ NearPftArgs X)
{ const int lane = threadIdx.x & 31;
  double v[PFT_NQ]; for (int q = 0; q < PFT_NQ; q++) v[q] = 0.0;
  // grid-stride over the pairs: a fixed number of CTAs (and partial sums), one warp per pair at a time
  for (int wid = blockIdx.x * (PFT_THREADS / 32) + (threadIdx.x >> 5); wid < X.n; wid += gridDim.x * (PFT_THREADS / 32))
   { const NearBF P = X.list[wid];
     const EdgeRec &EA = X.ea[P.ea], &EB = X.eb[P.eb];
     const int rowA = X.offA + (X.pecA ? P.ea : 2 * P.ea), rowB = X.offB + (X.pecB ? P.eb : 2 * P.eb);
     const double kar = X.kn[2 * rowA], kai = -X.kn[2 * rowA + 1];          // conj(k_a)
     const double kbr = X.kn[2 * rowB], kbi = X.kn[2 * rowB + 1];
     double nar = 0, nai = 0, nbr = 0, nbi = 0;
     if (HASN) { nar = -ZVAC * X.kn[2 * rowA + 2]; nai = ZVAC * X.kn[2 * rowA + 3]; nbr = -ZVAC * X.kn[2 * rowB + 2]; nbi = -ZVAC * X.kn[2 * rowB + 3]; }
     const int nPMa = EA.hasM ? 2 : 1, nPMb = EB.hasM ? 2 : 1;
     for (int pass = 0; pass < 2; pass++)
      { const double *tcr = pass ? X.tcr4 : X.tcr9; const int npts = pass ? 6 : SC_TCR9_NPTS; const double sgn = pass ? -1.0 : 1.0;
        const int total = nPMa * nPMb * npts * npts;
        for (int idx = lane; idx < total; idx += 32)
         { int t = idx; const int pb = t % npts; t /= npts; const int pa = t % npts; t /= npts; const int pmb = t % nPMb; const int pma = t / nPMb;
           double xa[3], fa[3], da, ja, xb[3], fb[3], db, jb;
           bf_point(EA, pma, tcr[3 * pa], tcr[3 * pa + 1], xa, fa, da, ja);
           bf_point(EB, pmb, tcr[3 * pb], tcr[3 * pb + 1], xb, fb, db, jb);
           CurPt b; double aKr[3], aKi[3], aNr[3], aNi[3];
           for (int m = 0; m < 3; m++)
            { b.x[m] = xb[m]; b.Kr[m] = kbr * fb[m]; b.Ki[m] = kbi * fb[m]; b.Nr[m] = nbr * fb[m]; b.Ni[m] = nbi * fb[m];
              aKr[m] = kar * fa[m]; aKi[m] = kai * fa[m]; aNr[m] = nar * fa[m]; aNi[m] = nai * fa[m]; }
           b.dKr = kbr * db; b.dKi = kbi * db; b.dNr = nbr * db; b.dNi = nbi * db;
           b.w = jb * tcr[3 * pb + 2];
           double acc[7] = { 0, 0, 0, 0, 0, 0, 0 };
           pft_point_pair<HASN, SAME>(X.Q, xa, aKr, aKi, aNr, aNi, kar * da, kai * da, nar * da, nai * da, b, acc);
           const double wa = sgn * ja * tcr[3 * pa + 2];
           for (int q = 0; q < 7; q++) { acc[q] *= wa; v[q] += acc[q]; }
           const double xt[3] = { xa[0] - X.Q.origin[0], xa[1] - X.Q.origin[1], xa[2] - X.Q.origin[2] };
           v[7] += xt[1] * acc[3] - xt[2] * acc[2]; v[8] += xt[2] * acc[1] - xt[0] * acc[3]; v[9] += xt[0] * acc[2] - xt[1] * acc[1]; } } }
  block_reduce_store(v, X.part + (size_t)blockIdx.x * PFT_NQ);
}

// extinction (GetExtinctionPFTT, EMTPFT.cc:389-531): one thread per 21-point-rule point of a panel
struct ExtArgs
{ const CurPt *A; int nA; PlaneWaveDev pw; int havePW; const double *tab;   // tab: 48 doubles per point (EH[6], dEH[3][6]) or null
  double origin[3]; double *part; };

__global__ void __launch_bounds__(PFT_THREADS) pft_ext_kernel(const __grid_constant__ ExtArgs X)
{ const int i = blockIdx.x * PFT_THREADS + threadIdx.x;
  double v[PFT_NQ]; for (int q = 0; q < PFT_NQ; q++) v[q] = 0.0;
  if (i < X.nA)
   { const CurPt a = X.A[i];
     cd EH[6], dEH[3][6];
     for (int q = 0; q < 6; q++) { EH[q] = mk(0.0, 0.0); for (int m = 0; m < 3; m++) dEH[m][q] = mk(0.0, 0.0); }
     if (X.havePW)
      { cd E[3], H[3]; planewave_fields(&X.pw, 1, a.x, E, H);
        for (int m = 0; m < 3; m++) { EH[m] = E[m]; EH[3 + m] = H[m]; }
        // dEH[i][j] = i K nHat[i] EH[j] (PlaneWave::GetFieldGradients, PlaneWave.cc:94-103)
        const cd iK = mk(-X.pw.K.y, X.pw.K.x);
        for (int m = 0; m < 3; m++) for (int q = 0; q < 6; q++) dEH[m][q] = iK * EH[q] * X.pw.n[m]; }
     if (X.tab)
      { const double *t = X.tab + 48 * (size_t)i;
        for (int q = 0; q < 6; q++) EH[q] += mk(t[2 * q], t[2 * q + 1]);
        for (int m = 0; m < 3; m++) for (int q = 0; q < 6; q++) dEH[m][q] += mk(t[12 + 12 * m + 2 * q], t[13 + 12 * m + 2 * q]); }
     cd Ks[3], Ns[3];                     // conj(K), conj(N)
     for (int m = 0; m < 3; m++) { Ks[m] = mk(a.Kr[m], -a.Ki[m]); Ns[m] = mk(a.Nr[m], -a.Ni[m]); }
     cd pabs = mk(0.0, 0.0);
     for (int m = 0; m < 3; m++) pabs += Ks[m] * EH[m] + Ns[m] * EH[3 + m];
     v[0] = a.w * pabs.x;
     double F[3];
     for (int mu = 0; mu < 3; mu++)
      { cd f = mk(0.0, 0.0);
        for (int m = 0; m < 3; m++) f += Ks[m] * dEH[mu][m] + Ns[m] * dEH[mu][3 + m];
        F[mu] = a.w * f.y; v[1 + mu] = F[mu];
        const int m1 = (mu + 1) % 3, m2 = (mu + 2) % 3;
        const cd t = (Ks[m1] * EH[m2] - Ks[m2] * EH[m1]) + (Ns[m1] * EH[3 + m2] - Ns[m2] * EH[3 + m1]);
        v[4 + mu] = a.w * t.y; }
     const double xt[3] = { a.x[0] - X.origin[0], a.x[1] - X.origin[1], a.x[2] - X.origin[2] };
     v[7] = xt[1] * F[2] - xt[2] * F[1]; v[8] = xt[2] * F[0] - xt[0] * F[2]; v[9] = xt[0] * F[1] - xt[1] * F[0]; }
  block_reduce_store(v, X.part + (size_t)blockIdx.x * PFT_NQ);
}

// ------------------------------------------------------------------------------------------------------------------
// scattered fields at evaluation points (GetRFMatrix / GetFields, GetFields.cc:141-330, 336-395)
// ------------------------------------------------------------------------------------------------------------------
// One CTA per evaluation point, threads stride over the edges of one surface.  Per (edge, point): the reduced fields
//   g = int f G - div f grad G / k^2 ,   c = int f x grad G / (-ik)        (RFIntegrand :86-105, GetG CalcGC.cc:194-240)
// by the 16-point rule if the point is >= 4 basis-function radii from the edge centroid, the 78-point rule if >= 1
// (GetBFCubature2, PanelCubature.cc:713-787: points Q + u(V1-Q) + v(V2-Q) on both panels); closer points need the
// reference's desingularised near-field integrals, which are not on the device: they are counted and reported.
// KN != null: contract with the solution on the fly (E, H accumulate into F); RF != null: also store the reduced-field
// matrix entries (RFMatrix[bf, 6 nx + mu], :307-325).
struct FieldArgs
{ const EdgeRec *edges; int NE; const double *X; int NX; const double *tcr7, *tcr20;
  cd k, zrel; double sign;
  const PanelRec *panels; const double *tcr9;
  const double *kn; int offset, isPEC;
  double *F;                       // 12 doubles per point, accumulated
  double *RF; long ldRF; const int *colOf;   // RFMatrix (complex, column-major, leading dimension ldRF), column block 6*colOf[nx]
  unsigned long long *nearCount;
};

__global__ void __launch_bounds__(PFT_THREADS) fields_kernel(const __grid_constant__ FieldArgs A)
{ __shared__ double red[PFT_THREADS / 32][12];
  const int nx = blockIdx.x;
  const double X0[3] = { A.X[3 * nx], A.X[3 * nx + 1], A.X[3 * nx + 2] };
  const cd ik = mk(-A.k.y, A.k.x), k2 = A.k * A.k, ook2 = cinv(k2), oomik = cinv(mk(A.k.y, -A.k.x));   // 1/(-ik)
  const cd ikz = ik * A.sign;
  const cd EK = ikz * A.zrel * ZVAC, HK = -ikz, EN = -(ikz * ZVAC), HN = -cdiv(ikz, A.zrel);
  cd acc[6]; for (int q = 0; q < 6; q++) acc[q] = mk(0.0, 0.0);
  for (int ne = threadIdx.x; ne < A.NE; ne += PFT_THREADS)
   { const EdgeRec &E = A.edges[ne];
     double c[3], rad = 0.0, d2 = 0.0;
     for (int m = 0; m < 3; m++) { c[m] = (E.V1[m] + E.V2[m]) / 2.0; d2 += (X0[m] - c[m]) * (X0[m] - c[m]); }
     { double a = 0, b = 0, q = 0, qm = 0;
       for (int m = 0; m < 3; m++) { a += (E.QP[m] - c[m]) * (E.QP[m] - c[m]); b += (E.V1[m] - c[m]) * (E.V1[m] - c[m]); q += (E.V2[m] - c[m]) * (E.V2[m] - c[m]); qm += (E.QM[m] - c[m]) * (E.QM[m] - c[m]); }
       rad = fmax(sqrt(a), fmax(sqrt(b), sqrt(q))); if (E.hasM) rad = fmax(rad, sqrt(qm)); }
     const double rRel = sqrt(d2) / rad;
     cd GC[6]; for (int q = 0; q < 6; q++) GC[q] = mk(0.0, 0.0);
     if (rRel < 1.0)
      { // closer than one basis-function radius: closed-form static integrals + desingularised cubature (nearfield.cuh)
        atomicAdd(A.nearCount, 1ull);
        cd e[3], h[3];
        nf_reduced_fields_panel(A.panels[E.pP], E.iP, X0, A.k, A.tcr9, e, h);
        if (E.hasM)
         { cd eM[3], hM[3];
           nf_reduced_fields_panel(A.panels[E.pM], E.iM, X0, A.k, A.tcr9, eM, hM);
           for (int m = 0; m < 3; m++) { e[m] -= eM[m]; h[m] -= hM[m]; } }
        for (int m = 0; m < 3; m++) { GC[m] = e[m]; GC[3 + m] = h[m] * oomik; } }
     const double *tcr = rRel >= 4.0 ? A.tcr7 : A.tcr20; const int npts = rRel < 1.0 ? 0 : (rRel >= 4.0 ? SC_TCR7_NPTS : SC_TCR20_NPTS);
     for (int np = 0; np < npts; np++)
      { const double u = tcr[3 * np], v = tcr[3 * np + 1], w = tcr[3 * np + 2];
        for (int pm = 0; pm < (E.hasM ? 2 : 1); pm++)
         { double Xs[3], b[3], div, jac;
           bf_point(E, pm, u, v, Xs, b, div, jac);
           const double W = jac * w;
           double R[3]; for (int m = 0; m < 3; m++) R[m] = Xs[m] - X0[m];
           const double r2 = R[0] * R[0] + R[1] * R[1] + R[2] * R[2], r = sqrt(r2);
           // Phi = exp(ikr)/(4 pi r), Psi = Phi (ikr - 1)/r^2, dG = R Psi
           double sn, cs; sincos(A.k.x * r, &sn, &cs);
           const double e = (A.k.y == 0.0 ? 1.0 : exp(-A.k.y * r)) / (4.0 * M_PI * r);
           const cd Phi = mk(e * cs, e * sn);
           const cd Psi = cmul_helper(Phi, mk(ik.x * r - 1.0, ik.y * r), 1.0 / r2);
           const cd dco = Psi * ook2 * div;                 // Divb dG / k^2 = R (Divb Psi / k^2)
           for (int m = 0; m < 3; m++) GC[m] += (Phi * b[m] - dco * R[m]) * W;
           const double bxR[3] = { b[1] * R[2] - b[2] * R[1], b[2] * R[0] - b[0] * R[2], b[0] * R[1] - b[1] * R[0] };
           const cd pc = Psi * oomik * W;
           for (int m = 0; m < 3; m++) GC[3 + m] += pc * bxR[m]; } }
     const int row = A.offset + (A.isPEC ? ne : 2 * ne);
     cd rf[12];
     for (int m = 0; m < 3; m++) { rf[m] = EK * GC[m]; rf[3 + m] = HK * GC[3 + m]; rf[6 + m] = EN * GC[3 + m]; rf[9 + m] = HN * GC[m]; }
     if (A.RF)
      { double *col = A.RF + 2 * ((size_t)6 * A.colOf[nx] * A.ldRF);
        for (int q = 0; q < 6; q++)
         { double *p = col + 2 * ((size_t)q * A.ldRF + row);
           p[0] = rf[q].x; p[1] = rf[q].y;
           if (!A.isPEC) { p[2] = rf[6 + q].x; p[3] = rf[6 + q].y; } } }
     if (A.kn)
      { const cd kc = mk(A.kn[2 * row], A.kn[2 * row + 1]);
        for (int q = 0; q < 6; q++) acc[q] += kc * rf[q];
        if (!A.isPEC) { const cd nc = mk(A.kn[2 * row + 2], A.kn[2 * row + 3]); for (int q = 0; q < 6; q++) acc[q] += nc * rf[6 + q]; } } }
  if (!A.kn) return;
  double v[12]; for (int q = 0; q < 6; q++) { v[2 * q] = acc[q].x; v[2 * q + 1] = acc[q].y; }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int q = 0; q < 12; q++) for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
  if (lane == 0) for (int q = 0; q < 12; q++) red[warp][q] = v[q];
  __syncthreads();
  if (threadIdx.x < 12)
   { double sum = 0.0; for (int wq = 0; wq < PFT_THREADS / 32; wq++) sum += red[wq][threadIdx.x];
     A.F[12 * (size_t)nx + threadIdx.x] += sum; }
}

// dyadic Green's functions (GetDyadicGFs.cc:150-172): for every evaluation point nx the 3x3 contractions over the basis functions
//   GE[i][j] = sum_bf D[bf, 6 nx + i] S[bf, 6 nx + j] ,  GM[i][j] = sum_bf D[bf, 6 nx + 3 + i] S[bf, 6 nx + 3 + j]
// of the destination reduced-field matrix D with the LU-solved source reduced-field matrix S.  One CTA per point.
__global__ void __launch_bounds__(PFT_THREADS) dgf_contract_kernel(const double *D, const double *S, int NBF, double *out)
{ __shared__ double red[PFT_THREADS / 32][36];
  const int nx = blockIdx.x;
  const double *Dc = D + 2 * (size_t)6 * nx * NBF, *Sc = S + 2 * (size_t)6 * nx * NBF;
  double acc[36]; for (int q = 0; q < 36; q++) acc[q] = 0.0;
  for (int bf = threadIdx.x; bf < NBF; bf += PFT_THREADS)
   { cd d[6], sv[6];
     for (int c = 0; c < 6; c++) { d[c] = mk(Dc[2 * ((size_t)c * NBF + bf)], Dc[2 * ((size_t)c * NBF + bf) + 1]); sv[c] = mk(Sc[2 * ((size_t)c * NBF + bf)], Sc[2 * ((size_t)c * NBF + bf) + 1]); }
     for (int half = 0; half < 2; half++)
      for (int i = 0; i < 3; i++)
       for (int j = 0; j < 3; j++)
        { const cd p = d[3 * half + i] * sv[3 * half + j]; acc[2 * (9 * half + 3 * i + j)] += p.x; acc[2 * (9 * half + 3 * i + j) + 1] += p.y; } }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int q = 0; q < 36; q++) for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
  if (lane == 0) for (int q = 0; q < 36; q++) red[warp][q] = acc[q];
  __syncthreads();
  if (threadIdx.x < 36)
   { double sum = 0.0; for (int wq = 0; wq < PFT_THREADS / 32; wq++) sum += red[wq][threadIdx.x];
     out[36 * (size_t)nx + threadIdx.x] = sum; }
}

// ---- host side: near basis-function pairs of a surface pair (AssessBFPair, AssessPanelPair.cc:200-241) -----------
static int build_near_list(scuffcuda_context *ctx, int nsa, int nsb, PftNearList &L)
{ Surface &SA = ctx->surfaces[nsa], &SB = ctx->surfaces[nsb];
  if (L.d && L.genA == SA.generation && L.genB == SB.generation) return 0;
  auto radius = [](const Surface &S, int ne)
   { const int *t = &S.edgeTopo[8 * (size_t)ne]; double c[3], r = 0.0;
     for (int m = 0; m < 3; m++) c[m] = 0.5 * (S.V[3 * (size_t)t[0] + m] + S.V[3 * (size_t)t[1] + m]);
     for (int q = 0; q < 4; q++)
      { const int vi = t[q]; if (vi < 0) continue; double d = 0.0;
        for (int m = 0; m < 3; m++) { const double dd = S.V[3 * (size_t)vi + m] - c[m]; d += dd * dd; }
        r = fmax(r, sqrt(d)); }
     return r; };
  std::vector<double> ra(SA.NE), rb(SB.NE); double rmax = 0.0;
  for (int ne = 0; ne < SA.NE; ne++) { ra[ne] = radius(SA, ne); rmax = fmax(rmax, ra[ne]); }
  for (int ne = 0; ne < SB.NE; ne++) { rb[ne] = radius(SB, ne); rmax = fmax(rmax, rb[ne]); }
  const double tolMax = 3.0 * CVTHRESHOLD * rmax;
  // vertices of b sorted by x; edges of b touching each vertex of b
  std::vector<int> order(SB.NV); for (int i = 0; i < SB.NV; i++) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int p, int q) { return SB.V[3 * (size_t)p] < SB.V[3 * (size_t)q]; });
  std::vector<double> xs(SB.NV); for (int i = 0; i < SB.NV; i++) xs[i] = SB.V[3 * (size_t)order[i]];
  std::vector<std::vector<int>> vedges(SB.NV);
  for (int ne = 0; ne < SB.NE; ne++)
   { const int *t = &SB.edgeTopo[8 * (size_t)ne];
     for (int q = 0; q < 4; q++) if (t[q] >= 0 && (q < 3 || t[5] >= 0)) vedges[t[q]].push_back(ne); }
  std::vector<NearBF> list; std::vector<int> stamp(SB.NE, -1);
  for (int nea = 0; nea < SA.NE; nea++)
   { const int *ta = &SA.edgeTopo[8 * (size_t)nea];
     const int iVa[4] = { ta[2], ta[0], ta[1], ta[3] }; const int NVA = (ta[3] == -1 ? 3 : 4);
     for (int qa = 0; qa < NVA; qa++)
      { const double *va = &SA.V[3 * (size_t)iVa[qa]];
        auto lo = std::lower_bound(xs.begin(), xs.end(), va[0] - tolMax), hi = std::upper_bound(xs.begin(), xs.end(), va[0] + tolMax);
        for (auto it = lo; it != hi; ++it)
         { const int vb = order[it - xs.begin()];
           const double *pb = &SB.V[3 * (size_t)vb];
           if (fabs(va[0] - pb[0]) + fabs(va[1] - pb[1]) + fabs(va[2] - pb[2]) > tolMax) continue;
           for (int neb : vedges[vb])
            { if (stamp[neb] == nea) continue;
              stamp[neb] = nea;
              // the reference's count with the pair's own threshold
              const int *tb = &SB.edgeTopo[8 * (size_t)neb];
              const int iVb[4] = { tb[2], tb[0], tb[1], tb[3] }; const int NVB = (tb[3] == -1 ? 3 : 4);
              const double thr = 3.0 * fabs(CVTHRESHOLD * fmin(ra[nea], rb[neb]));
              int ncv = 0;
              for (int i = 0; i < NVA; i++) for (int j = 0; j < NVB; j++)
               { const double *p = &SA.V[3 * (size_t)iVa[i]], *q = &SB.V[3 * (size_t)iVb[j]];
                 if (fabs(p[0] - q[0]) + fabs(p[1] - q[1]) + fabs(p[2] - q[2]) <= thr) ncv++; }
              if (ncv > 0) { NearBF P; P.ea = nea; P.eb = neb; list.push_back(P); } } } } }
  if (L.d) { cudaFree(L.d); L.d = nullptr; }
  L.n = (int)list.size();
  PCK(cudaMalloc(&L.d, std::max<size_t>(1, list.size()) * sizeof(NearBF)));
  if (L.n) PCK(cudaMemcpyAsync(L.d, list.data(), list.size() * sizeof(NearBF), cudaMemcpyHostToDevice, ctx->stream));
  PCK(cudaStreamSynchronize(ctx->stream));
  ctx->h2dBytes += (unsigned long long)list.size() * sizeof(NearBF);
  L.genA = SA.generation; L.genB = SB.generation;
  return 0;
}

static int sum_partials(scuffcuda_context *ctx, const double *dpart, size_t nblocks, double out[PFT_NQ])
{ std::vector<double> h(nblocks * PFT_NQ);
  PCK(cudaMemcpyAsync(h.data(), dpart, h.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  PCK(cudaStreamSynchronize(ctx->stream));
  ctx->d2hBytes += h.size() * sizeof(double);
  for (int q = 0; q < PFT_NQ; q++) { double s = 0.0; for (size_t b = 0; b < nblocks; b++) s += h[b * PFT_NQ + q]; out[q] = s; }
  return 0;
}

}  // namespace scuffcuda

using namespace scuffcuda;

extern "C" {

int scuffcuda_rhs_points(scuffcuda_context *ctx, int ns, double *X, long *npts)
{ if (ns < 0 || ns >= (int)ctx->surfaces.size()) return pfail(ctx, "rhs_points: invalid surface index");
  Surface &S = ctx->surfaces[ns];
  const long n = (long)S.NE * 2 * SC_TCR20_NPTS;
  if (npts) *npts = n;
  if (!X || n == 0) return 0;
  PCK(cudaSetDevice(ctx->device));
  if (pft_tables(ctx) || ensure_edges(ctx, S)) return 1;
  PftState *P = pft_state(ctx);
  if (grow(ctx, P->dtab, P->tabCap, (size_t)3 * n)) return 1;
  rhs_points_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>((const EdgeRec *)S.dedges, S.NE, P->dtcr20, P->dtab);
  ctx->launches++;
  PCK(cudaMemcpyAsync(X, P->dtab, (size_t)3 * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  PCK(cudaStreamSynchronize(ctx->stream));
  ctx->d2hBytes += (size_t)3 * n * sizeof(double);
  return 0;
}

int scuffcuda_assemble_rhs(scuffcuda_context *ctx, int ns, const double omega[2], int num_pw, const scuffcuda_planewave *pw,
                           const double *tabulated_eh, int vec_handle, int offset)
{ if (ns < 0 || ns >= (int)ctx->surfaces.size()) return pfail(ctx, "assemble_rhs: invalid surface index");
  Matrix *V = find_matrix(ctx, vec_handle);
  if (!V || V->NC != 1) return pfail(ctx, "assemble_rhs: invalid vector handle");
  Surface &S = ctx->surfaces[ns];
  const int nbf = S.isPEC ? S.NE : 2 * S.NE;
  if (offset < 0 || offset + nbf > V->NR) return pfail(ctx, "assemble_rhs: surface does not fit the vector");
  if (S.NE == 0) return 0;
  PCK(cudaSetDevice(ctx->device));
  if (pft_tables(ctx) || ensure_edges(ctx, S)) return 1;
  PftState *P = pft_state(ctx);
  RhsArgs A;
  A.edges = (const EdgeRec *)S.dedges; A.NE = S.NE; A.tcr20 = P->dtcr20; A.npw = num_pw;
  if (make_pw(ctx, omega, num_pw, pw, A.pw)) return 1;
  A.tab = nullptr;
  if (tabulated_eh)
   { const size_t n = (size_t)S.NE * 2 * SC_TCR20_NPTS * 12;
     if (grow(ctx, P->dtab, P->tabCap, n)) return 1;
     PCK(cudaMemcpyAsync(P->dtab, tabulated_eh, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
     ctx->h2dBytes += n * sizeof(double);
     A.tab = P->dtab; }
  A.vec = V->d; A.isReal = V->isReal; A.offset = offset; A.isPEC = S.isPEC;
  rhs_kernel<<<(S.NE + 3) / 4, 128, 0, ctx->stream>>>(A);
  ctx->launches++;
  PCK(cudaGetLastError());
  if (tabulated_eh) PCK(cudaStreamSynchronize(ctx->stream));    // the caller may reuse its host buffer
  return 0;
}

int scuffcuda_pft_points(scuffcuda_context *ctx, int ns, double *X, long *npts)
{ if (ns < 0 || ns >= (int)ctx->surfaces.size()) return pfail(ctx, "pft_points: invalid surface index");
  Surface &S = ctx->surfaces[ns];
  const long n = (long)S.NP * SC_TCR9_NPTS;
  if (npts) *npts = n;
  if (!X) return 0;
  // the 21-point rule in the panel's own vertex order (GetPanelCubature, PanelCubature.cc:155-236)
  for (int np = 0; np < S.NP; np++)
   { const PanelRec &P = S.hpanels[np];
     for (int p = 0; p < SC_TCR9_NPTS; p++)
      for (int m = 0; m < 3; m++)
       X[3 * ((size_t)np * SC_TCR9_NPTS + p) + m] = P.V[m] + SC_TCR9_H[3 * p] * (P.V[3 + m] - P.V[m]) + SC_TCR9_H[3 * p + 1] * (P.V[6 + m] - P.V[m]); }
  return 0;
}

int scuffcuda_emtpft(scuffcuda_context *ctx, int kn_handle, const scuffcuda_pft_args *a, double *scattered, double *extinction)
{ const int NS = (int)ctx->surfaces.size();
  if (!a || a->num_surfaces != NS) return pfail(ctx, "emtpft: num_surfaces does not match the device geometry");
  Matrix *KN = find_matrix(ctx, kn_handle);
  if (!KN || KN->NC != 1 || KN->isReal) return pfail(ctx, "emtpft: the solution vector must be a complex device vector");
  if (a->omega[1] != 0.0 || !(a->omega[0] > 0.0)) return pfail(ctx, "emtpft: power/force/torque needs a real positive frequency (EMTPFT.cc:86-89 takes real parts)");
  PCK(cudaSetDevice(ctx->device));
  if (pft_tables(ctx)) return 1;
  PftState *P = pft_state(ctx);
  const int which = a->interior ? 1 : 0;
  for (int i = 0; i < NS * NS * 11; i++) scattered[i] = 0.0;
  for (int i = 0; i < NS * 11; i++) extinction[i] = 0.0;
  // current points: order 4 (all surfaces), then order 9 (extinction)
  std::vector<size_t> off4(NS + 1, 0), off9(NS + 1, 0);
  for (int ns = 0; ns < NS; ns++) { off4[ns + 1] = off4[ns] + (size_t)ctx->surfaces[ns].NP * 6; off9[ns + 1] = off9[ns] + (size_t)ctx->surfaces[ns].NP * SC_TCR9_NPTS; }
  if (grow(ctx, P->dcur, P->curCap, 2 * off4[NS] + off9[NS])) return 1;
  for (int ns = 0; ns < NS; ns++)
   { Surface &S = ctx->surfaces[ns];
     const int nbf = S.isPEC ? S.NE : 2 * S.NE;
     if (a->bf_offsets[ns] < 0 || a->bf_offsets[ns] + nbf > KN->NR) return pfail(ctx, "emtpft: basis-function offsets do not fit the vector");
     if (ensure_edges(ctx, S)) return 1; }
  double extent = 0.0;
  { double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
    for (int ns = 0; ns < NS; ns++) { const Surface &S = ctx->surfaces[ns]; for (int v = 0; v < S.NV; v++) for (int m = 0; m < 3; m++) { lo[m] = fmin(lo[m], S.V[3 * (size_t)v + m]); hi[m] = fmax(hi[m], S.V[3 * (size_t)v + m]); } }
    for (int m = 0; m < 3; m++) extent += (hi[m] - lo[m]) * (hi[m] - lo[m]);
    extent = sqrt(extent); }
  CurPt *cur4 = P->dcur, *cur4k = P->dcur + off4[NS], *cur9 = P->dcur + 2 * off4[NS];
  for (int ns = 0; ns < NS; ns++)
   { Surface &S = ctx->surfaces[ns]; const int n4 = S.NP * 6, n9 = S.NP * SC_TCR9_NPTS;
     // with N (pairs of two dielectric surfaces, extinction) and without (pairs involving a PEC surface: GetKNBilinears drops N)
     pft_currents_kernel<<<(n4 + 127) / 128, 128, 0, ctx->stream>>>(S.dpanels, S.NP, P->dtcr4, 6, KN->d, a->bf_offsets[ns], S.isPEC, 1, cur4 + off4[ns]);
     pft_currents_kernel<<<(n4 + 127) / 128, 128, 0, ctx->stream>>>(S.dpanels, S.NP, P->dtcr4, 6, KN->d, a->bf_offsets[ns], S.isPEC, 0, cur4k + off4[ns]);
     pft_currents_kernel<<<(n9 + 127) / 128, 128, 0, ctx->stream>>>(S.dpanels, S.NP, P->dtcr9, SC_TCR9_NPTS, KN->d, a->bf_offsets[ns], S.isPEC, 1, cur9 + off9[ns]);
     ctx->launches += 3; }
  PCK(cudaGetLastError());
  // ---- scattered part: ordered surface pairs (sa, sb) gated as the reference's loop over nea <= neb does (EMTPFT.cc:640-668)
  for (int sa = 0; sa < NS; sa++)
   for (int sb = 0; sb < NS; sb++)
    { const int lo = std::min(sa, sb), hi = std::max(sa, sb);
      const int *rl = a->region_indices + 2 * lo, *rh = a->region_indices + 2 * hi;
      const int region = rl[which];
      if (region == -1) continue;                       // no interior PFT for PEC bodies
      double sign = 0.0;
      if (lo == hi) sign = 1.0;
      else if (rl[0] == rh[0]) sign = a->interior ? 0.0 : 1.0;
      else if (rl[0] == rh[1]) sign = a->interior ? 0.0 : -1.0;
      else if (rl[1] == rh[0]) sign = a->interior ? 1.0 : 0.0;
      if (sign == 0.0) continue;
      if (region < 0 || region >= a->num_regions) return pfail(ctx, "emtpft: invalid region index");
      Surface &SA = ctx->surfaces[sa], &SB = ctx->surfaces[sb];
      if (SA.NE == 0 || SB.NE == 0) continue;
      typedef std::complex<double> cx;
      const cx eps(a->eps_mu[4 * region], a->eps_mu[4 * region + 1]), mu(a->eps_mu[4 * region + 2], a->eps_mu[4 * region + 3]);
      const cx kc = a->omega[0] * std::sqrt(eps * mu);
      PftParams Q;
      if (fabs(kc.real()) * extent >= SCUFF_MAX_KR) return pfail(ctx, "emtpft: |k| x (extent of the geometry) exceeds the range of the device sin/cos (1e5)");
      Q.k = kc.real(); Q.omega = a->omega[0]; Q.zmu = ZVAC * mu.real(); Q.eoz = eps.real() / ZVAC; Q.ook2 = 1.0 / (Q.k * Q.k);
      for (int m = 0; m < 3; m++) Q.origin[m] = a->origins[3 * sa + m];
      const bool hasN = !SA.isPEC && !SB.isPEC, same = (sa == sb);
      const CurPt *CA = (hasN ? cur4 : cur4k) + off4[sa], *CB = (hasN ? cur4 : cur4k) + off4[sb];
      const int nA = SA.NP * 6, nB = SB.NP * 6;
      PftNearList &L = P->nearLists[std::make_pair(sa, sb)];
      if (build_near_list(ctx, sa, sb, L)) return 1;
      const int gx = (nA + PFT_THREADS - 1) / PFT_THREADS;
      int gy = std::max(1, std::min((4 * ctx->smCount + gx - 1) / gx, (nB + PFT_TILE - 1) / PFT_TILE));
      int chunk = (nB + gy - 1) / gy; chunk = ((chunk + PFT_TILE - 1) / PFT_TILE) * PFT_TILE; gy = (nB + chunk - 1) / chunk;
      const int nearBlocks = std::min((L.n + PFT_THREADS / 32 - 1) / (PFT_THREADS / 32), 8 * ctx->smCount);
      const size_t nblocks = (size_t)gx * gy + nearBlocks;
      if (grow(ctx, P->dpart, P->partCap, nblocks * PFT_NQ)) return 1;
      PairArgs X; X.A = CA; X.nA = nA; X.B = CB; X.nB = nB; X.bChunk = chunk; X.Q = Q; X.part = P->dpart;
      const dim3 grid(gx, gy);
      if (hasN) { if (same) pft_pair_kernel<true, true><<<grid, PFT_THREADS, 0, ctx->stream>>>(X); else pft_pair_kernel<true, false><<<grid, PFT_THREADS, 0, ctx->stream>>>(X); }
      else      { if (same) pft_pair_kernel<false, true><<<grid, PFT_THREADS, 0, ctx->stream>>>(X); else pft_pair_kernel<false, false><<<grid, PFT_THREADS, 0, ctx->stream>>>(X); }
      ctx->launches++;
      if (nearBlocks)
       { NearPftArgs Y; Y.list = L.d; Y.n = L.n; Y.ea = (const EdgeRec *)SA.dedges; Y.eb = (const EdgeRec *)SB.dedges; Y.kn = KN->d;
         Y.offA = a->bf_offsets[sa]; Y.offB = a->bf_offsets[sb]; Y.pecA = SA.isPEC; Y.pecB = SB.isPEC;
         Y.tcr9 = P->dtcr9; Y.tcr4 = P->dtcr4; Y.Q = Q; Y.part = P->dpart + (size_t)gx * gy * PFT_NQ;
         if (hasN) { if (same) pft_near_kernel<true, true><<<nearBlocks, PFT_THREADS, 0, ctx->stream>>>(Y); else pft_near_kernel<true, false><<<nearBlocks, PFT_THREADS, 0, ctx->stream>>>(Y); }
         else      { if (same) pft_near_kernel<false, true><<<nearBlocks, PFT_THREADS, 0, ctx->stream>>>(Y); else pft_near_kernel<false, false><<<nearBlocks, PFT_THREADS, 0, ctx->stream>>>(Y); }
         ctx->launches++; }
      PCK(cudaGetLastError());
      double q[PFT_NQ];
      if (sum_partials(ctx, P->dpart, nblocks, q)) return 1;
      // DeltaPFTT[sa][sb] += 0.5 Sign dPFTT, dPFTT itself carrying Sign (EMTPFT.cc:689-725): Sign^2 = 1
      double *out = scattered + ((size_t)sa * NS + sb) * 11;
      out[1] = 0.5 * q[0];
      for (int m = 0; m < 9; m++) out[2 + m] = 0.5 * q[1 + m]; }
  // ---- extinction (only the head of the incident-field chain, EMTPFT.cc:499-511)
  if (a->if_region >= 0 && (a->num_pw > 0 || a->tabulated_ehd))
   { ExtArgs E; E.havePW = 0;
     if (a->num_pw > 0) { if (make_pw(ctx, a->omega, 1, a->pw, &E.pw)) return 1; E.pw.sign = 1.0; E.havePW = 1; }
     size_t tabOff = 0;
     for (int ns = 0; ns < NS; ns++)
      { Surface &S = ctx->surfaces[ns]; const int nA = S.NP * SC_TCR9_NPTS;
        const size_t myTab = tabOff; tabOff += (size_t)nA * 48;
        if (a->if_region != a->region_indices[2 * ns + which] || nA == 0) continue;
        E.A = cur9 + off9[ns]; E.nA = nA; E.tab = nullptr;
        if (a->tabulated_ehd)
         { if (grow(ctx, P->dtab, P->tabCap, (size_t)nA * 48)) return 1;
           PCK(cudaMemcpyAsync(P->dtab, a->tabulated_ehd + myTab, (size_t)nA * 48 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
           ctx->h2dBytes += (size_t)nA * 48 * sizeof(double);
           E.tab = P->dtab; }
        for (int m = 0; m < 3; m++) E.origin[m] = a->origins[3 * ns + m];
        const int nb = (nA + PFT_THREADS - 1) / PFT_THREADS;
        if (grow(ctx, P->dpart, P->partCap, (size_t)nb * PFT_NQ)) return 1;
        E.part = P->dpart;
        pft_ext_kernel<<<nb, PFT_THREADS, 0, ctx->stream>>>(E);
        ctx->launches++;
        PCK(cudaGetLastError());
        double q[PFT_NQ];
        if (sum_partials(ctx, P->dpart, nb, q)) return 1;
        const double PFactor = 0.5, FTFactor = 0.5 * TENTHIRDS / a->omega[0];
        double *out = extinction + (size_t)ns * 11;
        out[0] = PFactor * q[0];
        for (int m = 0; m < 9; m++) out[2 + m] = FTFactor * q[1 + m]; } }
  return 0;
}

int scuffcuda_get_fields(scuffcuda_context *ctx, int ns, const double k[2], const double zrel[2], double sign, int npts, const double *X,
                         int kn_handle, int bf_offset, double *F, int rf_handle, const int *rf_columns, long *near_pairs)
{ if (ns < 0 || ns >= (int)ctx->surfaces.size()) return pfail(ctx, "get_fields: invalid surface index");
  if (near_pairs) *near_pairs = 0;
  if (npts <= 0) return 0;
  Surface &S = ctx->surfaces[ns];
  if (S.NE == 0) return 0;
  Matrix *KN = kn_handle ? find_matrix(ctx, kn_handle) : nullptr, *RF = rf_handle ? find_matrix(ctx, rf_handle) : nullptr;
  if (kn_handle && (!KN || KN->NC != 1 || KN->isReal)) return pfail(ctx, "get_fields: the solution vector must be a complex device vector");
  if (rf_handle && (!RF || RF->isReal || !rf_columns)) return pfail(ctx, "get_fields: the reduced-field matrix must be a complex device matrix");
  const int nbf = S.isPEC ? S.NE : 2 * S.NE;
  if (KN && (bf_offset < 0 || bf_offset + nbf > KN->NR)) return pfail(ctx, "get_fields: basis-function offset does not fit the vector");
  if (RF && (bf_offset < 0 || bf_offset + nbf > RF->NR)) return pfail(ctx, "get_fields: basis-function offset does not fit the reduced-field matrix");
  if (RF) for (int i = 0; i < npts; i++) if (rf_columns[i] < 0 || 6 * (rf_columns[i] + 1) > RF->NC) return pfail(ctx, "get_fields: reduced-field matrix has too few columns");
  if (!KN && !RF) return pfail(ctx, "get_fields: nothing to compute");
  PCK(cudaSetDevice(ctx->device));
  if (pft_tables(ctx) || ensure_edges(ctx, S)) return 1;
  PftState *P = pft_state(ctx);
  // layout of the scratch: X[3 npts] | F[12 npts] | counter (8 B) | columns (npts ints)
  const size_t need = (size_t)15 * npts + 1 + ((size_t)npts + 1) / 2 + 1;
  if (grow(ctx, P->dfx, P->fxCap, need)) return 1;
  double *dX = P->dfx, *dF = dX + (size_t)3 * npts; unsigned long long *dcnt = (unsigned long long *)(dF + (size_t)12 * npts); int *dcol = (int *)(dcnt + 1);
  PCK(cudaMemcpyAsync(dX, X, (size_t)3 * npts * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  PCK(cudaMemsetAsync(dF, 0, ((size_t)12 * npts + 1) * sizeof(double), ctx->stream));
  if (RF) PCK(cudaMemcpyAsync(dcol, rf_columns, (size_t)npts * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  ctx->h2dBytes += (size_t)3 * npts * sizeof(double);
  FieldArgs A;
  A.edges = (const EdgeRec *)S.dedges; A.NE = S.NE; A.X = dX; A.NX = npts; A.tcr7 = P->dtcr7; A.tcr20 = P->dtcr20;
  A.k = mk(k[0], k[1]); A.zrel = mk(zrel[0], zrel[1]); A.sign = sign;
  A.panels = S.dpanels; A.tcr9 = P->dtcr9;
  A.kn = KN ? KN->d : nullptr; A.offset = bf_offset; A.isPEC = S.isPEC; A.F = dF;
  A.RF = RF ? RF->d : nullptr; A.ldRF = RF ? RF->NR : 0; A.colOf = dcol; A.nearCount = dcnt;
  fields_kernel<<<npts, PFT_THREADS, 0, ctx->stream>>>(A);
  ctx->launches++;
  PCK(cudaGetLastError());
  unsigned long long cnt = 0;
  PCK(cudaMemcpyAsync(&cnt, dcnt, sizeof cnt, cudaMemcpyDeviceToHost, ctx->stream));
  if (KN && F)
   { std::vector<double> h((size_t)12 * npts);
     PCK(cudaMemcpyAsync(h.data(), dF, h.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
     PCK(cudaStreamSynchronize(ctx->stream));
     ctx->d2hBytes += h.size() * sizeof(double);
     for (size_t i = 0; i < h.size(); i++) F[i] += h[i]; }
  else PCK(cudaStreamSynchronize(ctx->stream));
  if (near_pairs) *near_pairs = (long)cnt;
  return 0;
}

int scuffcuda_dgf_contract(scuffcuda_context *ctx, int rf_dest_handle, int rf_source_handle, int npts, double *out)
{ Matrix *D = find_matrix(ctx, rf_dest_handle), *S = find_matrix(ctx, rf_source_handle);
  if (!D || !S || D->isReal || S->isReal) return pfail(ctx, "dgf_contract: invalid reduced-field matrix handle");
  if (D->NR != S->NR || D->NC != S->NC || npts <= 0 || D->NC != 6 * npts) return pfail(ctx, "dgf_contract: the matrices must be TotalBFs x 6 NX");
  PCK(cudaSetDevice(ctx->device));
  PftState *P = pft_state(ctx);
  if (grow(ctx, P->dpart, P->partCap, (size_t)36 * npts)) return 1;
  dgf_contract_kernel<<<npts, PFT_THREADS, 0, ctx->stream>>>(D->d, S->d, D->NR, P->dpart);
  ctx->launches++;
  PCK(cudaGetLastError());
  PCK(cudaMemcpyAsync(out, P->dpart, (size_t)36 * npts * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  PCK(cudaStreamSynchronize(ctx->stream));
  ctx->d2hBytes += (size_t)36 * npts * sizeof(double);
  return 0;
}

}  // extern "C"
