// Far-field panel-pair kernel: every panel pair with rRel > 4 (the reference's PPIALG_LOCUBATURE branch,
// PanelPanelInteractions.cc:420-424) is integrated ONCE with the 6x6 order-4 rule and scattered, with RWG signs and
// edge lengths, into the up-to-nine edge-pair entries it feeds (the reference re-evaluates the same panel pair for
// every (Q,Q') combination and every edge pair: 4.5-9x redundant).
//
// Formulation.  With X = V0 + u A + v B (A = V1-V0, B = V2-V0) the RWG factor of vertex i is
//     X - V_i = (u - [i==1]) A + (v - [i==2]) B ,
// so all nine (Q_a,Q_b) combinations of a panel pair follow from the 3x3 barycentric moments
//     MPhi[m][n] = sum_{p,p'} w_p w_p' {1,u,v}_m {1,u',v'}_n Phi(r_pp') ,   MPsi likewise with Psi:
//     H0_ij = sum_{X in {A,B}, Y in {A',B'}} (X.Y) <(x-d_i)(y-d_j) Phi> + 4/(ik)^2 <Phi>
//     H1_ij = sum_{X,Y} ((X x Y).E_ij) <(x-d_i)(y-d_j) Psi> ,   E_ij = V_i - V'_j (- L)
// (identical, term by term, to GetPPIs_Cubature's  (F.F' + 4/(ik)^2) Phi  and  ((F x F').R) Psi  because
// (F x F').R = (F x F').(R - F + F') and R - F + F' = V_i - V'_j.)  The inner loop per point pair is: R, r^2,
// 1/r (MUFU seed + FP64 Newton), sincos / exp with FP64 range reduction (fastmath.cuh), Phi (and Psi), and three
// complex accumulations whose multipliers u', v' are immediates.  PEC/PEC blocks need only H0 (the stamp uses GC[0]
// only, SurfaceSurfaceInteractions.cc:205-217), so the Psi half is compiled out (template PSI=false).
//
// Mapping: one thread per panel pair.  A CTA owns a tile of TA a-panels — brought into shared memory with ONE 1-D TMA
// bulk copy (cp.async.bulk + mbarrier; a tile of PanelRec is contiguous) and read by broadcast — x 128 b-panels, one
// per lane, whose cubature points stay in registers for the whole tile.  FP64-pipe bound.
#include <stdint.h>
#include <algorithm>
#include <cstdlib>
#include "context.h"
#include "classify.cuh"
#include "stamp.cuh"
#include "kernels.h"
#include "fastmath.cuh"
#include "sw_kernels.h"
#include "generated_tcr20.h"

namespace scuffcuda {

#define FAR_THREADS 128

// w_p * {1, u_p, v_p} of the 6-point rule (TCR.cc:83-91 literals)
__constant__ double c_far_wm[6][3] = {
  { 0.0549758718276, 0.0549758718276 * 0.0915762135098, 0.0549758718276 * 0.0915762135098 },
  { 0.0549758718276, 0.0549758718276 * 0.8168475729805, 0.0549758718276 * 0.0915762135098 },
  { 0.0549758718276, 0.0549758718276 * 0.0915762135098, 0.0549758718276 * 0.8168475729805 },
  { 0.1116907948390, 0.1116907948390 * 0.1081030181681, 0.1116907948390 * 0.4459484909160 },
  { 0.1116907948390, 0.1116907948390 * 0.4459484909160, 0.1116907948390 * 0.1081030181681 },
  { 0.1116907948390, 0.1116907948390 * 0.4459484909160, 0.1116907948390 * 0.4459484909160 } };

__device__ __forceinline__ constexpr double tcr_u(int p)
{ return p == 1 ? 0.8168475729805 : (p == 0 || p == 2) ? 0.0915762135098 : (p == 3) ? 0.1081030181681 : 0.4459484909160; }
__device__ __forceinline__ constexpr double tcr_v(int p)
{ return p == 2 ? 0.8168475729805 : (p == 0 || p == 1) ? 0.0915762135098 : (p == 4) ? 0.1081030181681 : 0.4459484909160; }
__device__ __forceinline__ constexpr double tcr_w(int p) { return p < 3 ? 0.0549758718276 : 0.1116907948390; }

// Phi * w' (and Psi * w') at one point pair; cw = w'/(4 pi)
template <int KMODE, bool PSI, typename S>
__device__ __forceinline__ void helm_fast(double r2, double cw, const KParams &K, S &Phi, S &Psi);

template <>
__device__ __forceinline__ void helm_fast<KMODE_REAL, false, cd>(double r2, double cw, const KParams &K, cd &Phi, cd &Psi)
{ const double rinv = rsqrt_fast<false>(r2), r = r2 * rinv;
  double s, c; sincos_fast(K.kr * r, s, c);
  const double f = cw * rinv;
  Phi = mk(f * c, f * s);
}
template <>
__device__ __forceinline__ void helm_fast<KMODE_REAL, true, cd>(double r2, double cw, const KParams &K, cd &Phi, cd &Psi)
{ const double rinv = rsqrt_fast<false>(r2), r = r2 * rinv;
  double s, c; sincos_fast(K.kr * r, s, c);
  const double f = cw * rinv;
  Phi = mk(f * c, f * s);
  const double a = -rinv * rinv, b = K.kr * rinv;          // Psi = Phi (ik - 1/r)/r
  Psi = mk(fma(Phi.x, a, -Phi.y * b), fma(Phi.x, b, Phi.y * a));
}
template <>
__device__ __forceinline__ void helm_fast<KMODE_IMAG, false, double>(double r2, double cw, const KParams &K, double &Phi, double &Psi)
{ const double rinv = rsqrt_fast<false>(r2), r = r2 * rinv;
  Phi = cw * rinv * exp_neg_fast(-K.ki * r);
}
template <>
__device__ __forceinline__ void helm_fast<KMODE_IMAG, true, double>(double r2, double cw, const KParams &K, double &Phi, double &Psi)
{ const double rinv = rsqrt_fast<false>(r2), r = r2 * rinv;
  Phi = cw * rinv * exp_neg_fast(-K.ki * r);
  Psi = Phi * (-K.ki - rinv) * rinv;
}
template <>
__device__ __forceinline__ void helm_fast<KMODE_CPLX, false, cd>(double r2, double cw, const KParams &K, cd &Phi, cd &Psi)
{ const double rinv = rsqrt_fast<false>(r2), r = r2 * rinv;
  double s, c; sincos_fast(K.kr * r, s, c);
  const double y = -K.ki * r;
  const double f = cw * rinv * (y <= 0.0 ? exp_neg_fast(y) : exp(y));
  Phi = mk(f * c, f * s);
}
template <>
__device__ __forceinline__ void helm_fast<KMODE_CPLX, true, cd>(double r2, double cw, const KParams &K, cd &Phi, cd &Psi)
{ const double rinv = rsqrt_fast<false>(r2), r = r2 * rinv;
  double s, c; sincos_fast(K.kr * r, s, c);
  const double y = -K.ki * r;
  const double f = cw * rinv * (y <= 0.0 ? exp_neg_fast(y) : exp(y));
  Phi = mk(f * c, f * s);
  const double a = -K.ki * rinv - rinv * rinv, b = K.kr * rinv;
  Psi = mk(fma(Phi.x, a, -Phi.y * b), fma(Phi.x, b, Phi.y * a));
}

// "Small-phase" variant for fine meshes: every point pair of a far panel pair has r = r0 + delta with r0 = |C_a - C_b| and
// |delta| <= R_a + R_b.  When |k| (R_a + R_b) <= 1/8 for all panels (the launcher checks), exp(ik r) = exp(ik r0) exp(ik delta):
// the pair factor is applied ONCE to the accumulated moments (they are linear in Phi and Psi), and exp(ik delta) needs no
// range reduction, no quadrant logic and only degree-9/10 Taylor polynomials (truncation < 3e-18 at |t| <= 1/8).
template <int KMODE, bool PSI, typename S>
__device__ __forceinline__ void helm_small(double r2, double cw, const KParams &K, double r0, S &Phi, S &Psi);

__device__ __forceinline__ void sincos_small(double t, double &s, double &c)
{ const double z = t * t;
  double ps = fma(z, 2.7557319223985893e-06, -1.9841269841269841e-04);       // 1/9!, -1/7!
  ps = fma(z, ps, 8.3333333333333332e-03); ps = fma(z, ps, -1.6666666666666666e-01);
  s = fma(t * z, ps, t);
  double pc = fma(z, -2.7557319223985888e-07, 2.4801587301587302e-05);        // -1/10!, 1/8!
  pc = fma(z, pc, -1.3888888888888889e-03); pc = fma(z, pc, 4.1666666666666664e-02); pc = fma(z, pc, -0.5);
  c = fma(z, pc, 1.0);
}
__device__ __forceinline__ double exp_small(double s)                          // |s| <= 1/8, Taylor to s^10
{ double p = 2.7557319223985888e-07;
  p = fma(p, s, 2.7557319223985893e-06); p = fma(p, s, 2.4801587301587302e-05); p = fma(p, s, 1.9841269841269841e-04);
  p = fma(p, s, 1.3888888888888889e-03); p = fma(p, s, 8.3333333333333332e-03); p = fma(p, s, 4.1666666666666664e-02);
  p = fma(p, s, 1.6666666666666666e-01); p = fma(p, s, 0.5); p = fma(p, s, 1.0);
  return fma(p, s, 1.0);
}
template <>
__device__ __forceinline__ void helm_small<KMODE_REAL, false, cd>(double r2, double cw, const KParams &K, double r0, cd &Phi, cd &Psi)
{ const double rinv = rsqrt_fast<false>(r2), r = r2 * rinv;
  double s, c; sincos_small(K.kr * (r - r0), s, c);
  const double f = cw * rinv;
  Phi = mk(f * c, f * s);
}
template <>
__device__ __forceinline__ void helm_small<KMODE_REAL, true, cd>(double r2, double cw, const KParams &K, double r0, cd &Phi, cd &Psi)
{ const double rinv = rsqrt_fast<false>(r2), r = r2 * rinv;
  double s, c; sincos_small(K.kr * (r - r0), s, c);
  const double f = cw * rinv;
  Phi = mk(f * c, f * s);
  const double a = -rinv * rinv, b = K.kr * rinv;
  Psi = mk(fma(Phi.x, a, -Phi.y * b), fma(Phi.x, b, Phi.y * a));
}
template <>
__device__ __forceinline__ void helm_small<KMODE_IMAG, false, double>(double r2, double cw, const KParams &K, double r0, double &Phi, double &Psi)
{ const double rinv = rsqrt_fast<false>(r2), r = r2 * rinv;
  Phi = cw * rinv * exp_small(-K.ki * (r - r0));
}
template <>
__device__ __forceinline__ void helm_small<KMODE_IMAG, true, double>(double r2, double cw, const KParams &K, double r0, double &Phi, double &Psi)
{ const double rinv = rsqrt_fast<false>(r2), r = r2 * rinv;
  Phi = cw * rinv * exp_small(-K.ki * (r - r0));
  Psi = Phi * (-K.ki - rinv) * rinv;
}
// general complex k never takes the small-phase path (the launcher does not select it)
template <>
__device__ __forceinline__ void helm_small<KMODE_CPLX, false, cd>(double r2, double cw, const KParams &K, double r0, cd &Phi, cd &Psi)
{ helm_fast<KMODE_CPLX, false, cd>(r2, cw, K, Phi, Psi); }
template <>
__device__ __forceinline__ void helm_small<KMODE_CPLX, true, cd>(double r2, double cw, const KParams &K, double r0, cd &Phi, cd &Psi)
{ helm_fast<KMODE_CPLX, true, cd>(r2, cw, K, Phi, Psi); }

__device__ __forceinline__ void smul(double &a, double f) { a *= f; }
__device__ __forceinline__ void smul(cd &a, cd f) { a = mk(fma(a.x, f.x, -a.y * f.y), fma(a.x, f.y, a.y * f.x)); }

__device__ __forceinline__ void sadd(double &a, double b) { a += b; }
__device__ __forceinline__ void sadd(cd &a, cd b) { a.x += b.x; a.y += b.y; }
__device__ __forceinline__ double ssub(double a, double b) { return a - b; }
__device__ __forceinline__ cd ssub(cd a, cd b) { return mk(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double dot3(const double *a, const double *b) { return fma(a[2], b[2], fma(a[1], b[1], a[0] * b[0])); }
__device__ __forceinline__ void cross3(const double *a, const double *b, double *c)
{ c[0] = fma(a[1], b[2], -a[2] * b[1]); c[1] = fma(a[2], b[0], -a[0] * b[2]); c[2] = fma(a[0], b[1], -a[1] * b[0]); }

// ---- 1-D TMA bulk copy global -> shared, completion on an mbarrier ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_load_1d(void *dstSmem, const void *srcGlobal, uint32_t bytes, uint64_t *bar)
{ const uint32_t b = smem_u32(bar), d = smem_u32(dstSmem);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(d), "l"(srcGlobal), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase)
{ const uint32_t b = smem_u32(bar); uint32_t ok;
  do { asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                    : "=r"(ok) : "r"(b), "r"(phase) : "memory"); } while (!ok);
}

// (an out-of-line scatter to shrink the 74 KB dielectric kernel was tried: the calls at 240 live registers made it 2x slower)
// Contraction of the 3x3 moments to the nine (Q_a,Q_b) combinations and scatter into the matrix.
// Va0/Aa/Ba: vertex 0 and edge vectors of panel a; Vb0 (already displaced by L), Ab, Bb of panel b.
template <typename S, bool PSI>
__device__ __forceinline__ void contract_scatter(const S (&mPhi)[3][3], const S (&mPsi)[3][3],
                                                 const double *Va0, const double *Aa, const double *Ba,
                                                 const double *Vb0, const double *Ab, const double *Bb,
                                                 const int *eia, const double *sla3, const int *eib, const double *slb,
                                                 int pa, int pb, const StampParams &SP, const KParams &K, int symmetric,
                                                 unsigned long long &nPPI)
{ double E00[3];
#pragma unroll
  for (int m = 0; m < 3; m++) E00[m] = Va0[m] - Vb0[m];
  const double dAA = dot3(Aa, Ab), dAB = dot3(Aa, Bb), dBA = dot3(Ba, Ab), dBB = dot3(Ba, Bb);
  // rows of the a-side shifted moments: row 1 ~ (u - [i==1]), row 2 ~ (v - [i==2])
  S r1s[3], r2s[3], q1s[3], q2s[3];
#pragma unroll
  for (int n = 0; n < 3; n++)
   { r1s[n] = ssub(mPhi[1][n], mPhi[0][n]); r2s[n] = ssub(mPhi[2][n], mPhi[0][n]);
     if (PSI) { q1s[n] = ssub(mPsi[1][n], mPsi[0][n]); q2s[n] = ssub(mPsi[2][n], mPsi[0][n]); } }
  const cd c4m = K.c4 * tocd(mPhi[0][0]);
  // triple products for H1
  double tAA = 0, tAB = 0, tBA = 0, tBB = 0, AbN = 0, BbN = 0, AaNp = 0, BaNp = 0;
  if (PSI)
   { double Na[3], Nb[3], eA[3], eB[3];
     cross3(Aa, Ba, Na); cross3(Ab, Bb, Nb);
     cross3(E00, Aa, eA); cross3(E00, Ba, eB);         // (X x Y).E = Y.(E x X)
     tAA = dot3(Ab, eA); tAB = dot3(Bb, eA); tBA = dot3(Ab, eB); tBB = dot3(Bb, eB);
     AbN = dot3(Ab, Na); BbN = dot3(Bb, Na); AaNp = dot3(Aa, Nb); BaNp = dot3(Ba, Nb); }

#pragma unroll
  for (int i = 0; i < 3; i++)
   { const int ea = eia[i];
     if (ea < 0) continue;
     const double sla = sla3[i];
     const S *U = (i == 1) ? r1s : mPhi[1], *V = (i == 2) ? r2s : mPhi[2];
     const S *UQ = (i == 1) ? q1s : mPsi[1], *VQ = (i == 2) ? q2s : mPsi[2];
#pragma unroll
     for (int j = 0; j < 3; j++)
      { if (eib[j] < 0) continue;
        // b-side shift: column 1 ~ (u' - [j==1]), column 2 ~ (v' - [j==2])
        const S Uu = (j == 1) ? ssub(U[1], U[0]) : U[1], Uv = (j == 2) ? ssub(U[2], U[0]) : U[2];
        const S Vu = (j == 1) ? ssub(V[1], V[0]) : V[1], Vv = (j == 2) ? ssub(V[2], V[0]) : V[2];
        S h0; szero(h0);
        fma_acc(h0, dAA, Uu); fma_acc(h0, dAB, Uv); fma_acc(h0, dBA, Vu); fma_acc(h0, dBB, Vv);
        cd H0 = tocd(h0) + c4m, H1 = mk(0.0, 0.0);
        if (PSI)
         { const S Pu = (j == 1) ? ssub(UQ[1], UQ[0]) : UQ[1], Pv = (j == 2) ? ssub(UQ[2], UQ[0]) : UQ[2];
           const S Qu = (j == 1) ? ssub(VQ[1], VQ[0]) : VQ[1], Qv = (j == 2) ? ssub(VQ[2], VQ[0]) : VQ[2];
           // (X x Y).E_ij with E_ij = E00 + {0,A,B}_i - {0,A',B'}_j
           const double cAA = tAA - (i == 2 ? AbN : 0.0) - (j == 2 ? AaNp : 0.0);
           const double cAB = tAB - (i == 2 ? BbN : 0.0) + (j == 1 ? AaNp : 0.0);
           const double cBA = tBA + (i == 1 ? AbN : 0.0) - (j == 2 ? BaNp : 0.0);
           const double cBB = tBB + (i == 1 ? BbN : 0.0) + (j == 1 ? BaNp : 0.0);
           S h1; szero(h1);
           fma_acc(h1, cAA, Pu); fma_acc(h1, cAB, Pv); fma_acc(h1, cBA, Qu); fma_acc(h1, cBB, Qv);
           H1 = tocd(h1); }
        scatter_ppi<PSI>(SP, K, pa, pb, ea, eib[j], sla, slb[j], H0, H1);
        nPPI += (symmetric && pa != pb && ea == eib[j]) ? 2 : 1;
      }
   }
}

// exact-decision far test without sqrt/divide on the common path: rRel > 4  <=>  sqrt(r2)/Rmax > 4.  Away from the
// threshold by more than 1e-9 relative the comparison of squares decides; inside that band the reference's own
// expression (classify.cuh pair_rrel) is evaluated.
__device__ __forceinline__ bool pair_is_far(const double *Ca, double Ra, const double *Cb, double Rb, bool useDisp, const double *L, double D[3])
{ if (useDisp)
   { D[0] = SC_SUB(SC_SUB(Ca[0], Cb[0]), L[0]); D[1] = SC_SUB(SC_SUB(Ca[1], Cb[1]), L[1]); D[2] = SC_SUB(SC_SUB(Ca[2], Cb[2]), L[2]); }
  else
   { D[0] = SC_SUB(Ca[0], Cb[0]); D[1] = SC_SUB(Ca[1], Cb[1]); D[2] = SC_SUB(Ca[2], Cb[2]); }
  double r2 = SC_MUL(D[0], D[0]);
  r2 = SC_ADD(r2, SC_MUL(D[1], D[1]));
  r2 = SC_ADD(r2, SC_MUL(D[2], D[2]));
  const double Rm = fmax(Ra, Rb), t = 16.0 * Rm * Rm;
  if (r2 > t * (1.0 + 1.0e-9)) return true;
  if (r2 < t * (1.0 - 1.0e-9)) return false;
  return sqrt(r2) / Rm > SCUFF_DESING_RADIUS;
}

// shared-memory layout (doubles): b-tile [TB][36] | lane (a) cubature points [18][128] | lane stash [16][128] | lane edge ids (int)[3][128]
#define FAR_SMEM_BYTES(TA) ((size_t)((TA) * 36 + 34 * FAR_THREADS) * sizeof(double) + 6 * FAR_THREADS * sizeof(int))

// Thread mapping: lane <-> a-panel (matrix ROW side), tile <-> b-panels (column side).  The matrix is column-major, so
// the RED.ADD.F64 of one warp for entry (e_a(lane), e_b) go to neighbouring rows of ONE column (edges of consecutive
// panels are nearly consecutive) instead of 32 different columns: ~4-8x fewer L2 atomic sectors per warp instruction.
template <int KMODE, bool PSI, int MINB, bool SMALL>
__global__ void __launch_bounds__(FAR_THREADS, MINB)
far_kernel(const __grid_constant__ FarArgs A)
{ extern __shared__ __align__(128) double sm[];
  __shared__ __align__(8) uint64_t bar;
  typedef typename KScalar<KMODE>::type S;
  const int TB = A.TA;
  const int b0 = blockIdx.y * TB;
  const int aBase = blockIdx.x * FAR_THREADS;
  const int nB = min(TB, A.NPb - b0);
  // symmetric: only pa <= pb is visited; skip CTAs whose b-tile lies entirely below their a-lanes
  if (A.symmetric && b0 + nB - 1 < aBase) return;
  const int tid = threadIdx.x;
  double *sXa = sm + TB * 36 + tid;                    // element k at sXa[k * FAR_THREADS]
  double *sSt = sm + TB * 36 + 18 * FAR_THREADS + tid;
  int *sEi = (int *)(sm + TB * 36 + 34 * FAR_THREADS) + tid;
  unsigned int *sCnt = (unsigned int *)sEi + 3 * FAR_THREADS;   // [0] far pairs, [1] reference-PPI count (thread-private slots)
  sCnt[0] = 0; sCnt[FAR_THREADS] = 0;

  // stage the b-tile (nB contiguous PanelRec) with one TMA bulk copy
  if (tid == 0)
   { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
     asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (tid == 0) tma_load_1d(sm, A.pb + b0, (uint32_t)(nB * sizeof(PanelRec)), &bar);

  const int pa = aBase + tid;
  const bool live = pa < A.NPa && !(A.symmetric && pa > b0 + nB - 1);
  // this lane's a-panel: cubature points, centroid/radius and the epilogue data in shared memory (thread-private columns)
  const PanelRec *Pa = A.pa + (live ? pa : 0);
#pragma unroll
  for (int i = 0; i < 18; i++) sXa[i * FAR_THREADS] = Pa->x[i];
#pragma unroll
  for (int m = 0; m < 3; m++)
   { const double v0 = Pa->V[m];
     sSt[m * FAR_THREADS] = v0;                                      // V_0
     sSt[(3 + m) * FAR_THREADS] = Pa->V[3 + m] - v0;                 // A
     sSt[(6 + m) * FAR_THREADS] = Pa->V[6 + m] - v0;                 // B
     sSt[(9 + m) * FAR_THREADS] = Pa->sl[m];
     sSt[(12 + m) * FAR_THREADS] = Pa->C[m];
     sEi[m * FAR_THREADS] = Pa->ei[m]; }
  sSt[15 * FAR_THREADS] = Pa->R;

  mbar_wait(&bar, 0);
  if (!live) return;

  for (int ib = 0; ib < nB; ib++)
   { const int pb = b0 + ib;
     if (A.symmetric && pb < pa) continue;
     const double *Pb = sm + ib * 36;
     double D[3];
     { const double Ca[3] = { sSt[12 * FAR_THREADS], sSt[13 * FAR_THREADS], sSt[14 * FAR_THREADS] };
       if (!pair_is_far(Ca, sSt[15 * FAR_THREADS], Pb, Pb[3], A.useDisp != 0, A.L, D)) continue; }
     if (A.S.colHi != 0x7fffffff)
      { // column-block assembly: skip pairs none of whose (<= 9) entries falls into the owned column range.  The test is
        // exact per entry (an interval test on the edge indices is loose: the three edges of a panel next to an icosahedron
        // face boundary are thousands of indices apart, and such pairs were integrated by every block).
        const int *eb3 = (const int *)(Pb + 34);
        const int w = A.S.bPEC ? 1 : 2;
        bool any = false;
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
         for (int j = 0; j < 3; j++)
          { const int e = sEi[i * FAR_THREADS], f = eb3[j];
            // symmetric blocks store entry (min(e,f), max(e,f)): its column is max(e,f); otherwise the column is f
            const int c0 = A.S.colOff + w * (A.symmetric ? max(e, f) : f);
            any |= (e >= 0 && f >= 0 && c0 + w - 1 >= A.S.colLo && c0 < A.S.colHi); }
        if (!any) continue; }
     sCnt[0] += 1;

     double r0 = 0.0;
     if (SMALL) { const double d2 = fma(D[2], D[2], fma(D[1], D[1], D[0] * D[0])); r0 = d2 * rsqrt_fast<true>(d2); }
     // ---- 3x3 barycentric moments m[m over a (outer loop: this lane's points)][n over b (inner loop: tile panel's points,
     //      broadcast shared-memory reads)] ----
     S mPhi[3][3], mPsi[3][3];
#pragma unroll
     for (int m = 0; m < 3; m++)
#pragma unroll
      for (int n = 0; n < 3; n++) { szero(mPhi[m][n]); szero(mPsi[m][n]); }

#pragma unroll 1
     for (int p = 0; p < 6; p += 2)
      { // two outer points per iteration share the broadcast loads of the tile panel's points and give the scheduler two
        // independent dependency chains per inner point.  R = X_a - X_b = (x_a(lane) + D) - x_b(tile)
        const double y0 = sXa[(3 * p) * FAR_THREADS] + D[0], y1 = sXa[(3 * p + 1) * FAR_THREADS] + D[1], y2 = sXa[(3 * p + 2) * FAR_THREADS] + D[2];
        const double z0 = sXa[(3 * p + 3) * FAR_THREADS] + D[0], z1 = sXa[(3 * p + 4) * FAR_THREADS] + D[1], z2 = sXa[(3 * p + 5) * FAR_THREADS] + D[2];
        S a0_, au, av, b0_, bu, bv, c0_, cu, cv, d0_, du, dv;
        szero(a0_); szero(au); szero(av); szero(b0_); szero(bu); szero(bv);
        szero(c0_); szero(cu); szero(cv); szero(d0_); szero(du); szero(dv);
#pragma unroll
        for (int pp = 0; pp < 6; pp++)
         { const double xb0 = Pb[13 + 3 * pp], xb1 = Pb[14 + 3 * pp], xb2 = Pb[15 + 3 * pp];
           const double R0 = y0 - xb0, R1 = y1 - xb1, R2 = y2 - xb2;
           const double Q0 = z0 - xb0, Q1 = z1 - xb1, Q2 = z2 - xb2;
           const double r2 = fma(R2, R2, fma(R1, R1, R0 * R0));
           const double q2 = fma(Q2, Q2, fma(Q1, Q1, Q0 * Q0));
           S Phi, Psi, Phj, Psj;
           if (SMALL) { helm_small<KMODE, PSI, S>(r2, SCUFF_OO4PI * tcr_w(pp), A.K, r0, Phi, Psi);
                        helm_small<KMODE, PSI, S>(q2, SCUFF_OO4PI * tcr_w(pp), A.K, r0, Phj, Psj); }
           else { helm_fast<KMODE, PSI, S>(r2, SCUFF_OO4PI * tcr_w(pp), A.K, Phi, Psi);
                  helm_fast<KMODE, PSI, S>(q2, SCUFF_OO4PI * tcr_w(pp), A.K, Phj, Psj); }
           sadd(a0_, Phi); fma_acc(au, tcr_u(pp), Phi); fma_acc(av, tcr_v(pp), Phi);
           sadd(c0_, Phj); fma_acc(cu, tcr_u(pp), Phj); fma_acc(cv, tcr_v(pp), Phj);
           if (PSI) { sadd(b0_, Psi); fma_acc(bu, tcr_u(pp), Psi); fma_acc(bv, tcr_v(pp), Psi);
                      sadd(d0_, Psj); fma_acc(du, tcr_u(pp), Psj); fma_acc(dv, tcr_v(pp), Psj); }
         }
        { const double w0 = c_far_wm[p][0], w1 = c_far_wm[p][1], w2 = c_far_wm[p][2];
          fma_acc(mPhi[0][0], w0, a0_); fma_acc(mPhi[0][1], w0, au); fma_acc(mPhi[0][2], w0, av);
          fma_acc(mPhi[1][0], w1, a0_); fma_acc(mPhi[1][1], w1, au); fma_acc(mPhi[1][2], w1, av);
          fma_acc(mPhi[2][0], w2, a0_); fma_acc(mPhi[2][1], w2, au); fma_acc(mPhi[2][2], w2, av);
          if (PSI)
           { fma_acc(mPsi[0][0], w0, b0_); fma_acc(mPsi[0][1], w0, bu); fma_acc(mPsi[0][2], w0, bv);
             fma_acc(mPsi[1][0], w1, b0_); fma_acc(mPsi[1][1], w1, bu); fma_acc(mPsi[1][2], w1, bv);
             fma_acc(mPsi[2][0], w2, b0_); fma_acc(mPsi[2][1], w2, bu); fma_acc(mPsi[2][2], w2, bv); } }
        { const double w0 = c_far_wm[p + 1][0], w1 = c_far_wm[p + 1][1], w2 = c_far_wm[p + 1][2];
          fma_acc(mPhi[0][0], w0, c0_); fma_acc(mPhi[0][1], w0, cu); fma_acc(mPhi[0][2], w0, cv);
          fma_acc(mPhi[1][0], w1, c0_); fma_acc(mPhi[1][1], w1, cu); fma_acc(mPhi[1][2], w1, cv);
          fma_acc(mPhi[2][0], w2, c0_); fma_acc(mPhi[2][1], w2, cu); fma_acc(mPhi[2][2], w2, cv);
          if (PSI)
           { fma_acc(mPsi[0][0], w0, d0_); fma_acc(mPsi[0][1], w0, du); fma_acc(mPsi[0][2], w0, dv);
             fma_acc(mPsi[1][0], w1, d0_); fma_acc(mPsi[1][1], w1, du); fma_acc(mPsi[1][2], w1, dv);
             fma_acc(mPsi[2][0], w2, d0_); fma_acc(mPsi[2][1], w2, du); fma_acc(mPsi[2][2], w2, dv); } }
      }

     if (SMALL)
      { // the pair factor exp(ik r0), once for all moments
        S E0;
        if (KMODE == KMODE_IMAG) { const double e = exp_neg_fast(-A.K.ki * r0); E0 = *reinterpret_cast<const S *>(&e); }
        else { double s0, c0; sincos_fast(A.K.kr * r0, s0, c0); const cd e = mk(c0, s0); E0 = *reinterpret_cast<const S *>(&e); }
#pragma unroll
        for (int m = 0; m < 3; m++)
#pragma unroll
         for (int n = 0; n < 3; n++) { smul(mPhi[m][n], E0); if (PSI) smul(mPsi[m][n], E0); }
      }
     // ---- contraction to the nine (Q_a,Q_b) combinations and scatter ----
     {
       int eia[3]; double sla[3], Aa[3], Ba[3], Ab[3], Bb[3], Va0[3], Vb0[3];
#pragma unroll
       for (int m = 0; m < 3; m++)
        { eia[m] = sEi[m * FAR_THREADS]; sla[m] = sSt[(9 + m) * FAR_THREADS];
          Va0[m] = sSt[m * FAR_THREADS]; Aa[m] = sSt[(3 + m) * FAR_THREADS]; Ba[m] = sSt[(6 + m) * FAR_THREADS];
          Vb0[m] = Pb[4 + m] + (A.useDisp ? A.L[m] : 0.0);
          Ab[m] = Pb[7 + m] - Pb[4 + m]; Bb[m] = Pb[10 + m] - Pb[4 + m]; }
       unsigned long long nPPI = 0;
       contract_scatter<S, PSI>(mPhi, mPsi, Va0, Aa, Ba, Vb0, Ab, Bb, eia, sla, (const int *)(Pb + 34), Pb + 31,
                                pa, pb, A.S, A.K, A.symmetric, nPPI);
       sCnt[FAR_THREADS] += (unsigned int)nPPI;
     }
   }
  // algorithm counters (reference log line SurfaceSurfaceInteractions.cc:550-558)
  if (sCnt[FAR_THREADS]) atomicAdd(A.counters, (unsigned long long)sCnt[FAR_THREADS]);
  if (sCnt[0]) atomicAdd(A.counters + 5, (unsigned long long)sCnt[0]);
}

// ------------------------------------------------------------------------------------------------------------------
// High-order cubature (78 x 78 points, GetTCR(20)) for near pairs without common vertices in the short-wavelength
// regime (PPIALG_HOCUBATURE, PanelPanelInteractions.cc:437-441).  One WARP per panel pair: lanes split the 78 outer
// points, every lane sweeps the 78 inner points (broadcast shared-memory reads), the 3x3 complex moments are reduced with
// warp shuffles and lane 0 contracts and scatters.
// ------------------------------------------------------------------------------------------------------------------
#define HO_NPTS SC_TCR20_NPTS
#define HO_WARPS 4

__device__ __forceinline__ double wsum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v; }
__device__ __forceinline__ cd wsum(cd v) { return mk(wsum(v.x), wsum(v.y)); }

template <int KMODE, bool PSI>
__global__ void __launch_bounds__(32 * HO_WARPS)
ho_kernel(const NearArgs A, const int *idx, int n, const double *tcr)
{ extern __shared__ __align__(16) double sm[];
  typedef typename KScalar<KMODE>::type S;
  double *sT = sm;
  for (int i = threadIdx.x; i < 3 * HO_NPTS; i += blockDim.x) sT[i] = tcr[i];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * HO_WARPS + warp;
  double *sX = sm + 3 * HO_NPTS * (1 + warp);
  const bool live = w < n;
  const NearPair N = A.near[idx[live ? w : 0]];
  const PanelRec &Pa = A.pa[N.pa], &Pb = A.pb[N.pb];
  // everything relative to panel a's centroid
  double Va0[3], Aa[3], Ba[3], Vb0[3], Ab[3], Bb[3];
#pragma unroll
  for (int m = 0; m < 3; m++)
   { Va0[m] = Pa.V[m]; Aa[m] = Pa.V[3 + m] - Pa.V[m]; Ba[m] = Pa.V[6 + m] - Pa.V[m];
     Vb0[m] = Pb.V[m] + (A.useDisp ? A.L[m] : 0.0); Ab[m] = Pb.V[3 + m] - Pb.V[m]; Bb[m] = Pb.V[6 + m] - Pb.V[m]; }
  __syncthreads();
  for (int q = lane; q < HO_NPTS; q += 32)
#pragma unroll
   for (int m = 0; m < 3; m++) sX[3 * q + m] = (Vb0[m] - Pa.C[m]) + sT[3 * q] * Ab[m] + sT[3 * q + 1] * Bb[m];
  __syncwarp();

  S mPhi[3][3], mPsi[3][3];
#pragma unroll
  for (int m = 0; m < 3; m++)
#pragma unroll
   for (int k = 0; k < 3; k++) { szero(mPhi[m][k]); szero(mPsi[m][k]); }
  if (live)
   for (int p = lane; p < HO_NPTS; p += 32)
    { const double u = sT[3 * p], v = sT[3 * p + 1], wp = sT[3 * p + 2];
      const double y0 = (Va0[0] - Pa.C[0]) + u * Aa[0] + v * Ba[0], y1 = (Va0[1] - Pa.C[1]) + u * Aa[1] + v * Ba[1],
                   y2 = (Va0[2] - Pa.C[2]) + u * Aa[2] + v * Ba[2];
      S a0_, au, av, b0_, bu, bv;
      szero(a0_); szero(au); szero(av); szero(b0_); szero(bu); szero(bv);
#pragma unroll 2
      for (int pp = 0; pp < HO_NPTS; pp++)
       { const double R0 = y0 - sX[3 * pp], R1 = y1 - sX[3 * pp + 1], R2 = y2 - sX[3 * pp + 2];
         const double r2 = fma(R2, R2, fma(R1, R1, R0 * R0));
         S Phi, Psi; szero(Psi);
         helm_fast<KMODE, PSI, S>(r2, SCUFF_OO4PI * sT[3 * pp + 2], A.K, Phi, Psi);
         const double up = sT[3 * pp], vp = sT[3 * pp + 1];
         sadd(a0_, Phi); fma_acc(au, up, Phi); fma_acc(av, vp, Phi);
         if (PSI) { sadd(b0_, Psi); fma_acc(bu, up, Psi); fma_acc(bv, vp, Psi); }
       }
      const double w0 = wp, w1 = wp * u, w2 = wp * v;
      fma_acc(mPhi[0][0], w0, a0_); fma_acc(mPhi[0][1], w0, au); fma_acc(mPhi[0][2], w0, av);
      fma_acc(mPhi[1][0], w1, a0_); fma_acc(mPhi[1][1], w1, au); fma_acc(mPhi[1][2], w1, av);
      fma_acc(mPhi[2][0], w2, a0_); fma_acc(mPhi[2][1], w2, au); fma_acc(mPhi[2][2], w2, av);
      if (PSI)
       { fma_acc(mPsi[0][0], w0, b0_); fma_acc(mPsi[0][1], w0, bu); fma_acc(mPsi[0][2], w0, bv);
         fma_acc(mPsi[1][0], w1, b0_); fma_acc(mPsi[1][1], w1, bu); fma_acc(mPsi[1][2], w1, bv);
         fma_acc(mPsi[2][0], w2, b0_); fma_acc(mPsi[2][1], w2, bu); fma_acc(mPsi[2][2], w2, bv); }
    }
#pragma unroll
  for (int m = 0; m < 3; m++)
#pragma unroll
   for (int k = 0; k < 3; k++) { mPhi[m][k] = wsum(mPhi[m][k]); if (PSI) mPsi[m][k] = wsum(mPsi[m][k]); }
  if (!live || lane != 0) return;
  int eia[3], eib[3]; double sla[3], slb[3];
#pragma unroll
  for (int m = 0; m < 3; m++) { eia[m] = Pa.ei[m]; eib[m] = Pb.ei[m]; sla[m] = Pa.sl[m]; slb[m] = Pb.sl[m]; }
  unsigned long long nPPI = 0;
  contract_scatter<S, PSI>(mPhi, mPsi, Va0, Aa, Ba, Vb0, Ab, Bb, eia, sla, eib, slb, N.pa, N.pb, A.S, A.K, A.S.symmetric, nPPI);
  if (nPPI) atomicAdd(A.counters + 1, nPPI);
}

template <int KMODE>
static void launch_ho_mode(scuffcuda_context *ctx, const NearArgs &A, const int *didx, int n, const double *dtcr, bool psi)
{ const size_t smem = (size_t)3 * HO_NPTS * (1 + HO_WARPS) * sizeof(double);
  const int blocks = (n + HO_WARPS - 1) / HO_WARPS;
  if (psi) ho_kernel<KMODE, true><<<blocks, 32 * HO_WARPS, smem, ctx->stream>>>(A, didx, n, dtcr);
  else ho_kernel<KMODE, false><<<blocks, 32 * HO_WARPS, smem, ctx->stream>>>(A, didx, n, dtcr);
}

int run_ho(scuffcuda_context *ctx, const NearArgs &A, const int *hoIdx, int n, int kmode)
{ if (n <= 0) return 0;
  static double *dtcr[64] = { nullptr };
  double *&T = dtcr[ctx->device & 63];
  if (!T)
   { if (cudaMalloc(&T, sizeof SC_TCR20_H) != cudaSuccess) { ctx->err = "high-order cubature table allocation failed"; return 1; }
     cudaMemcpy(T, SC_TCR20_H, sizeof SC_TCR20_H, cudaMemcpyHostToDevice); cudaDeviceSynchronize(); }
  if ((size_t)n > ctx->hoIdxCap)
   { if (ctx->dHoIdx) { cudaDeviceSynchronize(); cudaFree(ctx->dHoIdx); ctx->dHoIdx = nullptr; ctx->hoIdxCap = 0; }
     const size_t cap = std::max<size_t>(2 * (size_t)n, 4096);
     if (cudaMalloc(&ctx->dHoIdx, cap * sizeof(int)) != cudaSuccess) { ctx->err = "high-order cubature index allocation failed"; return 1; }
     ctx->hoIdxCap = cap; }
  int *didx = ctx->dHoIdx;                                      // persistent per context, grown on demand
  cudaMemcpyAsync(didx, hoIdx, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
  const bool psi = !(A.S.aPEC && A.S.bPEC);
  ctx->launches++;
  if (kmode == KMODE_REAL) launch_ho_mode<KMODE_REAL>(ctx, A, didx, n, T, psi);
  else if (kmode == KMODE_IMAG) launch_ho_mode<KMODE_IMAG>(ctx, A, didx, n, T, psi);
  else launch_ho_mode<KMODE_CPLX>(ctx, A, didx, n, T, psi);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);           // hoIdx is the caller's; the index buffer may be reused next call
  if (e != cudaSuccess) { ctx->err = std::string("ho_kernel: ") + cudaGetErrorString(e); return 1; }
  return 0;
}

// ---- self-test of fastmath.cuh (tests/test_fastmath_gpu.py) ----
__global__ void fastmath_selftest_kernel(int n, const double *x, double *out)
{ int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double a = fabs(x[i]) + 1e-300;
  double s, c; sincos_fast(x[i], s, c);
  out[6 * i + 0] = rsqrt_fast<true>(a); out[6 * i + 1] = rsqrt_fast<false>(a);
  out[6 * i + 2] = s; out[6 * i + 3] = c;
  out[6 * i + 4] = exp_neg_fast(-fabs(x[i])); out[6 * i + 5] = 0.0;
}
int run_fastmath_selftest(scuffcuda_context *ctx, int n, const double *xh, double *outh)
{ double *dx = nullptr, *dout = nullptr;
  if (cudaMalloc(&dx, n * sizeof(double)) != cudaSuccess || cudaMalloc(&dout, 6 * (size_t)n * sizeof(double)) != cudaSuccess)
   { cudaFree(dx); ctx->err = "selftest allocation failed"; return 1; }
  cudaMemcpyAsync(dx, xh, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  ctx->launches++;
  fastmath_selftest_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(n, dx, dout);
  cudaMemcpyAsync(outh, dout, 6 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  cudaFree(dx); cudaFree(dout);
  if (e != cudaSuccess) { ctx->err = std::string("fastmath selftest: ") + cudaGetErrorString(e); return 1; }
  return 0;
}

template <int KMODE, bool PSI, int MINB, bool SMALL>
static void launch_far_inst(scuffcuda_context *ctx, const FarArgs &A, dim3 grid, size_t smem)
{ // ask for the shared-memory carve-out that lets MINB CTAs (~46 KB each) share an SM
  cudaFuncSetAttribute(far_kernel<KMODE, PSI, MINB, SMALL>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  far_kernel<KMODE, PSI, MINB, SMALL><<<grid, FAR_THREADS, smem, ctx->stream>>>(A);
}
// resident CTAs per SM: 4 (128 registers) for the PEC variant, 3 (168) with the Psi half — measured best (3 / 4 / 5 tried)
template <int KMODE>
static void launch_far_mode(scuffcuda_context *ctx, const FarArgs &A, dim3 grid, size_t smem, bool psi, bool small)
{ static int psiMinb = -1;
  if (psiMinb < 0) { const char *e = getenv("SCUFF_CUDA_FAR_MINB_PSI"); psiMinb = e ? atoi(e) : 2; }   // 2 CTAs/SM (242 registers, no spills) measured 5 % faster than 3
  if (psi)
   { if (small) launch_far_inst<KMODE, true, 3, true>(ctx, A, grid, smem);
     else if (psiMinb == 2) launch_far_inst<KMODE, true, 2, false>(ctx, A, grid, smem);
     else launch_far_inst<KMODE, true, 3, false>(ctx, A, grid, smem); }
  else
   { static int pecMinb = -1;
     if (pecMinb < 0) { const char *e = getenv("SCUFF_CUDA_FAR_MINB"); pecMinb = e ? atoi(e) : 3; }   // 3 CTAs/SM (165 registers, no spills) measured 2-4 % faster than 4 with the two-point loop
     if (small) launch_far_inst<KMODE, false, 4, true>(ctx, A, grid, smem);
     else if (pecMinb == 3) launch_far_inst<KMODE, false, 3, false>(ctx, A, grid, smem);
     else launch_far_inst<KMODE, false, 4, false>(ctx, A, grid, smem); }
}

// largest panel radius of the surface whose device panels are `d`
static double surface_rmax(scuffcuda_context *ctx, const PanelRec *d)
{ for (const Surface &S : ctx->surfaces)
   if (S.dpanels == d) { double r = 0.0; for (const PanelRec &P : S.hpanels) r = std::max(r, P.R); return r; }
  return 1.0e300;
}

void launch_far(scuffcuda_context *ctx, const FarArgs &Ain, int kmode)
{ FarArgs A = Ain;
  // lanes: 128 a-panels (rows); tile: TB b-panels (columns) in shared memory.  Tile height: enough CTAs to fill the
  // machine on small meshes, 32 on large ones
  int aTiles = (A.NPa + FAR_THREADS - 1) / FAR_THREADS;
  const bool psi = !(A.S.aPEC && A.S.bPEC);     // PEC/PEC blocks stamp GC[0] only
  int TB = 32;
  while (TB > 2 && (long)aTiles * ((A.NPb + TB - 1) / TB) < 4L * ctx->smCount) TB >>= 1;
  A.TA = TB;
  dim3 grid(aTiles, (A.NPb + TB - 1) / TB);
  size_t smem = FAR_SMEM_BYTES(TB);
  // small-phase variant when |k| (R_a + R_b) <= 1/8 for every panel pair (fine meshes).  It executes 22 % fewer
  // instructions but measured the same duration on B200 (the kernel is latency-, not issue-bound at 4 CTAs/SM), so it is
  // opt-in: SCUFF_CUDA_FAR_SMALL=1
  static int smallEnv = -1;
  if (smallEnv < 0) { const char *e = getenv("SCUFF_CUDA_FAR_SMALL"); smallEnv = e ? atoi(e) : 0; }
  const double kk = (kmode == KMODE_REAL) ? fabs(A.K.kr) : fabs(A.K.ki);
  const bool small = smallEnv && kmode != KMODE_CPLX && kk * (surface_rmax(ctx, A.pa) + surface_rmax(ctx, A.pb)) <= 0.125;
  ctx->launches++;
  if (kmode == KMODE_REAL) launch_far_mode<KMODE_REAL>(ctx, A, grid, smem, psi, small);
  else if (kmode == KMODE_IMAG) launch_far_mode<KMODE_IMAG>(ctx, A, grid, smem, psi, small);
  else launch_far_mode<KMODE_CPLX>(ctx, A, grid, smem, psi, small);
}

}  // namespace scuffcuda
