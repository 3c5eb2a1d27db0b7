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
#include <cstdio>
#include <map>
#include <mutex>
#include <vector>
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
template <int KMODE, bool PSI, int DEG, typename S> struct HelmSmall;

// DEG = 1: |t| <= 1/8 (sin to t^9, cos to t^10; truncation < 3e-18), DEG = 2: |t| <= 1/16 (sin to t^7, cos to t^8; < 4e-17)
template <int DEG>
__device__ __forceinline__ void sincos_small(double t, double &s, double &c)
{ const double z = t * t;
  if (DEG == 1)
   { double ps = fma(z, 2.7557319223985893e-06, -1.9841269841269841e-04);       // 1/9!, -1/7!
     ps = fma(z, ps, 8.3333333333333332e-03); ps = fma(z, ps, -1.6666666666666666e-01);
     s = fma(t * z, ps, t);
     double pc = fma(z, -2.7557319223985888e-07, 2.4801587301587302e-05);        // -1/10!, 1/8!
     pc = fma(z, pc, -1.3888888888888889e-03); pc = fma(z, pc, 4.1666666666666664e-02); pc = fma(z, pc, -0.5);
     c = fma(z, pc, 1.0); }
  else
   { double ps = fma(z, -1.9841269841269841e-04, 8.3333333333333332e-03);        // -1/7!, 1/5!
     ps = fma(z, ps, -1.6666666666666666e-01);
     s = fma(t * z, ps, t);
     double pc = fma(z, 2.4801587301587302e-05, -1.3888888888888889e-03);        // 1/8!, -1/6!
     pc = fma(z, pc, 4.1666666666666664e-02); pc = fma(z, pc, -0.5);
     c = fma(z, pc, 1.0); }
}
template <int DEG>
__device__ __forceinline__ double exp_small(double s)                          // DEG 1: |s| <= 1/8, Taylor to s^10; DEG 2: |s| <= 1/16, to s^8
{ double p;
  if (DEG == 1)
   { p = 2.7557319223985888e-07;
     p = fma(p, s, 2.7557319223985893e-06); p = fma(p, s, 2.4801587301587302e-05); }
  else p = 2.4801587301587302e-05;
  p = fma(p, s, 1.9841269841269841e-04);
  p = fma(p, s, 1.3888888888888889e-03); p = fma(p, s, 8.3333333333333332e-03); p = fma(p, s, 4.1666666666666664e-02);
  p = fma(p, s, 1.6666666666666666e-01); p = fma(p, s, 0.5); p = fma(p, s, 1.0);
  return fma(p, s, 1.0);
}
template <int DEG> struct HelmSmall<KMODE_REAL, false, DEG, cd>
{ static __device__ __forceinline__ void eval(double r2, double cw, const KParams &K, double r0, cd &Phi, cd &Psi)
  { const double rinv = rsqrt_fast<false>(r2), r = r2 * rinv;
  double s, c; sincos_small<DEG>(K.kr * (r - r0), s, c);
  const double f = cw * rinv;
  Phi = mk(f * c, f * s);
  } };
template <int DEG> struct HelmSmall<KMODE_REAL, true, DEG, cd>
{ static __device__ __forceinline__ void eval(double r2, double cw, const KParams &K, double r0, cd &Phi, cd &Psi)
  { const double rinv = rsqrt_fast<false>(r2), r = r2 * rinv;
  double s, c; sincos_small<DEG>(K.kr * (r - r0), s, c);
  const double f = cw * rinv;
  Phi = mk(f * c, f * s);
  const double a = -rinv * rinv, b = K.kr * rinv;
  Psi = mk(fma(Phi.x, a, -Phi.y * b), fma(Phi.x, b, Phi.y * a));
  } };
template <int DEG> struct HelmSmall<KMODE_IMAG, false, DEG, double>
{ static __device__ __forceinline__ void eval(double r2, double cw, const KParams &K, double r0, double &Phi, double &Psi)
  { const double rinv = rsqrt_fast<false>(r2), r = r2 * rinv;
  Phi = cw * rinv * exp_small<DEG>(-K.ki * (r - r0));
  } };
template <int DEG> struct HelmSmall<KMODE_IMAG, true, DEG, double>
{ static __device__ __forceinline__ void eval(double r2, double cw, const KParams &K, double r0, double &Phi, double &Psi)
  { const double rinv = rsqrt_fast<false>(r2), r = r2 * rinv;
  Phi = cw * rinv * exp_small<DEG>(-K.ki * (r - r0));
  Psi = Phi * (-K.ki - rinv) * rinv;
  } };
// general complex k never takes the small-phase path (the launcher does not select it)
template <int DEG> struct HelmSmall<KMODE_CPLX, false, DEG, cd>
{ static __device__ __forceinline__ void eval(double r2, double cw, const KParams &K, double r0, cd &Phi, cd &Psi)
  { helm_fast<KMODE_CPLX, false, cd>(r2, cw, K, Phi, Psi); } };
template <int DEG> struct HelmSmall<KMODE_CPLX, true, DEG, cd>
{ static __device__ __forceinline__ void eval(double r2, double cw, const KParams &K, double r0, cd &Phi, cd &Psi)
  { helm_fast<KMODE_CPLX, true, cd>(r2, cw, K, Phi, Psi); } };

// (the small-phase pair factor)
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
  // H0_ij = sum_{al,be} d_{al be} <(x_al - [i==al]) (y_be - [j==be]) Phi> + c4 <Phi>
  //       = T - [j>0] P_j - [i>0] Q_i + [i>0][j>0] d_ij m00 + c4 m00 ,   T = sum d_{al be} m_{al be} (common to all nine),
  //         P_j = sum_al d_{al j} m_{al 0} ,  Q_i = sum_be d_{i be} m_{0 be}
  S T0, P1, P2, Q1, Q2;
  szero(T0); szero(P1); szero(P2); szero(Q1); szero(Q2);
  fma_acc(T0, dAA, mPhi[1][1]); fma_acc(T0, dAB, mPhi[1][2]); fma_acc(T0, dBA, mPhi[2][1]); fma_acc(T0, dBB, mPhi[2][2]);
  fma_acc(P1, dAA, mPhi[1][0]); fma_acc(P1, dBA, mPhi[2][0]);
  fma_acc(P2, dAB, mPhi[1][0]); fma_acc(P2, dBB, mPhi[2][0]);
  fma_acc(Q1, dAA, mPhi[0][1]); fma_acc(Q1, dAB, mPhi[0][2]);
  fma_acc(Q2, dBA, mPhi[0][1]); fma_acc(Q2, dBB, mPhi[0][2]);
  // rows of the a-side shifted Psi moments: row 1 ~ (u - [i==1]), row 2 ~ (v - [i==2])
  S q1s[3], q2s[3];
  if (PSI)
   {
#pragma unroll
     for (int n = 0; n < 3; n++) { q1s[n] = ssub(mPsi[1][n], mPsi[0][n]); q2s[n] = ssub(mPsi[2][n], mPsi[0][n]); } }
  const cd c4m = K.c4 * tocd(mPhi[0][0]);
  // triple products for H1
  double tAA = 0, tAB = 0, tBA = 0, tBB = 0, AbN = 0, BbN = 0, AaNp = 0, BaNp = 0;
  if (PSI)
   { double Na[3], Nb[3], eA[3], eB[3];
     cross3(Aa, Ba, Na); cross3(Ab, Bb, Nb);
     cross3(E00, Aa, eA); cross3(E00, Ba, eB);         // (X x Y).E = Y.(E x X)
     tAA = dot3(Ab, eA); tAB = dot3(Bb, eA); tBA = dot3(Ab, eB); tBB = dot3(Bb, eB);
     AbN = dot3(Ab, Na); BbN = dot3(Bb, Na); AaNp = dot3(Aa, Nb); BaNp = dot3(Ba, Nb); }

  if (!PSI)
   { // PEC/PEC blocks (one entry per edge pair, SurfaceSurfaceInteractions.cc:205-217).  Two phases: first ALL nine values and
     // addresses into their own registers, then the atomics back to back.  Interleaving value computation and RED made every
     // entry wait for the previous RED to release its operand registers (ncu: half of all stall samples were scoreboard
     // waits in this epilogue, the kernel time did not move when the cubature loop lost 27 % of its instructions).
     cd val[9]; long idx[9];
#pragma unroll
     for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++)
       { const int n = 3 * i + j;
         int ea = eia[i], eb = eib[j];
         idx[n] = -1;
         S h0 = T0;
         if (j == 1) h0 = ssub(h0, P1); else if (j == 2) h0 = ssub(h0, P2);
         if (i == 1) h0 = ssub(h0, Q1); else if (i == 2) h0 = ssub(h0, Q2);
         if (i > 0 && j > 0) fma_acc(h0, (i == 1) ? (j == 1 ? dAA : dAB) : (j == 1 ? dBA : dBB), mPhi[0][0]);
         double sc = sla3[i] * slb[j];
         bool ok = (ea >= 0 && eb >= 0);
         if (SP.symmetric)
          { if (pa == pb) ok = ok && (ea <= eb);
            else if (ea == eb) sc *= 2.0;
            else if (ea > eb) { const int t = ea; ea = eb; eb = t; } }
         const int col = SP.colOff + eb;
         ok = ok && col >= SP.colLo && col < SP.colHi;
         val[n] = SP.c1 * (sc * (tocd(h0) + c4m));
         if (ok) { idx[n] = (long)(SP.rowOff + ea) + (long)col * SP.ld; nPPI += (symmetric && pa != pb && eia[i] == eib[j]) ? 2 : 1; } }
     // materialise the nine addresses in nine register pairs BEFORE the first RED (the compiler otherwise recomputes each address
     // into the same pair just in time, which serialises the REDs on that pair)
     double *ptr[9];
#pragma unroll
     for (int n = 0; n < 9; n++) ptr[n] = SP.M + (SP.isReal ? 1 : 2) * (idx[n] >= 0 ? idx[n] : 0);
     asm volatile("" : "+l"(ptr[0]), "+l"(ptr[1]), "+l"(ptr[2]), "+l"(ptr[3]), "+l"(ptr[4]), "+l"(ptr[5]), "+l"(ptr[6]), "+l"(ptr[7]), "+l"(ptr[8]));
     asm volatile("" : "+d"(val[0].x), "+d"(val[1].x), "+d"(val[2].x), "+d"(val[3].x), "+d"(val[4].x), "+d"(val[5].x), "+d"(val[6].x), "+d"(val[7].x), "+d"(val[8].x));
     if (SP.isReal)
      {
#pragma unroll
        for (int n = 0; n < 9; n++) if (idx[n] >= 0) atomicAdd(ptr[n], val[n].x); }
     else
      { asm volatile("" : "+d"(val[0].y), "+d"(val[1].y), "+d"(val[2].y), "+d"(val[3].y), "+d"(val[4].y), "+d"(val[5].y), "+d"(val[6].y), "+d"(val[7].y), "+d"(val[8].y));
#pragma unroll
        for (int n = 0; n < 9; n++) if (idx[n] >= 0) { atomicAdd(ptr[n], val[n].x); atomicAdd(ptr[n] + 1, val[n].y); } }
     return; }

#pragma unroll
  for (int i = 0; i < 3; i++)
   { const int ea = eia[i];
     if (ea < 0) continue;
     const double sla = sla3[i];
     const S *UQ = (i == 1) ? q1s : mPsi[1], *VQ = (i == 2) ? q2s : mPsi[2];
#pragma unroll
     for (int j = 0; j < 3; j++)
      { if (eib[j] < 0) continue;
        S h0 = T0;
        if (j == 1) h0 = ssub(h0, P1); else if (j == 2) h0 = ssub(h0, P2);
        if (i == 1) h0 = ssub(h0, Q1); else if (i == 2) h0 = ssub(h0, Q2);
        if (i > 0 && j > 0) fma_acc(h0, (i == 1) ? (j == 1 ? dAA : dAB) : (j == 1 ? dBA : dBB), mPhi[0][0]);
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

// ---- 3x3 barycentric moments of one far panel pair: m[m over a (outer loop: this lane's points, thread-private shared-memory
//      column sXa)][n over b (inner loop: the tile panel's points Pb, broadcast shared-memory reads)] ----
template <int KMODE, bool PSI, int SMALL, typename S>
__device__ __forceinline__ void far_moments(const double *sXa, const double *Pb, const double (&D)[3], const KParams &K, double r0,
                                            S (&mPhi)[3][3], S (&mPsi)[3][3])
{
     // ---- 3x3 barycentric moments m[m over a (outer loop: this lane's points)][n over b (inner loop: tile panel's points,
     //      broadcast shared-memory reads)] ----
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
           if (SMALL) { HelmSmall<KMODE, PSI, (SMALL > 0 ? SMALL : 1), S>::eval(r2, SCUFF_OO4PI * tcr_w(pp), K, r0, Phi, Psi);
                        HelmSmall<KMODE, PSI, (SMALL > 0 ? SMALL : 1), S>::eval(q2, SCUFF_OO4PI * tcr_w(pp), K, r0, Phj, Psj); }
           else { helm_fast<KMODE, PSI, S>(r2, SCUFF_OO4PI * tcr_w(pp), K, Phi, Psi);
                  helm_fast<KMODE, PSI, S>(q2, SCUFF_OO4PI * tcr_w(pp), K, Phj, Psj); }
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

}

// the small-phase pair factor exp(ik r0)
template <int KMODE, typename S>
__device__ __forceinline__ S far_pair_factor(const KParams &K, double r0)
{ S E0;
  if (KMODE == KMODE_IMAG) { const double e = exp_neg_fast(-K.ki * r0); E0 = *reinterpret_cast<const S *>(&e); }
  else { double s0, c0; sincos_fast(K.kr * r0, s0, c0); const cd e = mk(c0, s0); E0 = *reinterpret_cast<const S *>(&e); }
  return E0;
}

// shared-memory layout (doubles): b-tile [TB][36] | lane (a) cubature points [18][128] | lane stash [16][128] | lane edge ids (int)[3][128]
#define FAR_SMEM_BYTES(TA) ((size_t)((TA) * 36 + 34 * FAR_THREADS) * sizeof(double) + 6 * FAR_THREADS * sizeof(int))

// Thread mapping: lane <-> a-panel (matrix ROW side), tile <-> b-panels (column side).  The matrix is column-major, so
// the RED.ADD.F64 of one warp for entry (e_a(lane), e_b) go to neighbouring rows of ONE column (edges of consecutive
// panels are nearly consecutive) instead of 32 different columns: ~4-8x fewer L2 atomic sectors per warp instruction.
template <int KMODE, bool PSI, int MINB, int SMALL>
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
     S mPhi[3][3], mPsi[3][3];
     far_moments<KMODE, PSI, SMALL, S>(sXa, Pb, D, A.K, r0, mPhi, mPsi);
     if (SMALL)
      { // the pair factor exp(ik r0), once for all moments
        const S E0 = far_pair_factor<KMODE, S>(A.K, r0);
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
// Staged scatter for PEC/PEC blocks.  Round-1 profile: the nine complex RED.ADD.F64 pairs per panel pair (1.64 G scalar
// atomics at N = 20 280) occupy the LSU for as long as the cubature occupies the FP64 pipe (measured alone: 8.4 ms,
// tools/probes/bulkred_probe.cu), and the shared-memory loads of the cubature loop queue behind them, so the two phases
// did not overlap.  Here a warp (32 consecutive a-panels x one b-panel) first adds its <= 96 x 3 contributions inside a
// dense shared-memory window [rows rowLo .. rowLo+span) x [3 columns = the b-panel's edges]: the edges of 32 consecutive
// panels lie within ~60-90 consecutive matrix rows (the reference numbers an edge when it meets it the second time,
// InitEdgeList.cc:129-204), and the two panels of an interior edge that sit in the same warp are summed there.  One lane then
// sends each column of the window with ONE bulk reduction (cp.reduce.async.bulk .add.f64, SASS UBLKRED: the TMA unit
// streams the window to the L2 atomic units; nothing goes through the LSU).  Edges outside the window (5 % of the
// slots), columns at the diagonal of a symmetric block (where the (min,max) swap applies) and pruned columns keep the scalar
// RED path.  The window of every 32-panel group and the slot of every (panel, edge) are computed once per surface on the
// host (far_slots()).
// ------------------------------------------------------------------------------------------------------------------
#define FAR_SPAN 96
#define FAR_SLOT_SECOND 0x4000   /* the other panel of this edge sits in a lower lane of the same warp: add to its value */
#define FAR_SLOT_NONE (-1)       /* exterior edge: no basis function */
#define FAR_SLOT_OUTLIER (-2)    /* -2 - r: edge outside the warp's window, r = its rank among the warp's outlier slots */
#define FAR_EX_ROWS 16           /* outlier slots handled per flush round */
#define FAR_WIN_BYTES 4736       /* per warp: complex 3 x 96 x 16 = 4608 B; real 2 parities x 3 x 98 x 8 = 4704 B */
#define FAR_STAGE_BYTES (FAR_WIN_BYTES + FAR_EX_ROWS * 64)
struct __align__(16) FarSlot { short s[3]; short span; int rowLo; int nOut; };
struct __align__(16) FarExRow { cd v[3]; int edge; int pad[3]; };    // one outlier slot: its three column values and its edge (-1: nothing)
struct FarArgsX { FarArgs A; const FarSlot *slots; long nElem; int dbg; };

__device__ __forceinline__ void bulk_red_f64(double *dst, uint32_t srcSmem, uint32_t bytes)
{ asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;"
               ::"l"(dst), "r"(srcSmem), "r"(bytes) : "memory"); }
// a warp-uniform value as the compiler can prove it (lane 0's copy): operands of the bulk reduction live in uniform registers,
// per-lane copies made ptxas wrap every UBLKRED in a waterfall loop
__device__ __forceinline__ int uni(int v) { return __shfl_sync(0xffffffffu, v, 0); }

__device__ __forceinline__ cd cscale(cd c, double a) { return mk(c.x * a, c.y * a); }
__device__ __forceinline__ cd cscale(cd c, cd a) { return c * a; }

// the nine PEC/PEC values  c1 l_a l_b [ <(F.F') Phi> + 4/(ik)^2 <Phi> ]  from the 3x3 moments:
//   H0_ij = T - [j>0] P_j - [i>0] Q_i + [i>0][j>0] d_ij m00 + c4 m00 ,  T = sum_{al,be} d_{al be} m_{al be} (common to all nine),
//   P_j = sum_al d_{al j} m_{al 0} ,  Q_i = sum_be d_{i be} m_{0 be} ,  d = {A,B}.{A',B'}
template <typename S>
__device__ __forceinline__ void contract_pec(const S (&m)[3][3], const double *Aa, const double *Ba, const double *Ab, const double *Bb,
                                             const double *sla, const double *slb, cd c1, cd c4, cd (&val)[9])
{ const double dAA = dot3(Aa, Ab), dAB = dot3(Aa, Bb), dBA = dot3(Ba, Ab), dBB = dot3(Ba, Bb);
  S T0, P1, P2, Q1, Q2;
  szero(T0); szero(P1); szero(P2); szero(Q1); szero(Q2);
  fma_acc(T0, dAA, m[1][1]); fma_acc(T0, dAB, m[1][2]); fma_acc(T0, dBA, m[2][1]); fma_acc(T0, dBB, m[2][2]);
  fma_acc(P1, dAA, m[1][0]); fma_acc(P1, dBA, m[2][0]);
  fma_acc(P2, dAB, m[1][0]); fma_acc(P2, dBB, m[2][0]);
  fma_acc(Q1, dAA, m[0][1]); fma_acc(Q1, dAB, m[0][2]);
  fma_acc(Q2, dBA, m[0][1]); fma_acc(Q2, dBB, m[0][2]);
  const cd m00 = cscale(c1, m[0][0]);
  const cd cT = cscale(c1, T0) + c4 * m00;
  const cd cP1 = cscale(c1, P1), cP2 = cscale(c1, P2), cQ1 = cscale(c1, Q1), cQ2 = cscale(c1, Q2);
  const cd TP[3] = { cT, cT - cP1, cT - cP2 };
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
   for (int j = 0; j < 3; j++)
    { cd h = TP[j];
      if (i > 0)
       { h = h - (i == 1 ? cQ1 : cQ2);
         if (j > 0) { const double d = (i == 1) ? (j == 1 ? dAA : dAB) : (j == 1 ? dBA : dBB); h.x = fma(d, m00.x, h.x); h.y = fma(d, m00.y, h.y); } }
      val[3 * i + j] = (sla[i] * slb[j]) * h; }
}

// one entry through the scalar path, with the upper-triangle convention of Symmetric blocks (stamp.cuh scatter_ppi); returns
// the number of reference PPIs it stands for
__device__ __forceinline__ unsigned int far_scalar_entry(const StampParams &SP, int pa, int pb, int ea, int eb, cd v)
{ unsigned int cnt = 1;
  if (SP.symmetric)
   { if (pa == pb) { if (ea > eb) return 0; }
     else if (ea == eb) { v = 2.0 * v; cnt = 2; }
     else if (ea > eb) { const int t = ea; ea = eb; eb = t; } }
  const int col = SP.colOff + eb;
  if (col < SP.colLo || col >= SP.colHi) return 0;
  const size_t idx = (size_t)(SP.rowOff + ea) + (size_t)col * (size_t)SP.ld;
  if (SP.isReal) atomicAdd(SP.M + idx, v.x);
  else { atomicAdd(SP.M + 2 * idx, v.x); atomicAdd(SP.M + 2 * idx + 1, v.y); }
  return cnt;
}

#define FAR_STAGED_SMEM_BYTES(TA) (FAR_SMEM_BYTES(TA) + (FAR_THREADS / 32) * FAR_STAGE_BYTES)

template <int KMODE, int SMALL>
__global__ void __launch_bounds__(FAR_THREADS, 3)
far_staged_kernel(const __grid_constant__ FarArgsX X)
{ const FarArgs &A = X.A;
  extern __shared__ __align__(128) double sm[];
  __shared__ __align__(8) uint64_t bar;
  typedef typename KScalar<KMODE>::type S;
  const int TB = A.TA;
  const int b0 = blockIdx.y * TB;
  const int aBase = blockIdx.x * FAR_THREADS;
  const int nB = min(TB, A.NPb - b0);
  if (A.symmetric && b0 + nB - 1 < aBase) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = uni(tid >> 5);
  double *sXa = sm + TB * 36 + tid;                    // element k at sXa[k * FAR_THREADS]
  double *sSt = sm + TB * 36 + 18 * FAR_THREADS + tid;
  int *sEi = (int *)(sm + TB * 36 + 34 * FAR_THREADS) + tid;
  unsigned int *sCnt = (unsigned int *)sEi + 3 * FAR_THREADS;
  char *stage = (char *)(sm + TB * 36 + 34 * FAR_THREADS) + 6 * FAR_THREADS * sizeof(int) + warp * FAR_STAGE_BYTES;
  FarExRow *ex = reinterpret_cast<FarExRow *>(stage + FAR_WIN_BYTES);
  const uint32_t stageU = smem_u32(stage);
  sCnt[0] = 0; sCnt[FAR_THREADS] = 0;

  if (tid == 0)
   { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
     asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (tid == 0) tma_load_1d(sm, A.pb + b0, (uint32_t)(nB * sizeof(PanelRec)), &bar);

  const int pa = aBase + tid;
  const bool live = pa < A.NPa && !(A.symmetric && pa > b0 + nB - 1);
  const PanelRec *Pa = A.pa + (pa < A.NPa ? pa : 0);
#pragma unroll
  for (int i = 0; i < 18; i++) sXa[i * FAR_THREADS] = Pa->x[i];
#pragma unroll
  for (int m = 0; m < 3; m++)
   { const double v0 = Pa->V[m];
     sSt[(3 + m) * FAR_THREADS] = Pa->V[3 + m] - v0;                 // A
     sSt[(6 + m) * FAR_THREADS] = Pa->V[6 + m] - v0;                 // B
     sSt[(9 + m) * FAR_THREADS] = Pa->sl[m];
     sSt[(12 + m) * FAR_THREADS] = Pa->C[m];
     sEi[m * FAR_THREADS] = (pa < A.NPa) ? Pa->ei[m] : -1; }
  sSt[15 * FAR_THREADS] = Pa->R;
  // this lane's slots in the warp's window (the table is padded to a multiple of FAR_THREADS panels; every lane of a
  // warp, live or not, takes part in the staging so that a slot's first writer always refreshes it)
  const FarSlot fs = X.slots[pa];
  const int sl0 = fs.s[0], sl1 = fs.s[1], sl2 = fs.s[2];
  const int rowLo = uni(fs.rowLo), span = uni(fs.span), nOut = uni(fs.nOut);
  const int nIn = (sl0 >= 0) + (sl1 >= 0) + (sl2 >= 0);
  const bool hasOut = (sl0 <= FAR_SLOT_OUTLIER) || (sl1 <= FAR_SLOT_OUTLIER) || (sl2 <= FAR_SLOT_OUTLIER);
  for (int i = lane; i < FAR_STAGE_BYTES / 16; i += 32) reinterpret_cast<double2 *>(stage)[i] = make_double2(0.0, 0.0);
  __syncwarp();

  mbar_wait(&bar, 0);
  if (!__any_sync(0xffffffffu, live)) return;

  const long rowBase = (long)A.S.rowOff + rowLo;
  bool pending = false;                                    // warp-uniform: a bulk group of this warp may still read the stage
  // SCUFF_CUDA_FAR_DBG & 16: cycle counts of the phases of an iteration (lane 0 of every warp), summed into counters[8..13]
  const bool prof = (X.dbg & 16) != 0;
  long long tc[6] = { 0, 0, 0, 0, 0, 0 }, tPrev = prof ? clock64() : 0;
#define FAR_TICK(k) do { if (prof) { const long long t_ = clock64(); tc[k] += t_ - tPrev; tPrev = t_; } } while (0)
  // every warp walks the tile's b-panels from its own starting point: warps (and CTAs of one grid row) that march through the
  // same b-panels in step would send their windows — overlapping row ranges of the same three columns — to the same L2
  // sectors at the same moment, and same-address atomics serialise in the L2
  const int ibRot = (X.dbg & 8) ? 0 : (int)((blockIdx.x * 4u + (unsigned)warp) * 7u) % nB;
  for (int it = 0; it < nB; it++)
   { int ib = it + ibRot; if (ib >= nB) ib -= nB;
     const int pb = b0 + ib;
     const double *Pb = sm + ib * 36;
     double D[3];
     bool valid = live && !(A.symmetric && pb < pa);
     { const double Ca[3] = { sSt[12 * FAR_THREADS], sSt[13 * FAR_THREADS], sSt[14 * FAR_THREADS] };
       valid = pair_is_far(Ca, sSt[15 * FAR_THREADS], Pb, Pb[3], A.useDisp != 0, A.L, D) && valid; }
     const int *eb3 = (const int *)(Pb + 34);
     const int eb0 = uni(eb3[0]), eb1 = uni(eb3[1]), eb2 = uni(eb3[2]);
     if (A.S.colHi != 0x7fffffff && valid)
      { // column-block assembly: skip pairs none of whose (<= 9) entries falls into the owned column range (exact per entry)
        bool any = false;
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
         for (int j = 0; j < 3; j++)
          { const int e = sEi[i * FAR_THREADS], f = (j == 0) ? eb0 : (j == 1 ? eb1 : eb2);
            const int c0 = A.S.colOff + (A.symmetric ? max(e, f) : f);
            any |= (e >= 0 && f >= 0 && c0 >= A.S.colLo && c0 < A.S.colHi); }
        valid = any; }
     if (!__any_sync(0xffffffffu, valid)) continue;

     cd val[9];
#pragma unroll
     for (int n = 0; n < 9; n++) val[n] = mk(0.0, 0.0);
     if (valid && (X.dbg & 32))
      { sCnt[0] += 1;
#pragma unroll
        for (int n = 0; n < 9; n++) val[n] = mk(D[0] + n, D[1] - n); }      // timing experiment: scatter only, no cubature
     else if (valid)
      { sCnt[0] += 1;
        double r0 = 0.0;
        if (SMALL) { const double d2 = fma(D[2], D[2], fma(D[1], D[1], D[0] * D[0])); r0 = d2 * rsqrt_fast<true>(d2); }
        S mPhi[3][3], mPsi[3][3];
        far_moments<KMODE, false, SMALL, S>(sXa, Pb, D, A.K, r0, mPhi, mPsi);
        cd c1 = A.S.c1;
        if (SMALL) c1 = cscale(c1, far_pair_factor<KMODE, S>(A.K, r0));   // the pair factor exp(ik r0) rides on the prefactor
        double sla[3], Aa[3], Ba[3], Ab[3], Bb[3];
#pragma unroll
        for (int m = 0; m < 3; m++)
         { sla[m] = sSt[(9 + m) * FAR_THREADS]; Aa[m] = sSt[(3 + m) * FAR_THREADS]; Ba[m] = sSt[(6 + m) * FAR_THREADS];
           Ab[m] = Pb[7 + m] - Pb[4 + m]; Bb[m] = Pb[10 + m] - Pb[4 + m]; }
        contract_pec<S>(mPhi, Aa, Ba, Ab, Bb, sla, Pb + 31, c1, A.K.c4, val); }

     // ---- dense part: window rows x the b-panel's three columns (all decisions warp-uniform) ----
     bool dense[3]; int par[3]; long gcol[3];
#pragma unroll
     for (int j = 0; j < 3; j++)
      { const int eb = (j == 0) ? eb0 : (j == 1 ? eb1 : eb2);
        const int col = A.S.colOff + eb;
        dense[j] = span > 0 && eb >= 0 && col >= A.S.colLo && col < A.S.colHi && (!A.symmetric || rowLo + span - 1 < eb);
        gcol[j] = rowBase + (long)col * A.S.ld;           // element index of the window's first row in this column
        par[j] = 0;
        if (A.S.isReal)
         { // 8-byte entries: the bulk reduction needs 16-byte aligned ends, so the window is widened to even element indices
           // (the extra elements are zero padding of the stage) and must stay inside the matrix
           par[j] = (int)(gcol[j] & 1);
           dense[j] = dense[j] && (gcol[j] - par[j] + ((par[j] + span + 1) & ~1) <= X.nElem); } }
     const bool anyDense = dense[0] || dense[1] || dense[2];
     const bool flushOut = nOut > 0 && (eb0 >= 0 || eb1 >= 0 || eb2 >= 0) && !(X.dbg & 64);
     FAR_TICK(0);                                          // far test, cubature, contraction
     if (anyDense && pending) { if (lane < 3) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); __syncwarp(); pending = false; }
     FAR_TICK(1);                                          // wait for the previous bulk group's reads
     if (anyDense)
      { if (!A.S.isReal)
         { cd *st = reinterpret_cast<cd *>(stage);
#pragma unroll
           for (int i = 0; i < 3; i++)
            { const int s = (i == 0) ? sl0 : (i == 1 ? sl1 : sl2);
              if (s >= 0 && !(s & FAR_SLOT_SECOND))
               {
#pragma unroll
                 for (int j = 0; j < 3; j++) if (dense[j]) *reinterpret_cast<double2 *>(st + j * FAR_SPAN + s) = make_double2(val[3 * i + j].x, val[3 * i + j].y); } }
           __syncwarp();
#pragma unroll
           for (int i = 0; i < 3; i++)
            { const int s = (i == 0) ? sl0 : (i == 1 ? sl1 : sl2);
              if (s >= 0 && (s & FAR_SLOT_SECOND))
               { const int s2 = s & (FAR_SLOT_SECOND - 1);
#pragma unroll
                 for (int j = 0; j < 3; j++)
                  if (dense[j])
                   { double2 *q = reinterpret_cast<double2 *>(st + j * FAR_SPAN + s2);
                     double2 t = *q; t.x += val[3 * i + j].x; t.y += val[3 * i + j].y; *q = t; } } }
           asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
           __syncwarp();
           FAR_TICK(2);                                    // stage writes, fence
           if (X.dbg & 4)
            { if (lane < 3)
               { const int j = lane;
                 if (dense[j]) bulk_red_f64(A.S.M + 2 * gcol[j], stageU + (uint32_t)(j * FAR_SPAN * 16), (uint32_t)span * 16u);
                 asm volatile("cp.async.bulk.commit_group;" ::: "memory"); } }
           else if (lane == 0 && !(X.dbg & 1))
            {
#pragma unroll
              for (int j = 0; j < 3; j++)
               if (dense[j])
                { if (X.dbg & 2) asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                              ::"l"(A.S.M + 2 * gcol[j]), "r"(stageU + (uint32_t)(j * FAR_SPAN * 16)), "r"((uint32_t)span * 16u) : "memory");
                  else bulk_red_f64(A.S.M + 2 * gcol[j], stageU + (uint32_t)(j * FAR_SPAN * 16), (uint32_t)span * 16u); }
              asm volatile("cp.async.bulk.commit_group;" ::: "memory"); } }
        else
         { double *st = reinterpret_cast<double *>(stage);    // [parity][column][FAR_SPAN + 2], element of slot s at index parity + s
#pragma unroll
           for (int i = 0; i < 3; i++)
            { const int s = (i == 0) ? sl0 : (i == 1 ? sl1 : sl2);
              if (s >= 0 && !(s & FAR_SLOT_SECOND))
               {
#pragma unroll
                 for (int j = 0; j < 3; j++) if (dense[j]) st[(par[j] * 3 + j) * (FAR_SPAN + 2) + par[j] + s] = val[3 * i + j].x; } }
           __syncwarp();
#pragma unroll
           for (int i = 0; i < 3; i++)
            { const int s = (i == 0) ? sl0 : (i == 1 ? sl1 : sl2);
              if (s >= 0 && (s & FAR_SLOT_SECOND))
               { const int s2 = s & (FAR_SLOT_SECOND - 1);
#pragma unroll
                 for (int j = 0; j < 3; j++) if (dense[j]) st[(par[j] * 3 + j) * (FAR_SPAN + 2) + par[j] + s2] += val[3 * i + j].x; } }
           asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
           __syncwarp();
           if (lane == 0)
            {
#pragma unroll
              for (int j = 0; j < 3; j++)
               if (dense[j]) bulk_red_f64(A.S.M + (gcol[j] - par[j]), stageU + (uint32_t)((par[j] * 3 + j) * (FAR_SPAN + 2) * 8),
                                          (uint32_t)((par[j] + span + 1) & ~1) * 8u);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory"); } }
        pending = true; }

     FAR_TICK(3);                                          // bulk reductions issued
     unsigned int nPPI = 0;
     if (valid) nPPI = (unsigned int)nIn * ((unsigned int)dense[0] + (unsigned int)dense[1] + (unsigned int)dense[2]);
     // ---- outlier slots (edges outside the window): their values are collected in the warp's extra rows and flushed with all
     //      lanes busy, FAR_EX_ROWS slots per round (one lane per (slot, column) record) ----
     if (flushOut)
      { for (int r0 = 0; r0 < nOut; r0 += FAR_EX_ROWS)
         { if (hasOut)
            {
#pragma unroll
              for (int i = 0; i < 3; i++)
               { const int s = (i == 0) ? sl0 : (i == 1 ? sl1 : sl2);
                 const int r = FAR_SLOT_OUTLIER - s - r0;                  // rank of this slot inside the round
                 if (s <= FAR_SLOT_OUTLIER && r >= 0 && r < FAR_EX_ROWS)
                  { FarExRow *E = ex + r;
#pragma unroll
                    for (int j = 0; j < 3; j++) *reinterpret_cast<double2 *>(&E->v[j]) = make_double2(val[3 * i + j].x, val[3 * i + j].y);
                    E->edge = valid ? sEi[i * FAR_THREADS] : -1; } } }
           __syncwarp();
           const int nRec = 3 * min(FAR_EX_ROWS, nOut - r0);
           for (int t = lane; t < nRec; t += 32)
            { const int r = t / 3, j = t - 3 * r;
              const int ea = ex[r].edge, eb = (j == 0) ? eb0 : (j == 1 ? eb1 : eb2);
              if (ea < 0 || eb < 0) continue;
              const double2 v = *reinterpret_cast<const double2 *>(&ex[r].v[j]);
              // the pair's own panel indices only matter for pa == pb, which is never a far pair
              nPPI += far_scalar_entry(A.S, 0, 1, ea, eb, mk(v.x, v.y)); }
           __syncwarp(); } }
     // ---- columns that cannot take the dense path (at the diagonal of a symmetric block, the last rows of a real matrix):
     //      in-window slots through scalar atomics ----
     if ((eb0 >= 0 && !dense[0]) || (eb1 >= 0 && !dense[1]) || (eb2 >= 0 && !dense[2]))
      { if (valid)
         {
#pragma unroll
           for (int i = 0; i < 3; i++)
            { const int s = (i == 0) ? sl0 : (i == 1 ? sl1 : sl2);
              if (s < 0) continue;
              const int ea = sEi[i * FAR_THREADS];
#pragma unroll
              for (int j = 0; j < 3; j++)
               { const int eb = (j == 0) ? eb0 : (j == 1 ? eb1 : eb2);
                 if (eb < 0 || dense[j]) continue;
                 nPPI += far_scalar_entry(A.S, pa, pb, ea, eb, val[3 * i + j]); } } } }
     sCnt[FAR_THREADS] += nPPI;
     FAR_TICK(4);                                          // outlier flush, scalar leftovers
   }
  if (pending && lane < 3) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  FAR_TICK(5);
  if (prof && lane == 0) for (int k = 0; k < 6; k++) atomicAdd(A.counters + 8 + k, (unsigned long long)tc[k]);
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
  static std::mutex dtcrLock;
  std::unique_lock<std::mutex> tlock(dtcrLock);
  double *&T = dtcr[ctx->device & 63];
  if (!T)
   { if (cudaMalloc(&T, sizeof SC_TCR20_H) != cudaSuccess) { ctx->err = "high-order cubature table allocation failed"; return 1; }
     cudaMemcpy(T, SC_TCR20_H, sizeof SC_TCR20_H, cudaMemcpyHostToDevice); cudaDeviceSynchronize(); }
  tlock.unlock();
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
  out[6 * i + 4] = exp_neg_fast(-fabs(x[i])); out[6 * i + 5] = 0.0;   // column 5: td_div_selftest_kernel (td_kernel.cu)
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

template <int KMODE, bool PSI, int MINB, int SMALL>
static void launch_far_inst(scuffcuda_context *ctx, const FarArgs &A, dim3 grid, size_t smem)
{ // ask for the shared-memory carve-out that lets MINB CTAs (~46 KB each) share an SM
  cudaFuncSetAttribute(far_kernel<KMODE, PSI, MINB, SMALL>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  far_kernel<KMODE, PSI, MINB, SMALL><<<grid, FAR_THREADS, smem, ctx->stream>>>(A);
}
// resident CTAs per SM and small-phase degree: measured defaults, overridable for experiments with SCUFF_CUDA_FAR_MINB (PEC: 3|4),
// SCUFF_CUDA_FAR_MINB_PSI (2|3)
template <int KMODE, int SMALL>
static void launch_far_small(scuffcuda_context *ctx, const FarArgs &A, dim3 grid, size_t smem, bool psi)
{ static int psiMinb = -1, pecMinb = -1;
  if (psiMinb < 0) { const char *e = getenv("SCUFF_CUDA_FAR_MINB_PSI"); psiMinb = e ? atoi(e) : (SMALL ? 3 : 2); }
  if (pecMinb < 0) { const char *e = getenv("SCUFF_CUDA_FAR_MINB"); pecMinb = e ? atoi(e) : (SMALL ? 4 : 3); }
  if (psi)
   { if (psiMinb == 2) launch_far_inst<KMODE, true, 2, SMALL>(ctx, A, grid, smem);
     else launch_far_inst<KMODE, true, 3, SMALL>(ctx, A, grid, smem); }
  else
   { if (pecMinb == 3) launch_far_inst<KMODE, false, 3, SMALL>(ctx, A, grid, smem);
     else launch_far_inst<KMODE, false, 4, SMALL>(ctx, A, grid, smem); }
}
template <int KMODE>
static void launch_far_mode(scuffcuda_context *ctx, const FarArgs &A, dim3 grid, size_t smem, bool psi, int small)
{ if (small == 2) launch_far_small<KMODE, 2>(ctx, A, grid, smem, psi);
  else if (small == 1) launch_far_small<KMODE, 1>(ctx, A, grid, smem, psi);
  else launch_far_small<KMODE, 0>(ctx, A, grid, smem, psi);
}

// largest panel radius of the surface whose device panels are `d`
static double surface_rmax(scuffcuda_context *ctx, const PanelRec *d)
{ for (const Surface &S : ctx->surfaces)
   if (S.dpanels == d) { double r = 0.0; for (const PanelRec &P : S.hpanels) r = std::max(r, P.R); return r; }
  return 1.0e300;
}

// ------------------------------------------------------------------------------------------------------------------
// Two-pass far assembly (PEC/PEC blocks).  Measured (tools/far_dbg_probe.sh, profiles/r02_far_scatter_experiments.txt): whatever
// the scatter mechanism — 18 scalar RED per pair, or the staged windows above — 1.2-1.6 G FP64 atomics into a matrix that does
// not fit the L2 keep the far kernel at 14-17 ms; with the atomics replaced by plain stores the same kernel takes 10 ms.  So:
//   pass 1  far_pass1_kernel: every panel pair pa<pb (pa, pb any for a non-symmetric block) writes its nine contributions
//           (zeros when the pair is not far) with plain 16-byte stores to the panel-pair array Q[pb][j][pa][i] — no atomics, no
//           fences, nothing the cubature loop can queue behind;
//   pass 2  far_gather_kernel: one thread per matrix entry (e_a, e_b) adds the <= 4 panel-pair values of its two x two panels
//           (coalesced: consecutive edges <-> consecutive panels), adds them to what the near/singular kernels have stamped, and
//           writes the entry AND its mirror image (32 x 32 tiles through shared memory) — it replaces the symmetrize pass.
// HBM traffic of the detour: 9 x 16 B per panel pair written and read once (13.4 GB at N = 20 280: ~4 ms at 6.5 TB/s, the write
// half hidden under the FP64-bound pass 1).  Q is sized NP^2/2 x 144 B for a symmetric block; when it does not fit the budget
// (SCUFF_CUDA_FAR_QMAX_GB, default 40) the staged-atomic kernel is used instead.
// ------------------------------------------------------------------------------------------------------------------
// Layout of the panel-pair array.  The values of a-panel pa for b-panel pb and local b-index j sit at the SLOTS of pa's edges,
//   element (pb, j, slot, c) = row(pb) + (j * EX + slot) * nc + c ,   slot = 2 * edge + (0: pa is the edge's P panel, 1: its M panel),
// so that the gather thread of entry (e_a, e_b) finds both panels of e_a in ONE contiguous 2 * nc * 16 B piece and a warp
// (consecutive e_a) reads contiguous memory: every sector is fetched once (with a panel-indexed layout the three values of a panel
// were split between the warps of its M-role and P-role edges and every sector crossed HBM -> L2 more than twice).  The stores of
// pass 1 stay dense: the 32 consecutive panels of a warp cover ~50 consecutive edges in both roles.
// A symmetric block stores pa < pb only: row pb extends over EX = ext[pb / 128] slots (all edges of the panels below the end of
// pb's 128-tile), rows are packed with the prefix sums base[].
struct FarQ { cd *Q; const long *base; const int *ext; int symmetric; int extFull; int nc; };   // nc complex values per combination: 1 (PEC/PEC: g0), 2 (g0, g1)
__device__ __forceinline__ long farq_row(const FarQ &L, int pb, int &EX)
{ if (L.symmetric) { const int t = pb >> 7; EX = L.ext[t]; return L.nc * (L.base[t] + (long)(pb - (t << 7)) * 3 * EX); }
  EX = L.extFull; return (long)L.nc * pb * 3 * L.extFull;
}
struct FarArgsQ { FarArgs A; FarQ L; };
#ifndef FAR_PSI_MINB
#define FAR_PSI_MINB 2          /* resident CTAs per SM of the Phi + Psi first pass */
#endif

template <int KMODE, int SMALL>
__global__ void __launch_bounds__(FAR_THREADS, 3)
far_pass1_kernel(const __grid_constant__ FarArgsQ X)
{ const FarArgs &A = X.A;
  extern __shared__ __align__(128) double sm[];
  __shared__ __align__(8) uint64_t bar;
  typedef typename KScalar<KMODE>::type S;
  const int TB = A.TA;
  const int b0 = blockIdx.y * TB;
  const int aBase = blockIdx.x * FAR_THREADS;
  const int nB = min(TB, A.NPb - b0);
  if (A.symmetric && b0 + nB - 1 < aBase) return;
  const int tid = threadIdx.x;
  double *sXa = sm + TB * 36 + tid;                    // element k at sXa[k * FAR_THREADS]
  double *sSt = sm + TB * 36 + 18 * FAR_THREADS + tid;
  int *sEi = (int *)(sm + TB * 36 + 34 * FAR_THREADS) + tid;
  unsigned int *sCnt = (unsigned int *)sEi + 3 * FAR_THREADS;
  sCnt[0] = 0; sCnt[FAR_THREADS] = 0;

  if (tid == 0)
   { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
     asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (tid == 0) tma_load_1d(sm, A.pb + b0, (uint32_t)(nB * sizeof(PanelRec)), &bar);

  const int pa = aBase + tid;
  const bool live = pa < A.NPa && !(A.symmetric && pa >= b0 + nB - 1);
  const PanelRec *Pa = A.pa + (pa < A.NPa ? pa : 0);
#pragma unroll
  for (int i = 0; i < 18; i++) sXa[i * FAR_THREADS] = Pa->x[i];
  int na = 0, slot[3];
#pragma unroll
  for (int m = 0; m < 3; m++)
   { const double v0 = Pa->V[m];
     sSt[(3 + m) * FAR_THREADS] = Pa->V[3 + m] - v0;                 // A
     sSt[(6 + m) * FAR_THREADS] = Pa->V[6 + m] - v0;                 // B
     sSt[(9 + m) * FAR_THREADS] = Pa->sl[m];
     sSt[(12 + m) * FAR_THREADS] = Pa->C[m];
     slot[m] = (Pa->ei[m] >= 0) ? 2 * Pa->ei[m] + (Pa->sl[m] < 0.0 ? 1 : 0) : -1;
     na += (Pa->ei[m] >= 0); }
  sSt[15 * FAR_THREADS] = Pa->R;
  mbar_wait(&bar, 0);
  if (!live) return;

  for (int ib = 0; ib < nB; ib++)
   { const int pb = b0 + ib;
     if (A.symmetric && pb <= pa) continue;            // the gather reads pairs pa < pb only
     const double *Pb = sm + ib * 36;
     const int *eb3 = (const int *)(Pb + 34);
     int nEnt = na * ((eb3[0] >= 0) + (eb3[1] >= 0) + (eb3[2] >= 0));
     if (A.S.colHi != 0x7fffffff)
      { // column-block assembly: a pair none of whose (<= 9) entries lies in the owned columns is neither computed nor stored
        // (the gather of the owned columns never reads it); exact per entry
        nEnt = 0;
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
         for (int j = 0; j < 3; j++)
          { const int e = slot[i] >> 1, f = eb3[j];
            const int c0 = A.S.colOff + (A.symmetric ? max(e, f) : f);
            nEnt += (slot[i] >= 0 && f >= 0 && c0 >= A.S.colLo && c0 < A.S.colHi); }
        if (nEnt == 0) continue; }
     double D[3];
     bool far;
     { const double Ca[3] = { sSt[12 * FAR_THREADS], sSt[13 * FAR_THREADS], sSt[14 * FAR_THREADS] };
       far = pair_is_far(Ca, sSt[15 * FAR_THREADS], Pb, Pb[3], A.useDisp != 0, A.L, D); }
     cd val[9];
#pragma unroll
     for (int n = 0; n < 9; n++) val[n] = mk(0.0, 0.0);
     if (far)
      { sCnt[0] += 1;
        sCnt[FAR_THREADS] += (unsigned int)nEnt;
        double r0 = 0.0;
        if (SMALL) { const double d2 = fma(D[2], D[2], fma(D[1], D[1], D[0] * D[0])); r0 = d2 * rsqrt_fast<true>(d2); }
        S mPhi[3][3], mPsi[3][3];
        far_moments<KMODE, false, SMALL, S>(sXa, Pb, D, A.K, r0, mPhi, mPsi);
        cd c1 = A.S.c1;
        if (SMALL) c1 = cscale(c1, far_pair_factor<KMODE, S>(A.K, r0));   // the pair factor exp(ik r0) rides on the prefactor
        double sla[3], Aa[3], Ba[3], Ab[3], Bb[3];
#pragma unroll
        for (int m = 0; m < 3; m++)
         { sla[m] = sSt[(9 + m) * FAR_THREADS]; Aa[m] = sSt[(3 + m) * FAR_THREADS]; Ba[m] = sSt[(6 + m) * FAR_THREADS];
           Ab[m] = Pb[7 + m] - Pb[4 + m]; Bb[m] = Pb[10 + m] - Pb[4 + m]; }
        contract_pec<S>(mPhi, Aa, Ba, Ab, Bb, sla, Pb + 31, c1, A.K.c4, val); }
     int EX; const long row = farq_row(X.L, pb, EX);
     double2 *q = reinterpret_cast<double2 *>(X.L.Q + row);
#pragma unroll
     for (int i = 0; i < 3; i++)
      if (slot[i] >= 0)
       {
#pragma unroll
         for (int j = 0; j < 3; j++) q[(long)j * EX + slot[i]] = make_double2(val[3 * i + j].x, val[3 * i + j].y); }
   }
  // algorithm counters (reference log line SurfaceSurfaceInteractions.cc:550-558)
  if (sCnt[FAR_THREADS]) atomicAdd(A.counters, (unsigned long long)sCnt[FAR_THREADS]);
  if (sCnt[0]) atomicAdd(A.counters + 5, (unsigned long long)sCnt[0]);
}

// Pass 1 for blocks with a dielectric side: per (Q_a,Q_b) combination the two values the stamps need,
//   g0 = l_a l_b H0 ,  g1 = l_a l_b H1 / (ik)      (EdgeEdgeInteractions.cc:183-188, SurfaceSurfaceInteractions.cc:219-327)
// (the region prefactors c1, c2, c3 are applied by the gather).
template <int KMODE, int SMALL>
__global__ void __launch_bounds__(FAR_THREADS, FAR_PSI_MINB)
far_pass1_psi_kernel(const __grid_constant__ FarArgsQ X)
{ const FarArgs &A = X.A;
  extern __shared__ __align__(128) double sm[];
  __shared__ __align__(8) uint64_t bar;
  typedef typename KScalar<KMODE>::type S;
  const int TB = A.TA;
  const int b0 = blockIdx.y * TB;
  const int aBase = blockIdx.x * FAR_THREADS;
  const int nB = min(TB, A.NPb - b0);
  if (A.symmetric && b0 + nB - 1 < aBase) return;
  const int tid = threadIdx.x;
  double *sXa = sm + TB * 36 + tid;                    // element k at sXa[k * FAR_THREADS]
  double *sSt = sm + TB * 36 + 18 * FAR_THREADS + tid;
  int *sEi = (int *)(sm + TB * 36 + 34 * FAR_THREADS) + tid;
  unsigned int *sCnt = (unsigned int *)sEi + 3 * FAR_THREADS;
  sCnt[0] = 0; sCnt[FAR_THREADS] = 0;

  if (tid == 0)
   { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
     asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (tid == 0) tma_load_1d(sm, A.pb + b0, (uint32_t)(nB * sizeof(PanelRec)), &bar);

  const int pa = aBase + tid;
  const bool live = pa < A.NPa && !(A.symmetric && pa >= b0 + nB - 1);
  const PanelRec *Pa = A.pa + (pa < A.NPa ? pa : 0);
#pragma unroll
  for (int i = 0; i < 18; i++) sXa[i * FAR_THREADS] = Pa->x[i];
  int na = 0;
#pragma unroll
  for (int m = 0; m < 3; m++)
   { const double v0 = Pa->V[m];
     sSt[m * FAR_THREADS] = v0;                                      // V_0
     sSt[(3 + m) * FAR_THREADS] = Pa->V[3 + m] - v0;                 // A
     sSt[(6 + m) * FAR_THREADS] = Pa->V[6 + m] - v0;                 // B
     sSt[(9 + m) * FAR_THREADS] = Pa->sl[m];
     sSt[(12 + m) * FAR_THREADS] = Pa->C[m];
     sEi[m * FAR_THREADS] = (Pa->ei[m] >= 0) ? 2 * Pa->ei[m] + (Pa->sl[m] < 0.0 ? 1 : 0) : -1;     // slot of this panel in the edge's piece
     na += (Pa->ei[m] >= 0); }
  sSt[15 * FAR_THREADS] = Pa->R;
  mbar_wait(&bar, 0);
  if (!live) return;

  for (int ib = 0; ib < nB; ib++)
   { const int pb = b0 + ib;
     if (A.symmetric && pb <= pa) continue;            // the gather reads pairs pa < pb only
     const double *Pb = sm + ib * 36;
     const int *eb3 = (const int *)(Pb + 34);
     int nEnt = na * ((eb3[0] >= 0) + (eb3[1] >= 0) + (eb3[2] >= 0));
     if (A.S.colHi != 0x7fffffff)
      { // column-block assembly (see far_pass1_kernel); a dielectric b side owns two columns per edge
        const int wb = A.S.bPEC ? 1 : 2;
        nEnt = 0;
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
         for (int j = 0; j < 3; j++)
          { const int sl = sEi[i * FAR_THREADS], e = sl >> 1, f = eb3[j];
            const int c0 = A.S.colOff + wb * (A.symmetric ? max(e, f) : f);
            nEnt += (sl >= 0 && f >= 0 && c0 + wb - 1 >= A.S.colLo && c0 < A.S.colHi); }
        if (nEnt == 0) continue; }
     double D[3];
     bool far;
     { const double Ca[3] = { sSt[12 * FAR_THREADS], sSt[13 * FAR_THREADS], sSt[14 * FAR_THREADS] };
       far = pair_is_far(Ca, sSt[15 * FAR_THREADS], Pb, Pb[3], A.useDisp != 0, A.L, D); }
     int EX; const long row = farq_row(X.L, pb, EX);
     double2 *q = reinterpret_cast<double2 *>(X.L.Q + row);
     if (!far)
      {
#pragma unroll
        for (int i = 0; i < 3; i++)
         { const int sl = sEi[i * FAR_THREADS];
           if (sl < 0) continue;
#pragma unroll
           for (int j = 0; j < 3; j++) { q[((long)j * EX + sl) * 2] = make_double2(0.0, 0.0); q[((long)j * EX + sl) * 2 + 1] = make_double2(0.0, 0.0); } }
        continue; }
     sCnt[0] += 1;
     sCnt[FAR_THREADS] += (unsigned int)nEnt;
     double r0 = 0.0;
     if (SMALL) { const double d2 = fma(D[2], D[2], fma(D[1], D[1], D[0] * D[0])); r0 = d2 * rsqrt_fast<true>(d2); }
     S mPhi[3][3], mPsi[3][3];
     far_moments<KMODE, true, SMALL, S>(sXa, Pb, D, A.K, r0, mPhi, mPsi);
     cd e0 = mk(1.0, 0.0), ooik = A.K.ooik;
     if (SMALL) { e0 = tocd(far_pair_factor<KMODE, S>(A.K, r0)); ooik = ooik * e0; }      // the pair factor rides on the two prefactors
     double Aa[3], Ba[3], Ab[3], Bb[3], E00[3];
#pragma unroll
     for (int m = 0; m < 3; m++)
      { Aa[m] = sSt[(3 + m) * FAR_THREADS]; Ba[m] = sSt[(6 + m) * FAR_THREADS];
        Ab[m] = Pb[7 + m] - Pb[4 + m]; Bb[m] = Pb[10 + m] - Pb[4 + m];
        E00[m] = sSt[m * FAR_THREADS] - (Pb[4 + m] + (A.useDisp ? A.L[m] : 0.0)); }
     const double dAA = dot3(Aa, Ab), dAB = dot3(Aa, Bb), dBA = dot3(Ba, Ab), dBB = dot3(Ba, Bb);
     // H0_ij = T - [j>0] P_j - [i>0] Q_i + [i>0][j>0] d_ij m00 + c4 m00  (see contract_pec)
     S T0, P1, P2, Q1, Q2;
     szero(T0); szero(P1); szero(P2); szero(Q1); szero(Q2);
     fma_acc(T0, dAA, mPhi[1][1]); fma_acc(T0, dAB, mPhi[1][2]); fma_acc(T0, dBA, mPhi[2][1]); fma_acc(T0, dBB, mPhi[2][2]);
     fma_acc(P1, dAA, mPhi[1][0]); fma_acc(P1, dBA, mPhi[2][0]);
     fma_acc(P2, dAB, mPhi[1][0]); fma_acc(P2, dBB, mPhi[2][0]);
     fma_acc(Q1, dAA, mPhi[0][1]); fma_acc(Q1, dAB, mPhi[0][2]);
     fma_acc(Q2, dBA, mPhi[0][1]); fma_acc(Q2, dBB, mPhi[0][2]);
     const cd m00 = cscale(e0, mPhi[0][0]);
     const cd cT = cscale(e0, T0) + A.K.c4 * m00;
     const cd cP1 = cscale(e0, P1), cP2 = cscale(e0, P2), cQ1 = cscale(e0, Q1), cQ2 = cscale(e0, Q2);
     // H1_ij = sum_{X in {A,B}, Y in {A',B'}} ((X x Y).E_ij) <(x - d_i)(y - d_j) Psi> ,  E_ij = E00 + {0,A,B}_i - {0,A',B'}_j
     double tAA, tAB, tBA, tBB, AbN, BbN, AaNp, BaNp;
     { double Na[3], Nb[3], eA[3], eB[3];
       cross3(Aa, Ba, Na); cross3(Ab, Bb, Nb);
       cross3(E00, Aa, eA); cross3(E00, Ba, eB);         // (X x Y).E = Y.(E x X)
       tAA = dot3(Ab, eA); tAB = dot3(Bb, eA); tBA = dot3(Ab, eB); tBB = dot3(Bb, eB);
       AbN = dot3(Ab, Na); BbN = dot3(Bb, Na); AaNp = dot3(Aa, Nb); BaNp = dot3(Ba, Nb); }
     S q1s[3], q2s[3];
#pragma unroll
     for (int n = 0; n < 3; n++) { q1s[n] = ssub(mPsi[1][n], mPsi[0][n]); q2s[n] = ssub(mPsi[2][n], mPsi[0][n]); }
#pragma unroll
     for (int i = 0; i < 3; i++)
      { const double sla = sSt[(9 + i) * FAR_THREADS];
        const S *UQ = (i == 1) ? q1s : mPsi[1], *VQ = (i == 2) ? q2s : mPsi[2];
#pragma unroll
        for (int j = 0; j < 3; j++)
         { cd h0 = (j == 0) ? cT : (j == 1 ? cT - cP1 : cT - cP2);
           if (i > 0)
            { h0 = h0 - (i == 1 ? cQ1 : cQ2);
              if (j > 0) { const double d = (i == 1) ? (j == 1 ? dAA : dAB) : (j == 1 ? dBA : dBB); h0.x = fma(d, m00.x, h0.x); h0.y = fma(d, m00.y, h0.y); } }
           const S Pu = (j == 1) ? ssub(UQ[1], UQ[0]) : UQ[1], Pv = (j == 2) ? ssub(UQ[2], UQ[0]) : UQ[2];
           const S Qu = (j == 1) ? ssub(VQ[1], VQ[0]) : VQ[1], Qv = (j == 2) ? ssub(VQ[2], VQ[0]) : VQ[2];
           const double cAA = tAA - (i == 2 ? AbN : 0.0) - (j == 2 ? AaNp : 0.0);
           const double cAB = tAB - (i == 2 ? BbN : 0.0) + (j == 1 ? AaNp : 0.0);
           const double cBA = tBA + (i == 1 ? AbN : 0.0) - (j == 2 ? BaNp : 0.0);
           const double cBB = tBB + (i == 1 ? BbN : 0.0) + (j == 1 ? BaNp : 0.0);
           S h1; szero(h1);
           fma_acc(h1, cAA, Pu); fma_acc(h1, cAB, Pv); fma_acc(h1, cBA, Qu); fma_acc(h1, cBB, Qv);
           const double sc = sla * Pb[31 + j];
           const cd g0 = sc * h0, g1 = sc * cscale(ooik, h1);
           const int sl = sEi[i * FAR_THREADS];
           if (sl >= 0)
            { q[((long)j * EX + sl) * 2] = make_double2(g0.x, g0.y);
              q[((long)j * EX + sl) * 2 + 1] = make_double2(g1.x, g1.y); } } }
   }
  if (sCnt[FAR_THREADS]) atomicAdd(A.counters, (unsigned long long)sCnt[FAR_THREADS]);
  if (sCnt[0]) atomicAdd(A.counters + 5, (unsigned long long)sCnt[0]);
}

// 32-byte streaming load (sm_100 LDG.E.EF.ENL2.256): both role slots of an edge in one request
__device__ __forceinline__ void ld256_cs(const void *p, double &a, double &b, double &c, double &d)
{ asm volatile("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p)); }

// edge -> its two panels and the local index of the edge in each: {pP, iP, pM, iM}, -1 where there is none
struct GatherArgs
{ FarQ L; const int4 *edgeA, *edgeB; int NEa, NEb;
  double *M; long ld; int rowOff, colOff;
  int aPEC, bPEC; cd c1, c2, c3;
  int colLo, colHi;          // owned matrix columns (column-block assembly); 0, INT_MAX = all
  int bxOff; };              // first column tile of the launch

// Tile = 32 columns (b-edges) x GR*32 rows (a-edges); blockIdx.x walks DOWN the rows of one column tile so that CTAs in flight
// at the same time read neighbouring pieces of the same Q rows (DRAM pages, shared sectors) and write neighbouring
// segments of the same 32 matrix columns.  Q is read once (streaming loads), the matrix written once per triangle.
#define FAR_GATHER_GR 2
// RED = true: the sums are ADDED with atomics instead of read-modify-written, so the gather may run while the near/singular
// kernels are still stamping (it then overlaps the latency-bound Taylor-Duffy work with its HBM-bound streaming); the mirror
// image is left to the symmetrize pass.
template <int W, bool SYM, bool RED>
__global__ void __launch_bounds__(256)
far_gather_kernel(const GatherArgs G)
{ __shared__ double tile[RED ? 1 : 32][RED ? 1 : 33 * W];
  const int bx = blockIdx.y + G.bxOff;             // column tile
  const int tx = threadIdx.x, ty = threadIdx.y;    // 32 x 8
  for (int g = 0; g < FAR_GATHER_GR; g++)
   { const int by = blockIdx.x * FAR_GATHER_GR + g;     // row tile
     if (by * 32 >= G.NEa || (SYM && by > bx)) break;
     const int ea = by * 32 + tx;
     int4 Ea = make_int4(-1, 0, -1, 0);
     if (ea < G.NEa) Ea = G.edgeA[ea];
     for (int j = ty; j < 32; j += 8)
      { const int eb = bx * 32 + j;
        if (ea >= G.NEa || eb >= G.NEb || (SYM && ea >= eb)) continue;
        if (G.colOff + eb < G.colLo || G.colOff + eb >= G.colHi) continue;
        const int4 Eb = G.edgeB[eb];
        double sx = 0.0, sy = 0.0;
#pragma unroll
        for (int v = 0; v < 2; v++)
         { const int pb = v ? Eb.z : Eb.x, jb = v ? Eb.w : Eb.y;
           if (pb < 0) continue;
           // the two panels of e_a that lie below pb (any, in a non-symmetric block): their slots 2 e_a, 2 e_a + 1 of row pb are
           // one aligned 32-byte piece; a slot whose panel does not qualify is loaded but not used (it may never have been written)
           const bool okP = Ea.x >= 0 && (!SYM || Ea.x < pb), okM = Ea.z >= 0 && (!SYM || Ea.z < pb);
           if (okP || okM)
            { int EX; const long row = farq_row(G.L, pb, EX);
              double a, b, c, d; ld256_cs(G.L.Q + row + (long)jb * EX + 2 * ea, a, b, c, d);
              if (okP) { sx += a; sy += b; }
              if (okM) { sx += c; sy += d; } }
           if (SYM)
            { // a panel of e_a ABOVE pb: the pair is stored in that panel's row, at the slot of (e_b, role v)
#pragma unroll
              for (int u = 0; u < 2; u++)
               { const int pa = u ? Ea.z : Ea.x, ia = u ? Ea.w : Ea.y;
                 if (pa <= pb) continue;                  // (pa == pb: same panel, never a far pair)
                 int EX; const long row = farq_row(G.L, pa, EX);
                 const double2 q = __ldcs(reinterpret_cast<const double2 *>(G.L.Q + row + (long)ia * EX + 2 * eb + v));
                 sx += q.x; sy += q.y; } } }
        const size_t idx = ((size_t)(G.rowOff + ea) + (size_t)(G.colOff + eb) * (size_t)G.ld) * W;
        if (RED)
         { if (sx != 0.0) atomicAdd(G.M + idx, sx);
           if (W == 2 && sy != 0.0) atomicAdd(G.M + idx + 1, sy); }
        else if (W == 2)
         { double2 m = *reinterpret_cast<double2 *>(G.M + idx); m.x += sx; m.y += sy;
           __stcs(reinterpret_cast<double2 *>(G.M + idx), m); tile[j][2 * tx] = m.x; tile[j][2 * tx + 1] = m.y; }
        else { const double m = G.M[idx] + sx; __stcs(G.M + idx, m); tile[j][tx] = m; }
      }
     if (!SYM || RED) continue;
     __syncthreads();
     for (int j = ty; j < 32; j += 8)
      { const int r = bx * 32 + tx, c = by * 32 + j;   // mirror image: lower(r, c) = upper(c, r)
        if (r < G.NEa && c < G.NEa && r > c)
         { const size_t idx = ((size_t)(G.rowOff + r) + (size_t)(G.colOff + c) * (size_t)G.ld) * W;
           if (W == 2) __stcs(reinterpret_cast<double2 *>(G.M + idx), make_double2(tile[tx][2 * j], tile[tx][2 * j + 1]));
           else __stcs(G.M + idx, tile[tx][j]); } }
     __syncthreads();
   }
}

// Gather for blocks with a dielectric side: G0 = sum g0, G1 = sum g1 over the <= 4 panel pairs of the edge pair, stamped as
//   dielectric/dielectric: [ c1 G0  c2 G1 ; c2 G1  c3 G0 ] at (2 e_a, 2 e_b);  PEC/dielectric: [ c1 G0  c2 G1 ] at (e_a, 2 e_b);
//   dielectric/PEC: [ c1 G0 ; c2 G1 ] at (2 e_a, e_b)     (SurfaceSurfaceInteractions.cc:219-327).  The matrix is complex or real
// (W doubles per entry; a real matrix keeps the real parts, HMatrix.cc:419-447).  SYM (same dielectric surface): entries e_a < e_b,
// and the mirror image of each of the four component planes through shared memory.
template <int W, bool SYM, bool RED>
__global__ void __launch_bounds__(256)
far_gather_psi_kernel(const GatherArgs G)
{ __shared__ double tile[RED ? 1 : 32][RED ? 1 : 33 * W];
  const int bx = blockIdx.y + G.bxOff;
  const int tx = threadIdx.x, ty = threadIdx.y;    // 32 x 8
  const int wa = G.aPEC ? 1 : 2, wb = G.bPEC ? 1 : 2;
  for (int g = 0; g < FAR_GATHER_GR; g++)
   { const int by = blockIdx.x * FAR_GATHER_GR + g;
     if (by * 32 >= G.NEa || (SYM && by > bx)) break;
     const int ea = by * 32 + tx;
     int4 Ea = make_int4(-1, 0, -1, 0);
     if (ea < G.NEa) Ea = G.edgeA[ea];
     cd v[4][4];                                   // [column j of this thread: ty, ty+8, ..][component plane]
#pragma unroll
     for (int jj = 0; jj < 4; jj++)
      { const int j = ty + 8 * jj, eb = bx * 32 + j;
#pragma unroll
        for (int c = 0; c < 4; c++) v[jj][c] = mk(0.0, 0.0);
        if (ea >= G.NEa || eb >= G.NEb || (SYM && ea >= eb)) continue;
        const int4 Eb = G.edgeB[eb];
        double g0x = 0, g0y = 0, g1x = 0, g1y = 0;
#pragma unroll
        for (int w = 0; w < 2; w++)
         { const int pb = w ? Eb.z : Eb.x, jb = w ? Eb.w : Eb.y;
           if (pb < 0) continue;
           const bool okP = Ea.x >= 0 && (!SYM || Ea.x < pb), okM = Ea.z >= 0 && (!SYM || Ea.z < pb);
           if (okP || okM)
            { int EX; const long row = farq_row(G.L, pb, EX);
              const cd *q = G.L.Q + row + ((long)jb * EX + 2 * ea) * 2;          // {g0, g1} of role P, {g0, g1} of role M: 64 bytes
              double a, b, c, d;
              if (okP) { ld256_cs(q, a, b, c, d); g0x += a; g0y += b; g1x += c; g1y += d; }
              if (okM) { ld256_cs(q + 2, a, b, c, d); g0x += a; g0y += b; g1x += c; g1y += d; } }
           if (SYM)
            {
#pragma unroll
              for (int u = 0; u < 2; u++)
               { const int pa = u ? Ea.z : Ea.x, ia = u ? Ea.w : Ea.y;
                 if (pa <= pb) continue;
                 int EX; const long row = farq_row(G.L, pa, EX);
                 double a, b, c, d; ld256_cs(G.L.Q + row + ((long)ia * EX + 2 * eb + w) * 2, a, b, c, d);
                 g0x += a; g0y += b; g1x += c; g1y += d; } } }
        const cd G0 = mk(g0x, g0y), G1 = mk(g1x, g1y);
        const cd v11 = G.c1 * G0, v12 = G.c2 * G1, v22 = G.c3 * G0;
        // component planes: 0 (X,Y), 1 (X,Y+1), 2 (X+1,Y), 3 (X+1,Y+1)
#pragma unroll
        for (int c = 0; c < 4; c++)
         { const int dr = c >> 1, dc = c & 1;
           if (dr >= wa || dc >= wb) continue;
           { const int col = G.colOff + wb * eb + dc; if (col < G.colLo || col >= G.colHi) continue; }
           const cd val = (c == 0) ? v11 : (c == 3 ? v22 : v12);
           const size_t idx = ((size_t)(G.rowOff + wa * ea + dr) + (size_t)(G.colOff + wb * eb + dc) * (size_t)G.ld) * W;
           if (RED)
            { if (val.x != 0.0) atomicAdd(G.M + idx, val.x);
              if (W == 2 && val.y != 0.0) atomicAdd(G.M + idx + 1, val.y); }
           else if (W == 2)
            { double2 m = *reinterpret_cast<double2 *>(G.M + idx); m.x += val.x; m.y += val.y;
              __stcs(reinterpret_cast<double2 *>(G.M + idx), m); v[jj][c] = mk(m.x, m.y); }
           else { const double m = G.M[idx] + val.x; __stcs(G.M + idx, m); v[jj][c] = mk(m, 0.0); } } }
     if (!SYM || RED) continue;
     // inside the 2x2 block of an edge with itself (near contributions only): (2e+1, 2e) = (2e, 2e+1)
     if (bx == by && ty == 0 && ea < G.NEa)
      { const size_t up = ((size_t)(G.rowOff + 2 * ea) + (size_t)(G.colOff + 2 * ea + 1) * (size_t)G.ld) * W;
        const size_t lo = ((size_t)(G.rowOff + 2 * ea + 1) + (size_t)(G.colOff + 2 * ea) * (size_t)G.ld) * W;
        for (int q = 0; q < W; q++) G.M[lo + q] = G.M[up + q]; }
     // mirror image, one component plane at a time: lower(2 e_b + q, 2 e_a + p) = upper(2 e_a + p, 2 e_b + q)
#pragma unroll
     for (int c = 0; c < 4; c++)
      { __syncthreads();
#pragma unroll
        for (int jj = 0; jj < 4; jj++)
         { const int j = ty + 8 * jj;
           if (W == 2) { tile[j][2 * tx] = v[jj][c].x; tile[j][2 * tx + 1] = v[jj][c].y; } else tile[j][tx] = v[jj][c].x; }
        __syncthreads();
        const int dr = c >> 1, dc = c & 1;
#pragma unroll
        for (int jj = 0; jj < 4; jj++)
         { const int j = ty + 8 * jj;
           const int eb = bx * 32 + tx, eA = by * 32 + j;      // this thread now writes the image of (e_a = eA, e_b = eb)
           if (eb < G.NEa && eA < G.NEa && eb > eA)
            { const size_t idx = ((size_t)(G.rowOff + 2 * eb + dc) + (size_t)(G.colOff + 2 * eA + dr) * (size_t)G.ld) * W;
              if (W == 2) __stcs(reinterpret_cast<double2 *>(G.M + idx), make_double2(tile[tx][2 * j], tile[tx][2 * j + 1]));
              else __stcs(G.M + idx, tile[tx][j]); } } }
     __syncthreads();
   }
}

// ---- host side of the staged scatter: window and slots of every 32-panel group of a surface ----
struct FarSlotTable { uint64_t generation = 0; int NP = 0; FarSlot *d = nullptr; };
static std::map<std::pair<scuffcuda_context *, const PanelRec *>, FarSlotTable> g_farSlots;
static std::mutex g_farSlotsLock;

void build_far_slots(const std::vector<PanelRec> &P, std::vector<FarSlot> &T)
{ const int NP = (int)P.size(), NPpad = (NP + FAR_THREADS - 1) / FAR_THREADS * FAR_THREADS;
  T.assign(NPpad, FarSlot{ { FAR_SLOT_NONE, FAR_SLOT_NONE, FAR_SLOT_NONE }, 0, 0, 0 });   // s[3], span, rowLo, nOut
  for (int g0 = 0; g0 < NP; g0 += 32)
   { const int g1 = std::min(NP, g0 + 32);
     // the window [lo, lo + FAR_SPAN) that holds most of the group's (panel, edge) slots
     std::vector<int> e;
     for (int p = g0; p < g1; p++) for (int i = 0; i < 3; i++) if (P[p].ei[i] >= 0) e.push_back(P[p].ei[i]);
     std::sort(e.begin(), e.end());
     int bestLo = 0, bestCnt = 0, bestHi = 0;
     for (size_t a = 0, b = 0; a < e.size(); a++)
      { if (a > 0 && e[a] == e[a - 1]) continue;
        while (b < e.size() && e[b] < e[a] + FAR_SPAN) b++;
        if ((int)(b - a) > bestCnt) { bestCnt = (int)(b - a); bestLo = e[a]; bestHi = e[b - 1]; } }
     const int span = bestCnt ? bestHi - bestLo + 1 : 0;
     for (int p = g0; p < std::min(NPpad, g0 + 32); p++) { T[p].rowLo = bestLo; T[p].span = (short)span; }
     int nOut = 0;
     for (int p = g0; p < g1; p++)
      for (int i = 0; i < 3; i++)
       { const int ed = P[p].ei[i];
         if (ed < 0) continue;
         if (ed < bestLo || ed >= bestLo + span) { T[p].s[i] = (short)(FAR_SLOT_OUTLIER - nOut++); continue; }
         bool second = false;
         for (int q = g0; q < p && !second; q++) for (int k = 0; k < 3; k++) if (P[q].ei[k] == ed) second = true;
         T[p].s[i] = (short)((ed - bestLo) | (second ? FAR_SLOT_SECOND : 0)); }
     for (int p = g0; p < std::min(NPpad, g0 + 32); p++) T[p].nOut = nOut; }
}

static const FarSlot *far_slots(scuffcuda_context *ctx, const PanelRec *d)
{ const Surface *S = nullptr;
  for (const Surface &Q : ctx->surfaces) if (Q.dpanels == d) S = &Q;
  if (!S) return nullptr;
  std::lock_guard<std::mutex> lock(g_farSlotsLock);
  FarSlotTable &T = g_farSlots[std::make_pair(ctx, d)];
  if (T.d && T.generation == S->generation && T.NP == S->NP) return T.d;
  std::vector<FarSlot> h;
  build_far_slots(S->hpanels, h);
  if (T.d) { cudaStreamSynchronize(ctx->stream); cudaFree(T.d); T.d = nullptr; }
  if (cudaMalloc(&T.d, h.size() * sizeof(FarSlot)) != cudaSuccess) { T.d = nullptr; return nullptr; }
  cudaMemcpyAsync(T.d, h.data(), h.size() * sizeof(FarSlot), cudaMemcpyHostToDevice, ctx->stream);
  cudaStreamSynchronize(ctx->stream);                   // h is a local
  T.generation = S->generation; T.NP = S->NP;
  return T.d;
}
void far_slots_release(scuffcuda_context *ctx)
{ std::lock_guard<std::mutex> lock(g_farSlotsLock);
  for (auto it = g_farSlots.begin(); it != g_farSlots.end();)
   if (it->first.first == ctx) { if (it->second.d) cudaFree(it->second.d); it = g_farSlots.erase(it); } else ++it;
}

// ---- host side of the two-pass path ----
struct FarTwoPassState
{ cd *Q = nullptr; size_t qBytes = 0;                       // panel-pair array, grown on demand, kept for the life of the context
  struct SurfTab { uint64_t generation = 0; int4 *edges = nullptr; long *base = nullptr; int *ext = nullptr; long symElems = 0; int extFull = 0; int NE = 0; };
  std::map<const PanelRec *, SurfTab> edges;                // surface -> edge table and row layout of the panel-pair array
  bool pending = false; GatherArgs G; bool sym = false; int isReal = 0; int NEa = 0, NEb = 0;
  cudaEvent_t zeroEvent = nullptr;
  cudaEvent_t g0 = nullptr, g1 = nullptr; bool timed = false; };   // gather start / end
static std::map<scuffcuda_context *, FarTwoPassState> g_twoPass;

static const FarTwoPassState::SurfTab *far_surface_tables(scuffcuda_context *ctx, FarTwoPassState &T, const PanelRec *d)
{ const Surface *S = nullptr;
  for (const Surface &Q : ctx->surfaces) if (Q.dpanels == d) S = &Q;
  if (!S) return nullptr;
  FarTwoPassState::SurfTab &slot = T.edges[d];
  if (slot.edges && slot.generation == S->generation) return &slot;
  std::vector<int4> h(std::max(S->NE, 1), make_int4(-1, 0, -1, 0));
  const int tiles = (S->NP + 127) / 128;
  std::vector<int> ext(tiles, 0); std::vector<long> base(tiles + 1, 0);
  int maxEdge = -1;
  for (int p = 0; p < S->NP; p++)
   { for (int i = 0; i < 3; i++)
      { const int e = S->hpanels[p].ei[i];
        if (e < 0 || e >= S->NE) continue;
        maxEdge = std::max(maxEdge, e);
        if (S->hpanels[p].sl[i] >= 0.0) { h[e].x = p; h[e].y = i; } else { h[e].z = p; h[e].w = i; } }
     // row extent of the 128-tile this panel ends: all slots of the edges met so far, rounded to whole 128-byte lines
     if ((p & 127) == 127 || p == S->NP - 1) ext[p >> 7] = ((2 * (maxEdge + 1) + 7) / 8) * 8; }
  for (int t = 0; t < tiles; t++) base[t + 1] = base[t] + 128L * 3 * ext[t];
  if (slot.edges) { cudaStreamSynchronize(ctx->stream); cudaFree(slot.edges); cudaFree(slot.base); cudaFree(slot.ext); slot.edges = nullptr; }
  if (cudaMalloc(&slot.edges, h.size() * sizeof(int4)) != cudaSuccess || cudaMalloc(&slot.base, base.size() * sizeof(long)) != cudaSuccess
      || cudaMalloc(&slot.ext, ext.size() * sizeof(int)) != cudaSuccess) { cudaGetLastError(); slot.edges = nullptr; return nullptr; }
  cudaMemcpyAsync(slot.edges, h.data(), h.size() * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(slot.base, base.data(), base.size() * sizeof(long), cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(slot.ext, ext.data(), ext.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
  cudaStreamSynchronize(ctx->stream);                   // the vectors are locals
  slot.generation = S->generation; slot.symElems = base[tiles]; slot.extFull = ((2 * std::max(S->NE, 1) + 7) / 8) * 8; slot.NE = S->NE;
  return &slot;
}
void far_set_zero_event(scuffcuda_context *ctx, cudaEvent_t ev)
{ std::lock_guard<std::mutex> lock(g_farSlotsLock);
  g_twoPass[ctx].zeroEvent = ev;
}
void far_twopass_release(scuffcuda_context *ctx)
{ std::lock_guard<std::mutex> lock(g_farSlotsLock);
  auto it = g_twoPass.find(ctx);
  if (it == g_twoPass.end()) return;
  if (it->second.Q) cudaFree(it->second.Q);
  if (it->second.g0) { cudaEventDestroy(it->second.g0); cudaEventDestroy(it->second.g1); }
  for (auto &e : it->second.edges) if (e.second.edges) { cudaFree(e.second.edges); cudaFree(e.second.base); cudaFree(e.second.ext); }
  g_twoPass.erase(it);
}

template <int KMODE, int SMALL>
static void launch_far_pass1_inst(scuffcuda_context *ctx, const FarArgsQ &X, dim3 grid, size_t smem)
{ cudaFuncSetAttribute(far_pass1_kernel<KMODE, SMALL>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  far_pass1_kernel<KMODE, SMALL><<<grid, FAR_THREADS, smem, ctx->stream>>>(X);
}
template <int KMODE>
static void launch_far_pass1(scuffcuda_context *ctx, const FarArgsQ &X, dim3 grid, size_t smem, int small)
{ if (small == 2) launch_far_pass1_inst<KMODE, 2>(ctx, X, grid, smem);
  else if (small == 1) launch_far_pass1_inst<KMODE, 1>(ctx, X, grid, smem);
  else launch_far_pass1_inst<KMODE, 0>(ctx, X, grid, smem);
}

template <int KMODE, int SMALL>
static void launch_far_pass1_psi_inst(scuffcuda_context *ctx, const FarArgsQ &X, dim3 grid, size_t smem)
{ cudaFuncSetAttribute(far_pass1_psi_kernel<KMODE, SMALL>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  far_pass1_psi_kernel<KMODE, SMALL><<<grid, FAR_THREADS, smem, ctx->stream>>>(X);
}
template <int KMODE>
static void launch_far_pass1_psi(scuffcuda_context *ctx, const FarArgsQ &X, dim3 grid, size_t smem, int small)
{ if (small == 2) launch_far_pass1_psi_inst<KMODE, 2>(ctx, X, grid, smem);
  else if (small == 1) launch_far_pass1_psi_inst<KMODE, 1>(ctx, X, grid, smem);
  else launch_far_pass1_psi_inst<KMODE, 0>(ctx, X, grid, smem);
}

// Second pass of the two-pass path: called by assemble_block once the near/singular kernels of this region pass have finished
// stamping (the gather adds to their entries without atomics).  Returns 1 if it also wrote the mirror image of a symmetric block.
// gather mode: 1 = atomic adds, concurrent with the near kernels, mirror left to symmetrize (default); 0 = read-modify-write after
// the near kernels, mirror included (SCUFF_CUDA_FAR_GATHER_RED=0)
int far_gather_red()
{ const char *e = getenv("SCUFF_CUDA_FAR_GATHER_RED"); return e ? atoi(e) : 1; }
int far_twopass_pending(scuffcuda_context *ctx)
{ std::lock_guard<std::mutex> lock(g_farSlotsLock);
  auto it = g_twoPass.find(ctx);
  return it != g_twoPass.end() && it->second.pending;
}
int launch_far_gather(scuffcuda_context *ctx)
{ bool red = far_gather_red() != 0;
  FarTwoPassState *T;
  { std::lock_guard<std::mutex> lock(g_farSlotsLock);
    auto it = g_twoPass.find(ctx);
    if (it == g_twoPass.end() || !it->second.pending) return 0;
    T = &it->second; }
  T->pending = false;
  if (T->G.colHi != 0x7fffffff) red = true;         // column blocks are never mirrored: the read-modify-write variant writes mirror images
  // column tiles that intersect the owned columns (b-edges wb * eb + dc in [colLo - colOff, colHi - colOff))
  int bx0 = 0, bx1 = (T->NEb + 31) / 32;
  if (T->G.colHi != 0x7fffffff)
   { const int wb = (T->G.L.nc == 2 && !T->G.bPEC) ? 2 : 1;
     const int e0 = std::max(0, (T->G.colLo - T->G.colOff) / wb), e1 = std::min(T->NEb, (T->G.colHi - T->G.colOff + wb - 1) / wb);
     bx0 = e0 / 32; bx1 = std::max(bx0, (e1 + 31) / 32); }
  T->G.bxOff = bx0;
  if (bx1 <= bx0) { T->timed = false; return 0; }
  dim3 grid(((T->NEa + 31) / 32 + FAR_GATHER_GR - 1) / FAR_GATHER_GR, bx1 - bx0), block(32, 8);
  if (!T->g0) { cudaEventCreate(&T->g0); cudaEventCreate(&T->g1); }
  cudaEventRecord(T->g0, ctx->stream);
  ctx->launches++;
#define FAR_GATHER_LAUNCH(K, W, SYM, RED) K<W, SYM, RED><<<grid, block, 0, ctx->stream>>>(T->G)
#define FAR_GATHER_PICK(K) do { \
    if (T->sym) { if (T->isReal) { if (red) FAR_GATHER_LAUNCH(K, 1, true, true); else FAR_GATHER_LAUNCH(K, 1, true, false); } \
                  else { if (red) FAR_GATHER_LAUNCH(K, 2, true, true); else FAR_GATHER_LAUNCH(K, 2, true, false); } } \
    else { if (T->isReal) { if (red) FAR_GATHER_LAUNCH(K, 1, false, true); else FAR_GATHER_LAUNCH(K, 1, false, false); } \
           else { if (red) FAR_GATHER_LAUNCH(K, 2, false, true); else FAR_GATHER_LAUNCH(K, 2, false, false); } } } while (0)
  if (T->G.L.nc == 2) FAR_GATHER_PICK(far_gather_psi_kernel); else FAR_GATHER_PICK(far_gather_kernel);
  cudaEventRecord(T->g1, ctx->stream); T->timed = true;
  return (T->sym && !red) ? 1 : 0;
}
// duration of the last gather (blocks until it has finished)
double far_gather_elapsed_ms(scuffcuda_context *ctx)
{ std::lock_guard<std::mutex> lock(g_farSlotsLock);
  auto it = g_twoPass.find(ctx);
  if (it == g_twoPass.end() || !it->second.timed) return 0.0;
  it->second.timed = false;
  cudaEventSynchronize(it->second.g1);
  float t = 0; cudaEventElapsedTime(&t, it->second.g0, it->second.g1);
  return t;
}

template <int KMODE, int SMALL>
static void launch_far_staged_inst(scuffcuda_context *ctx, const FarArgsX &X, dim3 grid, size_t smem)
{ cudaFuncSetAttribute(far_staged_kernel<KMODE, SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(far_staged_kernel<KMODE, SMALL>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  far_staged_kernel<KMODE, SMALL><<<grid, FAR_THREADS, smem, ctx->stream>>>(X);
}
template <int KMODE>
static void launch_far_staged(scuffcuda_context *ctx, const FarArgsX &X, dim3 grid, size_t smem, int small)
{ if (small == 2) launch_far_staged_inst<KMODE, 2>(ctx, X, grid, smem);
  else if (small == 1) launch_far_staged_inst<KMODE, 1>(ctx, X, grid, smem);
  else launch_far_staged_inst<KMODE, 0>(ctx, X, grid, smem);
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
  // small-phase variant when |k| (R_a + R_b) <= 1/8 (or 1/16: two polynomial degrees less) for every panel pair — fine meshes,
  // i.e. every large problem: exp(ik r) = exp(ik r0) exp(ik (r - r0)) with the pair factor applied once per panel pair and a
  // short Taylor polynomial, no range reduction and no quadrant logic, for the point-pair factor.  SCUFF_CUDA_FAR_SMALL=0
  // forces the general path.
  // (read at every launch: the tests toggle them inside one process)
  int smallEnv = -1, stagedEnv = 1;
  { const char *e = getenv("SCUFF_CUDA_FAR_SMALL"); if (e) smallEnv = atoi(e); }
  { const char *e = getenv("SCUFF_CUDA_FAR_STAGED"); if (e) stagedEnv = atoi(e); }
  const double kk = (kmode == KMODE_REAL) ? fabs(A.K.kr) : fabs(A.K.ki);
  const double span = kk * (surface_rmax(ctx, A.pa) + surface_rmax(ctx, A.pb));
  int small = 0;
  if (smallEnv != 0 && kmode != KMODE_CPLX) { if (span <= 0.0625) small = 2; else if (span <= 0.125) small = 1; }
  if (smallEnv == 1 && small == 2) small = 1;
  ctx->launches++;
  int twoPassEnv = 1;
  { const char *e = getenv("SCUFF_CUDA_FAR_TWOPASS"); if (e) twoPassEnv = atoi(e); }
  if (twoPassEnv)
   { // two-pass path (plain stores to the panel-pair array, gathered beside / after the near kernels); with a column range only the
     // pairs and entries of the owned columns
     std::unique_lock<std::mutex> lock(g_farSlotsLock);
     FarTwoPassState &T = g_twoPass[ctx];
     lock.unlock();
     const FarTwoPassState::SurfTab *ta = far_surface_tables(ctx, T, A.pa), *tb = far_surface_tables(ctx, T, A.pb);
     bool ok = ta && tb;
     const size_t need = !ok ? 0 : sizeof(cd) * (size_t)(psi ? 2 : 1) * (size_t)(A.symmetric ? ta->symElems : (long)A.NPb * 3 * ta->extFull);
     double qmaxGB = 40.0;
     { const char *e = getenv("SCUFF_CUDA_FAR_QMAX_GB"); if (e) qmaxGB = atof(e); }
     ok = ok && (double)need <= qmaxGB * 1073741824.0;
     if (ok && need > T.qBytes)
      { if (T.Q) { cudaStreamSynchronize(ctx->stream); cudaFree(T.Q); T.Q = nullptr; T.qBytes = 0; }
        size_t freeB = 0, totalB = 0; cudaMemGetInfo(&freeB, &totalB);
        if (need + (size_t)(2ull << 30) > freeB || cudaMalloc(&T.Q, need) != cudaSuccess) { cudaGetLastError(); T.Q = nullptr; ok = false; }
        else T.qBytes = need; }
     if (ok)
      { FarArgsQ X; X.A = A; X.L.Q = T.Q; X.L.symmetric = A.symmetric; X.L.base = ta->base; X.L.ext = ta->ext; X.L.extFull = ta->extFull; X.L.nc = psi ? 2 : 1;
        const int4 *ea = ta->edges, *eb = tb->edges; const int NEa = ta->NE, NEb = tb->NE;
        if (psi)
         { if (kmode == KMODE_REAL) launch_far_pass1_psi<KMODE_REAL>(ctx, X, grid, smem, small);
           else if (kmode == KMODE_IMAG) launch_far_pass1_psi<KMODE_IMAG>(ctx, X, grid, smem, small);
           else launch_far_pass1_psi<KMODE_CPLX>(ctx, X, grid, smem, 0); }
        else if (kmode == KMODE_REAL) launch_far_pass1<KMODE_REAL>(ctx, X, grid, smem, small);
        else if (kmode == KMODE_IMAG) launch_far_pass1<KMODE_IMAG>(ctx, X, grid, smem, small);
        else launch_far_pass1<KMODE_CPLX>(ctx, X, grid, smem, 0);
        T.G.L = X.L; T.G.edgeA = ea; T.G.edgeB = eb; T.G.NEa = NEa; T.G.NEb = NEb;
        T.G.M = A.S.M; T.G.ld = A.S.ld; T.G.rowOff = A.S.rowOff; T.G.colOff = A.S.colOff;
        T.G.aPEC = A.S.aPEC; T.G.bPEC = A.S.bPEC; T.G.c1 = A.S.c1; T.G.c2 = A.S.c2; T.G.c3 = A.S.c3;
        T.G.colLo = A.S.colLo; T.G.colHi = A.S.colHi; T.G.bxOff = 0;
        T.sym = A.symmetric != 0; T.isReal = A.S.isReal; T.NEa = NEa; T.NEb = NEb; T.pending = true;
        return; } }
  { // every other path accumulates into the matrix with atomics: the block must have been cleared
    cudaEvent_t zev = nullptr;
    { std::lock_guard<std::mutex> lock(g_farSlotsLock); auto it = g_twoPass.find(ctx); if (it != g_twoPass.end()) zev = it->second.zeroEvent; }
    if (zev) cudaStreamWaitEvent(ctx->stream, zev, 0); }
  if (!psi && stagedEnv)
   { // PEC/PEC: staged scatter (dense shared-memory window + bulk reductions)
     FarArgsX X; X.A = A; X.slots = far_slots(ctx, A.pa); X.nElem = 0; X.dbg = 0;
     { const char *e = getenv("SCUFF_CUDA_FAR_DBG"); if (e) X.dbg = atoi(e); }      // timing experiments only (1: no bulk ops, 2: plain stores: wrong results)
     for (const auto &kv : ctx->matrices) if (kv.second.d == A.S.M) X.nElem = (long)kv.second.NR * kv.second.NC;
     if (X.slots && X.nElem > 0)
      { smem = FAR_STAGED_SMEM_BYTES(TB);
        if (kmode == KMODE_REAL) launch_far_staged<KMODE_REAL>(ctx, X, grid, smem, small);
        else if (kmode == KMODE_IMAG) launch_far_staged<KMODE_IMAG>(ctx, X, grid, smem, small);
        else launch_far_staged<KMODE_CPLX>(ctx, X, grid, smem, 0);
        if (X.dbg & 16)
         { unsigned long long h[16]; cudaStreamSynchronize(ctx->stream);
           cudaMemcpy(h, A.counters, sizeof h, cudaMemcpyDeviceToHost);
           const double tot = (double)(h[8] + h[9] + h[10] + h[11] + h[12] + h[13]);
           fprintf(stderr, "far_staged phases (warp cycles, %% of total): compute %.1f  wait-read %.1f  stage %.1f  bulk-issue %.1f  outliers %.1f  drain %.1f   [total %.3g cycles, pairs %llu]\n",
                   100 * h[8] / tot, 100 * h[9] / tot, 100 * h[10] / tot, 100 * h[11] / tot, 100 * h[12] / tot, 100 * h[13] / tot, tot, h[5]); }
        return; } }
  if (kmode == KMODE_REAL) launch_far_mode<KMODE_REAL>(ctx, A, grid, smem, psi, small);
  else if (kmode == KMODE_IMAG) launch_far_mode<KMODE_IMAG>(ctx, A, grid, smem, psi, small);
  else launch_far_mode<KMODE_CPLX>(ctx, A, grid, smem, psi, small);
}

}  // namespace scuffcuda
