// Panel-pair integral math shared by the CUDA kernels (device) and by the host-compiled
// unit harness in tests/ (host build of the same inline functions, used only to check the
// formulas on a machine without a GPU; the product path never calls the host instantiation).
//
// Reference semantics (HomerReid/scuff-em):
//   GetPPIs_Cubature / AssembleInnerPPIIntegrand  libs/libscuff/PanelPanelInteractions.cc:108-350
//   ExpRel                                        libs/libscuff/PanelPanelInteractions.cc:84-103
//   GetEdgeEdgeInteractions                       libs/libscuff/EdgeEdgeInteractions.cc:154-188
//
// Formulation (B200-first, not the reference's): the reference evaluates every panel pair once
// per (Q,Q') RWG-vertex combination and once per edge pair (4.5x-9x redundant).  Here each
// panel pair is integrated ONCE into 18 Q-independent complex moments in panel-local
// coordinates x = X - C_a, x' = X' - C_b (no cancellation), and the nine (Q_a,Q_b)
// combinations are contracted from the moments afterwards:
//   H0(Qa,Qb) = sum w w' [(x-qa).(x'-qb) + 4/(ik)^2] Phi
//             = Sab - qa.Sb - qb.Sa + (qa.qb - 4/k^2) S0
//   H1(Qa,Qb) = sum w w' ((x-qa)x(x'-qb)).R Psi ,   R = x - x' + D,  D = C_a - C_b
//             = (D+qa-qb).Txx - qb.(D x Tx) - qa.(Txp x D) + (qa x qb).(Tx - Txp + D T0)
// with S* the Phi-moments {1, x, x', x.x'} and T* the Psi-moments {1, x, x', x cross x'}.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define SCUFF_HD __host__ __device__ __forceinline__
#else
#define SCUFF_HD inline
#endif

namespace scuffcuda {

// ---------------------------------------------------------------------------------------
// minimal complex arithmetic (no thrust / cuComplex dependency; explicit operation order)
// ---------------------------------------------------------------------------------------
struct cd { double x, y; };
SCUFF_HD cd mk(double x, double y) { cd r; r.x = x; r.y = y; return r; }
SCUFF_HD cd operator+(cd a, cd b) { return mk(a.x + b.x, a.y + b.y); }
SCUFF_HD cd operator-(cd a, cd b) { return mk(a.x - b.x, a.y - b.y); }
SCUFF_HD cd operator-(cd a) { return mk(-a.x, -a.y); }
SCUFF_HD cd operator*(cd a, cd b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
SCUFF_HD cd operator*(double s, cd a) { return mk(s * a.x, s * a.y); }
SCUFF_HD cd operator*(cd a, double s) { return mk(s * a.x, s * a.y); }
SCUFF_HD cd &operator+=(cd &a, cd b) { a.x += b.x; a.y += b.y; return a; }
SCUFF_HD cd &operator-=(cd &a, cd b) { a.x -= b.x; a.y -= b.y; return a; }
SCUFF_HD double norm2(cd a) { return a.x * a.x + a.y * a.y; }
// z1/z2 (Smith-free straightforward form; inputs here are never near overflow)
SCUFF_HD cd cdiv(cd a, cd b)
{ double d = b.x * b.x + b.y * b.y;
  return mk((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}
SCUFF_HD cd cinv(cd b) { double d = b.x * b.x + b.y * b.y; return mk(b.x / d, -b.y / d); }
// generic scalar helpers so the kernels can be instantiated for a REAL kernel
// (pure-imaginary frequency: exp(-kappa r)/4 pi r, real moments) or a COMPLEX one
SCUFF_HD double sreal(double a) { return a; }
SCUFF_HD double sreal(cd a) { return a.x; }
SCUFF_HD void fma_acc(double &acc, double s, double v) { acc += s * v; }
SCUFF_HD void fma_acc(cd &acc, double s, cd v) { acc.x += s * v.x; acc.y += s * v.y; }
SCUFF_HD void szero(double &a) { a = 0.0; }
SCUFF_HD void szero(cd &a) { a.x = 0.0; a.y = 0.0; }
SCUFF_HD cd tocd(double a) { return mk(a, 0.0); }
SCUFF_HD cd tocd(cd a) { return a; }

// ---------------------------------------------------------------------------------------
// 6-point order-4 triangle rule: the literals of libs/libTriInt/TCR.cc:83-91 (13-digit
// table; the numbers are the contract, the rule is not regenerated)
// ---------------------------------------------------------------------------------------
#define SCUFF_TCR4_NPTS 6
#define SCUFF_TCR4_U { 0.0915762135098, 0.8168475729805, 0.0915762135098, 0.1081030181681, 0.4459484909160, 0.4459484909160 }
#define SCUFF_TCR4_V { 0.0915762135098, 0.0915762135098, 0.8168475729805, 0.4459484909160, 0.1081030181681, 0.4459484909160 }
#define SCUFF_TCR4_W { 0.0549758718276, 0.0549758718276, 0.0549758718276, 0.1116907948390, 0.1116907948390, 0.1116907948390 }
SCUFF_HD double tcr4_w(int p) { return p < 3 ? 0.0549758718276 : 0.1116907948390; }

// kernel flavour
enum { KMODE_REAL = 0,    // k real            : exp(ikr) = cos + i sin
       KMODE_IMAG = 1,    // k = i*kappa       : exp(ikr) = exp(-kappa r), all-real arithmetic
       KMODE_CPLX = 2 };  // general complex k

struct KParams
{ double kr, ki;       // k = kr + i ki
  cd ik;               // i*k
  cd ooik;             // 1/(ik)
  cd c4;               // 4/(ik)^2
};

SCUFF_HD KParams make_kparams(double kr, double ki)
{ KParams K; K.kr = kr; K.ki = ki;
  K.ik = mk(-ki, kr);
  K.ooik = cinv(K.ik);
  cd ik2 = K.ik * K.ik;
  K.c4 = cdiv(mk(4.0, 0.0), ik2);
  return K;
}

template <int KMODE> struct KScalar { typedef cd type; };
template <> struct KScalar<KMODE_IMAG> { typedef double type; };

// ---------------------------------------------------------------------------------------
// ExpRel(x,n): sum_{m>=n} x^m/m! with the reference's truncation rule (stop when
// |term|^2 < 1e-16 |sum|^2, at most 100 terms) — PanelPanelInteractions.cc:84-103.
// The truncation is part of the contract (it is only converged to ~1e-8 relative).
// ---------------------------------------------------------------------------------------
SCUFF_HD cd exprel4(cd x)
{ // Term = x^3/3! after the first loop (m=1..3)
  cd Term = mk(1.0, 0.0);
  int m;
  for (m = 1; m < 4; m++)
   { cd q = mk(x.x / ((double)m), x.y / ((double)m)); Term = Term * q; }
  cd Sum = mk(0.0, 0.0);
  for (; m < 100; m++)
   { cd q = mk(x.x / ((double)m), x.y / ((double)m));
     Term = Term * q;
     Sum += Term;
     if (norm2(Term) < 1.0e-16 * norm2(Sum)) break;
   }
  return Sum;
}

// Phi (weight excluded) and Psi = Phi*(ik - 1/r)/r for one point pair.
// 1/(4 pi r) folded in.  DESING: Phi = ExpRel(ikr,4)/(4 pi r)  (PanelPanelInteractions.cc:158-171)
#define SCUFF_OO4PI 0.079577471545947673  /* 1/(4 pi) */

template <int KMODE, bool DESING> struct HelmTag {};

SCUFF_HD void helm(HelmTag<KMODE_REAL, false>, double r2, const KParams &K, cd &Phi, cd &Psi)
{ double r = sqrt(r2), rinv = 1.0 / r;
  double s, c; sincos(K.kr * r, &s, &c);
  double f = SCUFF_OO4PI * rinv;
  Phi = mk(f * c, f * s);
  // Psi = Phi * (i kr - 1/r)/r
  double a = -rinv * rinv, b = K.kr * rinv;
  Psi = mk(Phi.x * a - Phi.y * b, Phi.x * b + Phi.y * a);
}
SCUFF_HD void helm(HelmTag<KMODE_IMAG, false>, double r2, const KParams &K, double &Phi, double &Psi)
{ double r = sqrt(r2), rinv = 1.0 / r;
  Phi = SCUFF_OO4PI * rinv * exp(-K.ki * r);
  Psi = Phi * (-K.ki - rinv) * rinv;
}
SCUFF_HD void helm(HelmTag<KMODE_CPLX, false>, double r2, const KParams &K, cd &Phi, cd &Psi)
{ double r = sqrt(r2), rinv = 1.0 / r;
  double s, c; sincos(K.kr * r, &s, &c);
  double f = SCUFF_OO4PI * rinv * exp(-K.ki * r);
  Phi = mk(f * c, f * s);
  double a = -K.ki * rinv - rinv * rinv, b = K.kr * rinv;
  Psi = mk(Phi.x * a - Phi.y * b, Phi.x * b + Phi.y * a);
}
// desingularised flavour (near pairs only; always complex arithmetic; the pure-imaginary
// case yields a purely real result up to exact zeros).  Operation order follows
// PanelPanelInteractions.cc:156-171.
template <int KMODE>
SCUFF_HD void helm(HelmTag<KMODE, true>, double r2, const KParams &K, cd &Phi, cd &Psi)
{ double r = sqrt(r2);
  cd ikr = mk(K.ik.x * r, K.ik.y * r);
  cd E = exprel4(ikr);
  double f = 4.0 * M_PI * r;
  Phi = mk(E.x / f, E.y / f);
  if (!isfinite(Phi.x)) Phi = mk(0.0, 0.0);
  cd t = mk(K.ik.x - 1.0 / r, K.ik.y);
  cd p = Phi * t;
  Psi = mk(p.x / r, p.y / r);
}

// ---------------------------------------------------------------------------------------
// Q-independent moments of one panel pair
// ---------------------------------------------------------------------------------------
template <typename S> struct Moments
{ S S0, Sa[3], Sb[3], Sab;      // Phi moments: 1, x, x', x.x'
  S T0, Ta[3], Tb[3], Tx[3];    // Psi moments: 1, x, x', x cross x'
};

// xa, xb: the 6 cubature points of each panel in LOCAL coordinates (relative to own centroid);
// D = C_a - C_b (minus lattice displacement).
template <int KMODE, bool DESING, typename XA, typename XB>
SCUFF_HD void panel_pair_moments(const XA &xa, const XB &xb, const double D[3], const KParams &K,
                                 Moments<typename KScalar<DESING ? KMODE_CPLX : KMODE>::type> &M)
{ typedef typename KScalar<DESING ? KMODE_CPLX : KMODE>::type S;
  szero(M.S0); szero(M.Sab); szero(M.T0);
#pragma unroll
  for (int i = 0; i < 3; i++) { szero(M.Sa[i]); szero(M.Sb[i]); szero(M.Ta[i]); szero(M.Tb[i]); szero(M.Tx[i]); }

#pragma unroll 1
  for (int p = 0; p < SCUFF_TCR4_NPTS; p++)
   { const double x0 = xa[3 * p + 0], x1 = xa[3 * p + 1], x2 = xa[3 * p + 2];
     const double y0 = x0 + D[0], y1 = x1 + D[1], y2 = x2 + D[2];
     S a0, a1[3], b0, b1[3];
     szero(a0); szero(b0);
#pragma unroll
     for (int i = 0; i < 3; i++) { szero(a1[i]); szero(b1[i]); }
#pragma unroll
     for (int pp = 0; pp < SCUFF_TCR4_NPTS; pp++)
      { const double xp0 = xb[3 * pp + 0], xp1 = xb[3 * pp + 1], xp2 = xb[3 * pp + 2];
        const double R0 = y0 - xp0, R1 = y1 - xp1, R2 = y2 - xp2;
        const double r2 = R0 * R0 + R1 * R1 + R2 * R2;
        S Phi, Psi;
        helm(HelmTag<KMODE, DESING>(), r2, K, Phi, Psi);
        const double wp = tcr4_w(pp);
        fma_acc(a0, wp, Phi);
        fma_acc(b0, wp, Psi);
        const double wx0 = wp * xp0, wx1 = wp * xp1, wx2 = wp * xp2;
        fma_acc(a1[0], wx0, Phi); fma_acc(a1[1], wx1, Phi); fma_acc(a1[2], wx2, Phi);
        fma_acc(b1[0], wx0, Psi); fma_acc(b1[1], wx1, Psi); fma_acc(b1[2], wx2, Psi);
      }
     const double w = tcr4_w(p);
     const double wx0 = w * x0, wx1 = w * x1, wx2 = w * x2;
     fma_acc(M.S0, w, a0);
     fma_acc(M.Sb[0], w, a1[0]); fma_acc(M.Sb[1], w, a1[1]); fma_acc(M.Sb[2], w, a1[2]);
     fma_acc(M.Sa[0], wx0, a0); fma_acc(M.Sa[1], wx1, a0); fma_acc(M.Sa[2], wx2, a0);
     fma_acc(M.Sab, wx0, a1[0]); fma_acc(M.Sab, wx1, a1[1]); fma_acc(M.Sab, wx2, a1[2]);
     fma_acc(M.T0, w, b0);
     fma_acc(M.Tb[0], w, b1[0]); fma_acc(M.Tb[1], w, b1[1]); fma_acc(M.Tb[2], w, b1[2]);
     fma_acc(M.Ta[0], wx0, b0); fma_acc(M.Ta[1], wx1, b0); fma_acc(M.Ta[2], wx2, b0);
     // Tx += w * (x cross b1)
     fma_acc(M.Tx[0], wx1, b1[2]); fma_acc(M.Tx[0], -wx2, b1[1]);
     fma_acc(M.Tx[1], wx2, b1[0]); fma_acc(M.Tx[1], -wx0, b1[2]);
     fma_acc(M.Tx[2], wx0, b1[1]); fma_acc(M.Tx[2], -wx1, b1[0]);
   }
}

// contraction of the moments for one (Qa,Qb) pair: qa = Qa - C_a, qb = Qb - C_b (local)
template <typename S>
SCUFF_HD void contract_moments(const Moments<S> &M, const double qa[3], const double qb[3],
                               const double D[3], const KParams &K, cd &H0, cd &H1)
{ // H0
  double qq = qa[0] * qb[0] + qa[1] * qb[1] + qa[2] * qb[2];
  S h0 = M.Sab;
  fma_acc(h0, -qa[0], M.Sb[0]); fma_acc(h0, -qa[1], M.Sb[1]); fma_acc(h0, -qa[2], M.Sb[2]);
  fma_acc(h0, -qb[0], M.Sa[0]); fma_acc(h0, -qb[1], M.Sa[1]); fma_acc(h0, -qb[2], M.Sa[2]);
  fma_acc(h0, qq, M.S0);
  H0 = tocd(h0) + K.c4 * tocd(M.S0);
  // H1
  double e0 = D[0] + qa[0] - qb[0], e1 = D[1] + qa[1] - qb[1], e2 = D[2] + qa[2] - qb[2];
  S h1; szero(h1);
  fma_acc(h1, e0, M.Tx[0]); fma_acc(h1, e1, M.Tx[1]); fma_acc(h1, e2, M.Tx[2]);
  // - qb.(D x Ta) = -(qb x D).Ta ;  u = qb x D
  double u0 = qb[1] * D[2] - qb[2] * D[1], u1 = qb[2] * D[0] - qb[0] * D[2], u2 = qb[0] * D[1] - qb[1] * D[0];
  fma_acc(h1, -u0, M.Ta[0]); fma_acc(h1, -u1, M.Ta[1]); fma_acc(h1, -u2, M.Ta[2]);
  // - qa.(Tb x D) = -(D x qa).Tb ;  v = D x qa
  double v0 = D[1] * qa[2] - D[2] * qa[1], v1 = D[2] * qa[0] - D[0] * qa[2], v2 = D[0] * qa[1] - D[1] * qa[0];
  fma_acc(h1, -v0, M.Tb[0]); fma_acc(h1, -v1, M.Tb[1]); fma_acc(h1, -v2, M.Tb[2]);
  // (qa x qb).(Ta - Tb + D T0)
  double c0 = qa[1] * qb[2] - qa[2] * qb[1], c1 = qa[2] * qb[0] - qa[0] * qb[2], c2 = qa[0] * qb[1] - qa[1] * qb[0];
  fma_acc(h1, c0, M.Ta[0]); fma_acc(h1, c1, M.Ta[1]); fma_acc(h1, c2, M.Ta[2]);
  fma_acc(h1, -c0, M.Tb[0]); fma_acc(h1, -c1, M.Tb[1]); fma_acc(h1, -c2, M.Tb[2]);
  fma_acc(h1, c0 * D[0] + c1 * D[1] + c2 * D[2], M.T0);
  H1 = tocd(h1);
}

}  // namespace scuffcuda
