// TEST INFRASTRUCTURE: host build (g++) of the SCUFF_HD inline math used by the CUDA kernels
// (scuff-em_b200/csrc/*.cuh), so that formulas can be checked against the reference on a machine
// without a GPU.  Never loaded by the product path.
#include <cstring>
#include <cmath>
#include "../scuff-em_b200/csrc/ppi_math.cuh"
#include "../scuff-em_b200/csrc/classify.cuh"
#include "../scuff-em_b200/csrc/fippi_math.cuh"
using namespace scuffcuda;

static void panel_local(const double *V, double C[3], double &R, double x[18])
{ static const double TU[6] = SCUFF_TCR4_U, TV[6] = SCUFF_TCR4_V;
  for (int i = 0; i < 3; i++) C[i] = (V[i] + V[3 + i] + V[6 + i]) / 3.0;
  R = 0;
  for (int v = 0; v < 3; v++)
   { double d = 0; for (int i = 0; i < 3; i++) d += (C[i] - V[3 * v + i]) * (C[i] - V[3 * v + i]);
     R = fmax(R, sqrt(d)); }
  for (int p = 0; p < 6; p++) for (int m = 0; m < 3; m++)
   { double X = V[m] + TU[p] * (V[3 + m] - V[m]); X = X + TV[p] * (V[6 + m] - V[m]); x[3 * p + m] = X - C[m]; }
}

extern "C" {

// One PPI (H0,H1) for panels Va,Vb (9 doubles each, original vertex order) with RWG vertices iQa,iQb.
// rec: 36-double FIPPI record (used when the pair is near); info[0]=ncv, info[1]=flipped, info[2]=algorithm
// keyout: 15-float FIPPI key; OV: 18 doubles canonical vertices
int hh_ppi(const double *Va, const double *Vb, int iQa, int iQb, double kr, double ki, const double *rec,
           double *out4, int *info, float *keyout, double *OV)
{ double Ca[3], Cb[3], Ra, Rb, xa[18], xb[18], D[3], L[3] = { 0, 0, 0 };
  panel_local(Va, Ca, Ra, xa); panel_local(Vb, Cb, Rb, xb);
  double rRel = pair_rrel(Ca, Ra, Cb, Rb, false, L, D);
  KParams K = make_kparams(kr, ki);
  double qa[3], qb[3];
  for (int m = 0; m < 3; m++) { qa[m] = Va[3 * iQa + m] - Ca[m]; qb[m] = Vb[3 * iQb + m] - Cb[m]; }
  cd H0, H1;
  unsigned char ia[3] = { 0, 1, 2 }, ib[3] = { 0, 1, 2 }, oa[3], ob[3];
  int ncv = 0;
  if (!(rRel > 2.0)) ncv = assess_vertices(Va, Vb, fmax(Ra, Rb), ia, ib);
  const double *VaO[3] = { Va + 3 * ia[0], Va + 3 * ia[1], Va + 3 * ia[2] };
  const double *VbO[3] = { Vb + 3 * ib[0], Vb + 3 * ib[1], Vb + 3 * ib[2] };
  int flipped = canonical_order(VaO, VbO, ncv, oa, ob);
  const double *const *PA = flipped ? VbO : VaO, *const *PB = flipped ? VaO : VbO;
  const double *OVa[3] = { PA[oa[0]], PA[oa[1]], PA[oa[2]] };
  const double *OVb[3] = { PB[ob[0]], PB[ob[1]], PB[ob[2]] };
  if (keyout) fippi_key(OVa, OVb, keyout);
  if (OV) for (int v = 0; v < 3; v++) for (int m = 0; m < 3; m++) { OV[3 * v + m] = OVa[v][m]; OV[9 + 3 * v + m] = OVb[v][m]; }
  info[0] = ncv; info[1] = flipped;
  if (rRel > SCUFF_DESING_RADIUS)
   { info[2] = 0;
     if (ki == 0.0) { Moments<cd> M; panel_pair_moments<KMODE_REAL, false>(xa, xb, D, K, M); contract_moments(M, qa, qb, D, K, H0, H1); }
     else if (kr == 0.0) { Moments<double> M; panel_pair_moments<KMODE_IMAG, false>(xa, xb, D, K, M); contract_moments(M, qa, qb, D, K, H0, H1); }
     else { Moments<cd> M; panel_pair_moments<KMODE_CPLX, false>(xa, xb, D, K, M); contract_moments(M, qa, qb, D, K, H0, H1); }
   }
  else
   { info[2] = 4;
     Moments<cd> M; panel_pair_moments<KMODE_CPLX, true>(xa, xb, D, K, M); contract_moments(M, qa, qb, D, K, H0, H1);
     if (rec)
      { const double *const *CV = flipped ? VbO : VaO; double Cen[3];
        for (int m = 0; m < 3; m++) Cen[m] = (CV[0][m] + CV[1][m] + CV[2][m]) / 3.0;
        const double *Qa = Va + 3 * iQa, *Qb = Vb + 3 * iQb;
        QDFIPPI q; qdfippi(OVa, flipped ? Qb : Qa, OVb, flipped ? Qa : Qb, Qa, Qb, Cen, flipped, rec, q);
        cd F0 = mk(0, 0), F1 = mk(0, 0); add_fippi_terms(q, K, F0, F1); H0 += F0; H1 += F1;
      }
   }
  out4[0] = H0.x; out4[1] = H0.y; out4[2] = H1.x; out4[3] = H1.y;
  return 0;
}

}

// ---- Taylor-Duffy common-triangle check ----
#include "../scuff-em_b200/csrc/td_math.cuh"
#include "../scuff-em_b200/csrc/generated_tables.h"
extern "C" int hh_td_ct(const double *V, int iQa, int iQb, double kr, double ki, double *out4)
{ static_assert(sizeof(SCUpsEntry) == sizeof(UpsEntry), "layout");
  TDTables T; T.ccx = SC_CC_X_H; T.ccw = SC_CC_W_H; T.ups = (const UpsEntry *)SC_UPS_H;
  for (int i = 0; i < 5; i++) T.upsStart[i] = SC_UPS_START_H[i];
  static double cache[TD_CACHE_DOUBLES];
  cd H0, H1;
  int st = td_common_triangle(T, V, V + 3, V + 6, V + 3 * iQa, V + 3 * iQb, mk(kr, ki), cache, 1, H0, H1);
  out4[0] = H0.x; out4[1] = H0.y; out4[2] = H1.x; out4[3] = H1.y;
  return st;
}

// ---- FIPPI builder check (adaptive cubature replica) ----
#include "../scuff-em_b200/csrc/fippi_integrands.cuh"
#include <vector>
extern "C" int hh_fippi(const double *V, int ncv, double *out36, int *nregions)
{ static_assert(sizeof(SCPXEntry) == sizeof(PXEntry) && sizeof(SCPCEntry) == sizeof(PCEntry), "layout");
  FIPPITables T;
  T.cePX = (const PXEntry *)SC_CE_PX_H; T.cePXn = SC_CE_PX_N; T.cePC = (const PCEntry *)SC_CE_PC_H; T.cePCn = SC_CE_PC_N;
  T.cvPX = (const PXEntry *)SC_CV_PX_H; T.cvPXn = SC_CV_PX_N; T.cvPC = (const PCEntry *)SC_CV_PC_H; T.cvPCn = SC_CV_PC_N;
  const int cap = 16000;
  std::vector<char> pool((size_t)cap * 608); std::vector<int> heap(cap);
  const double *OVa[3] = { V, V + 3, V + 6 }, *OVb[3] = { V + 9, V + 12, V + 15 };
  double val[HC_FDIM], err[HC_FDIM]; int st;
  if (ncv == 0) { CFD3D f; f.setup(OVa, OVb); st = hcubature_run<3>(f, 1000000, 1.0e-10, 1.0e-6, (HCRegion<3> *)pool.data(), heap.data(), cap, val, err); }
  else
   { TDFIPPI f; f.setup(&T, OVa[0], OVa[1], OVa[2], OVb[1], OVb[2]);
     if (f.commonEdge) st = hcubature_run<1>(f, 10000, 0.0, 1.0e-5, (HCRegion<1> *)pool.data(), heap.data(), cap, val, err);
     else st = hcubature_run<2>(f, 10000, 0.0, 1.0e-5, (HCRegion<2> *)pool.data(), heap.data(), cap, val, err); }
  for (int j = 0; j < 33; j++) out36[j] = val[j];
  for (int m = 0; m < 3; m++) out36[33 + m] = (OVa[0][m] + OVa[1][m] + OVa[2][m]) / 3.0;
  (void)nregions;
  return st;
}

// ---- periodic Green's function (Ewald) check: host build of csrc/gbar.cuh ----
#include "../scuff-em_b200/csrc/gbar.cuh"
extern "C" int hh_gbar(double kr, double ki, const double *kBloch, const double *LBV4, int exclude, int npts, const double *R, double *out16)
{ GBarParams P; P.k = mk(kr, ki); P.kBloch[0] = kBloch[0]; P.kBloch[1] = kBloch[1];
  P.LBV[0][0] = LBV4[0]; P.LBV[0][1] = LBV4[1]; P.LBV[1][0] = LBV4[2]; P.LBV[1][1] = LBV4[3];
  P.LDim = 2; P.excludeInner = exclude;
  for (int n = 0; n < npts; n++)
   { cd G[8]; gbar_vd_ewald(R + 3 * n, P, G);
     for (int i = 0; i < 8; i++) { out16[16 * n + 2 * i] = G[i].x; out16[16 * n + 2 * i + 1] = G[i].y; } }
  return 0;
}
extern "C" void hh_faddeeva(int n, const double *z, double *w)
{ for (int i = 0; i < n; i++) { cd r = faddeeva_w(mk(z[2 * i], z[2 * i + 1])); w[2 * i] = r.x; w[2 * i + 1] = r.y; } }

// ---- whole table chain on the host (same inline source as gbar.cu): decisions, coefficients, GetGBar ----
#include "../scuff-em_b200/csrc/gbar_table.cuh"
// info[0..2] = N, info[3..5] = XMin, info[6..8] = DX, info[9] = LMax, info[10] = full; coeffs (may be null) receives 128 doubles per cell
extern "C" int hh_gbar_table(double kr, double ki, const double *kBloch, const double *LBV4, double RhoMin, double RhoMax, double RelTol,
                             double *info, double *coeffs, int maxCoeffs, int npts, const double *R, double *out8)
{ GBarParams P; P.k = mk(kr, ki); P.kBloch[0] = kBloch[0]; P.kBloch[1] = kBloch[1];
  P.LBV[0][0] = LBV4[0]; P.LBV[0][1] = LBV4[1]; P.LBV[1][0] = LBV4[2]; P.LBV[1][1] = LBV4[3];
  P.LDim = 2; P.excludeInner = 1;
  GBarGrid Gr; gbar_grid_lmax(P, RhoMin, Gr);
  GBarTableView T; T.P = P; T.RhoMin = RhoMin; T.RhoMax = RhoMax; T.LMax = Gr.LMax; T.forceFull = RhoMin > RhoMax; T.C = 0;
  std::vector<double> C;
  if (!T.forceFull)
   { double probes[8][3], MinDelta[3] = { 1e300, 1e300, 1e300 };
     const int nProbe = gbar_grid_probes(Gr, RhoMin, RhoMax, probes);
     for (int i = 0; i < nProbe; i++)
      { double Delta[3], DX[3], Pts[11][3], Opt[3]; cd V[11][SC_GB_NSUM];
        spacing_probe_setup(P, RhoMin, RhoMax, probes[i], Delta, DX, Pts);
        for (int n = 0; n < 11; n++) gbar_vd_ewald(Pts[n], P, V[n]);
        spacing_probe_finish(probes[i], Delta, DX, V, RelTol, Opt);
        for (int m = 0; m < 3; m++) MinDelta[m] = fmin(MinDelta[m], Opt[m]); }
     gbar_grid_dims(Gr, RhoMin, RhoMax, MinDelta, 0);
     const int N1 = Gr.N[0], N2 = Gr.N[1], N3 = Gr.N[2];
     std::vector<double> Phi((size_t)N1 * N2 * N3 * 16);
     for (int n1 = 0; n1 < N1; n1++) for (int n2 = 0; n2 < N2; n2++) for (int n3 = 0; n3 < N3; n3++)
      { const double X[3] = { Gr.XMin[0] + n1 * Gr.DX[0], Gr.XMin[1] + n2 * Gr.DX[1], Gr.XMin[2] + n3 * Gr.DX[2] };
        cd V[SC_GB_NSUM]; gbar_vd_ewald(X, P, V);
        const size_t i = n3 + (size_t)N3 * (n2 + (size_t)N2 * n1);
        for (int s = 0; s < 8; s++) { Phi[16 * i + s] = V[s].x; Phi[16 * i + 8 + s] = V[s].y; } }
     const int nCells = (N1 - 1) * (N2 - 1) * (N3 - 1);
     C.resize((size_t)nCells * 128);
     for (int cell = 0; cell < nCells; cell++) for (int nf = 0; nf < 2; nf++)
      { const int n3 = cell % (N3 - 1), n2 = (cell / (N3 - 1)) % (N2 - 1), n1 = cell / ((N3 - 1) * (N2 - 1));
        double D[8][8];
        for (int c = 0; c < 8; c++)
         { const size_t p = (n3 + (c & 1)) + (size_t)N3 * ((n2 + ((c >> 1) & 1)) + (size_t)N2 * (n1 + (c >> 2)));
           for (int d = 0; d < 8; d++) D[c][d] = Phi[16 * p + 8 * nf + d]; }
        tricubic_fit(D, Gr.DX, &C[(size_t)64 * (2 * cell + nf)]); }
     for (int m = 0; m < 3; m++) { T.N[m] = Gr.N[m]; T.XMin[m] = Gr.XMin[m]; T.DX[m] = Gr.DX[m]; info[m] = Gr.N[m]; info[3 + m] = Gr.XMin[m]; info[6 + m] = Gr.DX[m]; }
     T.C = C.data();
     if (coeffs) for (size_t i = 0; i < C.size() && (int)i < maxCoeffs; i++) coeffs[i] = C[i]; }
  info[9] = Gr.LMax; info[10] = T.forceFull;
  for (int n = 0; n < npts; n++)
   { cd G, dG[3]; gbar_get(T, R + 3 * n, G, dG);
     out8[8 * n] = G.x; out8[8 * n + 1] = G.y;
     for (int m = 0; m < 3; m++) { out8[8 * n + 2 + 2 * m] = dG[m].x; out8[8 * n + 3 + 2 * m] = dG[m].y; } }
  return 0;
}
