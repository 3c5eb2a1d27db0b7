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
