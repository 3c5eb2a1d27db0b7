// FIPPI builder kernel, one template instantiation per integration case so that the three heavy adaptive-cubature
// instantiations compile in parallel (fippi_case{0,1,2}.cu; each takes minutes in ptxas):
//   CASE 0: no common vertex   -> 3-D hcubature of CFDIntegrand3D        (QIFIPPICubature.cc:330-452)
//   CASE 1: common edge        -> 1-D hcubature (GK15) of the Taylor-Duffy reduced integrand (QIFIPPITaylorDuffyV2P0.cc)
//   CASE 2: common vertex      -> 2-D hcubature of the same
// Compiled with -fmad=false (see fippi_kernel.cu).
#pragma once
#include "kernels.h"
#include "fippi_integrands.cuh"

namespace scuffcuda {

#define FIPPI_REGION_BYTES 608   /* >= sizeof(HCRegion<3>) */
static_assert(sizeof(HCRegion<3>) <= FIPPI_REGION_BYTES, "region size");

// verts: 18 doubles per item (OVa[0..2], OVb[0..2]); out: 36 doubles per item; status: per item
// (0 ok, 1 region pool exhausted, 3 case mismatch)
template <int CASE>
__global__ void __launch_bounds__(64)
fippi_build_kernel(const FIPPITables T, const double *verts, const int *items, int nItems,
                   char *pool, int capacity, double *out, int *status)
{ const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nItems) return;
  const int it = items[t];
  const double *V = verts + 18 * (size_t)it;
  const double *OVa[3] = { V, V + 3, V + 6 }, *OVb[3] = { V + 9, V + 12, V + 15 };
  char *mypool = pool + (size_t)t * ((size_t)capacity * (FIPPI_REGION_BYTES + sizeof(int)));
  int *heap = (int *)(mypool + (size_t)capacity * FIPPI_REGION_BYTES);
  double val[HC_FDIM], err[HC_FDIM];
  int st;
  if (CASE == 0)
   { CFD3D f; f.setup(OVa, OVb);
     st = hcubature_run<3>(f, 1000000, 1.0e-10, 1.0e-6, (HCRegion<3> *)mypool, heap, capacity, val, err);
   }
  else
   { TDFIPPI f; f.setup(&T, OVa[0], OVa[1], OVa[2], OVb[1], OVb[2]);
     if ((f.commonEdge != 0) != (CASE == 1)) st = 3;
     else if (CASE == 1) st = hcubature_run<1>(f, 10000, 0.0, 1.0e-5, (HCRegion<1> *)mypool, heap, capacity, val, err);
     else                st = hcubature_run<2>(f, 10000, 0.0, 1.0e-5, (HCRegion<2> *)mypool, heap, capacity, val, err);
   }
  status[t] = st;
  if (st) return;
  double *o = out + FIPPI_REC_DOUBLES * (size_t)it;
  for (int j = 0; j < HC_FDIM; j++) o[j] = val[j];
  for (int m = 0; m < 3; m++) o[33 + m] = (OVa[0][m] + OVa[1][m] + OVa[2][m]) / 3.0;
}

typedef void (*fippi_launch_fn)(cudaStream_t, const FIPPITables &, const double *, const int *, int, char *, int, double *, int *);
void fippi_launch_case0(cudaStream_t s, const FIPPITables &T, const double *verts, const int *items, int n, char *pool, int cap, double *out, int *status);
void fippi_launch_case1(cudaStream_t s, const FIPPITables &T, const double *verts, const int *items, int n, char *pool, int cap, double *out, int *status);
void fippi_launch_case2(cudaStream_t s, const FIPPITables &T, const double *verts, const int *items, int n, char *pool, int cap, double *out, int *status);

#define FIPPI_DEFINE_LAUNCH(CASE)                                                                                              \
  void fippi_launch_case##CASE(cudaStream_t s, const FIPPITables &T, const double *verts, const int *items, int n, char *pool, \
                               int cap, double *out, int *status)                                                              \
  { fippi_build_kernel<CASE><<<(n + 63) / 64, 64, 0, s>>>(T, verts, items, n, pool, cap, out, status); }

}  // namespace scuffcuda
