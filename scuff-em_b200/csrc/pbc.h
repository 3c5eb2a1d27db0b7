// Periodic-geometry pieces of libscuff_cuda (internal): Green's function tables (gbar.cu) and the outer-cell
// cubature kernel (pbc_kernel.cu).
#pragma once
#include "kernels.h"

namespace scuffcuda {

struct GBarTableView;
GBarTable *find_gbar(scuffcuda_context *ctx, int handle);
int gbar_create(scuffcuda_context *ctx, const GBarParams &P, double RhoMin, double RhoMax, double RelTol, int forceFull, int nMaxGrid, int *handle);
int gbar_destroy(scuffcuda_context *ctx, int handle);
int gbar_ewald_points(scuffcuda_context *ctx, const GBarParams &P, int n, const double *R, double *out16);
int gbar_eval_points(scuffcuda_context *ctx, int handle, int n, const double *R, double *out8);
int gbar_download_table(scuffcuda_context *ctx, int handle, double *coeffs);

// step 2 of the periodic assembly for one region (AssembleBEMMatrix.cc:526-557): every panel pair of (Sa, Sb) with the 6x6
// rule and GetGBar in place of the free-space kernel, accumulated into the matrix
int run_pbc_outer(scuffcuda_context *ctx, const Surface &Sa, const Surface &Sb, const GBarTable &T, const KParams &K, const StampParams &SP,
                  unsigned long long *counters);

}  // namespace scuffcuda
