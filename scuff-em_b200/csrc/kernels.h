// Kernel argument blocks and host-side launchers (internal; not part of the ABI).
#pragma once
#include "context.h"

namespace scuffcuda {

struct FarArgs
{ const PanelRec *pa; int NPa;
  const PanelRec *pb; int NPb;
  int TA;
  int symmetric;            // same surface, no displacement: visit pa<=pb only
  int useDisp; double L[3];
  KParams K;
  StampParams S;
  unsigned long long *counters;   // [0] += number of far PPIs the reference would have evaluated
};

void launch_far(scuffcuda_context *ctx, const FarArgs &A, int kmode);
int run_fastmath_selftest(scuffcuda_context *ctx, int n, const double *x, double *out);
#define SCUFF_MAX_KR 1.0e5   /* fast-path range of the FP64 sincos reduction (fastmath.cuh) */

#define SW_THRESHOLD (1.0 * M_PI)      // PanelPanelInteractions.cc:58
#define VERYSW_THRESHOLD 20.0          // PanelPanelInteractions.cc:59

struct NearArgs
{ const PanelRec *pa, *pb; const NearPair *near; int nNear;
  int useDisp; double L[3]; int forceTD; double kabs;
  const double *fippi;
  KParams K; StampParams S;
  unsigned long long *counters;
};

// which algorithm the reference would use for this near pair (PanelPanelInteractions.cc:426-491)
SCUFF_HD int near_algorithm(int ncv, double kabs, double Ra, double Rb, int forceTD, bool &verySW)
{ // kR = abs(k * fmax(Ra,Rb)): complex*double then abs() == |k|*R up to rounding; the thresholds
  // are far from any realistic tie
  double kR = kabs * fmax(Ra, Rb);
  bool SW = kR > SW_THRESHOLD; verySW = kR > VERYSW_THRESHOLD;
  if (SW && ncv == 0) return 1;                          // high-order cubature
  if (ncv == 3 || (ncv > 0 && (SW || forceTD))) return 2; // Taylor-Duffy (3 if very-short-wavelength)
  return 4;                                              // desingularisation
}


// Taylor-Duffy pairs of a plan (td_kernel.cu)
int run_td(scuffcuda_context *ctx, const Plan &P, const NearArgs &A, int kmode);
// device FIPPI builder (fippi_kernel.cu): computes records recIds[] from canonically ordered vertices
// (18 doubles per record) into ctx->fippi.hrec
int fippi_compute_records(scuffcuda_context *ctx, const std::vector<int> &recIds, const std::vector<double> &verts,
                          const std::vector<int> &ncvs);

// near / singular pairs (near.cu)
int run_near(scuffcuda_context *ctx, int nsa, int nsb, const scuffcuda_ssi_args *args,
             const KParams &K, const StampParams &SP, int kmode, double kabs);

// matrix utilities (matrix_kernels.cu)
void launch_zero_block(scuffcuda_context *ctx, const Matrix &M, int rowOff, int nRows, int colOff, int nCols);
void launch_zero_upper(scuffcuda_context *ctx, const Matrix &M, int off, int n);
void launch_copy_block(scuffcuda_context *ctx, const Matrix &M, int dstRow, int dstCol, int nRows, int nCols, int srcRow, int srcCol);
void launch_symmetrize(scuffcuda_context *ctx, const Matrix &M, int off, int n);
double run_fp64_peak(scuffcuda_context *ctx);

}  // namespace scuffcuda
