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

// near / singular pairs (near.cu)
int run_near(scuffcuda_context *ctx, int nsa, int nsb, const scuffcuda_ssi_args *args,
             const KParams &K, const StampParams &SP, int kmode, double kabs);

// matrix utilities (matrix_kernels.cu)
void launch_zero_block(scuffcuda_context *ctx, const Matrix &M, int rowOff, int nRows, int colOff, int nCols);
void launch_copy_block(scuffcuda_context *ctx, const Matrix &M, int dstRow, int dstCol, int nRows, int nCols, int srcRow, int srcCol);
void launch_symmetrize(scuffcuda_context *ctx, const Matrix &M, int off, int n);
double run_fp64_peak(scuffcuda_context *ctx);

}  // namespace scuffcuda
