// Short-wavelength-regime kernels (|k| R_panel > pi, PanelPanelInteractions.cc:426-491): launchers shared between
// near.cu / td_kernel.cu / far_kernel.cu.  Kept out of kernels.h so that fippi_kernel.cu (a 20-minute compile) does
// not depend on it.
#pragma once
#include "kernels.h"

namespace scuffcuda {
struct TDTables;

// high-order (78 x 78 point) cubature for near pairs without common vertices (PPIALG_HOCUBATURE); hoIdx: indices into
// the plan's near-pair list (host memory)
int run_ho(scuffcuda_context *ctx, const NearArgs &A, const int *hoIdx, int n, int kmode);

// Taylor-Duffy for common-vertex / common-edge pairs (PPIALG_TD), 3-D / 2-D p-adaptive Clenshaw-Curtis cubature;
// items[t] = 36*nearIndex + 18*highk + 9*swapped + 3*i + j  (host memory)
int run_td_general(scuffcuda_context *ctx, const NearArgs &A, const TDTables &T, const int *items, int nItems);

int launch_add_entries(scuffcuda_context *ctx, const Matrix &M, int n, const int *rows, const int *cols, const double *vals);
void launch_stamp_neighbor(scuffcuda_context *ctx, const Matrix &M, int rowOff, int colOff, const Matrix &B, double pr, double pi, int useSymmetry);

// Page-locked, device-mapped host memory that kernels write their per-item status words / counters into directly
// (stores over PCIe from the SMs).  No copy engine is involved, so reading them back never queues behind a bulk
// device->host matrix download in flight on the copy stream (scuffcuda_matrix_download_region_async).
void *mapped_host_scratch(scuffcuda_context *ctx, size_t bytes, int slot);

}  // namespace scuffcuda
