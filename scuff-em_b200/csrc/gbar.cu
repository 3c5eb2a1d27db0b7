// Periodic Green's function tables on the device: CreateGBarAccelerator (GBarAccelerator.cc:339-547) with every Ewald
// sum, the grid-spacing probes (GetOptimalGridSpacing2D :257-333), the table fill (Interp3D::ReInitialize
// libs/libMDInterp/Interp3D.cc:319-505) and the tricubic coefficient fit evaluated by CUDA kernels; the host only takes
// the minimum of the probed spacings, rounds it to grid dimensions and allocates.  No CPU evaluation of GBar exists.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "kernels.h"
#include "gbar_table.cuh"
#include "pbc.h"

namespace scuffcuda {

// GBarVDEwald at n points: one thread per point
__global__ void gbar_points_kernel(GBarParams P, int n, const double *R, double *out16)
{ const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double X[3] = { R[3 * i], R[3 * i + 1], R[3 * i + 2] };
  cd V[SC_GB_NSUM];
  gbar_vd_ewald(X, P, V);
  for (int s = 0; s < SC_GB_NSUM; s++) { out16[16 * i + 2 * s] = V[s].x; out16[16 * i + 2 * s + 1] = V[s].y; }
}

// one CTA per probe of GetOptimalGridSpacing2D: 11 lanes evaluate the Ewald sums of the trial cell's corners and of the
// three mid-edge points concurrently, lane 0 fits the one-cell interpolant and turns its error into spacings
__global__ void gbar_spacing_kernel(GBarParams P, double RhoMin, double RhoMax, double RelTol, const double *probes, double *opt)
{ __shared__ double sPts[11][3], sDelta[3], sDX[3];
  __shared__ cd sV[11][SC_GB_NSUM];
  const double X[3] = { probes[3 * blockIdx.x], probes[3 * blockIdx.x + 1], probes[3 * blockIdx.x + 2] };
  if (threadIdx.x == 0) spacing_probe_setup(P, RhoMin, RhoMax, X, sDelta, sDX, sPts);
  __syncthreads();
  if (threadIdx.x < 11)
   { cd V[SC_GB_NSUM];
     const double Y[3] = { sPts[threadIdx.x][0], sPts[threadIdx.x][1], sPts[threadIdx.x][2] };
     gbar_vd_ewald(Y, P, V);
     for (int s = 0; s < SC_GB_NSUM; s++) sV[threadIdx.x][s] = V[s]; }
  __syncthreads();
  if (threadIdx.x == 0)
   { double Opt[3];
     spacing_probe_finish(X, sDelta, sDX, sV, RelTol, Opt);
     for (int m = 0; m < 3; m++) opt[3 * blockIdx.x + m] = Opt[m]; }
}

// GBarVD at every grid point (GetPhiVD_Thread, Interp3D.cc:96-180): 16 doubles per point, point index n3 + N3 (n2 + N2 n1)
__global__ void gbar_fill_kernel(GBarParams P, int N1, int N2, int N3, double X1Min, double X2Min, double X3Min,
                                 double DX1, double DX2, double DX3, double *phiVD)
{ const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N1 * N2 * N3) return;
  const int n3 = i % N3, n2 = (i / N3) % N2, n1 = i / (N3 * N2);
  const double X[3] = { X1Min + n1 * DX1, X2Min + n2 * DX2, X3Min + n3 * DX3 };
  cd V[SC_GB_NSUM];
  gbar_vd_ewald(X, P, V);
  for (int s = 0; s < SC_GB_NSUM; s++) { phiVD[16 * i + s] = V[s].x; phiVD[16 * i + 8 + s] = V[s].y; }
}

// tricubic coefficients, one thread per (cell, function)
__global__ void gbar_coeff_kernel(int N1, int N2, int N3, double DX1, double DX2, double DX3, const double *phiVD, double *C)
{ const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int nCells = (N1 - 1) * (N2 - 1) * (N3 - 1);
  if (i >= 2 * nCells) return;
  const int nf = i & 1, cell = i >> 1;
  const int n3 = cell % (N3 - 1), n2 = (cell / (N3 - 1)) % (N2 - 1), n1 = cell / ((N3 - 1) * (N2 - 1));
  double D[8][8];
  for (int c = 0; c < 8; c++)
   { const int p = (n3 + (c & 1)) + N3 * ((n2 + ((c >> 1) & 1)) + N2 * (n1 + (c >> 2)));
     for (int d = 0; d < 8; d++) D[c][d] = phiVD[16 * p + 8 * nf + d]; }
  const double DX[3] = { DX1, DX2, DX3 };
  double Cf[64];
  tricubic_fit(D, DX, Cf);
  for (int n = 0; n < 64; n++) C[(size_t)64 * i + n] = Cf[n];
}

// GetGBar (value + gradient) at n points
__global__ void gbar_eval_kernel(GBarTableView T, int n, const double *R, double *out8)
{ const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double X[3] = { R[3 * i], R[3 * i + 1], R[3 * i + 2] };
  cd G, dG[3];
  gbar_get(T, X, G, dG);
  out8[8 * i] = G.x; out8[8 * i + 1] = G.y;
  for (int m = 0; m < 3; m++) { out8[8 * i + 2 + 2 * m] = dG[m].x; out8[8 * i + 3 + 2 * m] = dG[m].y; }
}

static int gb_fail(scuffcuda_context *ctx, const char *msg) { ctx->err = msg; return 1; }
#define GCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { ctx->err = std::string("gbar: ") + #call + ": " + cudaGetErrorString(e_); return 1; } } while (0)

GBarTable *find_gbar(scuffcuda_context *ctx, int handle)
{ auto it = ctx->gbarTables.find(handle); return it == ctx->gbarTables.end() ? nullptr : &it->second; }

GBarTableView gbar_view(const GBarTable &T)
{ GBarTableView V; V.P = T.P; V.RhoMin = T.RhoMin; V.RhoMax = T.RhoMax; V.LMax = T.LMax; V.forceFull = T.forceFull;
  for (int m = 0; m < 3; m++) { V.N[m] = T.N[m]; V.XMin[m] = T.XMin[m]; V.DX[m] = T.DX[m]; }
  V.C = T.dC; return V; }

int gbar_ewald_points(scuffcuda_context *ctx, const GBarParams &P, int n, const double *R, double *out16)
{ if (n <= 0) return 0;
  double *d = nullptr;
  GCK(cudaMalloc(&d, (size_t)n * 19 * sizeof(double)));
  cudaMemcpyAsync(d, R, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  ctx->launches++;
  gbar_points_kernel<<<(n + 63) / 64, 64, 0, ctx->stream>>>(P, n, d, d + 3 * (size_t)n);
  cudaMemcpyAsync(out16, d + 3 * (size_t)n, (size_t)n * 16 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  if (e != cudaSuccess) { ctx->err = std::string("gbar_points_kernel: ") + cudaGetErrorString(e); return 1; }
  return 0;
}

int gbar_create(scuffcuda_context *ctx, const GBarParams &P, double RhoMin, double RhoMax, double RelTol, int forceFull, int nMaxGrid, int *handle)
{ if (P.LDim != 2) return gb_fail(ctx, "periodic Green's function: only 2-D lattices are supported on the device path (1-D lattices need the "
                                       "Bessel-function spectral sum of GBarVDEwald.cc:336-460)");
  if (P.LBV[0][1] != 0.0 || P.LBV[1][0] != 0.0) return gb_fail(ctx, "non-rectangular 2D lattices not yet supported");   // CheckLattice, RWGGeometry.cc:113-116
  GBarTable T; T.P = P; T.RhoMin = RhoMin; T.RhoMax = RhoMax;
  GBarGrid Gr; gbar_grid_lmax(P, RhoMin, Gr);
  T.LMax = Gr.LMax;
  T.forceFull = (forceFull || RhoMin > RhoMax) ? 1 : 0;
  if (!T.forceFull)
   { // probe the interpolation error at the corners / centre of the Wigner-Seitz cell (GBarAccelerator.cc:504-514)
     double probes[8][3];
     const int nProbe = gbar_grid_probes(Gr, RhoMin, RhoMax, probes);
     double *d = nullptr;
     GCK(cudaMalloc(&d, (size_t)nProbe * 6 * sizeof(double)));
     cudaMemcpyAsync(d, &probes[0][0], (size_t)nProbe * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
     ctx->launches++;
     gbar_spacing_kernel<<<nProbe, 32, 0, ctx->stream>>>(P, RhoMin, RhoMax, RelTol, d, d + 3 * nProbe);
     std::vector<double> opt(3 * nProbe);
     cudaMemcpyAsync(opt.data(), d + 3 * nProbe, (size_t)nProbe * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
     cudaError_t e = cudaStreamSynchronize(ctx->stream);
     cudaFree(d);
     if (e != cudaSuccess) { ctx->err = std::string("gbar_spacing_kernel: ") + cudaGetErrorString(e); return 1; }
     double MinDelta[3] = { opt[0], opt[1], opt[2] };
     for (int i = 1; i < nProbe; i++) for (int m = 0; m < 3; m++) MinDelta[m] = fmin(MinDelta[m], opt[3 * i + m]);
     for (int m = 0; m < 3; m++) T.minDelta[m] = MinDelta[m];
     if (!(MinDelta[0] > 0.0) || !(MinDelta[1] > 0.0)) return gb_fail(ctx, "periodic Green's function: grid-spacing probe failed (non-finite interpolation error)");
     gbar_grid_dims(Gr, RhoMin, RhoMax, MinDelta, nMaxGrid);
     const int nx = Gr.N[0], ny = Gr.N[1], nRho = Gr.N[2];
     if ((double)nx * ny * nRho > 5.0e7) return gb_fail(ctx, "periodic Green's function: interpolation grid too large");
     for (int m = 0; m < 3; m++) { T.N[m] = Gr.N[m]; T.XMin[m] = Gr.XMin[m]; T.DX[m] = Gr.DX[m]; }
     const int nPts = nx * ny * nRho, nCells = (nx - 1) * (ny - 1) * (nRho - 1);
     double *dPhi = nullptr;
     GCK(cudaMalloc(&dPhi, (size_t)nPts * 16 * sizeof(double)));
     if (cudaMalloc(&T.dC, (size_t)nCells * 128 * sizeof(double)) != cudaSuccess) { cudaFree(dPhi); return gb_fail(ctx, "gbar: table allocation failed"); }
     ctx->launches += 2;
     gbar_fill_kernel<<<(nPts + 31) / 32, 32, 0, ctx->stream>>>(P, nx, ny, nRho, T.XMin[0], T.XMin[1], T.XMin[2], T.DX[0], T.DX[1], T.DX[2], dPhi);
     gbar_coeff_kernel<<<(2 * nCells + 63) / 64, 64, 0, ctx->stream>>>(nx, ny, nRho, T.DX[0], T.DX[1], T.DX[2], dPhi, T.dC);
     e = cudaStreamSynchronize(ctx->stream);
     cudaFree(dPhi);
     if (e != cudaSuccess) { cudaFree(T.dC); ctx->err = std::string("gbar table kernels: ") + cudaGetErrorString(e); return 1; }
   }
  const int h = ctx->nextHandle++;
  ctx->gbarTables[h] = T;
  *handle = h;
  return 0;
}

int gbar_destroy(scuffcuda_context *ctx, int handle)
{ GBarTable *T = find_gbar(ctx, handle);
  if (!T) return gb_fail(ctx, "invalid periodic Green's function table handle");
  cudaStreamSynchronize(ctx->stream);
  if (T->dC) cudaFree(T->dC);
  ctx->gbarTables.erase(handle);
  return 0;
}

int gbar_eval_points(scuffcuda_context *ctx, int handle, int n, const double *R, double *out8)
{ GBarTable *T = find_gbar(ctx, handle);
  if (!T) return gb_fail(ctx, "invalid periodic Green's function table handle");
  if (n <= 0) return 0;
  double *d = nullptr;
  GCK(cudaMalloc(&d, (size_t)n * 11 * sizeof(double)));
  cudaMemcpyAsync(d, R, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  ctx->launches++;
  gbar_eval_kernel<<<(n + 63) / 64, 64, 0, ctx->stream>>>(gbar_view(*T), n, d, d + 3 * (size_t)n);
  cudaMemcpyAsync(out8, d + 3 * (size_t)n, (size_t)n * 8 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  if (e != cudaSuccess) { ctx->err = std::string("gbar_eval_kernel: ") + cudaGetErrorString(e); return 1; }
  return 0;
}

int gbar_download_table(scuffcuda_context *ctx, int handle, double *coeffs)
{ GBarTable *T = find_gbar(ctx, handle);
  if (!T) return gb_fail(ctx, "invalid periodic Green's function table handle");
  if (T->forceFull || !T->dC) return 0;
  const size_t n = (size_t)(T->N[0] - 1) * (T->N[1] - 1) * (T->N[2] - 1) * 128;
  GCK(cudaMemcpyAsync(coeffs, T->dC, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GCK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // namespace scuffcuda
