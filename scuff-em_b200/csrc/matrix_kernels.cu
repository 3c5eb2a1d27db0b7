// HMatrix-side device utilities: ZeroBlock, InsertBlock (Mate reuse), upper->lower symmetric fill
// (libs/libhmat/HMatrix.cc:381-447; AssembleBEMMatrix.cc:731-737,753-758), and the DFMA
// microbenchmark used as the FP64 roofline denominator.
#include "kernels.h"

namespace scuffcuda {

__global__ void zero_block_kernel(double *M, long ld, int w, int rowOff, int nRows, int colOff, int nCols)
{ long n = (long)nRows * nCols;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
   { int r = (int)(i % nRows), c = (int)(i / nRows);
     size_t idx = ((size_t)(rowOff + r) + (size_t)(colOff + c) * ld) * w;
     M[idx] = 0.0; if (w == 2) M[idx + 1] = 0.0;
   }
}

// upper triangle (diagonal included) of the square block at (off, off): column c holds c + 1 entries to clear.  Used for symmetric
// blocks whose lower triangle the caller overwrites with the mirror image afterwards: half the bytes of a full clear
__global__ void zero_upper_kernel(double *M, long ld, int w, int off, int n)
{ for (int c = blockIdx.x; c < n; c += gridDim.x)
   { double *col = M + ((size_t)off + (size_t)(off + c) * ld) * w;
     if (w == 2)
      { double2 *c2 = (double2 *)col; const double2 z = make_double2(0.0, 0.0);
        for (int i = threadIdx.x; i <= c; i += blockDim.x) c2[i] = z; }
     else
      for (int i = threadIdx.x; i <= c; i += blockDim.x) col[i] = 0.0; }
}

__global__ void copy_block_kernel(double *M, long ld, int w, int dstRow, int dstCol, int nRows, int nCols, int srcRow, int srcCol)
{ long n = (long)nRows * nCols;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
   { int r = (int)(i % nRows), c = (int)(i / nRows);
     size_t s = ((size_t)(srcRow + r) + (size_t)(srcCol + c) * ld) * w;
     size_t d = ((size_t)(dstRow + r) + (size_t)(dstCol + c) * ld) * w;
     M[d] = M[s]; if (w == 2) M[d + 1] = M[s + 1];
   }
}

// sparse accumulate M[rows[i], cols[i]] += vals[i] (complex values; a real matrix takes the real part, as HMatrix::AddEntry)
__global__ void add_entries_kernel(double *M, long ld, int w, int n, const int *rows, const int *cols, const double *vals)
{ int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t d = ((size_t)rows[i] + (size_t)cols[i] * ld) * w;
  atomicAdd(M + d, vals[2 * i]); if (w == 2) atomicAdd(M + d + 1, vals[2 * i + 1]);
}
int launch_add_entries(scuffcuda_context *ctx, const Matrix &M, int n, const int *rows, const int *cols, const double *vals)
{ if (n <= 0) return 0;
  int *d = nullptr;
  const size_t bi = (size_t)n * sizeof(int), bv = (size_t)2 * n * sizeof(double);
  if (cudaMalloc(&d, 2 * bi + bv + 16) != cudaSuccess) { cudaGetLastError(); return 1; }
  double *dv = (double *)((char *)d + ((2 * bi + 15) / 16) * 16);
  cudaMemcpyAsync(d, rows, bi, cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(d + n, cols, bi, cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(dv, vals, bv, cudaMemcpyHostToDevice, ctx->stream);
  ctx->launches++;
  add_entries_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(M.d, M.NR, M.isReal ? 1 : 2, n, d, d + n, dv);
  cudaStreamSynchronize(ctx->stream);          // the host arrays are the caller's
  cudaFree(d);
  return cudaGetLastError() != cudaSuccess;
}

// Bloch-phase stamp of a neighbour-cell block (StampInNeighborBlock, AssembleBEMMatrix.cc:72-114):
// M[rowOff+r, colOff+c] += BPF * B[r,c]  (+ conj(BPF) * B[c,r] when useSymmetry); both matrices complex, column-major.
// 32x32 tiles: the transposed read of B goes through shared memory so that both global reads are coalesced.
// bw = 2: B complex; bw = 1: B real (the LHM_REAL cache blocks of CreateABMBAccelerator(..., PureImagFreq = true))
__global__ void stamp_neighbor_kernel(double *M, long ldM, int rowOff, int colOff, const double *B, long ldB, int bw, int nr, int nc,
                                      double pr, double pi, int useSymmetry)
{ __shared__ double tile[32][33 * 2];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  if (useSymmetry)
   for (int j = threadIdx.y; j < 32; j += blockDim.y)
    { // element B[c0 + threadIdx.x, r0 + j]: row index runs with threadIdx.x (contiguous)
      const int rr = c0 + threadIdx.x, cc = r0 + j;
      double x = 0.0, y = 0.0;
      if (rr < nr && cc < nc) { const size_t s = ((size_t)rr + (size_t)cc * ldB) * bw; x = B[s]; y = bw == 2 ? B[s + 1] : 0.0; }
      tile[j][2 * threadIdx.x] = x; tile[j][2 * threadIdx.x + 1] = y; }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y)
   { const int r = r0 + threadIdx.x, c = c0 + j;
     if (r >= nr || c >= nc) continue;
     const size_t s = ((size_t)r + (size_t)c * ldB) * bw;
     const double bx = B[s], by = bw == 2 ? B[s + 1] : 0.0;
     double vx = pr * bx - pi * by, vy = pr * by + pi * bx;
     if (useSymmetry)
      { const double tx = tile[threadIdx.x][2 * j], ty = tile[threadIdx.x][2 * j + 1];       // B[c, r]
        vx += pr * tx + pi * ty; vy += pr * ty - pi * tx; }
     const size_t d = ((size_t)(rowOff + r) + (size_t)(colOff + c) * ldM) * 2;
     M[d] += vx; M[d + 1] += vy; }
}
void launch_stamp_neighbor(scuffcuda_context *ctx, const Matrix &M, int rowOff, int colOff, const Matrix &B, double pr, double pi, int useSymmetry)
{ dim3 grid((B.NR + 31) / 32, (B.NC + 31) / 32), block(32, 8);
  ctx->launches++;
  stamp_neighbor_kernel<<<grid, block, 0, ctx->stream>>>(M.d, M.NR, rowOff, colOff, B.d, B.NR, B.isReal ? 1 : 2, B.NR, B.NC, pr, pi, useSymmetry);
}

// lower(r,c) = upper(c,r) for r>c, 32x32 tiles through shared memory so both the read and the
// write are coalesced (column-major: consecutive rows are contiguous)
template <int W>
__global__ void symmetrize_kernel(double *M, long ld, int off, int n)
{ __shared__ double tile[32][33 * W];
  int bx = blockIdx.x, by = blockIdx.y;      // tile (row block by, col block bx) of the UPPER part: by <= bx
  if (by > bx) return;
  int tx = threadIdx.x, ty = threadIdx.y;    // 32 x 8
  for (int j = ty; j < 32; j += 8)
   { int r = by * 32 + tx, c = bx * 32 + j;
     if (r < n && c < n)
      { size_t idx = ((size_t)(off + r) + (size_t)(off + c) * ld) * W;
        for (int q = 0; q < W; q++) tile[j][tx * W + q] = M[idx + q];
      }
   }
  __syncthreads();
  for (int j = ty; j < 32; j += 8)
   { int r = bx * 32 + tx, c = by * 32 + j;   // transposed position
     if (r < n && c < n && r > c)
      { size_t idx = ((size_t)(off + r) + (size_t)(off + c) * ld) * W;
        for (int q = 0; q < W; q++) M[idx + q] = tile[tx][j * W + q];
      }
   }
}

void launch_zero_block(scuffcuda_context *ctx, const Matrix &M, int rowOff, int nRows, int colOff, int nCols)
{ if (nRows <= 0 || nCols <= 0) return;
  int w = M.isReal ? 1 : 2;
  if (rowOff == 0 && colOff == 0 && nRows == M.NR && nCols == M.NC)
   { cudaMemsetAsync(M.d, 0, (size_t)M.NR * M.NC * w * sizeof(double), ctx->stream); return; }
  ctx->launches++;
  zero_block_kernel<<<ctx->smCount * 4, 256, 0, ctx->stream>>>(M.d, M.NR, w, rowOff, nRows, colOff, nCols);
}
void launch_zero_upper(scuffcuda_context *ctx, const Matrix &M, int off, int n)
{ if (n <= 0) return;
  ctx->launches++;
  zero_upper_kernel<<<ctx->smCount * 8, 256, 0, ctx->stream>>>(M.d, M.NR, M.isReal ? 1 : 2, off, n);
}
void launch_copy_block(scuffcuda_context *ctx, const Matrix &M, int dstRow, int dstCol, int nRows, int nCols, int srcRow, int srcCol)
{ if (nRows <= 0 || nCols <= 0) return;
  ctx->launches++;
  copy_block_kernel<<<ctx->smCount * 4, 256, 0, ctx->stream>>>(M.d, M.NR, M.isReal ? 1 : 2, dstRow, dstCol, nRows, nCols, srcRow, srcCol);
}
void launch_symmetrize(scuffcuda_context *ctx, const Matrix &M, int off, int n)
{ if (n <= 1) return;
  int t = (n + 31) / 32; dim3 grid(t, t), block(32, 8);
  ctx->launches++;
  if (M.isReal) symmetrize_kernel<1><<<grid, block, 0, ctx->stream>>>(M.d, M.NR, off, n);
  else symmetrize_kernel<2><<<grid, block, 0, ctx->stream>>>(M.d, M.NR, off, n);
}

// ---- FP64 peak: 8 independent DFMA chains per thread ----
__global__ void dfma_peak_kernel(double *out, int iters)
{ double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-7;
  for (int i = 0; i < iters; i++)
   { a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
     a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

double run_fp64_peak(scuffcuda_context *ctx)
{ int blocks = ctx->smCount * 8, threads = 256, iters = 20000;
  double *d; cudaMalloc(&d, (size_t)blocks * threads * sizeof(double));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  dfma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d, 1000);
  double best = 0;
  for (int rep = 0; rep < 3; rep++)
   { cudaEventRecord(e0, ctx->stream);
     dfma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d, iters);
     cudaEventRecord(e1, ctx->stream); cudaEventSynchronize(e1);
     float ms; cudaEventElapsedTime(&ms, e0, e1);
     double tf = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
     if (tf > best) best = tf;
   }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  return best;
}

}  // namespace scuffcuda
