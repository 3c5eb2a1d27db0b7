// Microbenchmark (developer tool, not part of the product): cost of the far kernel's scatter phase.
//   mode 0: 18 scalar RED.ADD.F64 per lane and iteration (what far_kernel did in round 1)
//   mode 1: lanes stage 9 complex values in shared memory, one lane issues 3 cp.reduce.async.bulk .add.f64 of ROWS*16 bytes
// Same destination pattern: warp w, iteration it -> rows [r0, r0+ROWS) of 3 columns of an N x N complex128 column-major matrix.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulkred_probe bulkred_probe.cu && ./bulkred_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE, int ROWS>
__global__ void __launch_bounds__(128) probe(double *M, int N, int iters, int fmaPerIter, double *sink, int pattern = 0)
{ extern __shared__ __align__(128) double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *stage = sm + warp * (ROWS * 3 * 2);
  const int gw = blockIdx.x * 4 + warp, nw = gridDim.x * 4;
  // lane's three "edge" rows inside the window: ~1.7 rows per lane, two lanes share some rows
  int s[3] = { (lane * 7 / 4) % ROWS, (lane * 7 / 4 + 1) % ROWS, (lane * 7 / 4 + 3) % ROWS };
  double acc = lane * 1e-3;
  for (int i = lane; i < ROWS * 6; i += 32) stage[i] = 0.0;
  if (threadIdx.x < 64) sm[4 * (ROWS * 3 * 2) + threadIdx.x] = 0.999;
  __syncthreads();
  __syncwarp();
  for (int it = 0; it < iters; it++)
   { // stand-in for the cubature loop: a dependent DFMA chain x4
     double a0 = acc, a1 = acc + 1, a2 = acc + 2, a3 = acc + 3;
     if (fmaPerIter & 1)
      { // odd count: the multiplicands come from shared memory (broadcast LDS every 4 DFMA), like the cubature loop's b-tile reads
        const double *cs = sm + 4 * (ROWS * 3 * 2);
        for (int k = 0; k < fmaPerIter; k += 4) { const double m = cs[(k >> 2) & 63]; a0 = fma(a0, m, 1e-9); a1 = fma(a1, m, 1e-9); a2 = fma(a2, m, 1e-9); a3 = fma(a3, m, 1e-9); } }
     else
     for (int k = 0; k < fmaPerIter; k += 4) { a0 = fma(a0, 0.999, 1e-9); a1 = fma(a1, 0.999, 1e-9); a2 = fma(a2, 0.999, 1e-9); a3 = fma(a3, 0.999, 1e-9); }
     acc = (a0 + a1) + (a2 + a3);
     const unsigned h = (unsigned)(gw * 2654435761u + it * 40503u);
     int r0 = (int)(h % (unsigned)(N - ROWS)) & ~1;
     int c0 = (int)((h >> 7) % (unsigned)(N - 8));
     // structured patterns of the far kernel: (1) every warp of the grid works on the same three columns at the same time, windows
     // of neighbouring warps overlap (stride 60 rows, span ROWS); (2) same columns, disjoint windows (stride ROWS); (3) random
     // columns, overlapping windows; (4) like 1 but the column changes only every second iteration (the two panels of a b-edge)
     if (pattern == 1 || pattern == 4) { r0 = (gw * 60) % (N - ROWS); c0 = (((pattern == 4 ? it / 2 : it) * 9) % (N - 8)); }
     if (pattern == 2) { r0 = (gw * ROWS) % (N - ROWS); c0 = ((it * 9) % (N - 8)); }
     if (pattern == 3) { r0 = (gw * 60) % (N - ROWS); }
     if (pattern == 5) { r0 = (gw * 60) % (N - ROWS); c0 = (((it + gw) * 9) % (N - 8)); }
     if (pattern == 6) { r0 = ((gw * 60) % (N - ROWS)) | 1; c0 = ((it * 9) % (N - 8)); }     // like 1, windows start at odd rows (16-byte aligned only)
     if (pattern == 7) { r0 = ((gw * 61) % (N - ROWS)); c0 = ((it * 9) % (N - 8)); }         // like 1, mixed parity
     if (pattern == 8) { r0 = (int)(h % (unsigned)(N - ROWS)) | 1; }                           // random windows at odd rows
     if (MODE == 0)
      {
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
         for (int j = 0; j < 3; j++)
          { double *p = M + 2 * ((size_t)(r0 + s[i]) + (size_t)(c0 + 2 * j) * N);
            atomicAdd(p, acc); atomicAdd(p + 1, -acc); }
      }
     else
      { if (lane < (MODE == 4 ? 3 : 1)) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
         for (int j = 0; j < 3; j++)
          *reinterpret_cast<double2 *>(stage + 2 * (j * ROWS + s[i])) = make_double2(acc, -acc);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (MODE == 1 && lane == 0)
         {
#pragma unroll
           for (int j = 0; j < 3; j++)
            { double *p = M + 2 * ((size_t)r0 + (size_t)(c0 + 2 * j) * N);
              asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;"
                           :: "l"(p), "r"(smem_u32(stage + 2 * j * ROWS)), "r"(ROWS * 16) : "memory"); }
           asm volatile("cp.async.bulk.commit_group;" ::: "memory");
         }
        if (MODE == 2 && lane == 0)
         { double *p0 = M + 2 * ((size_t)r0 + (size_t)(c0) * N), *p1 = p0 + 4 * (size_t)N, *p2 = p1 + 4 * (size_t)N;
           const uint32_t s0 = smem_u32(stage);
           asm volatile("{\n cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%3], %6;\n"
                        " cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%1], [%4], %6;\n"
                        " cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%2], [%5], %6;\n"
                        " cp.async.bulk.commit_group;\n}"
                        :: "l"(p0), "l"(p1), "l"(p2), "r"(s0), "r"(s0 + ROWS * 16), "r"(s0 + 2 * ROWS * 16), "r"(ROWS * 16) : "memory"); }
        if (MODE == 3 && lane == 0)
         {
#pragma unroll
           for (int j = 0; j < 3; j++)
            { double *p = M + 2 * ((size_t)r0 + (size_t)(c0 + 2 * j) * N);
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                           :: "l"(p), "r"(smem_u32(stage + 2 * j * ROWS)), "r"(ROWS * 16) : "memory"); }
           asm volatile("cp.async.bulk.commit_group;" ::: "memory");
         }
        if (MODE == 4 && lane < 3)
         { double *p = M + 2 * ((size_t)r0 + (size_t)(c0 + 2 * lane) * N);
           asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;"
                        :: "l"(p), "r"(smem_u32(stage + 2 * lane * ROWS)), "r"(ROWS * 16) : "memory");
           asm volatile("cp.async.bulk.commit_group;" ::: "memory");
         }
      }
   }
  if (MODE >= 1 && lane < 3) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (acc == 12345.678) sink[0] = acc;
}

template <int MODE, int ROWS>
static int run(double *M, int N, int ctas, int iters, int fma, double *sink, const char *label, int threads = 128, int pattern = 0)
{ cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  size_t smem = (4 * ROWS * 6 + 64) * sizeof(double);
  CK(cudaFuncSetAttribute(probe<MODE, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++)
   { cudaEventRecord(e0);
     probe<MODE, ROWS><<<ctas, threads, smem>>>(M, N, iters, fma, sink, pattern);
     cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
     float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
  const double warpIters = (double)ctas * (threads / 32) * iters;
  printf("%-34s ROWS %3d fma/iter %5d : %8.3f ms  (%.2f M warp-iterations, %.1f ns per warp-iteration per SM-slot)\n", label, ROWS, fma, best,
         warpIters / 1e6, best * 1e6 / warpIters);
  CK(cudaGetLastError());
  return 0;
}

int main()
{ const int N = 20280;
  double *M, *sink;
  CK(cudaMalloc(&M, (size_t)N * N * 16)); CK(cudaMemset(M, 0, (size_t)N * N * 16)); CK(cudaMalloc(&sink, 8));
  const int ctas = 148 * 3, iters = 1604;      // 2.85 M warp-iterations = 91.1 M panel pairs / 32
  const char *pn[9] = { "random windows", "same columns, overlapping windows", "same columns, disjoint windows", "random columns, overlapping windows",
                        "same columns twice, overlapping", "staggered columns, overlapping", "same columns, overlapping, odd rows", "same columns, overlapping, mixed parity",
                        "random windows, odd rows" };
  for (int pat : { 0, 1, 6, 7, 8 })
   { printf("-- pattern %d: %s\n", pat, pn[pat]);
     run<0, 96>(M, N, ctas, iters, 0, sink, "scalar RED x18 per lane", 128, pat);
     run<1, 96>(M, N, ctas, iters, 0, sink, "staged + 3 bulk reduce", 128, pat);
     run<3, 96>(M, N, ctas, iters, 0, sink, "staged + 3 plain bulk stores", 128, pat); }
  // correctness: total sum of the matrix must be ~0 (acc, -acc pairs) and non-trivial
  return 0;
}
