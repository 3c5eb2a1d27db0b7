// Device builder of the FIPPI table (QIFIPPIData records, libscuffInternals.h:282-293): one thread per
// record runs the reference's adaptive cubature decision-for-decision (hcubature.cuh) on the integrands of
// fippi_integrands.cuh.  Compiled with -fmad=false so that error estimates and threshold comparisons
// are rounded like the reference's SSE2 arithmetic.  The table is built once per geometry and stays
// resident in HBM across the frequency sweep (ComputeQIFIPPIData QIFIPPICubature.cc:465-487).
#include <vector>
#include <algorithm>
#include <thread>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include "kernels.h"
#include "fippi_build.cuh"
#include <mutex>
#include "generated_tables.h"

namespace scuffcuda {

// warp-cooperative builder for the no-common-vertex case (fippi_coop.cu)
bool fippi_launch_coop(cudaStream_t s, int c, const FIPPITables &T, const double *verts, const int *items, int n, char *pool, int cap, double *out, int *status);

static_assert(sizeof(SCPXEntry) == sizeof(PXEntry) && sizeof(SCPCEntry) == sizeof(PCEntry), "generated table layout");

struct FIPPIDeviceTables { PXEntry *cePX = nullptr, *cvPX = nullptr; PCEntry *cePC = nullptr, *cvPC = nullptr; };
static FIPPIDeviceTables g_ft[64];
static std::mutex g_ftLock;          // several contexts (one host thread each) may reach the first use together

static int fippi_tables(scuffcuda_context *ctx, FIPPITables &T)
{ std::lock_guard<std::mutex> lock(g_ftLock);
  FIPPIDeviceTables &G = g_ft[ctx->device & 63];
  if (!G.cePX)
   { if (cudaMalloc(&G.cePX, sizeof SC_CE_PX_H) != cudaSuccess || cudaMalloc(&G.cePC, sizeof SC_CE_PC_H) != cudaSuccess
         || cudaMalloc(&G.cvPX, sizeof SC_CV_PX_H) != cudaSuccess || cudaMalloc(&G.cvPC, sizeof SC_CV_PC_H) != cudaSuccess)
      { ctx->err = "FIPPI table allocation failed"; return 1; }
     cudaMemcpy(G.cePX, SC_CE_PX_H, sizeof SC_CE_PX_H, cudaMemcpyHostToDevice);
     cudaMemcpy(G.cePC, SC_CE_PC_H, sizeof SC_CE_PC_H, cudaMemcpyHostToDevice);
     cudaMemcpy(G.cvPX, SC_CV_PX_H, sizeof SC_CV_PX_H, cudaMemcpyHostToDevice);
     cudaMemcpy(G.cvPC, SC_CV_PC_H, sizeof SC_CV_PC_H, cudaMemcpyHostToDevice);
     cudaDeviceSynchronize();   // pageable H2D copies may return before the DMA lands; the kernels run on non-blocking streams
   }
  T.cePX = G.cePX; T.cePXn = SC_CE_PX_N; T.cePC = G.cePC; T.cePCn = SC_CE_PC_N;
  T.cvPX = G.cvPX; T.cvPXn = SC_CV_PX_N; T.cvPC = G.cvPC; T.cvPCn = SC_CV_PC_N;
  return 0;
}

int fippi_compute_records(scuffcuda_context *ctx, const std::vector<int> &recIds, const std::vector<double> &verts,
                          const std::vector<int> &ncvs)
{ const int n = (int)recIds.size();
  if (n == 0) return 0;
  FIPPITables T; if (fippi_tables(ctx, T)) return 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, ctx->stream);
  double *dverts = nullptr, *dout = nullptr; int *dncv = nullptr, *ditems = nullptr, *dstatus = nullptr; char *dpool = nullptr;
  auto cleanup = [&]() { cudaFree(dverts); cudaFree(dout); cudaFree(dncv); cudaFree(ditems); cudaFree(dstatus); cudaFree(dpool); cudaEventDestroy(e0); cudaEventDestroy(e1); };
  if (cudaMalloc(&dverts, (size_t)n * 18 * sizeof(double)) != cudaSuccess || cudaMalloc(&dout, (size_t)n * FIPPI_REC_DOUBLES * sizeof(double)) != cudaSuccess
      || cudaMalloc(&dncv, n * sizeof(int)) != cudaSuccess || cudaMalloc(&ditems, n * sizeof(int)) != cudaSuccess || cudaMalloc(&dstatus, n * sizeof(int)) != cudaSuccess)
   { ctx->err = "FIPPI builder allocation failed"; cleanup(); return 1; }
  cudaMemcpy(dverts, verts.data(), (size_t)n * 18 * sizeof(double), cudaMemcpyHostToDevice);
  cudaMemcpy(dncv, ncvs.data(), n * sizeof(int), cudaMemcpyHostToDevice);
  cudaDeviceSynchronize();     // the builder kernels run on their own non-blocking streams

  // integration case of every record: 0 no common vertex (3-D), 1 common edge (1-D), 2 common vertex (2-D); the
  // common-edge test is the float comparison of V2 and V2' that the integrand setup makes (fippi_integrands.cuh)
  std::vector<int> byCase[3];
  for (int i = 0; i < n; i++)
   { const double *V = &verts[18 * (size_t)i];
     int c = 0;
     if (ncvs[i] != 0)
      c = (((float)V[3] == (float)V[12]) && ((float)V[4] == (float)V[13]) && ((float)V[5] == (float)V[14])) ? 1 : 2;
     byCase[c].push_back(i); }
#ifdef SCUFF_SERIAL_FIPPI
  const fippi_launch_fn launchers[3] = { fippi_launch_case0, fippi_launch_case1, fippi_launch_case2 };
#else
  const fippi_launch_fn launchers[3] = { nullptr, nullptr, nullptr };   // built without the one-thread-per-record kernels
#endif
  // The three integration cases run concurrently (one host thread + stream each): the 1-D / 2-D cases are short grids of
  // long-running threads that would otherwise leave the GPU idle.  Within a case: pass 0 gives every record a small region
  // pool, the records that need more are redone with larger pools; the batch of each pass is sized from the free HBM
  // (a quarter each), so that even the rare 15 200-region records run thousands at a time.
  size_t freeB = 0, totalB = 0; cudaMemGetInfo(&freeB, &totalB);
  const size_t budget = std::max<size_t>((size_t)1 << 30, freeB / 4);
  std::string errs[3]; int rcs[3] = { 0, 0, 0 };
  auto run_case = [&](int c)
  { cudaSetDevice(ctx->device);
    std::vector<int> todo = byCase[c];
    if (todo.empty()) return;
    cudaStream_t st_ = nullptr; cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking);
    int *ditems_c = nullptr, *dstatus_c = nullptr; char *dpool_c = nullptr;
    auto fail = [&](const std::string &m) { errs[c] = m; rcs[c] = 1; };
    if (cudaMalloc(&ditems_c, todo.size() * sizeof(int)) != cudaSuccess || cudaMalloc(&dstatus_c, todo.size() * sizeof(int)) != cudaSuccess)
     fail("FIPPI builder allocation failed");
    const int caps[3] = { 384, 4096, 15200 };
    static int coopEnv = -1;
    if (coopEnv < 0) { const char *e = getenv("SCUFF_CUDA_FIPPI_COOP"); coopEnv = e ? atoi(e) : 1; }
    const bool coop = coopEnv == 1 || (coopEnv == 3 && c == 0) || !launchers[c];   // 1: all cases warp-cooperative, 3: only the 3-D case, 0: none
    for (int pass = 0; pass < 3 && !todo.empty() && !rcs[c]; pass++)
     { const int cap = caps[pass];
       const auto tPass = std::chrono::steady_clock::now();
       const size_t perThread = (size_t)cap * (FIPPI_REGION_BYTES + sizeof(int));
       const int batch = (int)std::min<size_t>(todo.size(), std::max<size_t>(64, budget / perThread));
       if (dpool_c) { cudaFree(dpool_c); dpool_c = nullptr; }
       if (cudaMalloc(&dpool_c, perThread * batch) != cudaSuccess) { cudaGetLastError(); fail("FIPPI builder pool allocation failed"); break; }
       std::vector<int> next, st(batch);
       for (size_t first = 0; first < todo.size() && !rcs[c]; first += batch)
        { int m = (int)std::min((size_t)batch, todo.size() - first);
          cudaMemcpyAsync(ditems_c, todo.data() + first, m * sizeof(int), cudaMemcpyHostToDevice, st_);
          __atomic_fetch_add(&ctx->launches, 1ull, __ATOMIC_RELAXED);
          // the 3-D case: one warp per record (same bits as one thread per record, far shorter latency per record)
          if (!(coop && fippi_launch_coop(st_, c, T, dverts, ditems_c, m, dpool_c, cap, dout, dstatus_c)))
           { if (!launchers[c]) { fail("FIPPI builder: the one-thread-per-record kernels were not built (make SERIAL_FIPPI=1)"); break; }
             launchers[c](st_, T, dverts, ditems_c, m, dpool_c, cap, dout, dstatus_c); }
          cudaMemcpyAsync(st.data(), dstatus_c, m * sizeof(int), cudaMemcpyDeviceToHost, st_);
          cudaError_t e = cudaStreamSynchronize(st_);
          if (e != cudaSuccess) { fail(std::string("fippi_build_kernel: ") + cudaGetErrorString(e)); break; }
          for (int i = 0; i < m; i++)
           { if (st[i] == 3) { fail("FIPPI builder: integration case mismatch"); break; }
             if (st[i]) next.push_back(todo[first + i]); }
        }
       if (getenv("SCUFF_CUDA_FIPPI_DEBUG"))
        fprintf(stderr, "[fippi] case %d pass %d (cap %d, batch %d%s): %zu records, %zu need a larger pool, %.1f ms\n", c, pass, cap, batch,
                coop ? ", warp-cooperative" : "", todo.size(), next.size(),
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tPass).count());
       todo.swap(next);
     }
    if (!rcs[c] && !todo.empty()) fail("FIPPI builder: region pool exhausted");
    cudaFree(ditems_c); cudaFree(dstatus_c); cudaFree(dpool_c); cudaStreamDestroy(st_);
  };
  cudaStreamSynchronize(ctx->stream);
  { std::thread t1(run_case, 1), t2(run_case, 2);
    run_case(0);
    t1.join(); t2.join(); }
  for (int c = 0; c < 3; c++) if (rcs[c]) { ctx->err = errs[c]; cleanup(); return 1; }
  std::vector<double> hout((size_t)n * FIPPI_REC_DOUBLES);
  cudaMemcpy(hout.data(), dout, hout.size() * sizeof(double), cudaMemcpyDeviceToHost);
  for (int i = 0; i < n; i++)
   std::copy(hout.begin() + (size_t)i * FIPPI_REC_DOUBLES, hout.begin() + (size_t)(i + 1) * FIPPI_REC_DOUBLES,
             ctx->fippi.hrec.begin() + (size_t)recIds[i] * FIPPI_REC_DOUBLES);
  cudaEventRecord(e1, ctx->stream); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1); ctx->timing[3] += ms;
  cleanup();
  return 0;
}

}  // namespace scuffcuda
