// Device builder of the FIPPI table (QIFIPPIData records, libscuffInternals.h:282-293): one thread per
// record runs the reference's adaptive cubature decision-for-decision (hcubature.cuh) on the integrands of
// fippi_integrands.cuh.  Compiled with -fmad=false so that error estimates and threshold comparisons
// are rounded like the reference's SSE2 arithmetic.  The table is built once per geometry and stays
// resident in HBM across the frequency sweep (ComputeQIFIPPIData QIFIPPICubature.cc:465-487).
#include <vector>
#include <algorithm>
#include "kernels.h"
#include "fippi_build.cuh"
#include "generated_tables.h"

namespace scuffcuda {

static_assert(sizeof(SCPXEntry) == sizeof(PXEntry) && sizeof(SCPCEntry) == sizeof(PCEntry), "generated table layout");

struct FIPPIDeviceTables { PXEntry *cePX = nullptr, *cvPX = nullptr; PCEntry *cePC = nullptr, *cvPC = nullptr; };
static FIPPIDeviceTables g_ft[64];

static int fippi_tables(scuffcuda_context *ctx, FIPPITables &T)
{ FIPPIDeviceTables &G = g_ft[ctx->device & 63];
  if (!G.cePX)
   { if (cudaMalloc(&G.cePX, sizeof SC_CE_PX_H) != cudaSuccess || cudaMalloc(&G.cePC, sizeof SC_CE_PC_H) != cudaSuccess
         || cudaMalloc(&G.cvPX, sizeof SC_CV_PX_H) != cudaSuccess || cudaMalloc(&G.cvPC, sizeof SC_CV_PC_H) != cudaSuccess)
      { ctx->err = "FIPPI table allocation failed"; return 1; }
     cudaMemcpy(G.cePX, SC_CE_PX_H, sizeof SC_CE_PX_H, cudaMemcpyHostToDevice);
     cudaMemcpy(G.cePC, SC_CE_PC_H, sizeof SC_CE_PC_H, cudaMemcpyHostToDevice);
     cudaMemcpy(G.cvPX, SC_CV_PX_H, sizeof SC_CV_PX_H, cudaMemcpyHostToDevice);
     cudaMemcpy(G.cvPC, SC_CV_PC_H, sizeof SC_CV_PC_H, cudaMemcpyHostToDevice);
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

  // integration case of every record: 0 no common vertex (3-D), 1 common edge (1-D), 2 common vertex (2-D); the
  // common-edge test is the float comparison of V2 and V2' that the integrand setup makes (fippi_integrands.cuh)
  std::vector<int> byCase[3];
  for (int i = 0; i < n; i++)
   { const double *V = &verts[18 * (size_t)i];
     int c = 0;
     if (ncvs[i] != 0)
      c = (((float)V[3] == (float)V[12]) && ((float)V[4] == (float)V[13]) && ((float)V[5] == (float)V[14])) ? 1 : 2;
     byCase[c].push_back(i); }
  const fippi_launch_fn launchers[3] = { fippi_launch_case0, fippi_launch_case1, fippi_launch_case2 };
  std::vector<int> todo;
  for (int c = 0; c < 3; c++)
   { todo = byCase[c];
  // pass 1: small region pools, many records in flight; pass 2: the few records that need more regions
  const int caps[3] = { 384, 4096, 15200 }; const int batches[3] = { 24576, 2048, 512 };
  for (int pass = 0; pass < 3 && !todo.empty(); pass++)
   { const int cap = caps[pass], batch = std::min(batches[pass], (int)todo.size());
     size_t perThread = (size_t)cap * (FIPPI_REGION_BYTES + sizeof(int));
     if (dpool) { cudaFree(dpool); dpool = nullptr; }
     if (cudaMalloc(&dpool, perThread * batch) != cudaSuccess) { ctx->err = "FIPPI builder pool allocation failed"; cleanup(); return 1; }
     std::vector<int> next, st(batch);
     for (size_t first = 0; first < todo.size(); first += batch)
      { int m = (int)std::min((size_t)batch, todo.size() - first);
        cudaMemcpyAsync(ditems, todo.data() + first, m * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
        ctx->launches++;
        launchers[c](ctx->stream, T, dverts, ditems, m, dpool, cap, dout, dstatus);
        cudaMemcpyAsync(st.data(), dstatus, m * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { ctx->err = std::string("fippi_build_kernel: ") + cudaGetErrorString(e); cleanup(); return 1; }
        for (int i = 0; i < m; i++)
         { if (st[i] == 3) { ctx->err = "FIPPI builder: integration case mismatch"; cleanup(); return 1; }
           if (st[i]) next.push_back(todo[first + i]); }
      }
     todo.swap(next);
   }
     if (!todo.empty()) break;
   }
  if (!todo.empty()) { ctx->err = "FIPPI builder: region pool exhausted"; cleanup(); return 1; }
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
