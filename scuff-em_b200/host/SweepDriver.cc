// SweepDriver: the points of a frequency / xi / kBloch sweep sharded over the GPUs of one node from C++ — one persistent host
// thread and one device context per worker, point j -> worker j mod NumWorkers, no data-path exchange between the devices
// (the C++ counterpart of scuff-em_b200/sweep.py).  Mirrors how the reference's applications loop over
// frequencies in one process (scuff-scatter.cc:373-472, scuff-cas3D/SumsIntegrals.cc): the loop body becomes the
// PointFunction.  The frequency-independent FIPPI table is built ONCE: worker w computes the records with index = w (mod
// NumWorkers) and the tables are summed in host memory before the first point.
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <condition_variable>
#include <map>
#include <string>
#include <vector>
#include <mutex>
#include <thread>
#include "host_internal.h"

namespace scuff {

struct SweepJob
{ int NumPoints = 0; const cdouble *Omegas = 0; const double *kBlochs = 0;
  SweepDriver::PointFunction F = 0; void *UserData = 0; bool Share = true; };

struct SweepPool
{ std::vector<std::thread> threads;
  std::mutex m; std::condition_variable cv, cvDone;
  long generation = 0; int running = 0; bool quit = false;
  SweepJob job;
  // a reusable barrier for the workers of one job
  int barCount = 0; long barPhase = 0;
  std::vector<double *> rec; std::vector<long> cnt;       // the workers' FIPPI tables during the shared build
  std::vector<char> planned;                              // worker has built its plans (and its share of the table)
  std::string error;
  void barrier(int n)
  { std::unique_lock<std::mutex> lk(m);
    const long phase = barPhase;
    if (++barCount == n) { barCount = 0; barPhase++; cv.notify_all(); }
    else cv.wait(lk, [&] { return barPhase != phase; });
  }
};

static std::map<SweepDriver *, SweepPool *> g_pools;
static std::mutex g_poolsLock;

static void WorkerMain(SweepDriver *D, SweepPool *P, int w)
{ SetCudaDevice(D->DeviceOf[w]);
  long seen = 0;
  for (;;)
   { SweepJob J;
     { std::unique_lock<std::mutex> lk(P->m);
       P->cv.wait(lk, [&] { return P->quit || P->generation != seen; });
       if (P->quit) break;
       seen = P->generation; J = P->job; }
     RWGGeometry *G = D->Geometries[w];
     const int NW = D->NumWorkers;
     // cold start: pair plans and this worker's share of the FIPPI table
     if (!P->planned[w])
      { const auto t0 = std::chrono::steady_clock::now();
        scuffcuda_context *ctx = GetCudaContext();
        const bool share = J.Share && NW > 1 && G->LBasis == 0;
        if (share && scuffcuda_fippi_set_partition(ctx, w, NW)) ErrExit("%s", scuffcuda_last_error(ctx));
        G->PrepareNearPlans();
        if (share)
         { scuffcuda_fippi_set_partition(ctx, 0, 1);
           if (scuffcuda_fippi_host_records(ctx, &P->rec[w], &P->cnt[w])) ErrExit("%s", scuffcuda_last_error(ctx));
           P->barrier(NW);                                          // every share is built
           for (int v = 0; v < NW; v++)
            if (P->cnt[v] != P->cnt[w]) ErrExit("SweepDriver: the workers' FIPPI tables differ in size (%li vs %li)", P->cnt[v], P->cnt[w]);
           std::vector<double> sum((size_t)P->cnt[w] * 36, 0.0);
           for (int v = 0; v < NW; v++) for (size_t q = 0; q < sum.size(); q++) sum[q] += P->rec[v][q];
           P->barrier(NW);                                          // everybody has read everybody's share
           if (!sum.empty()) memcpy(P->rec[w], sum.data(), sum.size() * sizeof(double));
           if (scuffcuda_fippi_commit(ctx)) ErrExit("%s", scuffcuda_last_error(ctx)); }
        P->planned[w] = 1;
        D->ColdSeconds[w] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
     if (!D->Matrices[w]) D->Matrices[w] = G->AllocateBEMMatrix();
     for (int j = w; j < J.NumPoints; j += NW)
      { double kB[2] = { 0.0, 0.0 }, *kBloch = 0;
        if (J.kBlochs) { kB[0] = J.kBlochs[2 * j]; kB[1] = J.kBlochs[2 * j + 1]; kBloch = kB; }
        G->AssembleBEMMatrixDevice(J.Omegas[j], kBloch, D->Matrices[w]);
        if (J.F) J.F(G, D->Matrices[w], J.Omegas[j], kBloch, j, w, J.UserData); }
     { std::unique_lock<std::mutex> lk(P->m);
       if (--P->running == 0) P->cvDone.notify_all(); }
   }
  // device twins belong to this thread's context
  delete D->Matrices[w]; D->Matrices[w] = 0;
  ReleaseCudaContext();
}

SweepDriver::SweepDriver(const char *GeoFileName, int NumDevices, const int *Devices)
{ if (NumDevices <= 0)
   { NumDevices = scuffcuda_device_count();
     if (NumDevices <= 0) ErrExit("SweepDriver: no CUDA device available (the device assembly path has no CPU fallback)"); }
  NumWorkers = NumDevices;
  for (int w = 0; w < NumWorkers; w++) DeviceOf.push_back(Devices ? Devices[w] : w);
  for (int w = 0; w < NumWorkers; w++) Geometries.push_back(new RWGGeometry(GeoFileName));    // parsed one after the other on this thread
  Matrices.assign(NumWorkers, (HMatrix *)0);
  ColdSeconds.assign(NumWorkers, 0.0);
  SweepPool *P = new SweepPool;
  P->rec.assign(NumWorkers, (double *)0); P->cnt.assign(NumWorkers, 0); P->planned.assign(NumWorkers, 0);
  { std::lock_guard<std::mutex> lk(g_poolsLock); g_pools[this] = P; }
  for (int w = 0; w < NumWorkers; w++) P->threads.emplace_back(WorkerMain, this, P, w);
}

void SweepDriver::Run(int NumPoints, const cdouble *Omegas, const double *kBlochs, PointFunction F, void *UserData, bool ShareFIPPITable)
{ SweepPool *P;
  { std::lock_guard<std::mutex> lk(g_poolsLock); P = g_pools[this]; }
  std::unique_lock<std::mutex> lk(P->m);
  P->job.NumPoints = NumPoints; P->job.Omegas = Omegas; P->job.kBlochs = kBlochs; P->job.F = F; P->job.UserData = UserData; P->job.Share = ShareFIPPITable;
  P->running = NumWorkers; P->generation++;
  P->cv.notify_all();
  P->cvDone.wait(lk, [&] { return P->running == 0; });
}

SweepDriver::~SweepDriver()
{ SweepPool *P;
  { std::lock_guard<std::mutex> lk(g_poolsLock); P = g_pools[this]; g_pools.erase(this); }
  { std::unique_lock<std::mutex> lk(P->m); P->quit = true; P->cv.notify_all(); }
  for (auto &t : P->threads) t.join();
  delete P;
  for (RWGGeometry *G : Geometries) delete G;
}

}  // namespace scuff
