// RWGGeometry::AssembleBEMMatrix / AssembleBEMMatrixBlock with the reference's signatures
// (libs/libscuff/libscuff.h:363-365,500-508) marshalled to the device through the C ABI.
// Control flow restated from AssembleBEMMatrix.cc:349-391 (compact block), :684-765 (full matrix,
// Mate reuse, symmetric fill) and GetSurfaceSurfaceInteractions SurfaceSurfaceInteractions.cc:
// 426-492 (common regions, Eps/Mu/Sign selection, OmitRegion, early returns).
#include <cstdlib>
#include <cmath>
#include <vector>
#include <atomic>
#include <thread>
#include <algorithm>
#include <cstring>
#include <mutex>
#include <string>
#include <emmintrin.h>
#include "host_internal.h"

namespace scuff {

// CountCommonRegions, SurfaceSurfaceInteractions.cc:62-90
static int CountCommonRegions(RWGSurface *Sa, RWGSurface *Sb, int CommonRegionIndices[2], double Signs[2])
{ int n = 0;
  if (Sa->RegionIndices[0] == Sb->RegionIndices[0]) { CommonRegionIndices[n] = Sa->RegionIndices[0]; Signs[n] = +1.0; n++; }
  else if (Sa->RegionIndices[0] == Sb->RegionIndices[1]) { CommonRegionIndices[n] = Sa->RegionIndices[0]; Signs[n] = -1.0; n++; }
  if (Sa->RegionIndices[1] == Sb->RegionIndices[0]) { CommonRegionIndices[n] = Sa->RegionIndices[1]; Signs[n] = -1.0; n++; }
  else if (!Sa->IsPEC && !Sb->IsPEC && Sa->RegionIndices[1] == Sb->RegionIndices[1]) { CommonRegionIndices[n] = Sa->RegionIndices[1]; Signs[n] = +1.0; n++; }
  return n;
}

HMatrix *RWGGeometry::AllocateBEMMatrix(bool PureImagFreq, bool Packed)
{ int Storage = Packed ? LHM_SYMMETRIC : LHM_NORMAL;
  int DataType = (!LBasis && PureImagFreq) ? LHM_REAL : LHM_COMPLEX;
  return new HMatrix(TotalBFs, TotalBFs, DataType, Storage);
}

// KBIMBCache, the 'kBloch-independent matrix-block cache' (AssembleBEMMatrix.cc:250-334): the 2/3/5/9 inner-cell blocks of
// one surface pair stay resident in HBM across a kBloch sweep at fixed Omega; only the Bloch-phase stamping and the
// outer-cell step are redone per kBloch.
struct KBIMBCache
{ cdouble Omega; int NumMatrices; bool NeedZDerivative; int nsa, nsb;
  HMatrix *B[9];
};

void *RWGGeometry::CreateABMBAccelerator(int nsa, int nsb, bool PureImagFreq, bool NeedZDerivative)
{ if (LBasis == 0) return 0;
  if (nsa < 0 || nsa >= NumSurfaces || nsb < 0 || nsb >= NumSurfaces) ErrExit("CreateABMBAccelerator: invalid surface indices (%i,%i)", nsa, nsb);
  if (NeedZDerivative) ErrExit("CreateABMBAccelerator: z-derivative blocks (GradM[2]) are not yet available on the device path");
  const bool OneDLattice = (LDim == 1), SameSurface = (nsa == nsb);
  KBIMBCache *Cache = new KBIMBCache;
  Cache->Omega = -1.0;            // initially dirty
  Cache->NumMatrices = OneDLattice ? (SameSurface ? 2 : 3) : (SameSurface ? 5 : 9);
  Cache->NeedZDerivative = NeedZDerivative; Cache->nsa = nsa; Cache->nsb = nsb;
  for (int nm = 0; nm < 9; nm++) Cache->B[nm] = 0;
  for (int nm = 0; nm < Cache->NumMatrices; nm++)
   { Cache->B[nm] = new HMatrix(Surfaces[nsa]->NumBFs, Surfaces[nsb]->NumBFs, PureImagFreq ? LHM_REAL : LHM_COMPLEX);
     Cache->B[nm]->EnsureDevice(); }
  return (void *)Cache;
}
void RWGGeometry::DestroyABMBAccelerator(void *pCache)
{ if (pCache == 0) return;
  KBIMBCache *Cache = (KBIMBCache *)pCache;
  for (int nm = 0; nm < Cache->NumMatrices; nm++) delete Cache->B[nm];
  delete Cache;
}
static bool EqualFloat(cdouble a, cdouble b)    // libs/libhrutil/Vector.cc:273-276
{ return float(real(a)) == float(real(b)) && float(imag(a)) == float(imag(b)); }

// device-side GetSurfaceSurfaceInteractions for one block of M (device twin), no host copy-back
static void AssembleBlockOnDevice(RWGGeometry *G, int nsa, int nsb, cdouble Omega, HMatrix *M,
                                  int RowOffset, int ColOffset, bool Symmetric, bool Accumulate,
                                  const double *Displacement, bool OmitRegion1, bool OmitRegion2,
                                  int ColBegin = 0, int ColEnd = 0, int GBA1 = 0, int GBA2 = 0)
{ scuffcuda_context *ctx = GetCudaContext();
  RWGSurface *Sa = G->Surfaces[nsa], *Sb = G->Surfaces[nsb];
  G->UpdateCachedEpsMuValues(Omega);
  G->SyncDeviceGeometry();
  M->PrepareDevice();            // host-side edits of other parts of M reach the device before the block is written

  scuffcuda_ssi_args A; memset(&A, 0, sizeof A);
  A.omega[0] = real(Omega); A.omega[1] = imag(Omega);
  A.symmetric = Symmetric; A.accumulate = Accumulate;
  A.row_offset = RowOffset; A.col_offset = ColOffset;
  if (Displacement) { A.use_displacement = 1; for (int i = 0; i < 3; i++) A.displacement[i] = Displacement[i]; }
  A.force_taylor_duffy = RWGGeometry::DisableCache ? 1 : 0;
  A.col_begin = ColBegin; A.col_end = ColEnd;
  A.gba_a = GBA1; A.gba_b = GBA2;
  const char *fo = getenv("SCUFF_CUDA_FAR_ONLY"); A.far_only = (fo && fo[0] == '1');

  double Signs[2]; int CR[2];
  int NumCommonRegions = CountCommonRegions(Sa, Sb, CR, Signs);
  cdouble EpsA = 0.0, MuA = 0.0, EpsB = 0.0, MuB = 0.0; double SignA = 0.0, SignB = 0.0;
  if (NumCommonRegions >= 1) { EpsA = G->EpsTF[CR[0]]; MuA = G->MuTF[CR[0]]; SignA = Signs[0]; }
  if (NumCommonRegions == 2) { EpsB = G->EpsTF[CR[1]]; MuB = G->MuTF[CR[1]]; SignB = Signs[1]; }
  if (OmitRegion1) { EpsA = MuA = 0.0; SignA = 0.0; }
  if (OmitRegion2) { EpsB = MuB = 0.0; SignB = 0.0; }
  // note: with zero common regions (or both omitted) the reference still zeroes the block
  // unless Accumulate is set; assemble_block does the same with eps_a = eps_b = 0
  A.eps_a[0] = real(EpsA); A.eps_a[1] = imag(EpsA); A.mu_a[0] = real(MuA); A.mu_a[1] = imag(MuA); A.sign_a = SignA;
  A.eps_b[0] = real(EpsB); A.eps_b[1] = imag(EpsB); A.mu_b[0] = real(MuB); A.mu_b[1] = imag(MuB); A.sign_b = SignB;

  if (scuffcuda_assemble_block(ctx, nsa, nsb, &A, M->DeviceHandle)) ErrExit("%s", scuffcuda_last_error(ctx));
  unsigned long long c[5]; scuffcuda_get_ppi_counts(ctx, c);
  for (int i = 0; i < 5; i++) G->LastPPICounts[i] += c[i];
  M->HostStale = true;
}

// GradM[3] / dMdT[NumTorqueAxes] of a block between two different surfaces (scuffcuda_assemble_block_gradients)
static void AssembleDerivativeBlocksOnDevice(RWGGeometry *G, int nsa, int nsb, cdouble Omega, HMatrix **GradM, int RowOffset, int ColOffset,
                                             int NumTorqueAxes, HMatrix **dMdT, double *GammaMatrix)
{ scuffcuda_context *ctx = GetCudaContext();
  RWGSurface *Sa = G->Surfaces[nsa], *Sb = G->Surfaces[nsb];
  G->UpdateCachedEpsMuValues(Omega);
  G->SyncDeviceGeometry();
  int gh[3] = { 0, 0, 0 }, th[3] = { 0, 0, 0 };
  for (int m = 0; m < 3; m++)
   { if (GradM && GradM[m]) { GradM[m]->PrepareDevice(); gh[m] = GradM[m]->DeviceHandle; }
     if (dMdT && m < NumTorqueAxes && dMdT[m]) { dMdT[m]->PrepareDevice(); th[m] = dMdT[m]->DeviceHandle; } }
  scuffcuda_ssi_args A; memset(&A, 0, sizeof A);
  A.omega[0] = real(Omega); A.omega[1] = imag(Omega);
  A.row_offset = RowOffset; A.col_offset = ColOffset;
  A.force_taylor_duffy = RWGGeometry::DisableCache ? 1 : 0;
  double Signs[2]; int CR[2];
  int NumCommonRegions = CountCommonRegions(Sa, Sb, CR, Signs);
  cdouble EpsA = 0.0, MuA = 0.0, EpsB = 0.0, MuB = 0.0; double SignA = 0.0, SignB = 0.0;
  if (NumCommonRegions >= 1) { EpsA = G->EpsTF[CR[0]]; MuA = G->MuTF[CR[0]]; SignA = Signs[0]; }
  if (NumCommonRegions == 2) { EpsB = G->EpsTF[CR[1]]; MuB = G->MuTF[CR[1]]; SignB = Signs[1]; }
  A.eps_a[0] = real(EpsA); A.eps_a[1] = imag(EpsA); A.mu_a[0] = real(MuA); A.mu_a[1] = imag(MuA); A.sign_a = SignA;
  A.eps_b[0] = real(EpsB); A.eps_b[1] = imag(EpsB); A.mu_b[0] = real(MuB); A.mu_b[1] = imag(MuB); A.sign_b = SignB;
  if (scuffcuda_assemble_block_gradients(ctx, nsa, nsb, &A, gh, dMdT ? NumTorqueAxes : 0, GammaMatrix, th)) ErrExit("%s", scuffcuda_last_error(ctx));
  for (int m = 0; m < 3; m++)
   { if (GradM && GradM[m]) { GradM[m]->HostStale = true; GradM[m]->SyncToHost(); }
     if (dMdT && m < NumTorqueAxes && dMdT[m]) { dMdT[m]->HostStale = true; dMdT[m]->SyncToHost(); } }
}

// AddSurfaceZetaContributionToBEMMatrix (SurfaceSurfaceInteractions.cc:339-421): a PEC surface with SURFACE_IMPEDANCE Zeta adds
// -Zeta(w, X) * \int f_a . f_b to the entries of basis functions that share a panel; X = the edge centroid on the diagonal, else
// the centroid of the (first) shared panel.  Only for the undisplaced diagonal block of a compact geometry.  Entries outside
// the owned columns [ColBegin,ColEnd) (ColEnd > 0: column-block mode, upper triangle only) are left to their owner.
static void AddSurfaceZetaContribution(RWGGeometry *G, int ns, cdouble Omega, HMatrix *M, int Offset, int ColBegin, int ColEnd)
{ RWGSurface *S = G->Surfaces[ns];
  if (!S->SurfaceZeta || !S->IsPEC) return;
  const bool Partial = ColEnd > 0;
  // panel -> edges with a basis function on it
  std::vector<std::vector<int> > PanelEdges(S->NumPanels);
  for (int ne = 0; ne < S->NumEdges; ne++)
   { PanelEdges[S->Edges[ne]->iPPanel].push_back(ne);
     if (S->Edges[ne]->iMPanel != -1) PanelEdges[S->Edges[ne]->iMPanel].push_back(ne); }
  std::vector<int> Rows, Cols; std::vector<double> Vals;
  CExpr *Z = (CExpr *)S->SurfaceZeta;
  for (int neA = 0; neA < S->NumEdges; neA++)
   { RWGEdge *EA = S->Edges[neA];
     // candidates: the edges of the (<= 2) panels of neA, each once, neB >= neA
     int Cand[8], nCand = 0;
     for (int side = 0; side < 2; side++)
      { int np = side == 0 ? EA->iPPanel : EA->iMPanel; if (np == -1) continue;
        for (int neB : PanelEdges[np])
         { if (neB < neA) continue;
           bool seen = false; for (int q = 0; q < nCand; q++) if (Cand[q] == neB) seen = true;
           if (!seen && nCand < 8) Cand[nCand++] = neB; } }
     for (int q = 0; q < nCand; q++)
      { const int neB = Cand[q]; RWGEdge *EB = S->Edges[neB];
        const double Overlap = S->GetOverlap(neA, neB);
        if (Overlap == 0.0) continue;
        const double *X = EA->Centroid;
        if (neA == neB) X = EA->Centroid;
        else if (EA->iPPanel == EB->iPPanel || EA->iPPanel == EB->iMPanel) X = S->Panels[EA->iPPanel]->Centroid;
        else if (EA->iMPanel != -1 && (EA->iMPanel == EB->iPPanel || EA->iMPanel == EB->iMPanel)) X = S->Panels[EA->iMPanel]->Centroid;
        std::lock_guard<std::recursive_mutex> exprLock(CExprMutex());
        CExprSetVar(Z, "x", X[0]); CExprSetVar(Z, "y", X[1]); CExprSetVar(Z, "z", X[2]);
        const cdouble V = -1.0 * CExprEval(Z, Omega * MatProp::FreqUnit) * Overlap;
        const int r = Offset + neA, c = Offset + neB;
        if (!Partial || (c >= ColBegin && c < ColEnd)) { Rows.push_back(r); Cols.push_back(c); Vals.push_back(real(V)); Vals.push_back(imag(V)); }
        if (!Partial && neA != neB) { Rows.push_back(c); Cols.push_back(r); Vals.push_back(real(V)); Vals.push_back(imag(V)); }
      }
   }
  scuffcuda_context *ctx = GetCudaContext();
  M->EnsureDevice();
  if (!Rows.empty() && scuffcuda_matrix_add_entries(ctx, M->DeviceHandle, (int)Rows.size(), Rows.data(), Cols.data(), Vals.data()))
   ErrExit("%s", scuffcuda_last_error(ctx));
  M->HostStale = true;
}

void RWGGeometry::AssembleBEMMatrixBlock(int nsa, int nsb, cdouble Omega, double *kBloch,
                                         HMatrix *M, HMatrix **GradM, int RowOffset, int ColOffset,
                                         void *Accelerator, bool TransposeAccelerator,
                                         int NumTorqueAxes, HMatrix **dMdT, double *GammaMatrix)
{ (void)GammaMatrix;
  if (TransposeAccelerator) ErrExit("%s:%i: TransposeAccelerator not implemented", __FILE__, __LINE__);
  if (nsa < 0 || nsa >= NumSurfaces || nsb < 0 || nsb >= NumSurfaces) ErrExit("AssembleBEMMatrixBlock: invalid surface indices (%i,%i)", nsa, nsb);
  const bool WantDerivatives = GradM || (NumTorqueAxes > 0 && dMdT);
  if (WantDerivatives && (LBasis != 0 || nsa == nsb))
   ErrExit("AssembleBEMMatrixBlock: GradM / dMdT derivative blocks are available on the device path for blocks between two different "
           "surfaces of a compact geometry only");
  if (WantDerivatives && NumTorqueAxes > 0 && (NumTorqueAxes > 3 || GammaMatrix == 0))
   ErrExit("AssembleBEMMatrixBlock: NumTorqueAxes must be 1..3 with a GammaMatrix");
  if (getenv("SCUFF_TBLOCK_PATH")) Warn("SCUFF_TBLOCK_PATH is ignored by the device assembly path");
  // the caller may have written into M on the host (e.g. other blocks): keep both sides coherent
  M->PrepareDevice();
  for (int i = 0; i < 5; i++) LastPPICounts[i] = 0;
  if (LBasis == 0)
   { AssembleBlockOnDevice(this, nsa, nsb, Omega, M, RowOffset, ColOffset, nsa == nsb, false, 0, false, false);
     if (WantDerivatives) AssembleDerivativeBlocksOnDevice(this, nsa, nsb, Omega, GradM, RowOffset, ColOffset, NumTorqueAxes, dMdT, GammaMatrix);
     if (nsa == nsb)
      { if (RowOffset != ColOffset && Surfaces[nsa]->SurfaceZeta) ErrExit("%s:%i: internal error", __FILE__, __LINE__);
        AddSurfaceZetaContribution(this, nsa, Omega, M, RowOffset, 0, 0); }
     M->SyncToHost();
     return; }
  AssembleBEMMatrixBlockPBC(nsa, nsb, Omega, kBloch, M, RowOffset, ColOffset, Accelerator);
  M->SyncToHost();
}

// The reference hands back a HOST matrix; for large N the copy-back over PCIe (16 N^2 bytes) takes several times longer
// than the assembly itself.  Overlap the two: assemble the upper triangle in K column blocks (column-block mode of
// scuffcuda_assemble_block), mirror each block's diagonal square as soon as it is complete and start the download of the
// finished trapezoid rows [0,c1) x columns [c0,c1) on the copy stream while the next block is assembled; after the last
// block the whole matrix is mirrored once and the remaining lower parts follow.  Same entries as the one-pass assembly
// (summation order inside the atomics aside).  SCUFF_CUDA_PIPELINE_BLOCKS=1 disables, =K forces K blocks.
static int PipelineBlocks(int N)
{ const char *e = getenv("SCUFF_CUDA_PIPELINE_BLOCKS");
  if (e && e[0]) { int k = atoi(e); return k < 1 ? 1 : (k > 64 ? 64 : k); }
  // every column-block call has a latency floor of ~6 ms (the adaptive Taylor-Duffy kernels of its diagonal pairs): the
  // overlap pays once the copy-back itself takes several times that (16 N^2 bytes at ~50 GB/s: N >~ 16 000)
  return N >= 16000 ? 4 : 1;
}

static void AssembleBEMMatrixPipelined(RWGGeometry *G, cdouble Omega, HMatrix *M, int K)
{ scuffcuda_context *ctx = GetCudaContext();
  const int N = G->TotalBFs;
  M->DeviceStale = false;        // every entry is rewritten
  M->EnsureDevice();
  void *host = M->RealComplex == LHM_REAL ? (void *)M->DM : (void *)M->ZM;
  std::vector<int> c(K + 1);
  for (int j = 0; j <= K; j++) c[j] = (int)(N * sqrt((double)j / K) + 0.5);    // equal upper-triangle area per block
  c[0] = 0; c[K] = N;
  unsigned long long Counts[5] = { 0, 0, 0, 0, 0 };
  for (int j = 0; j < K; j++)
   { if (c[j + 1] <= c[j]) continue;
     G->AssembleBEMMatrixDevice(Omega, 0, M, c[j], c[j + 1]);
     for (int i = 0; i < 5; i++) Counts[i] += G->LastPPICounts[i];
     if (scuffcuda_matrix_symmetrize(ctx, M->DeviceHandle, c[j], c[j + 1] - c[j])
         || scuffcuda_matrix_download_region_async(ctx, M->DeviceHandle, host, M->NR, 0, c[j + 1], c[j], c[j + 1] - c[j]))
      ErrExit("%s", scuffcuda_last_error(ctx)); }
  // the mirror pass writes below the diagonal squares only what they already hold, but keep it ordered after the copies
  if (scuffcuda_matrix_download_wait(ctx, 1) || scuffcuda_matrix_symmetrize(ctx, M->DeviceHandle, 0, N))
   ErrExit("%s", scuffcuda_last_error(ctx));
  for (int j = 0; j + 1 < K; j++)
   if (c[j + 1] > c[j]
       && scuffcuda_matrix_download_region_async(ctx, M->DeviceHandle, host, M->NR, c[j + 1], N - c[j + 1], c[j], c[j + 1] - c[j]))
    ErrExit("%s", scuffcuda_last_error(ctx));
  if (scuffcuda_matrix_download_wait(ctx, 0)) ErrExit("%s", scuffcuda_last_error(ctx));
  for (int i = 0; i < 5; i++) G->LastPPICounts[i] = Counts[i];
  M->HostStale = false;
}

// Copy-back of a symmetric matrix: only the upper triangle crosses PCIe (half the bytes); the lower triangle is written by
// host threads while the later column chunks are still in flight.  The matrix is assembled in ONE pass (the two-pass far path
// needs whole blocks), the K chunk copies [0,c1) x [c0,c1) are enqueued at once, and the host mirrors chunk j as soon as
// its ticket completes.  SCUFF_CUDA_PIPELINE_MODE=columns selects the older scheme above (assembly in column blocks overlapped
// with full-matrix downloads).
static inline void StreamStore(cdouble *dst, const cdouble &v) { _mm_stream_pd(reinterpret_cast<double *>(dst), _mm_loadu_pd(reinterpret_cast<const double *>(&v))); }
static inline void StreamStore(double *dst, const double &v) { long long b; memcpy(&b, &v, 8); _mm_stream_si64(reinterpret_cast<long long *>(dst), b); }

template <typename T>
static void MirrorChunkOnHost(T *A, long ld, int c0, int c1, int nThreads)
{ // lower(r, c) = upper(c, r) for columns c in [0, c0), rows r in [c0, c1): read the chunk's part above its diagonal square
  // (columns [c0,c1), rows [0,c0)) in 64 x 64 tiles, write them transposed with streaming stores (no read-for-ownership)
  const int TS = 64;
  const long tilesR = (c0 + TS - 1) / TS, tilesC = (c1 - c0 + TS - 1) / TS, nTiles = tilesR * tilesC;
  if (nTiles <= 0) return;
  std::atomic<long> next(0);
  auto work = [&]()
   { T buf[TS][TS + 1];
     for (;;)
      { const long t = next.fetch_add(1);
        if (t >= nTiles) break;
        const int r0 = (int)(t % tilesR) * TS, q0 = c0 + (int)(t / tilesR) * TS;
        const int nr = std::min(TS, c0 - r0), nc = std::min(TS, c1 - q0);
        for (int q = 0; q < nc; q++) { const T *src = A + (size_t)r0 + (size_t)(q0 + q) * ld; for (int r = 0; r < nr; r++) buf[q][r] = src[r]; }
        for (int r = 0; r < nr; r++) { T *dst = A + (size_t)q0 + (size_t)(r0 + r) * ld; for (int q = 0; q < nc; q++) StreamStore(dst + q, buf[q][r]); } }
     _mm_sfence(); };
  std::vector<std::thread> pool;
  for (int i = 1; i < nThreads; i++) pool.emplace_back(work);
  work();
  for (auto &th : pool) th.join();
}

static void AssembleBEMMatrixUpperCopyBack(RWGGeometry *G, cdouble Omega, HMatrix *M, int K)
{ scuffcuda_context *ctx = GetCudaContext();
  const int N = G->TotalBFs;
  G->AssembleBEMMatrixDevice(Omega, 0, M);                 // whole matrix (both triangles) resident in HBM
  void *host = M->RealComplex == LHM_REAL ? (void *)M->DM : (void *)M->ZM;
  std::vector<int> c(K + 1), ticket(K, 0);
  for (int j = 0; j <= K; j++) c[j] = (int)(N * sqrt((double)j / K) + 0.5);    // equal upper-triangle area per chunk
  c[0] = 0; c[K] = N;
  for (int j = 0; j < K; j++)
   { if (c[j + 1] > c[j] && scuffcuda_matrix_download_region_async(ctx, M->DeviceHandle, host, M->NR, 0, c[j + 1], c[j], c[j + 1] - c[j]))
      ErrExit("%s", scuffcuda_last_error(ctx));
     if (scuffcuda_matrix_download_mark(ctx, &ticket[j])) ErrExit("%s", scuffcuda_last_error(ctx)); }
  int nThreads = (int)std::thread::hardware_concurrency();
  if (const char *e = getenv("SCUFF_CUDA_MIRROR_THREADS")) nThreads = atoi(e);
  nThreads = nThreads < 1 ? 1 : (nThreads > 64 ? 64 : nThreads);
  // The lower triangle is produced from both ends: host threads mirror the chunks in ascending order (as their upper parts land),
  // the copy engine — idle once the upper triangle is through — downloads the lower parts in descending order, until they meet.
  std::mutex claimLock; int lo = 0, hi = K - 1;            // chunks [lo, hi] are not yet claimed
  auto claim = [&](bool fromTop) -> int
   { std::lock_guard<std::mutex> g(claimLock);
     if (lo > hi) return -1;
     return fromTop ? hi-- : lo++; };
  std::string mirrorErr;
  std::thread mirror([&]()
   { for (;;)
      { const int j = claim(false);
        if (j < 0) break;
        if (scuffcuda_matrix_download_wait_ticket(ctx, ticket[j])) { mirrorErr = scuffcuda_last_error(ctx); break; }
        ticket[j] = 0;
        if (c[j + 1] <= c[j]) continue;
        if (M->RealComplex == LHM_REAL) MirrorChunkOnHost<double>(M->DM, M->NR, c[j], c[j + 1], nThreads > 1 ? nThreads - 1 : 1);
        else MirrorChunkOnHost<cdouble>(M->ZM, M->NR, c[j], c[j + 1], nThreads > 1 ? nThreads - 1 : 1); } });
  const bool dmaHelps = !(getenv("SCUFF_CUDA_MIRROR_DMA") && getenv("SCUFF_CUDA_MIRROR_DMA")[0] == '0');
  std::string dmaErr;
  while (dmaHelps)
   { // one lower part at a time, so that the claim reflects what the copy engine has really reached
     const int j = claim(true);
     if (j < 0) break;
     int t = 0;
     if ((c[j + 1] > c[j] && c[j] > 0 && scuffcuda_matrix_download_region_async(ctx, M->DeviceHandle, host, M->NR, c[j], c[j + 1] - c[j], 0, c[j]))
         || scuffcuda_matrix_download_mark(ctx, &t) || scuffcuda_matrix_download_wait_ticket(ctx, t))
      { dmaErr = scuffcuda_last_error(ctx); break; } }
  mirror.join();
  for (int j = 0; j < K; j++) if (ticket[j]) scuffcuda_matrix_download_wait_ticket(ctx, ticket[j]);   // upper parts of chunks the copy engine took
  if (!mirrorErr.empty()) ErrExit("%s", mirrorErr.c_str());
  if (!dmaErr.empty()) ErrExit("%s", dmaErr.c_str());
  if (scuffcuda_matrix_download_wait(ctx, 0)) ErrExit("%s", scuffcuda_last_error(ctx));
  M->HostStale = false;
}

HMatrix *RWGGeometry::AssembleBEMMatrix(cdouble Omega, double *kBloch, HMatrix *M)
{ if (LBasis == 0 && kBloch != 0 && (kBloch[0] != 0.0 || kBloch[1] != 0.0))
   ErrExit("%s:%i: Bloch wavevector is undefined for compact geometries", __FILE__, __LINE__);
  if (LBasis != 0 && kBloch == 0)
   ErrExit("%s:%i: Bloch wavevector must be specified for PBC geometries", __FILE__, __LINE__);
  if (M == NULL) M = AllocateBEMMatrix();
  else if (M->NR != TotalBFs || M->NC != TotalBFs)
   { Warn("wrong-size matrix passed to AssembleBEMMatrix; reallocating...");
     M = AllocateBEMMatrix(); }
  // SCUFF_CUDA_LAZY_HOST=1: return with the matrix in HBM only (M->HostStale).  LUFactorize / LUSolve / InsertBlock / the RHS,
  // PFT and field routines work on the device copy; the HMatrix accessors (GetEntry, SetEntry, ...) fetch it on first use.  For
  // callers that never read M->ZM / M->DM directly — the assemble -> LUFactorize -> LUSolve flow of scuff-scatter
  // (scuff-scatter.cc:402-472) — nothing else changes and the 16 N^2 bytes never cross PCIe.
  { const char *lazy = getenv("SCUFF_CUDA_LAZY_HOST");
    if (lazy && lazy[0] == '1') { AssembleBEMMatrixDevice(Omega, kBloch, M); return M; } }
  int K = PipelineBlocks(TotalBFs);
  // "upper" halves the PCIe bytes but spends host memory bandwidth on the mirror (a read and a write per byte against one write
  // for a DMA copy): right for one process per host, wrong when several ranks share the host's memory system (measured on a
  // 2-GPU box: 460 ms per step and rank against 213 ms with DMA only) — those default to the all-DMA "columns" scheme
  const char *mode = getenv("SCUFF_CUDA_PIPELINE_MODE");
  if (!mode)
   { const char *lw = getenv("LOCAL_WORLD_SIZE"), *ww = getenv("WORLD_SIZE");
     mode = ((lw && atoi(lw) > 1) || (!lw && ww && atoi(ww) > 1)) ? "columns" : "upper"; }
  if (LBasis == 0 && K > 1 && M->StorageType == LHM_NORMAL && !strcmp(mode, "columns")) AssembleBEMMatrixPipelined(this, Omega, M, K);
  else if (LBasis == 0 && K > 1 && M->StorageType == LHM_NORMAL) AssembleBEMMatrixUpperCopyBack(this, Omega, M, 2 * K);
  else
   { AssembleBEMMatrixDevice(Omega, kBloch, M);
     M->SyncToHost(); }
  return M;
}

// same as AssembleBEMMatrix but leaves the result in HBM only (M->HostStale == true): the form
// the on-device LU consumer and the sweep drivers use
void RWGGeometry::AssembleBEMMatrixDevice(cdouble Omega, double *kBloch, HMatrix *M, int ColBegin, int ColEnd)
{ scuffcuda_context *ctx = GetCudaContext();
  const bool Partial = ColEnd > 0;
  // a compact-geometry assembly rewrites every entry (of the owned columns): host-side edits need not be uploaded first
  if (LBasis == 0 && !Partial) M->DeviceStale = false;
  if (M->DeviceStale) M->SyncToDevice(); else M->EnsureDevice();
  for (int i = 0; i < 5; i++) LastPPICounts[i] = 0;
  bool MatrixIsSymmetric = (!kBloch || (kBloch[0] == 0.0 && kBloch[1] == 0.0));
  int nspStart = MatrixIsSymmetric ? 1 : 0;
  for (int ns = 0; ns < NumSurfaces; ns++)
   for (int nsp = nspStart * ns; nsp < NumSurfaces; nsp++)
    { int nsm;
      if (Partial && (BFIndexOffset[nsp] >= ColEnd || BFIndexOffset[nsp] + Surfaces[nsp]->NumBFs <= ColBegin)) continue;
      // (column-block mode assembles mates directly: the mate's diagonal block may belong to another rank)
      if (!Partial && ns == nsp && (nsm = Mate[ns]) != -1)
       { int ThisOffset = BFIndexOffset[ns], MateOffset = BFIndexOffset[nsm], Dim = Surfaces[ns]->NumBFs;
         SyncDeviceGeometry();
         if (scuffcuda_matrix_copy_block(ctx, M->DeviceHandle, ThisOffset, ThisOffset, Dim, Dim, MateOffset, MateOffset))
          ErrExit("%s", scuffcuda_last_error(ctx));
         M->HostStale = true; }
      else if (LBasis == 0)
       { // column-block mode: a mated diagonal block takes its mate's values (as the Mate copy does; the two differ at the
         // 1e-9 level through the float-precision FIPPI keys), computed here from the mate's panels into this block
         int nsA = ns, nsB = nsp;
         if (Partial && ns == nsp && Mate[ns] != -1) nsA = nsB = Mate[ns];
         // the whole matrix is mirrored once at the end: tell the block not to mirror itself (all columns, no mirror)
         const int ce = (!Partial && MatrixIsSymmetric) ? 0x7fffffff : ColEnd;
         AssembleBlockOnDevice(this, nsA, nsB, Omega, M, BFIndexOffset[ns], BFIndexOffset[nsp], ns == nsp, false, 0, false, false, ColBegin, ce);
         if (ns == nsp) AddSurfaceZetaContribution(this, nsA, Omega, M, BFIndexOffset[ns], ColBegin, Partial ? ColEnd : 0); }
      else
       AssembleBEMMatrixBlockPBC(ns, nsp, Omega, kBloch, M, BFIndexOffset[ns], BFIndexOffset[nsp]);
    }
  if (MatrixIsSymmetric && !Partial)
   { if (scuffcuda_matrix_symmetrize(ctx, M->DeviceHandle, 0, TotalBFs)) ErrExit("%s", scuffcuda_last_error(ctx));
     M->HostStale = true; }
}

HMatrix *RWGGeometry::AssembleBEMMatrix(cdouble Omega, HMatrix *M) { return AssembleBEMMatrix(Omega, 0, M); }

// pair plans (and FIPPI records) of every block AssembleBEMMatrix will visit, without assembling: the loop of
// AssembleBEMMatrixDevice over surface pairs with its Mate shortcut
void RWGGeometry::PrepareNearPlans()
{ scuffcuda_context *ctx = GetCudaContext();
  if (LBasis != 0) return;                   // displaced blocks of periodic geometries are planned when first assembled
  SyncDeviceGeometry();
  for (int ns = 0; ns < NumSurfaces; ns++)
   for (int nsp = ns; nsp < NumSurfaces; nsp++)
    { if (ns == nsp && Mate[ns] != -1) continue;
      if (scuffcuda_prepare_block(ctx, ns, nsp, ns == nsp, 0, 0)) ErrExit("%s", scuffcuda_last_error(ctx)); }
}

// Step 1 of the periodic branch of AssembleBEMMatrixBlock (AssembleBEMMatrix.cc:465-519): the kBloch-independent blocks
// of the innermost lattice cells, each from the displaced-surface assembly on the device and stamped into M with its Bloch
// phase factor (StampInNeighborBlock, :72-114).  Leaves the block of M in HBM (M->HostStale).
void RWGGeometry::AssembleInnerCellBlocks(int nsa, int nsb, cdouble Omega, double *kBloch, HMatrix *M, int RowOffset, int ColOffset, void *Accelerator)
{ if (LBasis == 0) ErrExit("AssembleInnerCellBlocks: geometry is not periodic");
  if (!kBloch) ErrExit("%s:%i: Bloch wavevector must be specified for PBC geometries", __FILE__, __LINE__);
  if (M->RealComplex != LHM_COMPLEX) ErrExit("AssembleInnerCellBlocks: the BEM matrix of a periodic geometry is complex");
  scuffcuda_context *ctx = GetCudaContext();
  RWGSurface *Sa = Surfaces[nsa], *Sb = Surfaces[nsb];
  int CR[2]; double Signs[2];
  int NumCommonRegions = CountCommonRegions(Sa, Sb, CR, Signs);
  if (NumCommonRegions == 0) return;
  const int nr1 = CR[0], nr2 = NumCommonRegions == 2 ? CR[1] : -1;
  const bool UseSymmetry = (nsa == nsb), OneDLattice = (LDim == 1);
  double LBV[2][3] = { { 0, 0, 0 }, { 0, 0, 0 } };
  for (int i = 0; i < 3; i++) { LBV[0][i] = LBasis->GetEntryD(i, 0); if (LDim > 1) LBV[1][i] = LBasis->GetEntryD(i, 1); }

  KBIMBCache *Cache = (KBIMBCache *)Accelerator;
  const bool HaveCache = (Cache != 0);
  if (HaveCache && (Cache->nsa != nsa || Cache->nsb != nsb)) ErrExit("AssembleBEMMatrixBlock: accelerator was created for surfaces (%i,%i), not (%i,%i)", Cache->nsa, Cache->nsb, nsa, nsb);
  const bool HaveCleanCache = HaveCache && EqualFloat(Cache->Omega, Omega);       // AssembleBEMMatrix.cc:405-408
  if (HaveCache) Cache->Omega = Omega;

  M->EnsureDevice();
  // M->ZeroBlock(RowOffset, NBFA, ColOffset, NBFB) on the device: a block assembly with both regions omitted clears it
  AssembleBlockOnDevice(this, nsa, nsb, Omega, M, RowOffset, ColOffset, false, false, 0, true, true);
  HMatrix *BTemp = HaveCache ? 0 : new HMatrix(Sa->NumBFs, Sb->NumBFs, LHM_COMPLEX);
  unsigned long long Counts[5] = { 0, 0, 0, 0, 0 };
  bool done = false;
  int nb = 0;
  for (int n1 = +1; n1 >= -1 && !done; n1--)
   for (int n2 = +1; n2 >= -1 && !done; n2--)
    { if (OneDLattice && n2 != 0) continue;
      HMatrix &B = HaveCache ? *Cache->B[nb++] : *BTemp;
      double L[3] = { n1 * LBV[0][0] + n2 * LBV[1][0], n1 * LBV[0][1] + n2 * LBV[1][1], 0.0 };
      // contributions of regions that do not extend across the cell boundary are omitted (:487-499)
      bool Omit1, Omit2;
      if (n1 == 0 && n2 == 0) { Omit1 = false; Omit2 = (nr2 == -1); }
      else
       { Omit1 = Omit2 = true;
         if (n1 != 0 && RegionIsExtended[0][nr1]) Omit1 = false;
         if (n1 != 0 && nr2 != -1 && RegionIsExtended[0][nr2]) Omit2 = false;
         if (n2 != 0 && RegionIsExtended[1][nr1]) Omit1 = false;
         if (n2 != 0 && nr2 != -1 && RegionIsExtended[1][nr2]) Omit2 = false; }
      const bool Center = (n1 == 0 && n2 == 0);
      if (!HaveCleanCache)
       { AssembleBlockOnDevice(this, nsa, nsb, Omega, &B, 0, 0, UseSymmetry && Center, false, L, Omit1, Omit2);
         for (int i = 0; i < 5; i++) Counts[i] += LastPPICounts[i]; }
      else { SyncDeviceGeometry(); B.EnsureDevice(); }
      const double Arg = kBloch[0] * L[0] + kBloch[1] * L[1], BPF[2] = { cos(Arg), sin(Arg) };
      if (scuffcuda_matrix_stamp_neighbor_block(ctx, M->DeviceHandle, RowOffset, ColOffset, B.DeviceHandle, BPF, UseSymmetry && !Center))
       ErrExit("%s", scuffcuda_last_error(ctx));
      if (UseSymmetry && Center) done = true; }
  delete BTemp;
  for (int i = 0; i < 5; i++) LastPPICounts[i] = Counts[i];
  M->HostStale = true;
}

// CreateRegionGBA(nr, Omega, kBloch, ns1, ns2) (GBarAccelerator.cc:1012-1124,1205-1221): the periodic Green's function of region
// nr for R = x1 - x2, x1 on surface ns1 and x2 on ns2, as a device table; returns 0 if the region contributes nothing
static int CreateRegionGBA(RWGGeometry *G, int nr, cdouble Omega, double *kBloch, int ns1, int ns2)
{ scuffcuda_context *ctx = GetCudaContext();
  // the lattice with respect to which the region is periodic may be a sublattice of the geometry's
  int NumExtendedDimensions = 0;
  if (G->RegionIsExtended[0][nr]) NumExtendedDimensions++;
  if (G->LDim > 1 && G->RegionIsExtended[1][nr]) NumExtendedDimensions++;
  if (NumExtendedDimensions == 0) return 0;
  double lbv[4] = { 0, 0, 0, 0 }; int ldim = G->LDim;
  if (NumExtendedDimensions == G->LDim)
   { lbv[0] = G->LBasis->GetEntryD(0, 0); lbv[1] = G->LBasis->GetEntryD(1, 0);
     if (G->LDim > 1) { lbv[2] = G->LBasis->GetEntryD(0, 1); lbv[3] = G->LBasis->GetEntryD(1, 1); } }
  else
   { const int ExtendedDim = G->RegionIsExtended[0][nr] ? 0 : 1;
     lbv[0] = G->LBasis->GetEntryD(0, ExtendedDim); lbv[1] = G->LBasis->GetEntryD(1, ExtendedDim); ldim = 1; }
  const cdouble k = Omega * std::sqrt(G->EpsTF[nr] * G->MuTF[nr]);      // Omega * RegionMPs[nr]->GetRefractiveIndex(Omega)
  // no table where GBar is exponentially tiny at all evaluation points
  if (imag(k) > 0.0)
   { double minL = 2.0 * sqrt(lbv[0] * lbv[0] + lbv[1] * lbv[1]);
     if (ldim > 1) minL = fmin(minL, 2.0 * sqrt(lbv[2] * lbv[2] + lbv[3] * lbv[3]));
     if (imag(k) * minL > 14.0) return 0; }
  // range of the transverse coordinate Rho = |z| of R = x1 - x2 (from the surfaces' bounding boxes as the reference keeps them)
  double RMin[3], RMax[3], AbsRMin[3], AbsRMax[3];
  for (int Mu = 0; Mu < 3; Mu++)
   { RMax[Mu] = G->Surfaces[ns1]->RMax[Mu] - G->Surfaces[ns2]->RMin[Mu];
     RMin[Mu] = G->Surfaces[ns1]->RMin[Mu] - G->Surfaces[ns2]->RMax[Mu]; }
  double RhoMin, RhoMax;
  if (RMin[0] > RMax[0]) { RhoMin = 1.0; RhoMax = 0.0; }
  else
   { for (int Mu = 0; Mu < 3; Mu++)
      { if (RMin[Mu] >= 0.0) { AbsRMin[Mu] = fabs(RMin[Mu]); AbsRMax[Mu] = fabs(RMax[Mu]); }
        else if (RMin[Mu] < 0.0 && RMax[Mu] > 0.0) { AbsRMin[Mu] = 0.0; AbsRMax[Mu] = fmax(fabs(RMax[Mu]), fabs(RMin[Mu])); }
        else { AbsRMin[Mu] = fabs(RMax[Mu]); AbsRMax[Mu] = fabs(RMin[Mu]); } }
     if (ldim == 1) { RhoMin = sqrt(AbsRMin[1] * AbsRMin[1] + AbsRMin[2] * AbsRMin[2]); RhoMax = sqrt(AbsRMax[1] * AbsRMax[1] + AbsRMax[2] * AbsRMax[2]); }
     else { RhoMin = AbsRMin[2]; RhoMax = AbsRMax[2]; } }
  double RelTol = 1.0e-3;
  if (const char *str = getenv("SCUFF_INTERPOLATION_TOLERANCE")) { sscanf(str, "%le", &RelTol); Log("Setting interpolation tolerance to %e.", RelTol); }
  int ForceFull = 0, NMax = 0;
  if (const char *str = getenv("SCUFF_EWALD_FULL")) if (str[0] == '1') { Log("Forcing full Ewald summation."); ForceFull = 1; }
  if (const char *str = getenv("SCUFF_GBAR_NMAX")) NMax = atoi(str);
  if (getenv("SCUFF_GBAR_NDINTERP")) ErrExit("SCUFF_GBAR_NDINTERP (InterpND tables) is not supported by the device assembly path");
  const double kk[2] = { real(k), imag(k) };
  int Handle = 0;
  if (scuffcuda_gbar_create(ctx, kk, kBloch, lbv, ldim, RhoMin, RhoMax, RelTol, 1, ForceFull, NMax, &Handle)) ErrExit("%s", scuffcuda_last_error(ctx));
  return Handle;
}

void RWGGeometry::AssembleBEMMatrixBlockPBC(int nsa, int nsb, cdouble Omega, double *kBloch, HMatrix *M, int RowOffset, int ColOffset, void *Accelerator)
{ scuffcuda_context *ctx = GetCudaContext();
  int CR[2]; double Signs[2];
  const int NumCommonRegions = CountCommonRegions(Surfaces[nsa], Surfaces[nsb], CR, Signs);
  if (NumCommonRegions == 0) return;                                    // AssembleBEMMatrix.cc:413-414 (the block is left as it is)
  const int nr1 = CR[0], nr2 = NumCommonRegions == 2 ? CR[1] : -1;
  // step 1: innermost grid cells, Bloch-phased (:462-521)
  AssembleInnerCellBlocks(nsa, nsb, Omega, kBloch, M, RowOffset, ColOffset, Accelerator);
  unsigned long long Counts[5]; for (int i = 0; i < 5; i++) Counts[i] = LastPPICounts[i];
  // step 2: outer grid cells through the periodic Green's function of each common region (:526-560)
  UpdateCachedEpsMuValues(Omega);
  bool OmitRegion1 = false, OmitRegion2 = (nr2 == -1);
  const int GBA1 = CreateRegionGBA(this, nr1, Omega, kBloch, nsa, nsb);
  if (GBA1 == 0) OmitRegion1 = true;
  int GBA2 = 0;
  if (!OmitRegion2) { GBA2 = CreateRegionGBA(this, nr2, Omega, kBloch, nsa, nsb); if (GBA2 == 0) OmitRegion2 = true; }
  if (!OmitRegion1 || !OmitRegion2)
   { AssembleBlockOnDevice(this, nsa, nsb, Omega, M, RowOffset, ColOffset, false, true, 0, OmitRegion1, OmitRegion2, 0, 0, GBA1, GBA2);
     for (int i = 0; i < 5; i++) Counts[i] += LastPPICounts[i]; }
  if (GBA1) scuffcuda_gbar_destroy(ctx, GBA1);
  if (GBA2) scuffcuda_gbar_destroy(ctx, GBA2);
  for (int i = 0; i < 5; i++) LastPPICounts[i] = Counts[i];
  M->HostStale = true;
}

// PreloadCache / StoreCache (libscuff.h:662-663 -> GlobalFIPPICache.PreLoad/Store)
void PreloadCache(const char *FileName)
{ scuffcuda_context *ctx = GetCudaContext();
  if (scuffcuda_fippi_preload(ctx, FileName)) Warn("%s", scuffcuda_last_error(ctx)); }
void StoreCache(const char *FileName)
{ scuffcuda_context *ctx = GetCudaContext();
  if (scuffcuda_fippi_store(ctx, FileName)) Warn("%s", scuffcuda_last_error(ctx)); }

}  // namespace scuff

// CreateGammaMatrix (EdgeEdgeInteractions.cc:241-270): Gamma = Lambda^T Gamma0 Lambda, the generator of rotations about TorqueAxis
namespace scuff {
void CreateGammaMatrix(double *TorqueAxis, double *GammaMatrix)
{ double ax[3] = { TorqueAxis[0], TorqueAxis[1], TorqueAxis[2] };
  const double n = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
  for (int i = 0; i < 3; i++) ax[i] /= n;
  const double ct = ax[2], st = sqrt(1.0 - ct * ct);
  const double cp = (st < 1.0e-8) ? 1.0 : ax[0] / st, sp = (st < 1.0e-8) ? 0.0 : ax[1] / st;
  const double Lambda[3][3] = { { ct * cp, ct * sp, -st }, { -sp, cp, 0.0 }, { st * cp, st * sp, ct } };
  const double Gamma0[3][3] = { { 0.0, -1.0, 0.0 }, { 1.0, 0.0, 0.0 }, { 0.0, 0.0, 0.0 } };
  memset(GammaMatrix, 0, 9 * sizeof(double));
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++)
   GammaMatrix[i + 3 * j] += Lambda[k][i] * Gamma0[k][l] * Lambda[l][j];
}
}
