// HMatrix: the boundary container of the path (libs/libhmat/libhmat.h:165-383, HMatrix.cc:325-447).
// Column-major dense real/complex storage on the host, plus a device-resident twin in HBM that the
// assembly kernels fill and cuSOLVER factors (LBWrappers.cc:150-230 -> scuffcuda_lu_factor/solve).
#include <cstdlib>
#include "host_internal.h"

namespace scuff {

HMatrix::HMatrix(int NRows, int NCols, int pRealComplex, int pStorageType, void *data)
{ NR = NRows; NC = NCols; RealComplex = pRealComplex; StorageType = pStorageType;
  ipiv = 0; DM = 0; ZM = 0; ErrMsg = 0; DeviceHandle = 0; HostStale = false; DeviceStale = false; Pinned = false;
  if (StorageType != LHM_NORMAL)
   ErrExit("HMatrix: packed storage is not supported by the device assembly path");
  size_t n = (size_t)NR * NC;
  if (data) { ownsM = false; if (RealComplex == LHM_REAL) DM = (double *)data; else ZM = (cdouble *)data; }
  else
   { ownsM = true;
     // large matrices come back from HBM over PCIe: page-locked host memory when a device is present
     size_t bytes = n * (RealComplex == LHM_REAL ? sizeof(double) : sizeof(cdouble));
     void *p = 0;
     if (bytes >= ((size_t)32 << 20) && scuffcuda_device_count() > 0 && scuffcuda_host_alloc(GetCudaContext(), bytes, &p) == 0 && p)
      { Pinned = true; AddLiveDeviceObject(+1); if (RealComplex == LHM_REAL) DM = (double *)p; else ZM = (cdouble *)p; }
     else if (RealComplex == LHM_REAL) DM = (double *)calloc(n, sizeof(double));
     else ZM = (cdouble *)calloc(n, sizeof(cdouble));
     if (!DM && !ZM) ErrMsg = strdup("out of memory");
   }
}

HMatrix::~HMatrix()
{ if (DeviceHandle) { scuffcuda_matrix_destroy(GetCudaContext(), DeviceHandle); AddLiveDeviceObject(-1); }
  if (ownsM && Pinned) { scuffcuda_host_free(GetCudaContext(), DM ? (void *)DM : (void *)ZM); AddLiveDeviceObject(-1); }
  else if (ownsM) { if (DM) free(DM); if (ZM) free(ZM); }
  if (ipiv) free(ipiv);
}

// Coherence of the two copies.  HostStale: the device twin is newer (kernels wrote it); DeviceStale: the host array is newer
// (SetEntry/AddEntry/Zero/ZeroBlock/InsertBlock, or a caller that wrote DM/ZM directly and called MarkHostDirty()).  Never both.
// Host mutators pull the device copy first if it is newer, then mark the device stale; every device consumer
// (assembly, LUFactorize, LUSolve) pushes the host copy first if the device is stale.
void HMatrix::EnsureDevice()
{ if (DeviceHandle) return;
  scuffcuda_context *ctx = GetCudaContext();
  if (scuffcuda_matrix_create(ctx, NR, NC, RealComplex == LHM_REAL, &DeviceHandle))
   ErrExit("%s", scuffcuda_last_error(ctx));
  AddLiveDeviceObject(+1);
}
void HMatrix::SyncToHost()
{ if (!DeviceHandle || !HostStale) return;
  scuffcuda_context *ctx = GetCudaContext();
  void *p = RealComplex == LHM_REAL ? (void *)DM : (void *)ZM;
  if (scuffcuda_matrix_download(ctx, DeviceHandle, p, NR)) ErrExit("%s", scuffcuda_last_error(ctx));
  HostStale = false;
}
void HMatrix::SyncToDevice()
{ EnsureDevice();
  scuffcuda_context *ctx = GetCudaContext();
  void *p = RealComplex == LHM_REAL ? (void *)DM : (void *)ZM;
  if (scuffcuda_matrix_upload(ctx, DeviceHandle, p, NR)) ErrExit("%s", scuffcuda_last_error(ctx));
  HostStale = false; DeviceStale = false;
}
// make the device twin current before a kernel reads or partially updates it
void HMatrix::PrepareDevice()
{ if (!DeviceHandle || DeviceStale) SyncToDevice(); }
void HMatrix::MarkHostDirty()
{ SyncToHost(); if (DeviceHandle) DeviceStale = true; }

cdouble HMatrix::GetEntry(size_t nr, size_t nc)
{ SyncToHost();
  size_t i = nr + nc * (size_t)NR;
  return RealComplex == LHM_REAL ? cdouble(DM[i], 0.0) : ZM[i];
}
double HMatrix::GetEntryD(size_t nr, size_t nc) { return real(GetEntry(nr, nc)); }
void HMatrix::SetEntry(size_t nr, size_t nc, double Entry)
{ MarkHostDirty(); size_t i = nr + nc * (size_t)NR;
  if (RealComplex == LHM_REAL) DM[i] = Entry; else ZM[i] = Entry; }
void HMatrix::SetEntry(size_t nr, size_t nc, cdouble Entry)
{ MarkHostDirty(); size_t i = nr + nc * (size_t)NR;
  if (RealComplex == LHM_REAL) DM[i] = real(Entry); else ZM[i] = Entry; }
void HMatrix::AddEntry(size_t nr, size_t nc, double Entry)
{ MarkHostDirty(); size_t i = nr + nc * (size_t)NR;
  if (RealComplex == LHM_REAL) DM[i] += Entry; else ZM[i] += Entry; }
void HMatrix::AddEntry(size_t nr, size_t nc, cdouble Entry)
{ MarkHostDirty(); size_t i = nr + nc * (size_t)NR;
  if (RealComplex == LHM_REAL) DM[i] += real(Entry); else ZM[i] += Entry; }

void HMatrix::Zero()
{ // whole matrix: nothing to fetch from the device first
  HostStale = false;
  memset(RealComplex == LHM_REAL ? (void *)DM : (void *)ZM, 0, (size_t)NR * NC * (RealComplex == LHM_REAL ? sizeof(double) : sizeof(cdouble)));
  if (DeviceHandle)
   { scuffcuda_context *ctx = GetCudaContext();
     if (scuffcuda_matrix_zero(ctx, DeviceHandle)) ErrExit("%s", scuffcuda_last_error(ctx));
     DeviceStale = false; }
}
void HMatrix::ZeroBlock(int RowOffset, int NumRows, int ColOffset, int NumCols)
{ if (RowOffset == 0 && ColOffset == 0 && NumRows == NR && NumCols == NC) { Zero(); return; }
  MarkHostDirty();
  for (int nc = ColOffset; nc < ColOffset + NumCols; nc++)
   for (int nr = RowOffset; nr < RowOffset + NumRows; nr++)
    { size_t i = nr + nc * (size_t)NR; if (RealComplex == LHM_REAL) DM[i] = 0.0; else ZM[i] = 0.0; }
}
void HMatrix::InsertBlock(HMatrix *B, int RowOffset, int ColOffset, int NRB, int NCB, int BRowOffset, int BColOffset)
{ if (RowOffset < 0 || ColOffset < 0 || BRowOffset < 0 || BColOffset < 0 || RowOffset + NRB > NR || ColOffset + NCB > NC
      || BRowOffset + NRB > B->NR || BColOffset + NCB > B->NC)
   ErrExit("HMatrix::InsertBlock: block out of range");
  // blocks that live in HBM (the T/U blocks of scuff-cas3D / scuff-neq, CasimirIntegrand.cc:159-196) are stamped on the
  // device: no round trip of either matrix through the host
  if (B != this && B->DeviceHandle && !B->DeviceStale && B->HostStale && B->RealComplex == RealComplex)
   { scuffcuda_context *ctx = GetCudaContext();
     PrepareDevice();
     if (scuffcuda_matrix_insert_block(ctx, DeviceHandle, RowOffset, ColOffset, B->DeviceHandle, BRowOffset, BColOffset, NRB, NCB))
      ErrExit("%s", scuffcuda_last_error(ctx));
     HostStale = true;
     return; }
  MarkHostDirty(); B->SyncToHost();
  for (int nc = 0; nc < NCB; nc++)
   for (int nr = 0; nr < NRB; nr++)
    { const size_t d = (size_t)(RowOffset + nr) + (size_t)(ColOffset + nc) * NR, sB = (size_t)(BRowOffset + nr) + (size_t)(BColOffset + nc) * B->NR;
      const cdouble v = B->RealComplex == LHM_REAL ? cdouble(B->DM[sB], 0.0) : B->ZM[sB];
      if (RealComplex == LHM_REAL) DM[d] = real(v); else ZM[d] = v; }
}

int HMatrix::LUFactorize()
{ scuffcuda_context *ctx = GetCudaContext();
  PrepareDevice();                     // matrix was filled or edited on the host
  if (scuffcuda_lu_factor(ctx, DeviceHandle)) { Warn("LUFactorize: %s", scuffcuda_last_error(ctx)); return 1; }
  HostStale = true;                    // as in LAPACK, the matrix now holds its LU factors
  return 0;
}
int HMatrix::LUSolve(HMatrix *X)
{ scuffcuda_context *ctx = GetCudaContext();
  if (!DeviceHandle) ErrExit("LUSolve: matrix has not been LU-factorized");
  if (DeviceStale) ErrExit("LUSolve: the matrix was modified on the host after LUFactorize()");
  if (X->NR != NR || X->RealComplex != RealComplex) ErrExit("LUSolve: dimension/type mismatch");
  if (X->DeviceHandle && X->HostStale && !X->DeviceStale)      // right-hand sides that already live in HBM (e.g. a reduced-field matrix)
   { if (scuffcuda_lu_solve_device(ctx, DeviceHandle, X->DeviceHandle)) ErrExit("%s", scuffcuda_last_error(ctx));
     return 0; }
  X->SyncToHost();
  void *p = RealComplex == LHM_REAL ? (void *)X->DM : (void *)X->ZM;
  if (scuffcuda_lu_solve(ctx, DeviceHandle, p, X->NC)) ErrExit("%s", scuffcuda_last_error(ctx));
  if (X->DeviceHandle) X->DeviceStale = true;
  return 0;
}

}  // namespace scuff
