// HMatrix: the boundary container of the path (libs/libhmat/libhmat.h:165-383, HMatrix.cc:325-447).
// Column-major dense real/complex storage on the host, plus a device-resident twin in HBM that the
// assembly kernels fill and cuSOLVER factors (LBWrappers.cc:150-230 -> scuffcuda_lu_factor/solve).
#include <cstdlib>
#include "host_internal.h"

namespace scuff {

HMatrix::HMatrix(int NRows, int NCols, int pRealComplex, int pStorageType, void *data)
{ NR = NRows; NC = NCols; RealComplex = pRealComplex; StorageType = pStorageType;
  ipiv = 0; DM = 0; ZM = 0; ErrMsg = 0; DeviceHandle = 0; HostStale = false; Pinned = false;
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
      { Pinned = true; if (RealComplex == LHM_REAL) DM = (double *)p; else ZM = (cdouble *)p; }
     else if (RealComplex == LHM_REAL) DM = (double *)calloc(n, sizeof(double));
     else ZM = (cdouble *)calloc(n, sizeof(cdouble));
     if (!DM && !ZM) ErrMsg = strdup("out of memory");
   }
}

HMatrix::~HMatrix()
{ if (DeviceHandle) scuffcuda_matrix_destroy(GetCudaContext(), DeviceHandle);
  if (ownsM && Pinned) scuffcuda_host_free(GetCudaContext(), DM ? (void *)DM : (void *)ZM);
  else if (ownsM) { if (DM) free(DM); if (ZM) free(ZM); }
  if (ipiv) free(ipiv);
}

void HMatrix::EnsureDevice()
{ if (DeviceHandle) return;
  scuffcuda_context *ctx = GetCudaContext();
  if (scuffcuda_matrix_create(ctx, NR, NC, RealComplex == LHM_REAL, &DeviceHandle))
   ErrExit("%s", scuffcuda_last_error(ctx));
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
  HostStale = false;
}

cdouble HMatrix::GetEntry(size_t nr, size_t nc)
{ SyncToHost();
  size_t i = nr + nc * (size_t)NR;
  return RealComplex == LHM_REAL ? cdouble(DM[i], 0.0) : ZM[i];
}
double HMatrix::GetEntryD(size_t nr, size_t nc) { return real(GetEntry(nr, nc)); }
void HMatrix::SetEntry(size_t nr, size_t nc, double Entry)
{ SyncToHost(); size_t i = nr + nc * (size_t)NR;
  if (RealComplex == LHM_REAL) DM[i] = Entry; else ZM[i] = Entry; }
void HMatrix::SetEntry(size_t nr, size_t nc, cdouble Entry)
{ SyncToHost(); size_t i = nr + nc * (size_t)NR;
  if (RealComplex == LHM_REAL) DM[i] = real(Entry); else ZM[i] = Entry; }
void HMatrix::AddEntry(size_t nr, size_t nc, double Entry)
{ SyncToHost(); size_t i = nr + nc * (size_t)NR;
  if (RealComplex == LHM_REAL) DM[i] += Entry; else ZM[i] += Entry; }
void HMatrix::AddEntry(size_t nr, size_t nc, cdouble Entry)
{ SyncToHost(); size_t i = nr + nc * (size_t)NR;
  if (RealComplex == LHM_REAL) DM[i] += real(Entry); else ZM[i] += Entry; }

void HMatrix::Zero() { ZeroBlock(0, NR, 0, NC); }
void HMatrix::ZeroBlock(int RowOffset, int NumRows, int ColOffset, int NumCols)
{ SyncToHost();
  for (int nc = ColOffset; nc < ColOffset + NumCols; nc++)
   for (int nr = RowOffset; nr < RowOffset + NumRows; nr++)
    { size_t i = nr + nc * (size_t)NR; if (RealComplex == LHM_REAL) DM[i] = 0.0; else ZM[i] = 0.0; }
}
void HMatrix::InsertBlock(HMatrix *B, int RowOffset, int ColOffset, int NRB, int NCB, int BRowOffset, int BColOffset)
{ SyncToHost(); B->SyncToHost();
  for (int nc = 0; nc < NCB; nc++)
   for (int nr = 0; nr < NRB; nr++)
    SetEntry(RowOffset + nr, ColOffset + nc, B->GetEntry(BRowOffset + nr, BColOffset + nc));
}

int HMatrix::LUFactorize()
{ scuffcuda_context *ctx = GetCudaContext();
  if (!DeviceHandle) SyncToDevice();   // matrix was filled on the host
  if (scuffcuda_lu_factor(ctx, DeviceHandle)) { Warn("LUFactorize: %s", scuffcuda_last_error(ctx)); return 1; }
  HostStale = true;
  return 0;
}
int HMatrix::LUSolve(HMatrix *X)
{ scuffcuda_context *ctx = GetCudaContext();
  if (!DeviceHandle) ErrExit("LUSolve: matrix has not been LU-factorized");
  if (X->NR != NR || X->RealComplex != RealComplex) ErrExit("LUSolve: dimension/type mismatch");
  X->SyncToHost();
  void *p = RealComplex == LHM_REAL ? (void *)X->DM : (void *)X->ZM;
  if (scuffcuda_lu_solve(ctx, DeviceHandle, p, X->NC)) ErrExit("%s", scuffcuda_last_error(ctx));
  return 0;
}

}  // namespace scuff
