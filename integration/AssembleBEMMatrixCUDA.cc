// AssembleBEMMatrixCUDA.cc — the binding a SCUFF-EM maintainer adds to libs/libscuff/ so that the reference's own
// RWGGeometry::AssembleBEMMatrixBlock (and through it AssembleBEMMatrix and every application: scuff-scatter, scuff-cas3D,
// scuff-neq, ...) runs the surface-surface block on the GPU through the C ABI of libscuff_cuda (include/libscuff_cuda.h).
//
// It is compiled INTO the reference tree (oracle/Makefile, target _ref_cuda: the reference's sources from /root/reference, this
// file, and a three-line hook inserted at the top of the compact branch of AssembleBEMMatrixBlock, AssembleBEMMatrix.cc:371),
// uses only the reference's own types (RWGGeometry, RWGSurface, HMatrix, GetSSIArgStruct) and links with -lscuff_cuda.
// Opt-in at run time: SCUFF_CUDA=1 (device: SCUFF_CUDA_DEVICE, default 0).  Anything the device path does not cover
// (derivative blocks of a surface with itself, packed HMatrix storage, periodic lattices in this shim) returns false and the
// reference's CPU path runs, unchanged.  GradM / dMdT of blocks between different surfaces (what scuff-cas3D asks for,
// CasimirIntegrand.cc:443-447) go through scuffcuda_assemble_block_gradients.
//
// What it replaces: GetSurfaceSurfaceInteractions for Args->B = M at (RowOffset, ColOffset) with Symmetric = (nsa == nsb)
// (AssembleBEMMatrix.cc:373-387, SurfaceSurfaceInteractions.cc:426-585): material values are evaluated on the host exactly
// as the reference does (UpdateCachedEpsMuValues, CountCommonRegions), the block is assembled in HBM and copied into the
// caller's column-major HMatrix; the surface-impedance term stays the reference's own host routine.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <map>
#include <libhrutil.h>
#include "libscuff.h"
#include "libscuffInternals.h"
#include "libscuff_cuda.h"

namespace scuff {

static long BlocksOnDevice=0;
static void ReportAtExit()
{ if (getenv("SCUFF_CUDA_VERBOSE")) fprintf(stderr,"libscuff_cuda: %li BEM matrix blocks assembled on the device\n",BlocksOnDevice); }

static scuffcuda_context *CudaContext()
{ static scuffcuda_context *ctx=0;
  if (!ctx)
   { const char *s=getenv("SCUFF_CUDA_DEVICE");
     if (scuffcuda_create(s ? atoi(s) : 0, &ctx)) ErrExit("libscuff_cuda: %s",scuffcuda_last_error(0));
     atexit(ReportAtExit);
   }
  return ctx;
}

// flatten one RWGSurface (libscuff.h:174-328) into the arrays scuffcuda_set_surface takes.  Re-sending byte-identical
// panels is cheap and keeps the device-side pair plans and FIPPI records; moved surfaces (Transform/UnTransform) are
// re-planned by the library.
static void SendSurface(scuffcuda_context *ctx, int ns, RWGSurface *S)
{ std::vector<int> pvi(3*S->NumPanels), et(8*(size_t)S->NumEdges);
  std::vector<double> pc(3*S->NumPanels), pr(S->NumPanels), el(S->NumEdges);
  for(int np=0; np<S->NumPanels; np++)
   { RWGPanel *P=S->Panels[np];
     for(int i=0; i<3; i++) { pvi[3*np+i]=P->VI[i]; pc[3*np+i]=P->Centroid[i]; }
     pr[np]=P->Radius;
   }
  for(int ne=0; ne<S->NumEdges; ne++)
   { RWGEdge *E=S->Edges[ne]; int *e=&et[8*(size_t)ne];
     e[0]=E->iV1; e[1]=E->iV2; e[2]=E->iQP; e[3]=E->iQM; e[4]=E->iPPanel; e[5]=E->iMPanel; e[6]=E->PIndex; e[7]=E->MIndex;
     el[ne]=E->Length;
   }
  if (scuffcuda_set_surface(ctx,ns,S->NumVertices,S->Vertices,S->NumPanels,pvi.data(),pc.data(),pr.data(),
                            S->NumEdges,et.data(),el.data(),S->IsPEC ? 1 : 0))
   ErrExit("%s",scuffcuda_last_error(ctx));
}

// returns true if the block was assembled on the device
bool AssembleBEMMatrixBlockCUDA(RWGGeometry *G, int nsa, int nsb, cdouble Omega, HMatrix *M, HMatrix **GradM,
                                int RowOffset, int ColOffset, int NumTorqueAxes, HMatrix **dMdT, double *GammaMatrix)
{ const char *Env=getenv("SCUFF_CUDA");
  if (!Env || Env[0]!='1') return false;
  const bool WantDerivatives = GradM!=0 || (NumTorqueAxes>0 && dMdT!=0);
  if (G->LBasis!=0 || M->StorageType!=LHM_NORMAL) return false;
  if (WantDerivatives && (nsa==nsb || NumTorqueAxes>3 || (NumTorqueAxes>0 && GammaMatrix==0))) return false;
  if (RWGGeometry::UseHRWGFunctions) return false;          // half-RWG basis functions: not on the device path

  scuffcuda_context *ctx=CudaContext();
  static RWGGeometry *Resident=0;                            // one geometry lives on the device at a time
  if (Resident!=G)
   { if (scuffcuda_set_num_surfaces(ctx,0) || scuffcuda_set_num_surfaces(ctx,G->NumSurfaces)) ErrExit("%s",scuffcuda_last_error(ctx));
     Resident=G;
   }
  SendSurface(ctx,nsa,G->Surfaces[nsa]);
  if (nsb!=nsa) SendSurface(ctx,nsb,G->Surfaces[nsb]);

  // materials and signs as GetSurfaceSurfaceInteractions chooses them (SurfaceSurfaceInteractions.cc:437,462-480)
  G->UpdateCachedEpsMuValues(Omega);
  RWGSurface *Sa=G->Surfaces[nsa], *Sb=G->Surfaces[nsb];
  double Signs[2]; int CR[2];
  int NCR=CountCommonRegions(Sa,Sb,CR,Signs);
  scuffcuda_ssi_args A; memset(&A,0,sizeof(A));
  A.omega[0]=real(Omega); A.omega[1]=imag(Omega);
  if (NCR>=1)
   { cdouble e=G->EpsTF[CR[0]], m=G->MuTF[CR[0]];
     A.eps_a[0]=real(e); A.eps_a[1]=imag(e); A.mu_a[0]=real(m); A.mu_a[1]=imag(m); A.sign_a=Signs[0]; }
  if (NCR==2)
   { cdouble e=G->EpsTF[CR[1]], m=G->MuTF[CR[1]];
     A.eps_b[0]=real(e); A.eps_b[1]=imag(e); A.mu_b[0]=real(m); A.mu_b[1]=imag(m); A.sign_b=Signs[1]; }
  A.symmetric=(nsa==nsb) ? 1 : 0;
  A.force_taylor_duffy=RWGGeometry::DisableCache ? 1 : 0;

  // a scratch device matrix per block shape, kept across calls (frequency sweeps call this thousands of times)
  int NBFA=Sa->NumBFs, NBFB=Sb->NumBFs, IsReal=(M->RealComplex==LHM_REAL) ? 1 : 0;
  static std::map<std::vector<int>, int> Scratch;
  std::vector<int> Key; Key.push_back(NBFA); Key.push_back(NBFB); Key.push_back(IsReal);
  int h=0;
  if (Scratch.count(Key)) h=Scratch[Key];
  else
   { if (scuffcuda_matrix_create(ctx,NBFA,NBFB,IsReal,&h)) ErrExit("%s",scuffcuda_last_error(ctx));
     Scratch[Key]=h; }
  if (scuffcuda_assemble_block(ctx,nsa,nsb,&A,h)) ErrExit("%s",scuffcuda_last_error(ctx));
  BlocksOnDevice++;

  // into the caller's HMatrix at (RowOffset,ColOffset): column-major, leading dimension M->NR (HMatrix.cc:333)
  if (RowOffset<0 || ColOffset<0 || RowOffset+NBFA>M->NR || ColOffset+NBFB>M->NC)
   ErrExit("%s:%i: block exceeds matrix",__FILE__,__LINE__);
  void *dst = IsReal ? (void *)(M->DM + RowOffset + (size_t)ColOffset*M->NR)
                     : (void *)(M->ZM + RowOffset + (size_t)ColOffset*M->NR);
  if (scuffcuda_matrix_download(ctx,h,dst,M->NR)) ErrExit("%s",scuffcuda_last_error(ctx));

  // derivative blocks: same arguments, one scratch matrix per requested component, copied into the caller's matrices
  if (WantDerivatives)
   { int gh[3]={0,0,0}, th[3]={0,0,0};
     static std::map<std::vector<int>, std::vector<int> > GradScratch;
     std::vector<int> &hs=GradScratch[Key];
     while (hs.size()<6) { int hh=0; if (scuffcuda_matrix_create(ctx,NBFA,NBFB,IsReal,&hh)) ErrExit("%s",scuffcuda_last_error(ctx)); hs.push_back(hh); }
     for(int m=0; m<3; m++)
      { if (GradM && GradM[m]) gh[m]=hs[m];
        if (dMdT && m<NumTorqueAxes && dMdT[m]) th[m]=hs[3+m]; }
     if (scuffcuda_assemble_block_gradients(ctx,nsa,nsb,&A,gh,dMdT ? NumTorqueAxes : 0,GammaMatrix,th)) ErrExit("%s",scuffcuda_last_error(ctx));
     for(int m=0; m<3; m++)
      for(int which=0; which<2; which++)
       { HMatrix *D = which==0 ? ((GradM && GradM[m]) ? GradM[m] : 0) : ((dMdT && m<NumTorqueAxes && dMdT[m]) ? dMdT[m] : 0);
         if (!D) continue;
         if (D->NR!=M->NR || D->NC!=M->NC || D->RealComplex!=M->RealComplex) ErrExit("%s:%i: derivative matrices must have the shape and type of M",__FILE__,__LINE__);
         void *dd = IsReal ? (void *)(D->DM + RowOffset + (size_t)ColOffset*D->NR) : (void *)(D->ZM + RowOffset + (size_t)ColOffset*D->NR);
         if (scuffcuda_matrix_download(ctx,which==0 ? gh[m] : th[m],dd,D->NR)) ErrExit("%s",scuffcuda_last_error(ctx)); }
   }

  // SURFACE_IMPEDANCE term: the reference's own routine on the host matrix (SurfaceSurfaceInteractions.cc:339-421,563-564)
  if (nsa==nsb && Sa->SurfaceZeta!=0)
   { GetSSIArgStruct MyArgs, *Args=&MyArgs;
     InitGetSSIArgs(Args);
     Args->G=G; Args->Sa=Sa; Args->Sb=Sb; Args->Omega=Omega; Args->Symmetric=true;
     Args->B=M; Args->RowOffset=RowOffset; Args->ColOffset=ColOffset;
     AddSurfaceZetaContributionToBEMMatrix(Args);
   }
  return true;
}

// PreloadCache / StoreCache twins for the device-side FIPPI table (same file format, FIPPICache.cc:209-400)
void PreloadCacheCUDA(const char *FileName)
{ if (scuffcuda_fippi_preload(CudaContext(),FileName)) Warn("%s",scuffcuda_last_error(CudaContext())); }
void StoreCacheCUDA(const char *FileName)
{ if (scuffcuda_fippi_store(CudaContext(),FileName)) Warn("%s",scuffcuda_last_error(CudaContext())); }

} // namespace scuff
