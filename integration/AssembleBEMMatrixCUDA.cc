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

static long BlocksOnDevice=0, RHSOnDevice=0, PFTOnDevice=0;
static void ReportAtExit()
{ if (getenv("SCUFF_CUDA_VERBOSE"))
   fprintf(stderr,"libscuff_cuda: %li BEM matrix blocks assembled on the device; %li right-hand sides, %li EMTPFT evaluations on the device\n",
           BlocksOnDevice,RHSOnDevice,PFTOnDevice); }

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

static RWGGeometry *Resident=0;                            // one geometry lives on the device at a time
static void MakeResident(scuffcuda_context *ctx, RWGGeometry *G)
{ if (Resident==G) return;
  if (scuffcuda_set_num_surfaces(ctx,0) || scuffcuda_set_num_surfaces(ctx,G->NumSurfaces)) ErrExit("%s",scuffcuda_last_error(ctx));
  Resident=G;
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
  MakeResident(ctx,G);
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

// ---- right-hand side and power/force/torque of a scattering run ------------------------------------
// Hooks (inserted by oracle/Makefile into generated copies of the reference's sources):
//   RWGGeometry::AssembleRHSVector, after UpdateIncFields (AssembleRHSVector.cc:375):
//       if (AssembleRHSVectorCUDA(this,Omega,kBloch,IF,RHS)) return RHS;
//   GetEMTPFTMatrix, after the PFTMatrix size check (EMTPFT.cc:556):
//       if (GetEMTPFTMatrixCUDA(G,Omega,IF,KNVector,DRMatrix,PFTMatrix,Interior,EMTPFTIMethod)) return PFTMatrix;
// Plane waves are evaluated on the device; any other IncField by its own GetFields / GetFieldGradients at the device's
// cubature points.  The vectors are the caller's host HVectors (16 N bytes each way).
static void FillPlaneWave(scuffcuda_planewave &W, PlaneWave *PW, double Sign)
{ for(int m=0; m<3; m++) { W.e0[2*m]=real(PW->E0[m]); W.e0[2*m+1]=imag(PW->E0[m]); W.nhat[m]=PW->nHat[m]; }
  W.eps[0]=real(PW->Eps); W.eps[1]=imag(PW->Eps); W.mu[0]=real(PW->Mu); W.mu[1]=imag(PW->Mu); W.sign=Sign;
}
static int VectorOnDevice(scuffcuda_context *ctx, int N, bool IsReal)
{ static std::map<std::pair<int,int>,int> Vectors;
  std::pair<int,int> Key(N,IsReal ? 1 : 0);
  if (Vectors.count(Key)) return Vectors[Key];
  int h=0; if (scuffcuda_matrix_create(ctx,N,1,IsReal ? 1 : 0,&h)) ErrExit("%s",scuffcuda_last_error(ctx));
  Vectors[Key]=h; return h;
}

bool AssembleRHSVectorCUDA(RWGGeometry *G, cdouble Omega, double *kBloch, IncField *IFList, HVector *RHS)
{ const char *Env=getenv("SCUFF_CUDA");
  if (!Env || Env[0]!='1') return false;
  if (G->LBasis!=0 || kBloch!=0 || RWGGeometry::UseHRWGFunctions || IFList==0) return false;
  scuffcuda_context *ctx=CudaContext();
  MakeResident(ctx,G);
  for(int ns=0; ns<G->NumSurfaces; ns++) SendSurface(ctx,ns,G->Surfaces[ns]);
  const bool IsReal=(RHS->RealComplex==LHM_REAL);
  int hv=VectorOnDevice(ctx,RHS->N,IsReal);
  if (scuffcuda_matrix_zero(ctx,hv)) ErrExit("%s",scuffcuda_last_error(ctx));
  const double omega[2]={real(Omega),imag(Omega)};
  for(int ns=0; ns<G->NumSurfaces; ns++)
   { RWGSurface *S=G->Surfaces[ns];
     if (S->NumEdges==0) continue;
     std::vector<scuffcuda_planewave> PWs; std::vector<IncField *> Others; std::vector<double> OtherSigns;
     for(IncField *IF=IFList; IF; IF=IF->Next)
      { double Sign;
        if (S->RegionIndices[0]==IF->RegionIndex) Sign=-1.0;          // AssembleRHSVector.cc:297-306
        else if (S->RegionIndices[1]==IF->RegionIndex) Sign=+1.0;
        else continue;
        PlaneWave *PW=dynamic_cast<PlaneWave *>(IF);
        if (PW && PW->LBasis==0 && PWs.size()<8) { scuffcuda_planewave W; FillPlaneWave(W,PW,Sign); PWs.push_back(W); }
        else { Others.push_back(IF); OtherSigns.push_back(Sign); } }
     if (PWs.empty() && Others.empty()) continue;
     std::vector<double> Tab;
     if (!Others.empty())
      { long npts=0;
        if (scuffcuda_rhs_points(ctx,ns,0,&npts)) ErrExit("%s",scuffcuda_last_error(ctx));
        std::vector<double> X(3*(size_t)npts); Tab.assign(12*(size_t)npts,0.0);
        if (scuffcuda_rhs_points(ctx,ns,&X[0],&npts)) ErrExit("%s",scuffcuda_last_error(ctx));
#pragma omp parallel for schedule(static)
        for(long n=0; n<npts; n++)
         for(size_t i=0; i<Others.size(); i++)
          { cdouble EH[6]; Others[i]->GetFields(&X[3*n],EH);
            for(int q=0; q<6; q++) { Tab[12*n+2*q]+=OtherSigns[i]*real(EH[q]); Tab[12*n+2*q+1]+=OtherSigns[i]*imag(EH[q]); } } }
     if (scuffcuda_assemble_rhs(ctx,ns,omega,(int)PWs.size(),PWs.empty() ? 0 : &PWs[0],Tab.empty() ? 0 : &Tab[0],hv,G->BFIndexOffset[ns]))
      ErrExit("%s",scuffcuda_last_error(ctx)); }
  void *dst = IsReal ? (void *)RHS->DV : (void *)RHS->ZV;
  if (scuffcuda_matrix_download(ctx,hv,dst,RHS->N)) ErrExit("%s",scuffcuda_last_error(ctx));
  RHSOnDevice++;
  return true;
}

bool GetEMTPFTMatrixCUDA(RWGGeometry *G, cdouble Omega, IncField *IF, HVector *KN, HMatrix *DRMatrix, HMatrix *PFTMatrix,
                         bool Interior, int EMTPFTIMethod)
{ const char *Env=getenv("SCUFF_CUDA");
  if (!Env || Env[0]!='1') return false;
  // the device has the cubature integrals (SCUFF_EMTPFTI_EHDERIVATIVES1); everything else stays on the reference's CPU path
  if (EMTPFTIMethod!=SCUFF_EMTPFTI_EHDERIVATIVES1 || DRMatrix!=0 || KN==0 || KN->RealComplex!=LHM_COMPLEX || G->LBasis!=0) return false;
  if (imag(Omega)!=0.0 || real(Omega)<=0.0 || RWGGeometry::UseHRWGFunctions) return false;
  if (getenv("SCUFF_EMTPFT_NSAONLY") || getenv("SCUFF_EMTPFT_NSBONLY")) return false;
  scuffcuda_context *ctx=CudaContext();
  MakeResident(ctx,G);
  const int NS=G->NumSurfaces;
  for(int ns=0; ns<NS; ns++) SendSurface(ctx,ns,G->Surfaces[ns]);
  int hv=VectorOnDevice(ctx,KN->N,false);
  if (scuffcuda_matrix_upload(ctx,hv,KN->ZV,KN->N)) ErrExit("%s",scuffcuda_last_error(ctx));
  std::vector<int> Regions(2*NS), Offsets(NS); std::vector<double> Origins(3*NS), EpsMu(4*G->NumRegions);
  for(int ns=0; ns<NS; ns++)
   { Regions[2*ns]=G->Surfaces[ns]->RegionIndices[0]; Regions[2*ns+1]=G->Surfaces[ns]->RegionIndices[1]; Offsets[ns]=G->BFIndexOffset[ns];
     for(int m=0; m<3; m++) Origins[3*ns+m]=G->Surfaces[ns]->Origin[m]; }
  for(int nr=0; nr<G->NumRegions; nr++)
   { cdouble Eps, Mu; G->RegionMPs[nr]->GetEpsMu(Omega,&Eps,&Mu);
     EpsMu[4*nr]=real(Eps); EpsMu[4*nr+1]=imag(Eps); EpsMu[4*nr+2]=real(Mu); EpsMu[4*nr+3]=imag(Mu); }
  scuffcuda_pft_args A; memset(&A,0,sizeof(A));
  A.omega[0]=real(Omega); A.omega[1]=0.0; A.interior=Interior ? 1 : 0; A.num_surfaces=NS;
  A.region_indices=&Regions[0]; A.bf_offsets=&Offsets[0]; A.origins=&Origins[0]; A.num_regions=G->NumRegions; A.eps_mu=&EpsMu[0]; A.if_region=-1;
  scuffcuda_planewave W; std::vector<double> Tab;
  if (IF)
   { A.if_region=IF->RegionIndex;
     PlaneWave *PW=dynamic_cast<PlaneWave *>(IF);
     if (PW && PW->LBasis==0) { FillPlaneWave(W,PW,1.0); A.num_pw=1; A.pw=&W; }
     else
      { size_t total=0; std::vector<size_t> first(NS);
        for(int ns=0; ns<NS; ns++) { long n=0; if (scuffcuda_pft_points(ctx,ns,0,&n)) ErrExit("%s",scuffcuda_last_error(ctx)); first[ns]=total; total+=(size_t)n; }
        std::vector<double> X(3*total); Tab.assign(48*total,0.0);
        for(int ns=0; ns<NS; ns++) { long n=0; if (scuffcuda_pft_points(ctx,ns,&X[3*first[ns]],&n)) ErrExit("%s",scuffcuda_last_error(ctx)); }
#pragma omp parallel for schedule(static)
        for(long n=0; n<(long)total; n++)
         { cdouble EH[6], dEH[3][6]; IF->GetFields(&X[3*n],EH); IF->GetFieldGradients(&X[3*n],dEH);
           double *t=&Tab[48*n];
           for(int q=0; q<6; q++) { t[2*q]=real(EH[q]); t[2*q+1]=imag(EH[q]); }
           for(int m=0; m<3; m++) for(int q=0; q<6; q++) { t[12+12*m+2*q]=real(dEH[m][q]); t[13+12*m+2*q]=imag(dEH[m][q]); } }
        A.tabulated_ehd=&Tab[0]; } }
  const int NQ=11;
  std::vector<double> Scat((size_t)NS*NS*NQ), Ext((size_t)NS*NQ);
  if (scuffcuda_emtpft(ctx,hv,&A,&Scat[0],&Ext[0])) ErrExit("%s",scuffcuda_last_error(ctx));
  // Q_full = Q_extinction - Q_scattered, except the scattered power; the two torque terms add (EMTPFT.cc:826-856)
  PFTMatrix->Zero();
  for(int nsa=0; nsa<NS; nsa++)
   for(int nq=0; nq<NUMPFT; nq++)
    { double v=Ext[nsa*NQ+nq];
      for(int nsb=0; nsb<NS; nsb++)
       { const double *s=&Scat[((size_t)nsa*NS+nsb)*NQ];
         if (nq==PFT_PABS) v-=s[PFT_PSCAT]; else if (nq==PFT_PSCAT) v+=s[PFT_PSCAT]; else v-=s[nq]; }
      if (PFT_XTORQUE<=nq && nq<=PFT_ZTORQUE)
       { v+=Ext[nsa*NQ+nq+3];
         for(int nsb=0; nsb<NS; nsb++) v-=Scat[((size_t)nsa*NS+nsb)*NQ+nq+3]; }
      PFTMatrix->SetEntry(nsa,nq,v); }
  if (Interior) for(int ns=0; ns<NS; ns++) PFTMatrix->SetEntry(ns,PFT_PSCAT,0.0);
  PFTOnDevice++;
  return true;
}

// PreloadCache / StoreCache twins for the device-side FIPPI table (same file format, FIPPICache.cc:209-400)
void PreloadCacheCUDA(const char *FileName)
{ if (scuffcuda_fippi_preload(CudaContext(),FileName)) Warn("%s",scuffcuda_last_error(CudaContext())); }
void StoreCacheCUDA(const char *FileName)
{ if (scuffcuda_fippi_store(CudaContext(),FileName)) Warn("%s",scuffcuda_last_error(CudaContext())); }

} // namespace scuff
