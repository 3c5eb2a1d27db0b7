/* TEST INFRASTRUCTURE ONLY — driver around the UNMODIFIED reference libscuff (compiled from
 * /root/reference into oracle/_ref/libscuff_ref.so by oracle/Makefile).  It only calls the
 * reference's public/internal API (libscuff.h:363-365, libscuffInternals.h:67-392) and dumps
 * results as raw little-endian doubles so tests / bench.py can compare the CUDA path with the
 * reference's own CPU output.  Never linked into, imported by or timed as the product.
 *
 *   ref_driver geom   <geo> <out.txt>
 *   ref_driver matrix <geo> <wRe> <wIm> <out.bin> [--kbloch kx ky] [--imag] [--repeat n]
 *                     [--cache-in f] [--cache-out f] [--block nsa nsb] [--transfile f --trans n]
 *                     (--trans: apply the n-th transformation of a .trans file first, as scuff-cas3D does)
 *   ref_driver gradblock <geo> <wRe> <wIm> <nsa> <nsb> <out.bin> [--imag] [--transfile f --trans n] [--torque ax ay az]...
 *        AssembleBEMMatrixBlock (AssembleBEMMatrix.cc:349-391) with GradM[3] and dMdT[NumTorqueAxes] (CreateGammaMatrix,
 *        EdgeEdgeInteractions.cc:241-270): writes the value block, the three gradient blocks and the torque blocks one after
 *        the other (complex, or real with --imag as scuff-cas3D allocates them)
 *   ref_driver scatter <geo> <wRe> <wIm> <out.bin> --pw E0xr E0xi E0yr E0yi E0zr E0zi nx ny nz [--region label] [--ps x y z Pxr Pxi Pyr Pyi Pzr Pzi]
 *                      [--interior] [--method m] [--transfile f --trans n] [--nopft] [--dsi radius npoints] [--fields pts.bin [--rf]] [--kbloch kx ky]
 *        one frequency point of scuff-scatter (scuff-scatter.cc:402-520): AssembleBEMMatrix, AssembleRHSVector with the given
 *        incident-field chain (plane wave first, then an optional point source), LUFactorize, LUSolve, GetPFTMatrix by the
 *        energy/momentum-transfer method with integral method m (default 0 = SCUFF_EMTPFTI_EHDERIVATIVES1).  Writes RHS
 *        (TotalBFs complex), the solution KN (TotalBFs complex) and the PFT matrix (NumSurfaces x 8 doubles, row-major).
 *        --kbloch: periodic geometry (use with --nopft): AssembleBEMMatrix / AssembleRHSVector with the Bloch vector.
 *        --dsi: then the PFT matrix by the displaced-surface-integral method (GetDSIPFT, DSIPFT.cc:449-590; radius 0 = automatic), NS x 8.
 *        --fields: then GetFields (GetFields.cc:336-395) at the points of pts.bin (raw doubles x,y,z): NX x 6 complex, row-major,
 *        appended; --rf: then GetRFMatrix (:141-330), TotalBFs x 6NX complex column-major, appended.
 *   ref_driver dgf <geo> <wRe> <wIm> <pts.bin> <ncols> <out.bin> [--total]
 *        RWGGeometry::GetDyadicGFs (GetDyadicGFs.cc:58-190) at the rows of pts.bin (ncols = 3: one-point, 6: destination + source):
 *        AssembleBEMMatrix, LUFactorize, then NX x 18 complex (GE, GM row-major); --total adds the direct (homogeneous) contribution.
 *   ref_driver ssi    <geo> <wRe> <wIm> <nsa> <nsb> <Lx> <Ly> <Lz> <out.bin> [--cache-out f]
 *        GetSurfaceSurfaceInteractions (SurfaceSurfaceInteractions.cc:426-585) for one surface pair with surface b displaced
 *        by L (the neighbour-cell blocks of the periodic assembly, AssembleBEMMatrix.cc:72-114), not Symmetric
 *   ref_driver pbcinner <geo> <wRe> <wIm> <kx> <ky> <nsa> <nsb> <out.bin> [--cache-out f]
 *        periodic geometry: AssembleBEMMatrixBlock (AssembleBEMMatrix.cc:349-565) MINUS its step 2 (outer cells through the
 *        GBar tables, :530-560, recomputed here with the reference's own CreateRegionGBA + GetSurfaceSurfaceInteractions):
 *        what remains is step 1, the Bloch-phased inner-cell blocks.  Also writes <out.bin>.full (the complete block).
 *   ref_driver overlaps <geo> <ns> <out.txt>
 *        RWGSurface::GetOverlap (OPFT.cc:219-226) for every edge pair neA<=neB with a non-zero overlap: "neA neB value"
 *   ref_driver gbar <kRe> <kIm> <kBx> <kBy> <LDim> <L1x> <L1y> <L2x> <L2y> <exclude> <npts> <seed> <out.bin>
 *        GBarVDEwald (GBarVDEwald.cc:931-996, E=-1: automatic Ewald parameter) at npts seeded random points R in
 *        [-0.5,0.5]^2 x [-1,1]; record = R[3], then GBarVD[8] (value, d/dx, d/dy, d/dz, d2/dxdy, d2/dxdz, d2/dydz, d3/dxdydz)
 *        as (re,im) pairs: 19 doubles.  No geometry file is read.
 *   ref_driver gbarpts <kRe> <kIm> <kBx> <kBy> <L1x> <L1y> <L2x> <L2y> <exclude> <pts.bin> <out.bin>
 *        GBarVDEwald at the points of pts.bin (raw doubles x,y,z), 2-D lattice; out: 16 doubles per point
 *   ref_driver gba <geo> <wRe> <wIm> <kx> <ky> <nr> <nsa> <nsb> <pts.bin> <out.bin>
 *        CreateRegionGBA(nr, Omega, kBloch, nsa, nsb) (GBarAccelerator.cc:1205-1221): prints the table decisions as JSON
 *        {null | RhoMin, RhoMax, LMax, full, N1, N2, N3, X1Min, DX1, ...}, writes <out.bin>.ctable (the tricubic coefficients)
 *        and GetGBar (value + gradient, 8 doubles per point) at the points of pts.bin into <out.bin>
 *   ref_driver ppis   <geo> <kRe> <kIm> <pairs.txt> <out.bin>    (lines: nsa npa iQa nsb npb iQb)
 *   ref_driver fippi  <geo> <pairs.txt> <out.bin>                (lines: nsa npa nsb npb)
 *   ref_driver eeisample <geo> <wRe> <wIm> <npairs> <seed> <out.bin> [--repeat n] [--sweep] [--patch m] [--cache-out f]
 *        (--patch m: after the random pairs, ALL pairs nea<=neb<m — a compact patch of the mesh, mostly near/singular pairs;
 *         --cache-out: StoreCache after the last pass, whose values then all come from the stored records)
 *        GetEdgeEdgeInteractions (EdgeEdgeInteractions.cc:50-217) on a seeded random sample of edge pairs
 *        of surface 0 in the exterior medium (k=Omega): the per-element work of AssembleBEMMatrix on a
 *        bounded sample of a large mesh.  out.bin: records {nea, neb, Re GC0, Im GC0, Re GC1, Im GC1}.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>
#include <libhrutil.h>
#include "libscuff.h"
#include "libscuffInternals.h"
#include "GBarAccelerator.h"
#include <libMDInterp.h>
using namespace scuff;

static double Now()
{ struct timespec ts; clock_gettime(CLOCK_MONOTONIC,&ts); return ts.tv_sec+1e-9*ts.tv_nsec; }

static int DumpGeom(RWGGeometry *G, const char *Out)
{ FILE *f=fopen(Out,"w"); if(!f) return 1;
  fprintf(f,"NumSurfaces %i\nNumRegions %i\nTotalBFs %i\nLDim %i\n",G->NumSurfaces,G->NumRegions,G->TotalBFs,G->LDim);
  for(int nd=0; nd<G->LDim; nd++)
   fprintf(f,"LBV %.17g %.17g %.17g\n",G->LBasis->GetEntryD(0,nd),G->LBasis->GetEntryD(1,nd),G->LBasis->GetEntryD(2,nd));
  for(int nr=0; nr<G->NumRegions; nr++)
   { fprintf(f,"Region %i %s %s",nr,G->RegionLabels[nr],G->RegionMPs[nr]->Name);
     for(int nd=0; nd<G->LDim; nd++) fprintf(f," %i",(int)G->RegionIsExtended[nd][nr]);
     fprintf(f,"\n");
   }
  for(int ns=0; ns<G->NumSurfaces; ns++)
   { RWGSurface *S=G->Surfaces[ns];
     fprintf(f,"Surface %i %s IsPEC %i Regions %i %i BFOffset %i Mate %i NV %i NP %i NE %i NBF %i\n",ns,S->Label,
             (int)S->IsPEC,S->RegionIndices[0],S->RegionIndices[1],G->BFIndexOffset[ns],G->Mate[ns],
             S->NumVertices,S->NumPanels,S->NumEdges,S->NumBFs);
     for(int nv=0; nv<S->NumVertices; nv++)
      fprintf(f,"V %.17g %.17g %.17g\n",S->Vertices[3*nv],S->Vertices[3*nv+1],S->Vertices[3*nv+2]);
     for(int np=0; np<S->NumPanels; np++)
      { RWGPanel *P=S->Panels[np];
        fprintf(f,"P %i %i %i %.17g %.17g %.17g %.17g %.17g\n",P->VI[0],P->VI[1],P->VI[2],
                P->Centroid[0],P->Centroid[1],P->Centroid[2],P->Radius,P->Area);
      }
     for(int ne=0; ne<S->NumEdges; ne++)
      { RWGEdge *E=S->Edges[ne];
        fprintf(f,"E %i %i %i %i %i %i %i %i %.17g\n",E->iV1,E->iV2,E->iQP,E->iQM,E->iPPanel,E->iMPanel,
                E->PIndex,E->MIndex,E->Length);
      }
   }
  fclose(f); return 0;
}

int main(int argc, char **argv)
{ if(argc<3) { fprintf(stderr,"usage: see header\n"); return 2; }
  const char *Cmd=argv[1];
  SetLogFileName("/dev/null");
  if(!strcmp(argv[1],"gbar"))
   { if(argc<15) { fprintf(stderr,"gbar: 13 arguments\n"); return 2; }
     cdouble k(atof(argv[2]),atof(argv[3])); double kBloch[2]={atof(argv[4]),atof(argv[5])}; int LDim=atoi(argv[6]);
     double LBV[2][3]={{atof(argv[7]),atof(argv[8]),0.0},{atof(argv[9]),atof(argv[10]),0.0}};
     bool Exclude=atoi(argv[11])!=0; int npts=atoi(argv[12]); srand48(atol(argv[13]));
     FILE *f=fopen(argv[14],"wb"); if(!f) return 1;
     for(int n=0;n<npts;n++)
      { double R[3]={drand48()-0.5, LDim==2 ? drand48()-0.5 : 0.6*(drand48()-0.5), 2.0*drand48()-1.0};
        cdouble GBarVD[8];
        GBarVDEwald(R,k,kBloch,LBV,LDim,-1.0,Exclude,GBarVD);
        fwrite(R,sizeof(double),3,f); fwrite(GBarVD,sizeof(cdouble),8,f); }
     fclose(f); return 0;
   }
  if(!strcmp(argv[1],"gbarpts"))
   { if(argc<13) { fprintf(stderr,"gbarpts: 11 arguments\n"); return 2; }
     cdouble k(atof(argv[2]),atof(argv[3])); double kBloch[2]={atof(argv[4]),atof(argv[5])};
     double LBV[2][3]={{atof(argv[6]),atof(argv[7]),0.0},{atof(argv[8]),atof(argv[9]),0.0}};
     bool Exclude=atoi(argv[10])!=0;
     FILE *fi=fopen(argv[11],"rb"), *fo=fopen(argv[12],"wb"); if(!fi||!fo) return 1;
     double R[3];
     while(fread(R,sizeof(double),3,fi)==3)
      { cdouble GBarVD[8];
        GBarVDEwald(R,k,kBloch,LBV,2,-1.0,Exclude,GBarVD);
        fwrite(GBarVD,sizeof(cdouble),8,fo); }
     fclose(fi); fclose(fo); return 0;
   }
  RWGGeometry *G=new RWGGeometry(argv[2]);

  if(!strcmp(Cmd,"geom")) return DumpGeom(G,argv[3]);

  if(!strcmp(Cmd,"matrix"))
   { cdouble Omega(atof(argv[3]),atof(argv[4])); const char *Out=argv[5];
     double kBloch[2]; bool HavekBloch=false, Imag=false; int Repeat=1;
     const char *CacheIn=0,*CacheOut=0,*TransFile=0; int nsa=-1,nsb=-1,nTrans=-1;
     for(int n=6;n<argc;n++)
      { if(!strcmp(argv[n],"--kbloch")) { kBloch[0]=atof(argv[n+1]); kBloch[1]=atof(argv[n+2]); HavekBloch=true; n+=2; }
        else if(!strcmp(argv[n],"--imag")) Imag=true;
        else if(!strcmp(argv[n],"--repeat")) Repeat=atoi(argv[++n]);
        else if(!strcmp(argv[n],"--cache-in")) CacheIn=argv[++n];
        else if(!strcmp(argv[n],"--cache-out")) CacheOut=argv[++n];
        else if(!strcmp(argv[n],"--block")) { nsa=atoi(argv[n+1]); nsb=atoi(argv[n+2]); n+=2; }
        else if(!strcmp(argv[n],"--transfile")) TransFile=argv[++n];
        else if(!strcmp(argv[n],"--trans")) nTrans=atoi(argv[++n]);
      }
     if(TransFile && nTrans>=0)
      { GTCList GTCs=ReadTransFile(TransFile);
        if(nTrans>=(int)GTCs.size()) { fprintf(stderr,"no transformation %i in %s\n",nTrans,TransFile); return 2; }
        G->Transform(GTCs[nTrans]);
      }
     if(CacheIn) PreloadCache(CacheIn);
     HMatrix *M;
     if(nsa>=0) M=new HMatrix(G->Surfaces[nsa]->NumBFs,G->Surfaces[nsb]->NumBFs,Imag?LHM_REAL:LHM_COMPLEX);
     else M=G->AllocateBEMMatrix(Imag);
     printf("{\"NR\": %i, \"NC\": %i, \"real\": %i, \"threads\": %i, \"times\": [",M->NR,M->NC,(int)Imag,omp_get_max_threads());
     for(int r=0;r<Repeat;r++)
      { double t0=Now();
        if(nsa>=0) G->AssembleBEMMatrixBlock(nsa,nsb,Omega,HavekBloch?kBloch:0,M);
        else G->AssembleBEMMatrix(Omega,HavekBloch?kBloch:0,M);
        printf("%s%.6f",r?", ":"",Now()-t0);
      }
     printf("], \"eps\": [");
     for(int nr=0;nr<G->NumRegions;nr++)
      printf("%s[%.17g, %.17g]",nr?", ":"",real(G->EpsTF[nr]),imag(G->EpsTF[nr]));
     printf("], \"mu\": [");
     for(int nr=0;nr<G->NumRegions;nr++)
      printf("%s[%.17g, %.17g]",nr?", ":"",real(G->MuTF[nr]),imag(G->MuTF[nr]));
     printf("]}\n");
     if(CacheOut) StoreCache(CacheOut);
     FILE *f=fopen(Out,"wb"); if(!f) return 1;
     size_t N=(size_t)M->NR*M->NC;
     if(Imag) fwrite(M->DM,sizeof(double),N,f); else fwrite(M->ZM,sizeof(cdouble),N,f);
     fclose(f); return 0;
   }

  if(!strcmp(Cmd,"gradblock"))
   { cdouble Omega(atof(argv[3]),atof(argv[4])); int nsa=atoi(argv[5]), nsb=atoi(argv[6]); const char *Out=argv[7];
     bool Imag=false; const char *TransFile=0; int nTrans=0; int NTA=0; double Axes[9];
     for(int n=8;n<argc;n++)
      { if(!strcmp(argv[n],"--imag")) Imag=true;
        else if(!strcmp(argv[n],"--transfile")) TransFile=argv[++n];
        else if(!strcmp(argv[n],"--trans")) nTrans=atoi(argv[++n]);
        else if(!strcmp(argv[n],"--torque") && NTA<3) { Axes[3*NTA]=atof(argv[n+1]); Axes[3*NTA+1]=atof(argv[n+2]); Axes[3*NTA+2]=atof(argv[n+3]); NTA++; n+=3; } }
     if (TransFile)
      { GTCList GTCs=ReadTransFile(TransFile);
        if(nTrans<0||nTrans>=(int)GTCs.size()) { fprintf(stderr,"no transformation %i in %s\n",nTrans,TransFile); return 2; }
        G->Transform(GTCs[nTrans]); }
     int NBFA=G->Surfaces[nsa]->NumBFs, NBFB=G->Surfaces[nsb]->NumBFs, RC=Imag?LHM_REAL:LHM_COMPLEX;
     HMatrix *M=new HMatrix(NBFA,NBFB,RC), *GradM[3], *dMdT[3]={0,0,0};
     for(int m=0;m<3;m++) GradM[m]=new HMatrix(NBFA,NBFB,RC);
     for(int m=0;m<NTA;m++) dMdT[m]=new HMatrix(NBFA,NBFB,RC);
     double Gamma[27];
     for(int m=0;m<NTA;m++) CreateGammaMatrix(Axes+3*m,Gamma+9*m);
     double t0=Now();
     G->AssembleBEMMatrixBlock(nsa,nsb,Omega,0,M,GradM,0,0,0,false,NTA,NTA?dMdT:0,NTA?Gamma:0);
     printf("{\"NR\": %i, \"NC\": %i, \"real\": %i, \"ntorque\": %i, \"threads\": %i, \"times\": [%.6f]}\n",NBFA,NBFB,(int)Imag,NTA,omp_get_max_threads(),Now()-t0);
     FILE *f=fopen(Out,"wb"); if(!f) return 1;
     size_t N=(size_t)NBFA*NBFB;
     HMatrix *All[7]={M,GradM[0],GradM[1],GradM[2],dMdT[0],dMdT[1],dMdT[2]};
     for(int q=0;q<4+NTA;q++) { if(Imag) fwrite(All[q]->DM,sizeof(double),N,f); else fwrite(All[q]->ZM,sizeof(cdouble),N,f); }
     fclose(f); return 0;
   }

  if(!strcmp(Cmd,"scatter"))
   { cdouble Omega(atof(argv[3]),atof(argv[4])); const char *Out=argv[5];
     IncField *IFList=0, *Tail=0; bool Interior=false, NoPFT=false, WantRF=false, WantDSI=false, HaveKB=false; double kBloch[2]={0,0}; double DSIRadius=0.0; int DSIPoints=302; int Method=0; const char *TransFile=0, *PtsFile=0; int nTrans=0;
     for(int n=6;n<argc;n++)
      { IncField *New=0;
        if(!strcmp(argv[n],"--pw") && n+9<argc)
         { cdouble E0[3]; double nHat[3];
           for(int m=0;m<3;m++) { E0[m]=cdouble(atof(argv[n+1+2*m]),atof(argv[n+2+2*m])); nHat[m]=atof(argv[n+7+m]); }
           New=new PlaneWave(E0,nHat); n+=9; }
        else if(!strcmp(argv[n],"--ps") && n+9<argc)
         { double X0[3]; cdouble P[3];
           for(int m=0;m<3;m++) { X0[m]=atof(argv[n+1+m]); P[m]=cdouble(atof(argv[n+4+2*m]),atof(argv[n+5+2*m])); }
           New=new PointSource(X0,P); n+=9; }
        else if(!strcmp(argv[n],"--region") && Tail) Tail->SetRegionLabel(argv[++n]);      // region of the sources of the last field given
        else if(!strcmp(argv[n],"--interior")) Interior=true;
        else if(!strcmp(argv[n],"--nopft")) NoPFT=true;
        else if(!strcmp(argv[n],"--rf")) WantRF=true;
        else if(!strcmp(argv[n],"--kbloch")) { kBloch[0]=atof(argv[n+1]); kBloch[1]=atof(argv[n+2]); HaveKB=true; n+=2; }
        else if(!strcmp(argv[n],"--dsi")) { WantDSI=true; DSIRadius=atof(argv[n+1]); DSIPoints=atoi(argv[n+2]); n+=2; }
        else if(!strcmp(argv[n],"--fields")) PtsFile=argv[++n];
        else if(!strcmp(argv[n],"--method")) Method=atoi(argv[++n]);
        else if(!strcmp(argv[n],"--transfile")) TransFile=argv[++n];
        else if(!strcmp(argv[n],"--trans")) nTrans=atoi(argv[++n]);
        if(New) { if(Tail) Tail->Next=New; else IFList=New; Tail=New; } }
     if(!IFList) { fprintf(stderr,"scatter: no incident field\n"); return 2; }
     if (TransFile)
      { GTCList GTCs=ReadTransFile(TransFile);
        if(nTrans<0||nTrans>=(int)GTCs.size()) { fprintf(stderr,"no transformation %i in %s\n",nTrans,TransFile); return 2; }
        G->Transform(GTCs[nTrans]); }
     double t0=Now();
     HMatrix *M=G->AssembleBEMMatrix(Omega,HaveKB ? kBloch : 0,(HMatrix *)0);
     double t1=Now();
     HVector *KN=G->AssembleRHSVector(Omega,HaveKB ? kBloch : 0,IFList,0), *RHS=new HVector(KN->N,LHM_COMPLEX);
     RHS->Copy(KN);
     double t2=Now();
     M->LUFactorize(); M->LUSolve(KN);
     double t3=Now();
     HMatrix *PFT=new HMatrix(G->NumSurfaces,NUMPFT);
     if(!NoPFT)
      { PFTOptions Opts; InitPFTOptions(&Opts); Opts.PFTMethod=SCUFF_PFT_EMT; Opts.IF=IFList; Opts.Interior=Interior; Opts.EMTPFTIMethod=Method;
        G->GetPFTMatrix(KN,Omega,&Opts,PFT); }
     double t4=Now();
     printf("{\"N\": %i, \"NS\": %i, \"threads\": %i, \"assemble_s\": %.6f, \"rhs_s\": %.6f, \"lu_solve_s\": %.6f, \"pft_s\": %.6f}\n",
            KN->N,G->NumSurfaces,omp_get_max_threads(),t1-t0,t2-t1,t3-t2,t4-t3);
     FILE *f=fopen(Out,"wb"); if(!f) return 1;
     fwrite(RHS->ZV,sizeof(cdouble),RHS->N,f); fwrite(KN->ZV,sizeof(cdouble),KN->N,f);
     for(int ns=0;ns<G->NumSurfaces;ns++) for(int nq=0;nq<NUMPFT;nq++) { double v=PFT->GetEntryD(ns,nq); fwrite(&v,sizeof(double),1,f); }
     if(WantDSI)
      { PFTOptions Opts; InitPFTOptions(&Opts); Opts.PFTMethod=SCUFF_PFT_DSI; Opts.IF=IFList; Opts.DSIRadius=DSIRadius; Opts.DSIPoints=DSIPoints;
        HMatrix *P2=G->GetPFTMatrix(KN,Omega,&Opts,0);
        for(int ns=0;ns<G->NumSurfaces;ns++) for(int nq=0;nq<NUMPFT;nq++) { double v=P2->GetEntryD(ns,nq); fwrite(&v,sizeof(double),1,f); } }
     if(PtsFile)
      { FILE *fp=fopen(PtsFile,"rb"); if(!fp) return 1;
        fseek(fp,0,SEEK_END); long NX=ftell(fp)/(3*sizeof(double)); fseek(fp,0,SEEK_SET);
        HMatrix *XM=new HMatrix(NX,3);
        for(long nx=0;nx<NX;nx++) { double X[3]; if(fread(X,sizeof(double),3,fp)!=3) return 1; for(int m=0;m<3;m++) XM->SetEntry(nx,m,X[m]); }
        fclose(fp);
        HMatrix *FM=G->GetFields(IFList,KN,Omega,XM);
        for(long nx=0;nx<NX;nx++) for(int q=0;q<6;q++) { cdouble v=FM->GetEntry(nx,q); fwrite(&v,sizeof(cdouble),1,f); }
        if(WantRF) { HMatrix *RF=G->GetRFMatrix(Omega,0,XM,0); fwrite(RF->ZM,sizeof(cdouble),(size_t)RF->NR*RF->NC,f); } }
     fclose(f); return 0;
   }

  if(!strcmp(Cmd,"dgf"))
   { cdouble Omega(atof(argv[3]),atof(argv[4])); const char *Pts=argv[5]; int ncols=atoi(argv[6]); const char *Out=argv[7];
     bool Total=false; for(int n=8;n<argc;n++) if(!strcmp(argv[n],"--total")) Total=true;
     FILE *fp=fopen(Pts,"rb"); if(!fp) return 1;
     fseek(fp,0,SEEK_END); long NX=ftell(fp)/(ncols*sizeof(double)); fseek(fp,0,SEEK_SET);
     HMatrix *XM=new HMatrix(NX,ncols);
     for(long nx=0;nx<NX;nx++) for(int c=0;c<ncols;c++) { double v; if(fread(&v,sizeof(double),1,fp)!=1) return 1; XM->SetEntry(nx,c,v); }
     fclose(fp);
     HMatrix *M=G->AssembleBEMMatrix(Omega,(HMatrix *)0);
     M->LUFactorize();
     double t0=Now();
     HMatrix *GM=G->GetDyadicGFs(Omega,0,XM,M,0,!Total);
     printf("{\"NX\": %li, \"threads\": %i, \"dgf_s\": %.6f}\n",NX,omp_get_max_threads(),Now()-t0);
     FILE *f=fopen(Out,"wb"); if(!f) return 1;
     for(long nx=0;nx<NX;nx++) for(int q=0;q<18;q++) { cdouble v=GM->GetEntry(nx,q); fwrite(&v,sizeof(cdouble),1,f); }
     fclose(f); return 0;
   }

  if(!strcmp(Cmd,"ssi"))
   { cdouble Omega(atof(argv[3]),atof(argv[4])); int nsa=atoi(argv[5]), nsb=atoi(argv[6]);
     double L[3]={atof(argv[7]),atof(argv[8]),atof(argv[9])}; const char *Out=argv[10]; const char *CacheOut=0;
     for(int n=11;n<argc;n++) if(!strcmp(argv[n],"--cache-out")) CacheOut=argv[++n];
     G->UpdateCachedEpsMuValues(Omega);
     HMatrix *B=new HMatrix(G->Surfaces[nsa]->NumBFs,G->Surfaces[nsb]->NumBFs,LHM_COMPLEX);
     GetSSIArgStruct MyArgs, *Args=&MyArgs; InitGetSSIArgs(Args);
     Args->G=G; Args->Sa=G->Surfaces[nsa]; Args->Sb=G->Surfaces[nsb]; Args->Omega=Omega;
     Args->Symmetric=false; Args->B=B; Args->Displacement=L;
     double t0=Now();
     GetSurfaceSurfaceInteractions(Args);
     printf("{\"NR\": %i, \"NC\": %i, \"real\": 0, \"threads\": %i, \"times\": [%.6f]}\n",B->NR,B->NC,omp_get_max_threads(),Now()-t0);
     if(CacheOut) StoreCache(CacheOut);
     FILE *f=fopen(Out,"wb"); if(!f) return 1;
     fwrite(B->ZM,sizeof(cdouble),(size_t)B->NR*B->NC,f); fclose(f); return 0;
   }

  if(!strcmp(Cmd,"pbcinner"))
   { cdouble Omega(atof(argv[3]),atof(argv[4])); double kBloch[2]={atof(argv[5]),atof(argv[6])};
     int nsa=atoi(argv[7]), nsb=atoi(argv[8]); const char *Out=argv[9]; const char *CacheOut=0;
     for(int n=10;n<argc;n++) if(!strcmp(argv[n],"--cache-out")) CacheOut=argv[++n];
     if (G->LBasis==0) { fprintf(stderr,"pbcinner: geometry is not periodic\n"); return 1; }
     int NBFA=G->Surfaces[nsa]->NumBFs, NBFB=G->Surfaces[nsb]->NumBFs;
     HMatrix *Full=new HMatrix(NBFA,NBFB,LHM_COMPLEX), *Outer=new HMatrix(NBFA,NBFB,LHM_COMPLEX);
     double t0=Now();
     G->AssembleBEMMatrixBlock(nsa,nsb,Omega,kBloch,Full);
     double t1=Now();
     // step 2 alone, as :530-560 does it
     int CR[2]; double Signs[2];
     int NCR=CountCommonRegions(G->Surfaces[nsa],G->Surfaces[nsb],CR,Signs);
     Outer->Zero();
     if (NCR>0)
      { int nr1=CR[0], nr2=(NCR==2)?CR[1]:-1;
        GetSSIArgStruct MyArgs, *Args=&MyArgs; InitGetSSIArgs(Args);
        Args->G=G; Args->Sa=G->Surfaces[nsa]; Args->Sb=G->Surfaces[nsb]; Args->Omega=Omega;
        Args->Displacement=0; Args->Symmetric=false; Args->Accumulate=true; Args->B=Outer;
        Args->OmitRegion1=false; Args->OmitRegion2=(nr2==-1);
        Args->GBA1=G->CreateRegionGBA(nr1,Omega,kBloch,nsa,nsb);
        if (Args->GBA1==0) Args->OmitRegion1=true;
        if (!Args->OmitRegion2) { Args->GBA2=G->CreateRegionGBA(nr2,Omega,kBloch,nsa,nsb); if (Args->GBA2==0) Args->OmitRegion2=true; }
        else Args->GBA2=0;
        if (!Args->OmitRegion1 || !Args->OmitRegion2) GetSurfaceSurfaceInteractions(Args);
      }
     double nF=0, nO=0;
     for(size_t i=0;i<(size_t)NBFA*NBFB;i++) { nF+=norm(Full->ZM[i]); nO+=norm(Outer->ZM[i]); }
     printf("{\"NR\": %i, \"NC\": %i, \"real\": 0, \"threads\": %i, \"times\": [%.6f], \"norm_full\": %.17g, \"norm_outer\": %.17g}\n",
            NBFA,NBFB,omp_get_max_threads(),t1-t0,sqrt(nF),sqrt(nO));
     if(CacheOut) StoreCache(CacheOut);
     char fn[2000]; snprintf(fn,2000,"%s.full",Out);
     FILE *f=fopen(fn,"wb"); if(!f) return 1; fwrite(Full->ZM,sizeof(cdouble),(size_t)NBFA*NBFB,f); fclose(f);
     for(size_t i=0;i<(size_t)NBFA*NBFB;i++) Full->ZM[i]-=Outer->ZM[i];
     f=fopen(Out,"wb"); if(!f) return 1; fwrite(Full->ZM,sizeof(cdouble),(size_t)NBFA*NBFB,f); fclose(f); return 0;
   }

  if(!strcmp(Cmd,"gba"))
   { cdouble Omega(atof(argv[3]),atof(argv[4])); double kBloch[2]={atof(argv[5]),atof(argv[6])};
     int nr=atoi(argv[7]), nsa=atoi(argv[8]), nsb=atoi(argv[9]); const char *Pts=argv[10], *Out=argv[11];
     G->UpdateCachedEpsMuValues(Omega);
     GBarAccelerator *GBA=G->CreateRegionGBA(nr,Omega,kBloch,nsa,nsb);
     if(!GBA) { printf("{\"null\": 1}\n"); return 0; }
     printf("{\"null\": 0, \"kRe\": %.17g, \"kIm\": %.17g, \"RhoMin\": %.17g, \"RhoMax\": %.17g, \"LMax\": %.17g, \"full\": %i, \"LDim\": %i",
            real(GBA->k),imag(GBA->k),GBA->RhoMin,GBA->RhoMax,GBA->LMax,(int)GBA->ForceFullEwald,GBA->LDim);
     if(GBA->I3D)
      { Interp3D *I=GBA->I3D;
        printf(", \"N\": [%i, %i, %i], \"XMin\": [%.17g, %.17g, %.17g], \"DX\": [%.17g, %.17g, %.17g]",
               I->N1,I->N2,I->N3,I->X1Min,I->X2Min,I->X3Min,I->DX1,I->DX2,I->DX3);
        char fn[2000]; snprintf(fn,2000,"%s.ctable",Out);
        FILE *f=fopen(fn,"wb"); if(!f) return 1;
        fwrite(I->CTable,sizeof(double),(size_t)(I->N1-1)*(I->N2-1)*(I->N3-1)*2*64,f); fclose(f);
      }
     printf("}\n");
     FILE *fi=fopen(Pts,"rb"), *fo=fopen(Out,"wb"); if(!fi||!fo) return 1;
     double R[3];
     while(fread(R,sizeof(double),3,fi)==3)
      { cdouble GdG[4];
        GdG[0]=GetGBar(R,GBA,GdG+1,0,false);
        fwrite(GdG,sizeof(cdouble),4,fo); }
     fclose(fi); fclose(fo); return 0;
   }

  if(!strcmp(Cmd,"overlaps"))
   { int ns=atoi(argv[3]); RWGSurface *S=G->Surfaces[ns];
     FILE *f=fopen(argv[4],"w"); if(!f) return 1;
     for(int a=0;a<S->NumEdges;a++)
      for(int b=a;b<S->NumEdges;b++)
       { double O=S->GetOverlap(a,b); if (O!=0.0) fprintf(f,"%i %i %.17g\n",a,b,O); }
     fclose(f); return 0;
   }

  if(!strcmp(Cmd,"ppis"))
   { cdouble k(atof(argv[3]),atof(argv[4]));
     FILE *fi=fopen(argv[5],"r"), *fo=fopen(argv[6],"wb"); if(!fi||!fo) return 1;
     int nsa,npa,iQa,nsb,npb,iQb; int N=0;
     while(fscanf(fi,"%i %i %i %i %i %i",&nsa,&npa,&iQa,&nsb,&npb,&iQb)==6)
      { GetPPIArgStruct Args; InitGetPPIArgs(&Args);
        Args.Sa=G->Surfaces[nsa]; Args.Sb=G->Surfaces[nsb];
        Args.npa=npa; Args.npb=npb; Args.iQa=iQa; Args.iQb=iQb; Args.k=k;
        GetPanelPanelInteractions(&Args);
        double Rec[5]={real(Args.H[0]),imag(Args.H[0]),real(Args.H[1]),imag(Args.H[1]),(double)Args.WhichAlgorithm};
        fwrite(Rec,sizeof(double),5,fo); N++;
      }
     fclose(fi); fclose(fo); printf("{\"n\": %i}\n",N); return 0;
   }

  if(!strcmp(Cmd,"fippi"))
   { FILE *fi=fopen(argv[3],"r"), *fo=fopen(argv[4],"wb"); if(!fi||!fo) return 1;
     int nsa,npa,nsb,npb; int N=0;
     while(fscanf(fi,"%i %i %i %i",&nsa,&npa,&nsb,&npb)==4)
      { double *Va[3],*Vb[3],*OVa[3],*OVb[3]; double rRel;
        int ncv=AssessPanelPair(G->Surfaces[nsa],npa,G->Surfaces[nsb],npb,&rRel,Va,Vb);
        int Flipped=CanonicallyOrderVertices(Va,Vb,ncv,OVa,OVb);
        QIFIPPIData QIFD; memset(&QIFD,0,sizeof(QIFD));
        ComputeQIFIPPIData(OVa,OVb,ncv,&QIFD);
        double Hdr[3]={(double)ncv,rRel,(double)Flipped};
        fwrite(Hdr,sizeof(double),3,fo);
        for(int i=0;i<3;i++) fwrite(OVa[i],sizeof(double),3,fo);
        for(int i=0;i<3;i++) fwrite(OVb[i],sizeof(double),3,fo);
        fwrite(&QIFD,sizeof(double),36,fo); N++;
      }
     fclose(fi); fclose(fo); printf("{\"n\": %i}\n",N); return 0;
   }

  if(!strcmp(Cmd,"eeisample"))
   { cdouble Omega(atof(argv[3]),atof(argv[4])); long NPairs=atol(argv[5]); unsigned long long Seed=strtoull(argv[6],0,10);
     const char *Out=argv[7]; int Repeat=1; long Patch=0; const char *CacheOut=0;
     bool Sweep=false;   // --sweep: pass r runs at Omega*(1+0.01 r), the frequency points of a sweep (FIPPI cache stays warm)
     for(int n=8;n<argc;n++)
      { if(!strcmp(argv[n],"--repeat")) Repeat=atoi(argv[++n]);
        else if(!strcmp(argv[n],"--sweep")) Sweep=true;
        else if(!strcmp(argv[n],"--patch")) Patch=atol(argv[++n]);
        else if(!strcmp(argv[n],"--cache-out")) CacheOut=argv[++n]; }
     RWGSurface *S=G->Surfaces[0]; int NE=S->NumEdges;
     if(Patch>NE) Patch=NE;
     const long NRandom=NPairs; NPairs+=Patch*(Patch+1)/2;
     int *ea=(int *)malloc(NPairs*sizeof(int)), *eb=(int *)malloc(NPairs*sizeof(int));
     double *Rec=(double *)malloc(NPairs*6*sizeof(double));
     unsigned long long x=Seed*2862933555777941757ULL+3037000493ULL;   // 64-bit LCG (Knuth MMIX constants below)
     for(long n=0;n<NRandom;n++)
      { x=x*6364136223846793005ULL+1442695040888963407ULL; ea[n]=(int)((x>>33)%(unsigned long long)NE);
        x=x*6364136223846793005ULL+1442695040888963407ULL; eb[n]=(int)((x>>33)%(unsigned long long)NE);
        // the reference's assembly evaluates a pair in the order nea<=neb and mirrors (SurfaceSurfaceInteractions.cc:140-150,573-583); its
        // near-singular integrals are not symmetric under exchanging the panels, so the sample must ask for the same order
        if(ea[n]>eb[n]) { int t=ea[n]; ea[n]=eb[n]; eb[n]=t; } }
     { long n=NRandom; for(int a=0;a<Patch;a++) for(int b=a;b<Patch;b++,n++) { ea[n]=a; eb[n]=b; } }
     printf("{\"n\": %li, \"NE\": %i, \"threads\": %i, \"times\": [",NPairs,NE,omp_get_max_threads());
     for(int r=0;r<Repeat;r++)
      { cdouble OmegaR = Sweep ? Omega*(1.0+0.01*r) : Omega;
        double t0=Now();
#pragma omp parallel for schedule(dynamic,256)
        for(long n=0;n<NPairs;n++)
         { GetEEIArgStruct Args; InitGetEEIArgs(&Args);
           Args.Sa=S; Args.Sb=S; Args.nea=ea[n]; Args.neb=eb[n]; Args.k=OmegaR;
           GetEdgeEdgeInteractions(&Args);
           double *R=Rec+6*n; R[0]=ea[n]; R[1]=eb[n];
           R[2]=real(Args.GC[0]); R[3]=imag(Args.GC[0]); R[4]=real(Args.GC[1]); R[5]=imag(Args.GC[1]);
         }
        printf("%s%.6f",r?", ":"",Now()-t0);
      }
     printf("], \"n_random\": %li, \"n_patch\": %li}\n",NRandom,NPairs-NRandom);
     if(CacheOut) StoreCache(CacheOut);
     FILE *f=fopen(Out,"wb"); if(!f) return 1;
     fwrite(Rec,sizeof(double),6*NPairs,f); fclose(f); return 0;
   }

  fprintf(stderr,"unknown command %s\n",Cmd); return 2;
}
