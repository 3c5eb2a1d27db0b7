/* libscuff_b200.h — host-side C++ mirror of the reference's API for the BEM-matrix assembly path.
 *
 * Same class and member names, argument meaning and error behaviour as the reference
 * (HomerReid/scuff-em): libs/libscuff/libscuff.h:100-146 (RWGPanel, RWGEdge), :174-328
 * (RWGSurface), :348-634 (RWGGeometry; AssembleBEMMatrix :363-365, AssembleBEMMatrixBlock
 * :500-508), libs/libhmat/libhmat.h:165-383 (HMatrix), libs/libMatProp/libMatProp.h (MatProp),
 * PreloadCache/StoreCache libscuff.h:662-663.  The bodies marshal to the C ABI of
 * include/libscuff_cuda.h; nothing is computed on the CPU.
 *
 * Only what the path and its callers touch is mirrored; RHS vectors, fields,
 * PFT, visualisation are out of scope.
 */
#ifndef LIBSCUFF_B200_H
#define LIBSCUFF_B200_H

#include <complex>
#include <string>
#include <vector>
#include <cstdio>
#include "libscuff_cuda.h"

typedef std::complex<double> cdouble;

#define LHM_REAL    0
#define LHM_COMPLEX 1
#define LHM_NORMAL    0
#define LHM_SYMMETRIC 1   /* packed storage: not supported by the device path (ErrExit) */

#define SCUFF_NOLOGGING      0
#define SCUFF_TERSELOGGING   1
#define SCUFF_VERBOSELOGGING 2
#define SCUFF_VERBOSE2       3

#define MAXLDIM 2

namespace scuff {

/* error convention of the reference: ErrExit() logs and exits, Warn() continues
 * (libs/libhrutil/libhrutil.h:136-137).  Tests may install a handler that throws instead. */
void ErrExit(const char *fmt, ...);
void Warn(const char *fmt, ...);
void Log(const char *fmt, ...);
void SetLogFileName(const char *name);
typedef void (*ErrExitHandler)(const char *msg);
void SetErrExitHandler(ErrExitHandler h);

/* the device context of the calling HOST THREAD (created lazily on device SCUFF_CUDA_DEVICE, default 0, or LOCAL_RANK when
 * launched by torchrun; SetCudaDevice picks another one).  One context per thread: see SweepDriver below.                  */
scuffcuda_context *GetCudaContext();
void SetCudaDevice(int device);
void ReleaseCudaContext();      /* destroys the calling thread's context (worker threads call it before they end) */

/*--------------------------------------------------------------------------------------*/
class HMatrix
 {
  public:
   HMatrix(int NRows, int NCols, int RealComplex = LHM_REAL, int StorageType = LHM_NORMAL, void *data = NULL);
   ~HMatrix();

   cdouble GetEntry(size_t nr, size_t nc);
   double GetEntryD(size_t nr, size_t nc);
   void SetEntry(size_t nr, size_t nc, double Entry);
   void SetEntry(size_t nr, size_t nc, cdouble Entry);
   void AddEntry(size_t nr, size_t nc, double Entry);
   void AddEntry(size_t nr, size_t nc, cdouble Entry);   /* real matrix keeps real() (HMatrix.cc:429-430) */
   void Zero();
   void ZeroBlock(int RowOffset, int NumRows, int ColOffset, int NumCols);
   void InsertBlock(HMatrix *B, int RowOffset, int ColOffset, int NRB, int NCB, int BRowOffset, int BColOffset);

   /* downstream consumer, on the device (cuSOLVER getrf/getrs); LBWrappers.cc:150-230 */
   int LUFactorize();
   int LUSolve(HMatrix *X);

   /* device residency: the assembled matrix lives in HBM (DeviceHandle); host DM/ZM is
    * refreshed by SyncToHost() (AssembleBEMMatrix does it before returning)                */
   int  DeviceHandle;         /* 0 = none                                                     */
   bool HostStale;            /* device copy is newer than DM/ZM                              */
   bool DeviceStale;          /* DM/ZM is newer than the device copy (host-side edits)        */
   void SyncToHost();
   void SyncToDevice();
   void EnsureDevice();
   void PrepareDevice();      /* make the device copy current (uploads if it is missing or stale)            */
   void MarkHostDirty();      /* call after writing DM/ZM directly (SetEntry/AddEntry/... do it themselves)  */

   int NR, NC;
   int RealComplex;
   int StorageType;
   int *ipiv;
   double *DM;
   cdouble *ZM;
   bool ownsM;
   char *ErrMsg;
   bool Pinned;               /* DM/ZM is page-locked memory from scuffcuda_host_alloc        */
 };

/*--------------------------------------------------------------------------------------*/
class MatProp
 {
  public:
   MatProp(const char *MaterialName);
   void GetEpsMu(cdouble Omega, cdouble *pEps, cdouble *pMu);
   cdouble GetEps(cdouble Omega) { cdouble e, m; GetEpsMu(Omega, &e, &m); return e; }
   cdouble GetMu(cdouble Omega) { cdouble e, m; GetEpsMu(Omega, &e, &m); return m; }
   bool IsPEC() { return Type == 0; }
   void Zero() { Zeroed = 1; }
   void UnZero() { Zeroed = 0; }
   static void SetLengthUnit(double NewLengthUnit) { FreqUnit = 299792458.0 / NewLengthUnit; }
   static double FreqUnit;

   char *Name;
   int Type;        /* 0 PEC, 1 VACUUM, 2 CONSTANT, 4 PARSED (MP_INTERP data files unsupported) */
   int Zeroed;
   cdouble Eps, Mu;
   void *EpsExpr, *MuExpr;   /* parsed user expressions */
   char *ErrMsg;
 };
/* MATERIAL ... ENDMATERIAL section of a .scuffgeo file (libMatProp/UserDefinedMaterials.cc) */
char *AddMaterialToMatPropDataBase(FILE *f, const char *FileName, const char *MaterialName, int *LineNum);

/*--------------------------------------------------------------------------------------*/
typedef struct RWGPanel
 { int VI[3];
   int EI[3];
   double Centroid[3];
   double ZHat[3];
   bool ZHatFlipped;
   double Radius;
   double Area;
   int Index;
 } RWGPanel;

typedef struct RWGEdge
 { int iV1, iV2, iQP, iQM;
   double Centroid[3];
   double Length;
   double Radius;
   int iPPanel, iMPanel;
   int PIndex, MIndex;
   int Index;
   RWGEdge *Next;
 } RWGEdge;

/* EdgeEdgeInteractions.cc:241-270 (libscuff.h: CreateGammaMatrix) */
void CreateGammaMatrix(double *TorqueAxis, double *GammaMatrix);

class RWGGeometry;

class RWGSurface
 {
  public:
   RWGSurface(FILE *f, const char *Label, int *LineNum, char *Keyword);
   RWGSurface(const char *MeshFile, int MeshTag = -1);
   ~RWGSurface();

   void Transform(const char *format, ...);      /* "DISPLACED x y z" / "ROTATED t ABOUT x y z" (RWGSurface.cc:629-640) */
   void UnTransform();
   bool Contains(const double X[3]);
   bool Contains(RWGSurface *S);

   int NumVertices, NumPanels, NumEdges, NumHalfRWGEdges, NumExteriorEdges, NumBFs;
   double *Vertices;
   RWGPanel **Panels;
   RWGEdge **Edges;
   RWGEdge **HalfRWGEdges;
   int IsClosed;
   bool IsPEC, IsObject;
   int Index;
   int RegionIndices[2];
   char *RegionLabels[2];
   char *MaterialName;
   char *MeshFileName;
   int MeshTag;
   char *Label;
   char *ErrMsg;
   double RMax[3], RMin[3];
   double tolVecClose;
   int TotalStraddlers;
   /* PBCSetup.cc:142-302: promote exterior edges that have a lattice-translated partner to full RWG edges */
   void AddStraddlers(HMatrix *LBasis, int NumStraddlers[MAXLDIM]);

   /* internal */
   void *OTGT, *GT;             /* one-time and current rigid transformations */
   void *SurfaceZeta;           /* SURFACE_IMPEDANCE expression in (w,x,y,z) (RWGSurface.cc:188-210), or 0 */
   double GetOverlap(int neAlpha, int neBeta);   /* \int f_a . f_b  (OPFT.cc:81-226, Overlaps[0]) */
   char *InitRWGSurface(std::vector<double> &VertexCoordinates, std::vector<int> &PanelVertexIndices);
   char *InitEdgeList();
   void UpdatePanels();
 };

/*--------------------------------------------------------------------------------------*/
class RWGGeometry
 {
  public:
   RWGGeometry(const char *GeoFileName, int pLogLevel = SCUFF_NOLOGGING);
   ~RWGGeometry();

   /* the path: libscuff.h:363-365, :500-508 */
   HMatrix *AllocateBEMMatrix(bool PureImagFreq = false, bool Packed = false);
   HMatrix *AssembleBEMMatrix(cdouble Omega, double *kBloch, HMatrix *M = NULL);
   HMatrix *AssembleBEMMatrix(cdouble Omega, HMatrix *M = NULL);
   void AssembleBEMMatrixBlock(int nsa, int nsb, cdouble Omega, double *kBloch,
                               HMatrix *M, HMatrix **GradM = 0,
                               int RowOffset = 0, int ColOffset = 0,
                               void *ABMBCache = 0, bool CacheTranspose = false,
                               int NumTorqueAxes = 0, HMatrix **dMdT = 0, double *GammaMatrix = 0);
   void *CreateABMBAccelerator(int nsa, int nsb, bool PureImagFreq = false, bool NeedZDerivative = false);
   void DestroyABMBAccelerator(void *Accelerator);

   void UpdateCachedEpsMuValues(cdouble Omega);
   int GetRegionByLabel(const char *Label);
   RWGSurface *GetSurfaceByLabel(const char *Label, int *pns = 0);
   /* one line of a .trans file: "TRANS name SURFACE label DISP x y z [SURFACE ...]"
    * (GTransformation.cc:380-470); sets SurfaceMoved[] like RWGGeometry::Transform        */
   void Transform(const char *TransLine);
   void UnTransform();
   /* push (changed) surfaces to the device; called lazily by the assembly routines */
   void SyncDeviceGeometry();

   int NumSurfaces;
   RWGSurface **Surfaces;
   int AllSurfacesClosed;
   int NumRegions;
   char **RegionLabels;
   MatProp **RegionMPs;
   int TotalBFs, TotalEdges, TotalPanels;
   int *BFIndexOffset, *EdgeIndexOffset, *PanelIndexOffset;
   cdouble *EpsTF, *MuTF;
   cdouble StoredOmega;
   int *Mate;
   int *SurfaceMoved;
   int LDim;
   HMatrix *LBasis;
   bool *RegionIsExtended[MAXLDIM];
   int *NumStraddlers[MAXLDIM];   /* NumStraddlers[nd][ns] (PBCSetup.cc:318-333) */
   char *GeoFileName;
   int LogLevel;
   double tolVecClose;

   static bool UseHRWGFunctions, UseHighKTaylorDuffy, UseTaylorDuffyV2P0, DisableCache;

   /* AssembleBEMMatrix without the final device->host copy: the matrix stays in HBM
    * (M->HostStale) for the on-device LU consumer and the sweep drivers                   */
   /* ColBegin/ColEnd (ColEnd > 0): column-block assembly across GPUs — fill only the global columns [ColBegin, ColEnd)
    * of the UPPER triangle (symmetric case) and do not mirror; the caller gathers the column blocks and mirrors once   */
   void PrepareNearPlans();   /* device path: pair plans + FIPPI records of all blocks, no assembly (sweeps that share the table) */
  void AssembleBEMMatrixDevice(cdouble Omega, double *kBloch, HMatrix *M, int ColBegin = 0, int ColEnd = 0);

   /* internal */
   void InitPBCData();
   void AssembleBEMMatrixBlockPBC(int nsa, int nsb, cdouble Omega, double *kBloch, HMatrix *M, int RowOffset, int ColOffset, void *Accelerator = 0);
   /* step 1 of the periodic branch (AssembleBEMMatrix.cc:465-519): inner-cell blocks with Bloch phases, result in HBM */
   void AssembleInnerCellBlocks(int nsa, int nsb, cdouble Omega, double *kBloch, HMatrix *M, int RowOffset=0, int ColOffset=0, void *Accelerator = 0);
   std::vector<char> DeviceDirty;
   unsigned long long DeviceId;
   void AddRegion(const char *RegionLabel, const char *MaterialName, int LineNum);
   bool PointInRegion(int nr, const double X[3]);
   unsigned long long LastPPICounts[5];
 };

void PreloadCache(const char *FileName);
void StoreCache(const char *FileName);


/*--------------------------------------------------------------------------------------*/
/* Sweep sharding from C++: the frequency / imaginary-frequency / Bloch-vector points of a sweep are    */
/* independent, so point j goes to device j mod NumDevices — one host thread and one device context per GPU inside ONE */
/* process, each with its own copy of the geometry and its own matrix; the FIPPI table is built once, the devices      */
/* computing disjoint shares of the records (as sweep.share_fippi_table does across processes).  This is what the      */
/* frequency loops of the reference applications (scuff-scatter.cc:373-472, scuff-cas3D SumsIntegrals.cc) map onto.    */
/*--------------------------------------------------------------------------------------*/
class SweepDriver
 {
  public:
   /* Devices == 0: devices 0 .. NumDevices-1; NumDevices == 0: every visible device.  The same device may be listed more than once */
   SweepDriver(const char *GeoFileName, int NumDevices = 0, const int *Devices = 0);
   ~SweepDriver();

   /* called on the worker thread that owns G and M, after AssembleBEMMatrix(Omega, kBloch) has filled M in HBM (M->HostStale) */
   typedef void (*PointFunction)(RWGGeometry *G, HMatrix *M, cdouble Omega, double *kBloch, int PointIndex, int Worker, void *UserData);

   /* kBlochs: 2 doubles per point, or 0 for compact geometries */
   void Run(int NumPoints, const cdouble *Omegas, const double *kBlochs, PointFunction F, void *UserData, bool ShareFIPPITable = true);

   int NumWorkers;
   std::vector<int> DeviceOf;            /* device of worker w */
   std::vector<RWGGeometry *> Geometries;
   std::vector<HMatrix *> Matrices;      /* created by the workers on first use */
   std::vector<double> ColdSeconds;      /* per worker: plans + its share of the FIPPI table */
 };

}  /* namespace scuff */

#endif
