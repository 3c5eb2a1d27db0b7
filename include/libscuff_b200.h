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
class HVector;
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
   int LUSolve(HVector *X);        /* the vector's device copy is solved in place (no copy through the host) */

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
/* HVector (libs/libhmat/libhmat.h:95-160): here an N x 1 HMatrix, so that it has the same device twin and host/device    */
/* coherence rules; N, DV, ZV as in the reference.  RHS and solution vectors of a scattering run stay in HBM.             */
class HVector : public HMatrix
 {
  public:
   HVector(int N, int RealComplex = LHM_REAL, void *data = NULL);
   cdouble GetEntry(int n);
   double GetEntryD(int n);
   void SetEntry(int n, cdouble Entry);
   void SetEntry(int n, double Entry);
   void AddEntry(int n, cdouble Entry);
   void AddEntry(int n, double Entry);
   void Copy(HVector *V);          /* device-to-device when V lives in HBM */
   int N;
   double *DV;
   cdouble *ZV;
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

/*--------------------------------------------------------------------------------------*/
/* incident fields (libs/libIncField/libIncField.h:45-160)                                */
/*--------------------------------------------------------------------------------------*/
#define LIF_ELECTRIC_DIPOLE 0
#define LIF_MAGNETIC_DIPOLE 1
class IncField
 {
  public:
   cdouble Omega, Eps, Mu;
   IncField *Next;
   char *RegionLabel;
   int RegionIndex;
   IncField();
   virtual ~IncField();
   void SetFrequency(cdouble Omega, bool Traverse = true);
   void SetFrequencyAndEpsMu(cdouble Omega, cdouble Eps, cdouble Mu, bool Traverse = true);
   void SetRegionLabel(const char *Label = 0);
   virtual bool GetSourcePoint(double X[3]) const { (void)X; return false; }
   virtual void GetFields(const double X[3], cdouble EH[6]) = 0;
   void GetTotalFields(const double X[3], cdouble EH[6]);
   virtual void GetFieldGradients(const double X[3], cdouble dEH[3][6]);   /* finite differences unless overridden */
 };
void DeleteIncFieldChain(IncField *IF);
/* evaluated ON THE DEVICE by AssembleRHSVector / GetPFT; every other IncField subclass is evaluated by its own GetFields  */
/* at the device's cubature points and contracted there                                                                   */
class PlaneWave : public IncField
 {
  public:
   cdouble E0[3];
   double nHat[3];
   PlaneWave(const cdouble E0[3], const double nHat[3], const char *Label = 0);
   void SetE0(cdouble pE0[3]);
   void SetnHat(double pnHat[3]);
   void GetFields(const double X[3], cdouble EH[6]);
   void GetFieldGradients(const double X[3], cdouble dEH[3][6]);
 };
class PointSource : public IncField      /* compact geometries only */
 {
  public:
   double X0[3];
   cdouble P[3];
   int Type;
   PointSource(const double X0[3], const cdouble P[3], int Type = LIF_ELECTRIC_DIPOLE, const char *Label = 0);
   bool GetSourcePoint(double X[3]) const;
   void GetFields(const double X[3], cdouble EH[6]);
 };

/* power / force / torque (libs/libscuff/PFTOptions.h)                                    */
#define PFT_PABS 0
#define PFT_PSCAT 1
#define PFT_XFORCE 2
#define PFT_YFORCE 3
#define PFT_ZFORCE 4
#define PFT_XTORQUE 5
#define PFT_YTORQUE 6
#define PFT_ZTORQUE 7
#define NUMPFT 8
#define SCUFF_PFT_OVERLAP 0
#define SCUFF_PFT_DSI 1
#define SCUFF_PFT_EMT 5
#define SCUFF_PFT_EMT_EXTERIOR 5
#define SCUFF_PFT_EMT_INTERIOR 6
#define SCUFF_PFT_MOMENTS 7
#define SCUFF_PFT_DEFAULT SCUFF_PFT_EMT_EXTERIOR
#define SCUFF_EMTPFTI_EHDERIVATIVES1 0
#define SCUFF_EMTPFTI_EHDERIVATIVES2 1
#define SCUFF_EMTPFTI_DEFAULT SCUFF_EMTPFTI_EHDERIVATIVES2
typedef struct PFTOptions
 { int PFTMethod;
   IncField *IF;
   HMatrix *DRMatrix;
   double *kBloch;
   HVector *RHSVector;
   bool Interior;
   int EMTPFTIMethod;
   bool Itemize;
   bool GetRegionPFTs;
   /* displaced-surface-integral method (SCUFF_PFT_DSI): bounding mesh, or a Lebedev rule of DSIPoints points on a sphere of
    * DSIRadius (0 = 1.5 x the surface's extent about its origin, GetPFT.cc:73-86)                                           */
   char *DSIMesh;
   double DSIRadius;
   int DSIPoints;
   bool DSIFarField;
   char *FluxFileName;
 } PFTOptions;
PFTOptions *InitPFTOptions(PFTOptions *Options = 0);

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
   double Origin[3];            /* torque centre, moves with the surface (RWGSurface.cc:423-429, :609, :657) */
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

   /* scattering run on the device: libscuff.h:367-372, :409-418 */
   HVector *AllocateRHSVector(bool PureImagFreq = false);
   HVector *AssembleRHSVector(cdouble Omega, double *kBloch, IncField *IF, HVector *RHS = NULL);
   HVector *AssembleRHSVector(cdouble Omega, IncField *IF, HVector *RHS = NULL);
   int UpdateIncFields(IncField *IF, cdouble Omega, double *kBloch = 0);
   void GetPFT(int SurfaceIndex, HVector *KN, cdouble Omega, double PFT[NUMPFT], PFTOptions *Options = 0);
   HMatrix *GetPFTMatrix(HVector *KN, cdouble Omega, PFTOptions *Options = 0, HMatrix *PFTMatrix = 0);
   int GetRegionIndex(const double X[3]);
   /* scattered + incident fields at the rows of XMatrix (NX x 3): FMatrix NX x 6 complex (E, H) (libscuff.h:512-520, GetFields.cc:336-417);
    * GetRFMatrix: the reduced-field matrix TotalBFs x 6 NX whose columns contract with KN to the scattered fields (:141-330), incl. the
    * near-field branch (GetReducedFields_Nearby) for points closer than one basis-function radius to a surface.                      */
   HMatrix *GetFields(IncField *IF, HVector *KN, cdouble Omega, double *kBloch, HMatrix *XMatrix, HMatrix *FMatrix = NULL);
   HMatrix *GetFields(IncField *IF, HVector *KN, cdouble Omega, HMatrix *XMatrix, HMatrix *FMatrix = NULL);
   void GetFields(IncField *IF, HVector *KN, cdouble Omega, double *X, cdouble *EH);
   HMatrix *GetRFMatrix(cdouble Omega, double *kBloch, HMatrix *XMatrix, HMatrix *RFMatrix = 0, bool MinuskBloch = false, int ColumnOffset = 0);
   /* scattering parts of the electric and magnetic dyadic Green's functions at the rows of XMatrix (NX x 3: one-point; NX x 6: destination,
    * source) (libscuff.h:390-405, GetDyadicGFs.cc:58-190): GMatrix NX x 18 complex (GE row-major, then GM); M = the LU-factorised BEM matrix   */
   HMatrix *GetDyadicGFs(cdouble Omega, double *kBloch, HMatrix *XMatrix, HMatrix *M, HMatrix *GMatrix = 0, bool ScatteringOnly = true);

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

/* GetDSIPFT (DSIPFT.cc:449-590): Poynting vector and Maxwell stress tensor integrated over a surface that encloses the body; the
 * fields at its cubature points come from GetFields on the device.  GT-transformed like the surface (OTGT, then GT).       */
void GetDSIPFT(RWGGeometry *G, cdouble Omega, double *kBloch, HVector *KN, IncField *IF, double PFT[NUMPFT],
               char *DSIMesh, double DSIRadius, int DSIPoints, bool FarField, char *PlotFileName, void *GT1, void *GT2);

/* EMTPFT.cc:533-905 on the device.  Only the cubature integrals SCUFF_EMTPFTI_EHDERIVATIVES1 (Options->EMTPFTIMethod = 0 or
 * SCUFF_EMTPFTI_METHOD=0 in the environment) are available; the reference's default (EHDERIVATIVES2, which goes through
 * GetGCMatrixElements and the FIBBI cache of the 2018 assembler) is refused with a message.                              */
HMatrix *GetEMTPFTMatrix(RWGGeometry *G, cdouble Omega, IncField *IF, HVector *KNVector, HMatrix *DRMatrix = 0,
                         HMatrix *PFTMatrix = 0, bool Interior = false, int EMTPFTIMethod = SCUFF_EMTPFTI_DEFAULT, bool Itemize = false);

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
