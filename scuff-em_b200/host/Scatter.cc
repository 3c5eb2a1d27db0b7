// One frequency point of a scattering run on the device: HVector, the incident-field classes,
// AssembleRHSVector, LUSolve on a device-resident vector and GetPFT / GetPFTMatrix / GetEMTPFTMatrix — the calls
// scuff-scatter makes after AssembleBEMMatrix (applications/scuff-scatter/scuff-scatter.cc:466-520).  Same names, argument
// meaning and error behaviour as the reference (libs/libhmat/libhmat.h:95-160, libs/libIncField/libIncField.h,
// libs/libscuff/AssembleRHSVector.cc, GetPFT.cc, EMTPFT.cc); the arithmetic runs in csrc/rhs_pft.cu through the C ABI.
// Host work here is O(number of surfaces): which fields act on which surface with which sign, material values, and the
// final combination of the 11 sums per surface pair into the 8 PFT numbers per surface (EMTPFT.cc:826-856).
#include <cstdlib>
#include <vector>
#include <thread>
#include <functional>
#include "host_internal.h"
#include "generated_lebedev.h"

#define ZVAC 376.73031346177
#define II cdouble(0.0, 1.0)

namespace scuff {

// user-supplied incident fields are evaluated on host threads (IncField::GetFields is the caller's code)
static void ParallelFor(long n, const std::function<void(long)> &f)
{ int nt = (int)std::thread::hardware_concurrency(); if (nt < 1) nt = 1; if (nt > 32) nt = 32;
  if (n < 4096 || nt == 1) { for (long i = 0; i < n; i++) f(i); return; }
  std::vector<std::thread> pool;
  for (int t = 0; t < nt; t++) pool.emplace_back([=, &f]() { for (long i = t; i < n; i += nt) f(i); });
  for (auto &th : pool) th.join();
}

// ---- HVector -------------------------------------------------------------------------------------------------------------
HVector::HVector(int pN, int pRealComplex, void *data) : HMatrix(pN, 1, pRealComplex, LHM_NORMAL, data)
{ N = pN; DV = DM; ZV = ZM; }
cdouble HVector::GetEntry(int n) { return HMatrix::GetEntry((size_t)n, 0); }
double HVector::GetEntryD(int n) { return HMatrix::GetEntryD((size_t)n, 0); }
void HVector::SetEntry(int n, cdouble Entry) { HMatrix::SetEntry((size_t)n, 0, Entry); }
void HVector::SetEntry(int n, double Entry) { HMatrix::SetEntry((size_t)n, 0, Entry); }
void HVector::AddEntry(int n, cdouble Entry) { HMatrix::AddEntry((size_t)n, 0, Entry); }
void HVector::AddEntry(int n, double Entry) { HMatrix::AddEntry((size_t)n, 0, Entry); }
void HVector::Copy(HVector *V)
{ if (V->N != N || V->RealComplex != RealComplex) ErrExit("HVector::Copy: size/type mismatch");
  if (V->DeviceHandle && V->HostStale)                 // the source lives in HBM: copy there
   { scuffcuda_context *ctx = GetCudaContext();
     EnsureDevice();
     if (scuffcuda_matrix_copy(ctx, DeviceHandle, V->DeviceHandle)) ErrExit("%s", scuffcuda_last_error(ctx));
     HostStale = true; DeviceStale = false;
     return; }
  MarkHostDirty();
  if (RealComplex == LHM_REAL) memcpy(DV, V->DV, N * sizeof(double)); else memcpy(ZV, V->ZV, N * sizeof(cdouble));
}

int HMatrix::LUSolve(HVector *X)
{ scuffcuda_context *ctx = GetCudaContext();
  if (!DeviceHandle) ErrExit("LUSolve: matrix has not been LU-factorized");
  if (DeviceStale) ErrExit("LUSolve: the matrix was modified on the host after LUFactorize()");
  if (X->NR != NR || X->RealComplex != RealComplex) ErrExit("LUSolve: dimension/type mismatch");
  X->PrepareDevice();
  if (scuffcuda_lu_solve_device(ctx, DeviceHandle, X->DeviceHandle)) ErrExit("%s", scuffcuda_last_error(ctx));
  X->HostStale = true;
  return 0;
}

// ---- incident fields -----------------------------------------------------------------------------------------------------
IncField::IncField()
{ Eps = cdouble(1.0, 0.0); Mu = cdouble(1.0, 0.0); Omega = -1.0; RegionLabel = 0; RegionIndex = 0; Next = 0; }
IncField::~IncField() { if (RegionLabel) free(RegionLabel); }
void DeleteIncFieldChain(IncField *IF)
{ while (IF) { IncField *N = IF->Next; delete IF; IF = N; } }
void IncField::SetFrequency(cdouble pOmega, bool Traverse)
{ Omega = pOmega;
  if (Traverse) for (IncField *I = Next; I; I = I->Next) I->Omega = pOmega; }
void IncField::SetFrequencyAndEpsMu(cdouble pOmega, cdouble pEps, cdouble pMu, bool Traverse)
{ Omega = pOmega; Eps = pEps; Mu = pMu;
  if (Traverse) for (IncField *I = Next; I; I = I->Next) { I->Omega = pOmega; I->Eps = pEps; I->Mu = pMu; } }
void IncField::SetRegionLabel(const char *Label)
{ if (RegionLabel) free(RegionLabel);
  RegionLabel = Label ? strdup(Label) : 0;
  if (!Label) RegionIndex = 0; }
void IncField::GetTotalFields(const double X[3], cdouble EH[6])
{ for (int n = 0; n < 6; n++) EH[n] = 0.0;
  for (IncField *I = this; I; I = I->Next)
   { cdouble P[6]; I->GetFields(X, P); for (int n = 0; n < 6; n++) EH[n] += P[n]; } }
// central differences with the reference's step (IncField.cc:178-198)
void IncField::GetFieldGradients(const double X[3], cdouble dEH[3][6])
{ cdouble EHP[6], EHM[6];
  for (int i = 0; i < 3; i++)
   { double x[3] = { X[0], X[1], X[2] };
     const double Delta = (X[i] == 0.0 ? 1.0e-4 : 1.0e-4 * fabs(X[i]));
     x[i] += Delta; GetFields(x, EHP);
     x[i] -= 2.0 * Delta; GetFields(x, EHM);
     for (int j = 0; j < 6; j++) dEH[i][j] = (EHP[j] - EHM[j]) / (2.0 * Delta); } }

PlaneWave::PlaneWave(const cdouble pE0[3], const double pnHat[3], const char *Label)
{ memcpy(E0, pE0, 3 * sizeof(cdouble)); memcpy(nHat, pnHat, 3 * sizeof(double)); SetRegionLabel(Label); }
void PlaneWave::SetE0(cdouble pE0[3]) { memcpy(E0, pE0, 3 * sizeof(cdouble)); }
void PlaneWave::SetnHat(double pnHat[3]) { memcpy(nHat, pnHat, 3 * sizeof(double)); }
// host evaluation for callers that ask for the field itself (the device evaluates the same formula in rhs_pft.cu)
void PlaneWave::GetFields(const double X[3], cdouble EH[6])
{ const cdouble K = sqrt(Eps * Mu) * Omega, Z = ZVAC * sqrt(Mu / Eps);
  const cdouble ExpFac = exp(II * K * (nHat[0] * X[0] + nHat[1] * X[1] + nHat[2] * X[2]));
  for (int m = 0; m < 3; m++) EH[m] = E0[m] * ExpFac;
  EH[3] = (nHat[1] * EH[2] - nHat[2] * EH[1]) / Z;
  EH[4] = (nHat[2] * EH[0] - nHat[0] * EH[2]) / Z;
  EH[5] = (nHat[0] * EH[1] - nHat[1] * EH[0]) / Z; }
void PlaneWave::GetFieldGradients(const double X[3], cdouble dEH[3][6])
{ cdouble EH[6]; GetFields(X, EH);
  const cdouble K = sqrt(Eps * Mu) * Omega;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 6; j++) dEH[i][j] = II * K * nHat[i] * EH[j]; }

PointSource::PointSource(const double pX0[3], const cdouble pP[3], int pType, const char *Label)
{ memcpy(X0, pX0, 3 * sizeof(double)); memcpy(P, pP, 3 * sizeof(cdouble)); Type = pType; SetRegionLabel(Label); }
bool PointSource::GetSourcePoint(double X[3]) const { memcpy(X, X0, 3 * sizeof(double)); return true; }
// fields of a point dipole in a homogeneous medium (PointSource.cc:118-175); P = dipole moment / eps_0
void PointSource::GetFields(const double X[3], cdouble EH[6])
{ double RHat[3] = { X[0] - X0[0], X[1] - X0[1], X[2] - X0[2] };
  const double R = sqrt(RHat[0] * RHat[0] + RHat[1] * RHat[1] + RHat[2] * RHat[2]);
  for (int m = 0; m < 3; m++) RHat[m] /= R;
  const cdouble PDotR = P[0] * RHat[0] + P[1] * RHat[1] + P[2] * RHat[2];
  const cdouble RCrossP[3] = { RHat[1] * P[2] - RHat[2] * P[1], RHat[2] * P[0] - RHat[0] * P[2], RHat[0] * P[1] - RHat[1] * P[0] };
  const cdouble k = Omega * sqrt(Eps * Mu), ikr = II * k * R, ikr2 = ikr * ikr;
  cdouble ExpFac = k * k * exp(ikr) / (4.0 * M_PI * R);
  const cdouble Z = ZVAC * sqrt(Mu / Eps);
  const cdouble Term1 = 1.0 - 1.0 / ikr + 1.0 / ikr2, Term2 = (-1.0 + 3.0 / ikr - 3.0 / ikr2) * PDotR, Term3 = (1.0 - 1.0 / ikr);
  if (Type == LIF_ELECTRIC_DIPOLE)
   { ExpFac /= Eps;
     for (int m = 0; m < 3; m++) { EH[m] = ExpFac * (Term1 * P[m] + Term2 * RHat[m]); EH[3 + m] = ExpFac * Term3 * RCrossP[m] / Z; } }
  else
   { ExpFac /= Mu;
     for (int m = 0; m < 3; m++) { EH[m] = -1.0 * Z * ExpFac * Term3 * RCrossP[m]; EH[3 + m] = ExpFac * (Term1 * P[m] + Term2 * RHat[m]); } }
}

// ---- right-hand side ---------------------------------------------------------------------------------------------------
int RWGGeometry::GetRegionIndex(const double X[3])
{ if (LBasis) ErrExit("GetRegionIndex: point sources in periodic geometries are not supported on the device path");
  if (!AllSurfacesClosed) ErrExit("GetRegionIndex: geometries with open surfaces need an explicit region label for the incident field");
  for (int i = NumSurfaces - 1; i >= 0; --i) if (Surfaces[i]->Contains(X)) return Surfaces[i]->RegionIndices[1];
  return 0;
}

// AssembleRHSVector.cc:62-118
int RWGGeometry::UpdateIncFields(IncField *IFList, cdouble Omega, double *kBloch)
{ if (IFList == 0) return 0;
  if ((LBasis == 0 && kBloch != 0) || (LBasis != 0 && kBloch == 0)) ErrExit("%s:%i: internal error", __FILE__, __LINE__);
  // periodic geometries: a plane wave is its own Bloch-periodic field (PlaneWave::GetFields ignores the lattice), so the inner products
  // over the unit cell's basis functions — straddlers included, their second panel lies in the neighbour cell — are the same integrals;
  // lattice sums of localised sources (PointSource::GetFields_Periodic) are not on the device path
  if (LBasis)
   for (IncField *IF = IFList; IF; IF = IF->Next)
    if (!dynamic_cast<PlaneWave *>(IF)) ErrExit("AssembleRHSVector: in periodic geometries only plane-wave incident fields are on the device path");
  UpdateCachedEpsMuValues(Omega);
  int NIF = 0;
  for (IncField *IF = IFList; IF; NIF++, IF = IF->Next)
   { double X[3];
     if (!LBasis && IF->GetSourcePoint(X)) IF->RegionIndex = GetRegionIndex(X);
     else if (IF->RegionLabel) IF->RegionIndex = GetRegionByLabel(IF->RegionLabel);
     else IF->RegionIndex = 0;
     if (IF->RegionIndex < 0 || IF->RegionIndex > NumRegions) ErrExit("invalid region index %i", IF->RegionIndex);
     IF->SetFrequencyAndEpsMu(Omega, EpsTF[IF->RegionIndex], MuTF[IF->RegionIndex], false); }
  return NIF;
}

HVector *RWGGeometry::AllocateRHSVector(bool PureImagFreq)
{ return new HVector(TotalBFs, (PureImagFreq && LBasis == 0) ? LHM_REAL : LHM_COMPLEX); }

static void FillPlaneWave(scuffcuda_planewave &W, PlaneWave *PW, double Sign)
{ for (int m = 0; m < 3; m++) { W.e0[2 * m] = real(PW->E0[m]); W.e0[2 * m + 1] = imag(PW->E0[m]); W.nhat[m] = PW->nHat[m]; }
  W.eps[0] = real(PW->Eps); W.eps[1] = imag(PW->Eps); W.mu[0] = real(PW->Mu); W.mu[1] = imag(PW->Mu); W.sign = Sign; }

HVector *RWGGeometry::AssembleRHSVector(cdouble Omega, double *kBloch, IncField *IFList, HVector *RHS)
{ if (RHS == NULL) RHS = AllocateRHSVector();
  if (RHS->N != TotalBFs) ErrExit("AssembleRHSVector: wrong-size RHS vector");
  RHS->Zero();
  UpdateIncFields(IFList, Omega, kBloch);
  SyncDeviceGeometry();
  scuffcuda_context *ctx = GetCudaContext();
  RHS->PrepareDevice();
  const double omega[2] = { real(Omega), imag(Omega) };
  for (int ns = 0; ns < NumSurfaces; ns++)
   { RWGSurface *S = Surfaces[ns];
     if (S->NumEdges == 0) continue;
     // fields with sources in the surface's exterior region enter with a minus sign, interior region with a plus sign
     // (AssembleRHSVector.cc:297-306); plane waves go to the device as parameters, everything else as values at its points
     std::vector<scuffcuda_planewave> PWs; std::vector<std::pair<IncField *, double> > Others;
     for (IncField *IF = IFList; IF; IF = IF->Next)
      { double Sign = 0.0;
        if (S->RegionIndices[0] == IF->RegionIndex) Sign = -1.0;
        else if (S->RegionIndices[1] == IF->RegionIndex) Sign = +1.0;
        else continue;
        PlaneWave *PW = dynamic_cast<PlaneWave *>(IF);
        if (PW && PWs.size() < 8) { scuffcuda_planewave W; FillPlaneWave(W, PW, Sign); PWs.push_back(W); }
        else Others.push_back(std::make_pair(IF, Sign)); }
     if (PWs.empty() && Others.empty()) continue;
     std::vector<double> Tab;
     if (!Others.empty())
      { long npts = 0;
        if (scuffcuda_rhs_points(ctx, ns, 0, &npts)) ErrExit("%s", scuffcuda_last_error(ctx));
        std::vector<double> X(3 * (size_t)npts); Tab.assign(12 * (size_t)npts, 0.0);
        if (scuffcuda_rhs_points(ctx, ns, X.data(), &npts)) ErrExit("%s", scuffcuda_last_error(ctx));
        ParallelFor(npts, [&](long n)
         { for (size_t i = 0; i < Others.size(); i++)
            { cdouble EH[6]; Others[i].first->GetFields(&X[3 * n], EH);
              for (int q = 0; q < 6; q++) { Tab[12 * n + 2 * q] += Others[i].second * real(EH[q]); Tab[12 * n + 2 * q + 1] += Others[i].second * imag(EH[q]); } } }); }
     if (scuffcuda_assemble_rhs(ctx, ns, omega, (int)PWs.size(), PWs.data(), Tab.empty() ? 0 : Tab.data(), RHS->DeviceHandle, BFIndexOffset[ns]))
      ErrExit("%s", scuffcuda_last_error(ctx)); }
  RHS->HostStale = true; RHS->DeviceStale = false;
  if (UseHRWGFunctions) ErrExit("AssembleRHSVector: half-RWG basis functions are not supported on the device path");
  return RHS;
}
HVector *RWGGeometry::AssembleRHSVector(cdouble Omega, IncField *IF, HVector *RHS) { return AssembleRHSVector(Omega, 0, IF, RHS); }

// ---- power, force, torque ------------------------------------------------------------------------------------------------
PFTOptions *InitPFTOptions(PFTOptions *Options)
{ if (Options == 0) Options = (PFTOptions *)calloc(1, sizeof(PFTOptions));
  Options->PFTMethod = SCUFF_PFT_DEFAULT; Options->IF = 0; Options->DRMatrix = 0; Options->kBloch = 0; Options->RHSVector = 0;
  Options->Interior = false; Options->EMTPFTIMethod = SCUFF_EMTPFTI_DEFAULT; Options->Itemize = false; Options->GetRegionPFTs = false;
  Options->DSIMesh = 0; Options->DSIRadius = 0.0; Options->DSIPoints = 302; Options->DSIFarField = false; Options->FluxFileName = 0;
  return Options;
}

#define NUMPFTT 11
HMatrix *GetEMTPFTMatrix(RWGGeometry *G, cdouble Omega, IncField *IF, HVector *KN, HMatrix *DRMatrix, HMatrix *PFTMatrix,
                         bool Interior, int EMTPFTIMethod, bool Itemize)
{ const char *sss = getenv("SCUFF_EMTPFTI_METHOD");
  if (sss) sscanf(sss, "%i", &EMTPFTIMethod);
  const int NS = G->NumSurfaces;
  if (PFTMatrix == 0 || PFTMatrix->NR != NS || PFTMatrix->NC != NUMPFT) ErrExit("invalid PFTMatrix in GetEMTPFT");
  if (EMTPFTIMethod != SCUFF_EMTPFTI_EHDERIVATIVES1)
   ErrExit("GetEMTPFTMatrix: only the cubature integrals (EMTPFTIMethod = SCUFF_EMTPFTI_EHDERIVATIVES1 = 0, or SCUFF_EMTPFTI_METHOD=0) are "
           "available on the device path; method %i (GetGCMatrixElements + FIBBI cache) is not", EMTPFTIMethod);
  if (DRMatrix) ErrExit("GetEMTPFTMatrix: the Rytov-matrix (non-equilibrium) form is not available on the device path");
  if (KN == 0) ErrExit("GetEMTPFTMatrix: no solution vector");
  if (Itemize || (getenv("SCUFF_ITEMIZE_PFT") && getenv("SCUFF_ITEMIZE_PFT")[0] == '1')) Warn("GetEMTPFTMatrix: itemized output files are not written by the device path");
  if (G->LBasis) ErrExit("GetEMTPFTMatrix: periodic geometries are not supported");
  G->SyncDeviceGeometry();
  scuffcuda_context *ctx = GetCudaContext();
  KN->PrepareDevice();
  G->UpdateCachedEpsMuValues(Omega);
  std::vector<int> Regions(2 * NS), Offsets(NS); std::vector<double> Origins(3 * NS), EpsMu(4 * G->NumRegions);
  for (int ns = 0; ns < NS; ns++)
   { Regions[2 * ns] = G->Surfaces[ns]->RegionIndices[0]; Regions[2 * ns + 1] = G->Surfaces[ns]->RegionIndices[1];
     Offsets[ns] = G->BFIndexOffset[ns];
     for (int m = 0; m < 3; m++) Origins[3 * ns + m] = G->Surfaces[ns]->Origin[m]; }
  for (int nr = 0; nr < G->NumRegions; nr++)
   { cdouble Eps, Mu; G->RegionMPs[nr]->GetEpsMu(Omega, &Eps, &Mu);
     EpsMu[4 * nr] = real(Eps); EpsMu[4 * nr + 1] = imag(Eps); EpsMu[4 * nr + 2] = real(Mu); EpsMu[4 * nr + 3] = imag(Mu); }
  scuffcuda_pft_args A; memset(&A, 0, sizeof A);
  A.omega[0] = real(Omega); A.omega[1] = imag(Omega); A.interior = Interior ? 1 : 0; A.num_surfaces = NS;
  A.region_indices = Regions.data(); A.bf_offsets = Offsets.data(); A.origins = Origins.data();
  A.num_regions = G->NumRegions; A.eps_mu = EpsMu.data(); A.if_region = -1;
  scuffcuda_planewave W; std::vector<double> Tab;
  if (IF)
   { A.if_region = IF->RegionIndex;
     PlaneWave *PW = dynamic_cast<PlaneWave *>(IF);
     if (PW) { FillPlaneWave(W, PW, 1.0); A.num_pw = 1; A.pw = &W; }
     else
      { // any other incident field: its own GetFields / GetFieldGradients at the device's cubature points (EMTPFT.cc:389-436)
        size_t total = 0; std::vector<size_t> first(NS);
        for (int ns = 0; ns < NS; ns++) { long n = 0; if (scuffcuda_pft_points(ctx, ns, 0, &n)) ErrExit("%s", scuffcuda_last_error(ctx)); first[ns] = total; total += (size_t)n; }
        std::vector<double> X(3 * total); Tab.assign(48 * total, 0.0);
        for (int ns = 0; ns < NS; ns++) { long n = 0; if (scuffcuda_pft_points(ctx, ns, &X[3 * first[ns]], &n)) ErrExit("%s", scuffcuda_last_error(ctx)); }
        ParallelFor((long)total, [&](long n)
         { cdouble EH[6], dEH[3][6]; IF->GetFields(&X[3 * n], EH); IF->GetFieldGradients(&X[3 * n], dEH);
           double *t = &Tab[48 * n];
           for (int q = 0; q < 6; q++) { t[2 * q] = real(EH[q]); t[2 * q + 1] = imag(EH[q]); }
           for (int m = 0; m < 3; m++) for (int q = 0; q < 6; q++) { t[12 + 12 * m + 2 * q] = real(dEH[m][q]); t[13 + 12 * m + 2 * q] = imag(dEH[m][q]); } });
        A.tabulated_ehd = Tab.data(); } }
  std::vector<double> Scat((size_t)NS * NS * NUMPFTT), Ext((size_t)NS * NUMPFTT);
  if (scuffcuda_emtpft(ctx, KN->DeviceHandle, &A, Scat.data(), Ext.data())) ErrExit("%s", scuffcuda_last_error(ctx));
  // Q_full = Q_extinction - Q_scattered (except the scattered power itself); the two torque terms add (EMTPFT.cc:826-856)
  PFTMatrix->Zero();
  for (int nsa = 0; nsa < NS; nsa++)
   for (int nq = 0; nq < NUMPFT; nq++)
    { double v = Ext[nsa * NUMPFTT + nq];
      for (int nsb = 0; nsb < NS; nsb++)
       { const double *s = &Scat[((size_t)nsa * NS + nsb) * NUMPFTT];
         if (nq == PFT_PABS) v += -1.0 * s[PFT_PSCAT];
         else if (nq == PFT_PSCAT) v += s[PFT_PSCAT];
         else v += -1.0 * s[nq]; }
      if (PFT_XTORQUE <= nq && nq <= PFT_ZTORQUE)
       { v += Ext[nsa * NUMPFTT + nq + 3];
         for (int nsb = 0; nsb < NS; nsb++) v += -1.0 * Scat[((size_t)nsa * NS + nsb) * NUMPFTT + nq + 3]; }
      PFTMatrix->SetEntry(nsa, nq, v); }
  if (Interior) for (int ns = 0; ns < NS; ns++) PFTMatrix->SetEntry(ns, PFT_PSCAT, 0.0);
  return PFTMatrix;
}

// ---- displaced-surface-integral PFT (DSIPFT.cc:261-590) ---------------------------------------------------------------------
// cubature rule over the bounding surface: rows (x, y, z, nx, ny, nz, w) (GetSCRMatrix :261-347)
static std::vector<double> GetSCRMatrix(const char *DSIMesh, double DSIRadius, int DSIPoints, const GTrans *GT1, const GTrans *GT2)
{ std::vector<double> SCR;
  if (DSIMesh)
   { RWGSurface *BS = new RWGSurface(DSIMesh);
     if (BS->ErrMsg) ErrExit("%s", BS->ErrMsg);
     for (int np = 0; np < BS->NumPanels; np++)
      { const double *X0 = BS->Panels[np]->Centroid, *ZHat = BS->Panels[np]->ZHat;
        double Sign = 1.0, XP[3];
        for (int m = 0; m < 3; m++) XP[m] = X0[m] + 0.1 * BS->Panels[np]->Radius * ZHat[m];
        if (BS->Contains(XP)) Sign = -1.0;                  // Sign*ZHat is the outward normal
        for (int m = 0; m < 3; m++) SCR.push_back(X0[m]);
        for (int m = 0; m < 3; m++) SCR.push_back(Sign * ZHat[m]);
        SCR.push_back(BS->Panels[np]->Area); }
     delete BS; }
  else
   { const double *LRule = ScLebedevRule(DSIPoints);
     if (LRule == 0) ErrExit("no Lebedev rule with %i points", DSIPoints);
     const double R = DSIRadius, R2 = R * R;
     for (int np = 0; np < DSIPoints; np++)
      { for (int m = 0; m < 3; m++) SCR.push_back(R * LRule[4 * np + m]);
        for (int m = 0; m < 3; m++) SCR.push_back(LRule[4 * np + m]);
        SCR.push_back(R2 * LRule[4 * np + 3]); } }
  const GTrans *GTs[2] = { GT1, GT2 };
  for (int n = 0; n < 2; n++)
   { if (!GTs[n]) continue;
     for (size_t nr = 0; nr < SCR.size() / 7; nr++)
      { double *X = &SCR[7 * nr], *nHat = X + 3, XP[3];
        for (int m = 0; m < 3; m++) XP[m] = X[m] + nHat[m];
        GTs[n]->Apply(X, 1); GTs[n]->Apply(XP, 1);
        for (int m = 0; m < 3; m++) nHat[m] = XP[m] - X[m]; } }
  return SCR;
}

// the 3x3 matrices sandwiched between the field vectors: Poynting flux, Maxwell stress, torque (GetNMatrices :358-428)
static void GetNMatrices(const double nHat[3], const double X[3], const double XTorque[3], double NMatrix[NUMPFT][3][3])
{ static const double LeviCivita[3][3][3] = { { { 0, 0, 0 }, { 0, 0, 1 }, { 0, -1, 0 } }, { { 0, 0, -1 }, { 0, 0, 0 }, { 1, 0, 0 } }, { { 0, 1, 0 }, { -1, 0, 0 }, { 0, 0, 0 } } };
  memset(NMatrix, 0, NUMPFT * 9 * sizeof(double));
  double (*P)[3] = NMatrix[PFT_PABS];
  P[1][2] = nHat[0]; P[2][1] = -nHat[0]; P[2][0] = nHat[1]; P[0][2] = -nHat[1]; P[0][1] = nHat[2]; P[1][0] = -nHat[2];
  for (int i = 0; i < 3; i++)
   { double (*F)[3] = NMatrix[PFT_XFORCE + i];
     for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++)
       F[a][b] = (a == b ? -nHat[i] : 0.0) + (a == i ? nHat[b] : 0.0) + (b == i ? nHat[a] : 0.0);   // n_a d_ib + n_b d_ia - n_i d_ab
   }
  double D[3], DxN[3];
  for (int m = 0; m < 3; m++) D[m] = X[m] - XTorque[m];
  DxN[0] = D[1] * nHat[2] - D[2] * nHat[1]; DxN[1] = D[2] * nHat[0] - D[0] * nHat[2]; DxN[2] = D[0] * nHat[1] - D[1] * nHat[0];
  for (int i = 0; i < 3; i++)
   for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++)
     { double v = (a == b) ? -DxN[i] : 0.0;
       for (int j = 0; j < 3; j++) v += LeviCivita[i][j][a] * D[j] * nHat[b] + LeviCivita[i][j][b] * D[j] * nHat[a];
       NMatrix[PFT_XTORQUE + i][a][b] = v; }
}
static double HVMVP(const cdouble V1[3], const double M[3][3], const cdouble V2[3])
{ double Sum = 0.0;
  for (int Mu = 0; Mu < 3; Mu++) for (int Nu = 0; Nu < 3; Nu++) Sum += real(conj(V1[Mu]) * M[Mu][Nu] * V2[Nu]);
  return Sum;
}

void GetDSIPFT(RWGGeometry *G, cdouble Omega, double *kBloch, HVector *KN, IncField *IF, double PFT[NUMPFT],
               char *DSIMesh, double DSIRadius, int DSIPoints, bool FarField, char *PlotFileName, void *GT1, void *GT2)
{ (void)FarField;
  if (PlotFileName) Warn("GetDSIPFT: flux plots are not written by the device path");
  std::vector<double> SCR = GetSCRMatrix(DSIMesh, DSIRadius, DSIPoints, (const GTrans *)GT1, (const GTrans *)GT2);
  const int NX = (int)(SCR.size() / 7);
  // all cubature points are assumed to lie in one region: that of the first point (DSIPFT.cc:473-481)
  const int RegionIndex = G->GetRegionIndex(&SCR[0]);
  if (RegionIndex < 0) ErrExit("GetDSIPFT: the bounding surface lies inside a PEC body");
  cdouble EpsRel, MuRel; G->RegionMPs[RegionIndex]->GetEpsMu(Omega, &EpsRel, &MuRel);
  const double TENTHIRDS = 3.33333333333333333333333;
  const double EpsAbs = TENTHIRDS * real(EpsRel) / ZVAC, MuAbs = TENTHIRDS * real(MuRel) * ZVAC;
  double XTorque[3] = { 0.0, 0.0, 0.0 };
  if (GT1) ((const GTrans *)GT1)->Apply(XTorque, 1);
  if (GT2) ((const GTrans *)GT2)->Apply(XTorque, 1);
  HMatrix XMatrix(NX, 3, LHM_REAL);
  for (int nx = 0; nx < NX; nx++) for (int m = 0; m < 3; m++) XMatrix.SetEntry(nx, m, SCR[7 * nx + m]);
  HMatrix *FScat = G->GetFields(0, KN, Omega, kBloch, &XMatrix, 0);          // the O(N NX) part, on the device
  HMatrix *FInc = IF ? G->GetFields(IF, 0, Omega, kBloch, &XMatrix, 0) : 0;
  memset(PFT, 0, NUMPFT * sizeof(double));
  for (int nr = 0; nr < NX; nr++)
   { const double *X = &SCR[7 * nr], *nHat = X + 3, w = SCR[7 * nr + 6];
     double NMatrix[NUMPFT][3][3];
     GetNMatrices(nHat, X, XTorque, NMatrix);
     cdouble ES[3], HS[3], E[3], H[3];
     for (int m = 0; m < 3; m++)
      { ES[m] = FScat->GetEntry(nr, m); HS[m] = FScat->GetEntry(nr, 3 + m);
        E[m] = ES[m] + (FInc ? FInc->GetEntry(nr, m) : cdouble(0.0)); H[m] = HS[m] + (FInc ? FInc->GetEntry(nr, 3 + m) : cdouble(0.0)); }
     PFT[PFT_PABS] += -0.25 * w * (HVMVP(E, NMatrix[PFT_PABS], H) - HVMVP(H, NMatrix[PFT_PABS], E));
     PFT[PFT_PSCAT] += 0.25 * w * (HVMVP(ES, NMatrix[PFT_PABS], HS) - HVMVP(HS, NMatrix[PFT_PABS], ES));
     for (int nq = 2; nq < NUMPFT; nq++)
      PFT[nq] += 0.25 * w * (EpsAbs * HVMVP(E, NMatrix[nq], E) + MuAbs * HVMVP(H, NMatrix[nq], H)); }
  delete FScat; if (FInc) delete FInc;
}

void RWGGeometry::GetPFT(int SurfaceIndex, HVector *KN, cdouble Omega, double PFT[NUMPFT], PFTOptions *Options)
{ PFTOptions Default; if (Options == 0) { Options = &Default; InitPFTOptions(Options); }
  if (SurfaceIndex < 0 || SurfaceIndex >= NumSurfaces) ErrExit("GetPFT: invalid surface index %i", SurfaceIndex);
  if (Options->PFTMethod == SCUFF_PFT_EMT_EXTERIOR || Options->PFTMethod == SCUFF_PFT_EMT_INTERIOR)
   { HMatrix *PFTMatrix = new HMatrix(NumSurfaces, NUMPFT);
     GetEMTPFTMatrix(this, Omega, Options->IF, KN, Options->DRMatrix, PFTMatrix,
                     Options->Interior || Options->PFTMethod == SCUFF_PFT_EMT_INTERIOR, Options->EMTPFTIMethod);
     for (int nq = 0; nq < NUMPFT; nq++) PFT[nq] = PFTMatrix->GetEntryD(SurfaceIndex, nq);
     delete PFTMatrix; }
  else if (Options->PFTMethod == SCUFF_PFT_DSI)
   { if (Options->DRMatrix) ErrExit("GetPFT: the trace (Rytov-matrix) form of the DSI method is not available on the device path");
     RWGSurface *S = Surfaces[SurfaceIndex];
     double DSIRadius = Options->DSIRadius;
     if (DSIRadius == 0.0)                                   // AutodetectDSIRadius (GetPFT.cc:73-86): 1.5 x the extent about the origin
      { double Radius = 0.0;
        for (int np = 0; np < S->NumPanels; np++) Radius = fmax(Radius, VecDistance(S->Origin, S->Panels[np]->Centroid));
        DSIRadius = 1.5 * Radius; }
     GetDSIPFT(this, Omega, Options->kBloch, KN, Options->IF, PFT, Options->DSIMesh, DSIRadius, Options->DSIPoints,
               Options->DSIFarField, Options->FluxFileName, S->OTGT, S->GT); }
  else ErrExit("GetPFT: PFT method %i is not available on the device path (energy/momentum transfer SCUFF_PFT_EMT and displaced surface integral SCUFF_PFT_DSI are)", Options->PFTMethod);
}

HMatrix *RWGGeometry::GetPFTMatrix(HVector *KN, cdouble Omega, PFTOptions *Options, HMatrix *PFTMatrix)
{ PFTOptions Default; if (Options == 0) { Options = &Default; InitPFTOptions(Options); }
  if (PFTMatrix && (PFTMatrix->NR != NumSurfaces || PFTMatrix->NC != NUMPFT))
   { Warn("wrong-size PFTMatrix in GetPFTMatrix (reallocating...)"); delete PFTMatrix; PFTMatrix = 0; }
  if (PFTMatrix == 0) PFTMatrix = new HMatrix(NumSurfaces, NUMPFT);
  if (Options->PFTMethod == SCUFF_PFT_EMT_EXTERIOR || Options->PFTMethod == SCUFF_PFT_EMT_INTERIOR)
   GetEMTPFTMatrix(this, Omega, Options->IF, KN, Options->DRMatrix, PFTMatrix,
                   Options->Interior || Options->PFTMethod == SCUFF_PFT_EMT_INTERIOR, Options->EMTPFTIMethod);
  else
   for (int ns = 0; ns < NumSurfaces; ns++)
    { double PFT[NUMPFT]; GetPFT(ns, KN, Omega, PFT, Options);
      for (int nq = 0; nq < NUMPFT; nq++) PFTMatrix->SetEntry(ns, nq, PFT[nq]); }
  return PFTMatrix;
}

// ---- scattered fields at evaluation points -----------------------------------------------------------
// shared by GetFields (contracted with KN on the device) and GetRFMatrix (entries stored in a device matrix)
static void ReducedFieldsOnDevice(RWGGeometry *G, cdouble Omega, HMatrix *XMatrix, int ColumnOffset, HVector *KN, double *F, HMatrix *RFMatrix)
{ const int NX = XMatrix->NR;
  G->SyncDeviceGeometry();
  scuffcuda_context *ctx = GetCudaContext();
  if (KN) KN->PrepareDevice();
  if (RFMatrix) RFMatrix->PrepareDevice();
  std::vector<int> Region(NX);
  std::vector<double> XAll(3 * (size_t)NX);
  for (int nx = 0; nx < NX; nx++)
   { for (int m = 0; m < 3; m++) XAll[3 * (size_t)nx + m] = XMatrix->GetEntryD(nx, ColumnOffset + m);
     Region[nx] = G->GetRegionIndex(&XAll[3 * (size_t)nx]); }
  for (int nr = 0; nr < G->NumRegions; nr++)
   { std::vector<int> Idx; std::vector<double> X;
     for (int nx = 0; nx < NX; nx++) if (Region[nx] == nr) { Idx.push_back(nx); for (int m = 0; m < 3; m++) X.push_back(XAll[3 * (size_t)nx + m]); }
     if (Idx.empty()) continue;
     cdouble EpsRel, MuRel; G->RegionMPs[nr]->GetEpsMu(Omega, &EpsRel, &MuRel);
     const cdouble ZRel = sqrt(MuRel / EpsRel), k = sqrt(MuRel * EpsRel) * Omega;
     const double kk[2] = { real(k), imag(k) }, zz[2] = { real(ZRel), imag(ZRel) };
     std::vector<double> Fr(KN ? 12 * Idx.size() : 0, 0.0);
     for (int ns = 0; ns < G->NumSurfaces; ns++)
      { RWGSurface *S = G->Surfaces[ns];
        double Sign = 0.0;
        if (S->RegionIndices[0] == nr) Sign = +1.0; else if (S->RegionIndices[1] == nr) Sign = -1.0; else continue;
        long Near = 0;
        if (scuffcuda_get_fields(ctx, ns, kk, zz, Sign, (int)Idx.size(), X.data(), KN ? KN->DeviceHandle : 0, G->BFIndexOffset[ns],
                                 KN ? Fr.data() : 0, RFMatrix ? RFMatrix->DeviceHandle : 0, RFMatrix ? Idx.data() : 0, &Near))
         ErrExit("%s", scuffcuda_last_error(ctx));
        (void)Near; }                        // pairs closer than one basis-function radius: closed-form static integrals on the device
     if (KN) for (size_t i = 0; i < Idx.size(); i++) for (int q = 0; q < 12; q++) F[12 * (size_t)Idx[i] + q] += Fr[12 * i + q]; }
  if (RFMatrix) { RFMatrix->HostStale = true; RFMatrix->DeviceStale = false; }
}

HMatrix *RWGGeometry::GetRFMatrix(cdouble Omega, double *kBloch, HMatrix *XMatrix, HMatrix *RFMatrix, bool MinuskBloch, int ColumnOffset)
{ (void)MinuskBloch;
  if (LBasis || kBloch) ErrExit("GetRFMatrix: periodic geometries are not supported on the device path");
  const int NX = XMatrix->NR;
  if (RFMatrix == 0 || RFMatrix->NR != TotalBFs || RFMatrix->NC != 6 * NX)
   { if (RFMatrix) { Warn("wrong-size RFMatrix passed to GetRFMatrix; reallocating"); delete RFMatrix; }
     RFMatrix = new HMatrix(TotalBFs, 6 * NX, LHM_COMPLEX); }
  RFMatrix->Zero();
  ReducedFieldsOnDevice(this, Omega, XMatrix, ColumnOffset, 0, 0, RFMatrix);
  return RFMatrix;
}

HMatrix *RWGGeometry::GetFields(IncField *IFList, HVector *KN, cdouble Omega, double *kBloch, HMatrix *XMatrix, HMatrix *FMatrix)
{ if (XMatrix == 0 || XMatrix->NC < 3 || XMatrix->NR == 0) ErrExit("wrong-size XMatrix (%ix%i) passed to GetFields", XMatrix ? XMatrix->NR : 0, XMatrix ? XMatrix->NC : 0);
  if (LBasis || kBloch) ErrExit("GetFields: periodic geometries are not supported on the device path");
  const int NX = XMatrix->NR;
  if (FMatrix == 0 || FMatrix->NR != NX || FMatrix->NC != 6)
   { if (FMatrix) { Warn(" ** warning: wrong-size FMatrix passed to GetFields(); reallocating"); delete FMatrix; }
     FMatrix = new HMatrix(NX, 6, LHM_COMPLEX); }
  FMatrix->Zero();
  if (IFList) UpdateIncFields(IFList, Omega, kBloch);
  std::vector<double> F(12 * (size_t)NX, 0.0);
  if (KN) ReducedFieldsOnDevice(this, Omega, XMatrix, 0, KN, F.data(), 0);
  for (int nx = 0; nx < NX; nx++)
   { cdouble EH[6]; for (int q = 0; q < 6; q++) EH[q] = cdouble(F[12 * (size_t)nx + 2 * q], F[12 * (size_t)nx + 2 * q + 1]);
     if (IFList)
      { double X[3] = { XMatrix->GetEntryD(nx, 0), XMatrix->GetEntryD(nx, 1), XMatrix->GetEntryD(nx, 2) };
        const int RegionIndex = GetRegionIndex(X);
        if (RegionIndex != -1)
         for (IncField *IF = IFList; IF; IF = IF->Next)
          if (IF->RegionIndex == RegionIndex) { cdouble P[6]; IF->GetFields(X, P); for (int q = 0; q < 6; q++) EH[q] += P[q]; } }
     for (int q = 0; q < 6; q++) FMatrix->SetEntry(nx, q, EH[q]); }
  return FMatrix;
}
HMatrix *RWGGeometry::GetFields(IncField *IF, HVector *KN, cdouble Omega, HMatrix *XMatrix, HMatrix *FMatrix)
{ return GetFields(IF, KN, Omega, 0, XMatrix, FMatrix); }
void RWGGeometry::GetFields(IncField *IF, HVector *KN, cdouble Omega, double *X, cdouble *EH)
{ HMatrix XMatrix(1, 3, LHM_REAL, LHM_NORMAL, (void *)X), FMatrix(1, 6, LHM_COMPLEX, LHM_NORMAL, (void *)EH);
  GetFields(IF, KN, Omega, 0, &XMatrix, &FMatrix); }

// scattering parts of the dyadic Green's functions (GetDyadicGFs.cc:58-190): reduced-field matrices of the destination and source points
// in HBM, 6 NX right-hand sides solved there, 3x3 contractions over the basis functions on the device
HMatrix *RWGGeometry::GetDyadicGFs(cdouble Omega, double *kBloch, HMatrix *XMatrix, HMatrix *M, HMatrix *GMatrix, bool ScatteringOnly)
{ if (LBasis || kBloch) ErrExit("GetDyadicGFs: periodic geometries are not supported on the device path");
  const int NX = XMatrix->NR;
  if (GMatrix == 0 || GMatrix->NR != NX || GMatrix->NC != 18)
   { if (GMatrix) { Warn("wrong-size GMatrix passed to GetDyadicGFs (reallocating)"); delete GMatrix; }
     GMatrix = new HMatrix(NX, 18, LHM_COMPLEX); }
  const bool TwoPointDGF = (XMatrix->NC >= 6), AddDirectContribution = TwoPointDGF && !ScatteringOnly;
  HMatrix *RFDest = GetRFMatrix(Omega, 0, XMatrix, 0, true), *RFSource = 0;
  if (TwoPointDGF) RFSource = GetRFMatrix(Omega, 0, XMatrix, 0, false, 3);
  else
   { RFSource = new HMatrix(TotalBFs, 6 * NX, LHM_COMPLEX);
     RFSource->EnsureDevice();
     scuffcuda_context *c0 = GetCudaContext();
     if (scuffcuda_matrix_copy(c0, RFSource->DeviceHandle, RFDest->DeviceHandle)) ErrExit("%s", scuffcuda_last_error(c0));
     RFSource->HostStale = true; RFSource->DeviceStale = false; }
  M->LUSolve(RFSource);
  scuffcuda_context *ctx = GetCudaContext();
  std::vector<double> C(36 * (size_t)NX);
  if (scuffcuda_dgf_contract(ctx, RFDest->DeviceHandle, RFSource->DeviceHandle, NX, C.data())) ErrExit("%s", scuffcuda_last_error(ctx));
  delete RFDest; delete RFSource;
  GMatrix->Zero();
  cdouble P0[3] = { 0.0, 0.0, 0.0 }; double X00[3] = { 0.0, 0.0, 0.0 };
  PointSource PS(X00, P0);
  for (int nx = 0; nx < NX; nx++)
   { double XDest[3] = { XMatrix->GetEntryD(nx, 0), XMatrix->GetEntryD(nx, 1), XMatrix->GetEntryD(nx, 2) };
     const int nr = GetRegionIndex(XDest);
     if (nr == -1) continue;
     cdouble Eps, Mu; RegionMPs[nr]->GetEpsMu(Omega, &Eps, &Mu);
     const cdouble k = Omega * sqrt(Eps * Mu), ZRel = sqrt(Mu / Eps);
     const cdouble GEScatNormFac = -1.0 / (II * k * ZVAC * ZVAC * ZRel), GMScatNormFac = ZRel / (II * k);
     cdouble GEDirect[3][3], GMDirect[3][3];
     if (AddDirectContribution)
      { const cdouble GEDirectNormFac = k * k / Eps, GMDirectNormFac = k * k / Mu;
        for (int m = 0; m < 3; m++) PS.X0[m] = XMatrix->GetEntryD(nx, 3 + m);
        UpdateIncFields(&PS, Omega, 0);
        for (int i = 0; i < 3; i++)
         { cdouble EH[6];
           for (int m = 0; m < 3; m++) PS.P[m] = (m == i) ? 1.0 : 0.0;
           PS.Type = LIF_ELECTRIC_DIPOLE; PS.GetFields(XDest, EH);
           for (int j = 0; j < 3; j++) GEDirect[j][i] = EH[j] / GEDirectNormFac;
           PS.Type = LIF_MAGNETIC_DIPOLE; PS.GetFields(XDest, EH);
           for (int j = 0; j < 3; j++) GMDirect[j][i] = EH[3 + j] / GMDirectNormFac; } }
     for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
       { cdouble GE = GEScatNormFac * cdouble(C[36 * (size_t)nx + 2 * (3 * i + j)], C[36 * (size_t)nx + 2 * (3 * i + j) + 1]);
         cdouble GM = GMScatNormFac * cdouble(C[36 * (size_t)nx + 18 + 2 * (3 * i + j)], C[36 * (size_t)nx + 18 + 2 * (3 * i + j) + 1]);
         if (AddDirectContribution) { GE += GEDirect[i][j]; GM += GMDirect[i][j]; }
         GMatrix->SetEntry(nx, 3 * i + j, GE); GMatrix->SetEntry(nx, 9 + 3 * i + j, GM); } }
  return GMatrix;
}

}  // namespace scuff
