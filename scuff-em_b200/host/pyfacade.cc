// extern "C" facade over the C++ mirror (RWGGeometry / HMatrix) for ctypes callers (tests, bench.py).
// ErrExit() is converted into an error return + message instead of exit().
#include <stdexcept>
#include "host_internal.h"
using namespace scuff;

static std::string g_err;
static void Thrower(const char *msg) { throw std::runtime_error(msg); }
#define GUARD(body, failret) try { SetErrExitHandler(Thrower); body } catch (std::exception &e) { g_err = e.what(); return failret; }

extern "C" {

const char *scuffb200_last_error() { return g_err.c_str(); }
int scuffb200_set_device(int dev) { GUARD(SetCudaDevice(dev); return 0;, 1) }
void *scuffb200_context() { GUARD(return (void *)GetCudaContext();, 0) }

void *scuffb200_geometry_open(const char *geofile) { GUARD(return (void *)new RWGGeometry(geofile);, 0) }
void scuffb200_geometry_close(void *G) { delete (RWGGeometry *)G; }

// out[0..7] = NumSurfaces, TotalBFs, NumRegions, LDim, TotalPanels, TotalEdges, AllSurfacesClosed, 0
int scuffb200_geometry_info(void *g, int *out)
{ RWGGeometry *G = (RWGGeometry *)g;
  out[0] = G->NumSurfaces; out[1] = G->TotalBFs; out[2] = G->NumRegions; out[3] = G->LDim;
  out[4] = G->TotalPanels; out[5] = G->TotalEdges; out[6] = G->AllSurfacesClosed; out[7] = 0; return 0; }
// out[0..9] = NV, NP, NE, NBF, IsPEC, Region0, Region1, BFIndexOffset, Mate, NumHalfRWGEdges
int scuffb200_surface_info(void *g, int ns, int *out)
{ RWGGeometry *G = (RWGGeometry *)g; if (ns < 0 || ns >= G->NumSurfaces) return 1;
  RWGSurface *S = G->Surfaces[ns];
  out[0] = S->NumVertices; out[1] = S->NumPanels; out[2] = S->NumEdges; out[3] = S->NumBFs; out[4] = S->IsPEC;
  out[5] = S->RegionIndices[0]; out[6] = S->RegionIndices[1]; out[7] = G->BFIndexOffset[ns]; out[8] = G->Mate[ns]; out[9] = S->NumHalfRWGEdges; return 0; }
const char *scuffb200_surface_label(void *g, int ns) { return ((RWGGeometry *)g)->Surfaces[ns]->Label; }
const char *scuffb200_region_label(void *g, int nr) { return ((RWGGeometry *)g)->RegionLabels[nr]; }
const char *scuffb200_region_material(void *g, int nr) { return ((RWGGeometry *)g)->RegionMPs[nr]->Name; }
// V[3NV], PVI[3NP], PC[3NP], PR[NP], PA[NP], ET[8NE], EL[NE]
int scuffb200_surface_arrays(void *g, int ns, double *V, int *PVI, double *PC, double *PR, double *PA, int *ET, double *EL)
{ RWGGeometry *G = (RWGGeometry *)g; if (ns < 0 || ns >= G->NumSurfaces) return 1;
  RWGSurface *S = G->Surfaces[ns];
  memcpy(V, S->Vertices, 3 * S->NumVertices * sizeof(double));
  for (int np = 0; np < S->NumPanels; np++)
   { RWGPanel *P = S->Panels[np];
     for (int i = 0; i < 3; i++) { PVI[3 * np + i] = P->VI[i]; PC[3 * np + i] = P->Centroid[i]; }
     PR[np] = P->Radius; PA[np] = P->Area; }
  for (int ne = 0; ne < S->NumEdges; ne++)
   { RWGEdge *E = S->Edges[ne]; int *e = ET + 8 * ne;
     e[0] = E->iV1; e[1] = E->iV2; e[2] = E->iQP; e[3] = E->iQM; e[4] = E->iPPanel; e[5] = E->iMPanel; e[6] = E->PIndex; e[7] = E->MIndex;
     EL[ne] = E->Length; }
  return 0; }
// eps/mu per region at Omega: out[4*nr + {0,1,2,3}] = Re eps, Im eps, Re mu, Im mu
int scuffb200_eps_mu(void *g, double wr, double wi, double *out)
{ RWGGeometry *G = (RWGGeometry *)g;
  GUARD(G->UpdateCachedEpsMuValues(cdouble(wr, wi));
        for (int nr = 0; nr < G->NumRegions; nr++)
         { out[4 * nr] = real(G->EpsTF[nr]); out[4 * nr + 1] = imag(G->EpsTF[nr]); out[4 * nr + 2] = real(G->MuTF[nr]); out[4 * nr + 3] = imag(G->MuTF[nr]); }
        return 0;, 1) }
int scuffb200_transform(void *g, const char *line) { GUARD(((RWGGeometry *)g)->Transform(line); return 0;, 1) }
int scuffb200_untransform(void *g) { GUARD(((RWGGeometry *)g)->UnTransform(); return 0;, 1) }
// marks every surface as moved (what Transform()/UnTransform() do): the next assembly re-sends the panels to the device
int scuffb200_mark_geometry_dirty(void *g) { RWGGeometry *G = (RWGGeometry *)g; G->DeviceDirty.assign(G->NumSurfaces, 1); return 0; }
int scuffb200_zero_region(void *g, int nr, int zero)
{ RWGGeometry *G = (RWGGeometry *)g; if (nr < 0 || nr >= G->NumRegions) return 1;
  if (zero) G->RegionMPs[nr]->Zero(); else G->RegionMPs[nr]->UnZero(); return 0; }

void *scuffb200_matrix_alloc(void *g, int pure_imag) { GUARD(return (void *)((RWGGeometry *)g)->AllocateBEMMatrix(pure_imag != 0);, 0) }
void *scuffb200_matrix_new(int nr, int nc, int is_real) { GUARD(return (void *)new HMatrix(nr, nc, is_real ? LHM_REAL : LHM_COMPLEX);, 0) }
void scuffb200_matrix_free(void *m) { delete (HMatrix *)m; }
void *scuffb200_matrix_data(void *m) { HMatrix *M = (HMatrix *)m; return M->RealComplex == LHM_REAL ? (void *)M->DM : (void *)M->ZM; }
int scuffb200_matrix_shape(void *m, int *out) { HMatrix *M = (HMatrix *)m; out[0] = M->NR; out[1] = M->NC; out[2] = (M->RealComplex == LHM_REAL); return 0; }
int scuffb200_matrix_set_entry(void *m, int r, int c, double re, double im) { GUARD(((HMatrix *)m)->SetEntry(r, c, cdouble(re, im)); return 0;, 1) }
int scuffb200_matrix_add_entry(void *m, int r, int c, double re, double im) { GUARD(((HMatrix *)m)->AddEntry(r, c, cdouble(re, im)); return 0;, 1) }
int scuffb200_matrix_get_entry(void *m, int r, int c, double *out) { GUARD(cdouble v = ((HMatrix *)m)->GetEntry(r, c); out[0] = real(v); out[1] = imag(v); return 0;, 1) }
int scuffb200_matrix_zero(void *m) { GUARD(((HMatrix *)m)->Zero(); return 0;, 1) }
int scuffb200_matrix_zero_block(void *m, int ro, int nr, int co, int nc) { GUARD(((HMatrix *)m)->ZeroBlock(ro, nr, co, nc); return 0;, 1) }
int scuffb200_matrix_insert_block(void *m, void *b, int ro, int co, int nrb, int ncb, int bro, int bco)
{ GUARD(((HMatrix *)m)->InsertBlock((HMatrix *)b, ro, co, nrb, ncb, bro, bco); return 0;, 1) }
int scuffb200_matrix_mark_host_dirty(void *m) { GUARD(((HMatrix *)m)->MarkHostDirty(); return 0;, 1) }
// state[0] = HostStale, state[1] = DeviceStale, state[2] = DeviceHandle != 0
int scuffb200_matrix_state(void *m, int *state) { HMatrix *M = (HMatrix *)m; state[0] = M->HostStale; state[1] = M->DeviceStale; state[2] = M->DeviceHandle != 0; return 0; }
int scuffb200_matrix_sync_host(void *m) { GUARD(((HMatrix *)m)->SyncToHost(); return 0;, 1) }
int scuffb200_matrix_device_handle(void *m) { return ((HMatrix *)m)->DeviceHandle; }

// RWGGeometry::AssembleBEMMatrix(Omega, kBloch, M); to_host=0 leaves the result in HBM only
int scuffb200_assemble(void *g, void *m, double wr, double wi, const double *kBloch, int to_host)
{ RWGGeometry *G = (RWGGeometry *)g; HMatrix *M = (HMatrix *)m;
  GUARD(if (to_host) { HMatrix *R = G->AssembleBEMMatrix(cdouble(wr, wi), (double *)kBloch, M); if (R != M) { delete R; throw std::runtime_error("wrong-size matrix"); } }
        else
         { if (G->LBasis == 0 && kBloch != 0 && (kBloch[0] != 0.0 || kBloch[1] != 0.0)) ErrExit("Bloch wavevector is undefined for compact geometries");
           if (G->LBasis != 0 && kBloch == 0) ErrExit("Bloch wavevector must be specified for PBC geometries");
           G->AssembleBEMMatrixDevice(cdouble(wr, wi), (double *)kBloch, M); }
        return 0;, 1) }
// column-block assembly (multi-GPU): upper-triangle entries of the global columns [c0,c1) only, left in HBM, not mirrored
int scuffb200_assemble_columns(void *g, void *m, double wr, double wi, int c0, int c1)
{ RWGGeometry *G = (RWGGeometry *)g; HMatrix *M = (HMatrix *)m;
  GUARD(if (G->LBasis != 0) ErrExit("column-block assembly is not available for periodic geometries");
        G->AssembleBEMMatrixDevice(cdouble(wr, wi), 0, M, c0, c1); return 0;, 1) }
// mirror the upper triangle of the (gathered) device matrix into the lower one
int scuffb200_matrix_symmetrize(void *m)
{ HMatrix *M = (HMatrix *)m;
  GUARD(M->EnsureDevice();
        if (scuffcuda_matrix_symmetrize(GetCudaContext(), M->DeviceHandle, 0, M->NR)) ErrExit("%s", scuffcuda_last_error(GetCudaContext()));
        M->HostStale = true; return 0;, 1) }
int scuffb200_assemble_block(void *g, void *m, int nsa, int nsb, double wr, double wi, const double *kBloch, int row_off, int col_off)
{ RWGGeometry *G = (RWGGeometry *)g; HMatrix *M = (HMatrix *)m;
  GUARD(G->AssembleBEMMatrixBlock(nsa, nsb, cdouble(wr, wi), (double *)kBloch, M, 0, row_off, col_off); return 0;, 1) }
// value block + derivative blocks: grad[3] / dmdt[3] are HMatrix pointers or null, axes = 3 doubles per torque axis
int scuffb200_assemble_block_grad(void *g, void *m, int nsa, int nsb, double wr, double wi, void **grad, int ntorque, const double *axes, void **dmdt)
{ RWGGeometry *G = (RWGGeometry *)g; HMatrix *M = (HMatrix *)m;
  double Gamma[27];
  for (int n = 0; n < ntorque && n < 3; n++) { double ax[3]; ax[0] = axes[3 * n]; ax[1] = axes[3 * n + 1]; ax[2] = axes[3 * n + 2]; CreateGammaMatrix(ax, Gamma + 9 * n); }
  GUARD(G->AssembleBEMMatrixBlock(nsa, nsb, cdouble(wr, wi), 0, M, (HMatrix **)grad, 0, 0, 0, false, ntorque, (HMatrix **)dmdt, Gamma); return 0;, 1) }
int scuffb200_prepare_plans(void *g) { GUARD(((RWGGeometry *)g)->PrepareNearPlans(); return 0;, 1) }
void *scuffb200_create_abmb(void *g, int nsa, int nsb, int pure_imag) { GUARD(return ((RWGGeometry *)g)->CreateABMBAccelerator(nsa, nsb, pure_imag != 0, false);, 0) }
void scuffb200_destroy_abmb(void *g, void *acc) { ((RWGGeometry *)g)->DestroyABMBAccelerator(acc); }
int scuffb200_assemble_block_acc(void *g, void *m, int nsa, int nsb, double wr, double wi, const double *kBloch, int row_off, int col_off, void *acc)
{ RWGGeometry *G = (RWGGeometry *)g; HMatrix *M = (HMatrix *)m;
  GUARD(G->AssembleBEMMatrixBlock(nsa, nsb, cdouble(wr, wi), (double *)kBloch, M, 0, row_off, col_off, acc); return 0;, 1) }
int scuffb200_assemble_inner_cells(void *g, void *m, int nsa, int nsb, double wr, double wi, const double *kBloch, int row_off, int col_off)
{ RWGGeometry *G = (RWGGeometry *)g; HMatrix *M = (HMatrix *)m;
  GUARD(G->AssembleInnerCellBlocks(nsa, nsb, cdouble(wr, wi), (double *)kBloch, M, row_off, col_off); M->SyncToHost(); return 0;, 1) }
double scuffb200_get_overlap(void *g, int ns, int nea, int neb) { return ((RWGGeometry *)g)->Surfaces[ns]->GetOverlap(nea, neb); }
int scuffb200_surface_zeta(void *g, int ns, double wr, double wi, const double *X, double *out)
{ RWGSurface *S = ((RWGGeometry *)g)->Surfaces[ns];
  if (!S->SurfaceZeta) return 1;
  GUARD(CExpr *Z = (CExpr *)S->SurfaceZeta;
        CExprSetVar(Z, "x", X[0]); CExprSetVar(Z, "y", X[1]); CExprSetVar(Z, "z", X[2]);
        cdouble v = CExprEval(Z, cdouble(wr, wi) * MatProp::FreqUnit); out[0] = real(v); out[1] = imag(v); return 0;, 2) }
int scuffb200_ppi_counts(void *g, unsigned long long *out) { for (int i = 0; i < 5; i++) out[i] = ((RWGGeometry *)g)->LastPPICounts[i]; return 0; }
int scuffb200_lu_factorize(void *m) { GUARD(return ((HMatrix *)m)->LUFactorize();, 1) }
int scuffb200_lu_solve(void *m, void *x) { GUARD(return ((HMatrix *)m)->LUSolve((HMatrix *)x);, 1) }
int scuffb200_preload_cache(const char *f) { GUARD(PreloadCache(f); return 0;, 1) }
int scuffb200_store_cache(const char *f) { GUARD(StoreCache(f); return 0;, 1) }
int scuffb200_clear_cache() { GUARD(if (scuffcuda_fippi_clear(GetCudaContext())) ErrExit("%s", scuffcuda_last_error(GetCudaContext())); return 0;, 1) }
long scuffb200_cache_records() { long n = 0; GUARD(scuffcuda_fippi_num_records(GetCudaContext(), &n); return n;, -1) }
int scuffb200_set_disable_cache(int v) { RWGGeometry::DisableCache = (v != 0); return 0; }

/* ---- scattering run (Scatter.cc): vectors, incident fields, RHS, solve, PFT ---- */
void *scuffb200_vector_new(int n, int is_real) { GUARD(return (void *)(HMatrix *)new HVector(n, is_real ? LHM_REAL : LHM_COMPLEX);, 0) }
void scuffb200_vector_free(void *v) { delete (HVector *)(HMatrix *)v; }
void *scuffb200_planewave_new(const double *e0, const double *nhat, const char *label)
{ cdouble E0[3]; for (int m = 0; m < 3; m++) E0[m] = cdouble(e0[2 * m], e0[2 * m + 1]);
  GUARD(return (void *)(IncField *)new PlaneWave(E0, nhat, (label && label[0]) ? label : 0);, 0) }
void *scuffb200_pointsource_new(const double *x0, const double *p, int type, const char *label)
{ cdouble P[3]; for (int m = 0; m < 3; m++) P[m] = cdouble(p[2 * m], p[2 * m + 1]);
  GUARD(return (void *)(IncField *)new PointSource(x0, P, type, (label && label[0]) ? label : 0);, 0) }
void scuffb200_incfield_set_next(void *a, void *b) { ((IncField *)a)->Next = (IncField *)b; }
void scuffb200_incfield_free(void *a) { delete (IncField *)a; }
int scuffb200_assemble_rhs(void *g, double wr, double wi, void *incfield, void *vec)
{ GUARD(((RWGGeometry *)g)->AssembleRHSVector(cdouble(wr, wi), (IncField *)incfield, (HVector *)(HMatrix *)vec); return 0;, 1) }
int scuffb200_assemble_rhs_kbloch(void *g, double wr, double wi, const double *kBloch, void *incfield, void *vec)
{ GUARD(((RWGGeometry *)g)->AssembleRHSVector(cdouble(wr, wi), (double *)kBloch, (IncField *)incfield, (HVector *)(HMatrix *)vec); return 0;, 1) }
int scuffb200_lu_solve_vector(void *m, void *vec) { GUARD(return ((HMatrix *)m)->LUSolve((HVector *)(HMatrix *)vec);, 1) }
int scuffb200_vector_copy(void *dst, void *src) { GUARD(((HVector *)(HMatrix *)dst)->Copy((HVector *)(HMatrix *)src); return 0;, 1) }
/* out: NumSurfaces x 8, row-major */
int scuffb200_get_pft_matrix(void *g, void *kn, double wr, double wi, void *incfield, int interior, int method, double *out)
{ GUARD(RWGGeometry *G = (RWGGeometry *)g;
        PFTOptions Opts; InitPFTOptions(&Opts); Opts.IF = (IncField *)incfield; Opts.Interior = (interior != 0); Opts.EMTPFTIMethod = method;
        HMatrix *P = G->GetPFTMatrix((HVector *)(HMatrix *)kn, cdouble(wr, wi), &Opts, 0);
        for (int ns = 0; ns < G->NumSurfaces; ns++) for (int nq = 0; nq < NUMPFT; nq++) out[ns * NUMPFT + nq] = P->GetEntryD(ns, nq);
        delete P; return 0;, 1) }
/* X: NX x 3 row-major; out: NX x 6 complex row-major (E, H) */
int scuffb200_get_fields(void *g, void *incfield, void *kn, double wr, double wi, int nx, const double *X, double *out)
{ GUARD(RWGGeometry *G = (RWGGeometry *)g;
        HMatrix XM(nx, 3, LHM_REAL); for (int i = 0; i < nx; i++) for (int m = 0; m < 3; m++) XM.SetEntry(i, m, X[3 * i + m]);
        HMatrix *F = G->GetFields((IncField *)incfield, (HVector *)(HMatrix *)kn, cdouble(wr, wi), &XM, 0);
        for (int i = 0; i < nx; i++) for (int q = 0; q < 6; q++) { cdouble v = F->GetEntry(i, q); out[12 * i + 2 * q] = real(v); out[12 * i + 2 * q + 1] = imag(v); }
        delete F; return 0;, 1) }
void *scuffb200_get_rf_matrix(void *g, double wr, double wi, int nx, const double *X)
{ GUARD(RWGGeometry *G = (RWGGeometry *)g;
        HMatrix XM(nx, 3, LHM_REAL); for (int i = 0; i < nx; i++) for (int m = 0; m < 3; m++) XM.SetEntry(i, m, X[3 * i + m]);
        return (void *)G->GetRFMatrix(cdouble(wr, wi), 0, &XM, 0);, 0) }
/* PFT by the displaced-surface-integral method (radius 0 = automatic; mesh may be NULL); out: NumSurfaces x 8 */
int scuffb200_get_dsipft_matrix(void *g, void *kn, double wr, double wi, void *incfield, double radius, int npoints, const char *mesh, double *out)
{ GUARD(RWGGeometry *G = (RWGGeometry *)g;
        PFTOptions Opts; InitPFTOptions(&Opts); Opts.PFTMethod = SCUFF_PFT_DSI; Opts.IF = (IncField *)incfield; Opts.DSIRadius = radius; Opts.DSIPoints = npoints;
        Opts.DSIMesh = (mesh && mesh[0]) ? (char *)mesh : 0;
        HMatrix *P = G->GetPFTMatrix((HVector *)(HMatrix *)kn, cdouble(wr, wi), &Opts, 0);
        for (int ns = 0; ns < G->NumSurfaces; ns++) for (int nq = 0; nq < NUMPFT; nq++) out[ns * NUMPFT + nq] = P->GetEntryD(ns, nq);
        delete P; return 0;, 1) }
/* IncField::GetFields / GetFieldGradients of the host classes at npts points (the caller's own evaluation; no device involved) */
int scuffb200_incfield_eval(void *incfield, double wr, double wi, const double *epsmu, int npts, const double *X, double *eh, double *deh)
{ GUARD(IncField *IF = (IncField *)incfield;
        IF->SetFrequencyAndEpsMu(cdouble(wr, wi), cdouble(epsmu[0], epsmu[1]), cdouble(epsmu[2], epsmu[3]), false);
        for (int n = 0; n < npts; n++)
         { cdouble EH[6]; IF->GetFields(X + 3 * n, EH);
           for (int q = 0; q < 6; q++) { eh[12 * n + 2 * q] = real(EH[q]); eh[12 * n + 2 * q + 1] = imag(EH[q]); }
           if (deh) { cdouble dEH[3][6]; IF->GetFieldGradients(X + 3 * n, dEH);
                      for (int m = 0; m < 3; m++) for (int q = 0; q < 6; q++) { deh[36 * n + 12 * m + 2 * q] = real(dEH[m][q]); deh[36 * n + 12 * m + 2 * q + 1] = imag(dEH[m][q]); } } }
        return 0;, 1) }
/* X: NX x ncols (3 or 6) row-major; out: NX x 18 complex row-major; m = the LU-factorised BEM matrix */
int scuffb200_get_dyadic_gfs(void *g, void *m, double wr, double wi, int nx, int ncols, const double *X, int scattering_only, double *out)
{ GUARD(RWGGeometry *G = (RWGGeometry *)g;
        HMatrix XM(nx, ncols, LHM_REAL); for (int i = 0; i < nx; i++) for (int c = 0; c < ncols; c++) XM.SetEntry(i, c, X[ncols * i + c]);
        HMatrix *GM = G->GetDyadicGFs(cdouble(wr, wi), 0, &XM, (HMatrix *)m, 0, scattering_only != 0);
        for (int i = 0; i < nx; i++) for (int q = 0; q < 18; q++) { cdouble v = GM->GetEntry(i, q); out[36 * i + 2 * q] = real(v); out[36 * i + 2 * q + 1] = imag(v); }
        delete GM; return 0;, 1) }
int scuffb200_surface_origin(void *g, int ns, double *out) { for (int m = 0; m < 3; m++) out[m] = ((RWGGeometry *)g)->Surfaces[ns]->Origin[m]; return 0; }

}  // extern "C"
