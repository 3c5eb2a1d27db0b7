/* libscuff_cuda — C ABI of the B200 (sm_100a) implementation of SCUFF-EM's dense BEM
 * system-matrix assembly.
 *
 * This is the drop-in boundary: plain pointers and sizes only.  Each entry point names the
 * reference interface (HomerReid/scuff-em, paths relative to the reference root) it replaces
 * or feeds.  The host-side C++ mirror of the reference API (include/libscuff_b200.h:
 * RWGGeometry::AssembleBEMMatrix / AssembleBEMMatrixBlock, HMatrix) marshals to these calls;
 * INTEGRATION.md shows the shim a maintainer would add to the reference's libscuff.
 *
 * All functions return 0 on success and a non-zero status on failure;
 * scuffcuda_last_error() returns the message (the C++ shim turns it into ErrExit(), the
 * reference's error convention, libs/libhrutil/libhrutil.h:136).
 * There is NO CPU fallback: without a CUDA device scuffcuda_create() fails.
 */
#ifndef LIBSCUFF_CUDA_H
#define LIBSCUFF_CUDA_H

#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct scuffcuda_context scuffcuda_context;

/* ---- context ------------------------------------------------------------------------ */
/* One context per device (one host process/thread per GPU for sweep sharding,
 * SURVEY §8e).  Owns the device copy of the geometry, the FIPPI table
 * (GlobalFIPPICache, libs/libscuff/FIPPICache.cc:405) and device matrices.            */
int  scuffcuda_create(int device, scuffcuda_context **ctx);
void scuffcuda_destroy(scuffcuda_context *ctx);
const char *scuffcuda_last_error(scuffcuda_context *ctx);   /* ctx may be NULL        */
int  scuffcuda_device_count(void);

/* ---- geometry ----------------------------------------------------------------------- */
/* Flattened RWGSurface (libs/libscuff/libscuff.h:174-328): what the hot path reads from
 * Sa->Vertices, Sa->Panels[np]->{VI,Centroid,Radius}, Sa->Edges[ne]->{iPPanel,iMPanel,
 * PIndex,MIndex,Length}.  Must be re-sent after RWGSurface::Transform/UnTransform
 * (RWGSurface.cc:588-668) because vertices, centroids move.
 *   vertices      [3*num_vertices]  xyz
 *   panel_vi      [3*num_panels]    vertex indices
 *   panel_centroid[3*num_panels], panel_radius[num_panels]  (as InitRWGPanel computed them)
 *   edge_topo     [8*num_edges]     iV1,iV2,iQP,iQM,iPPanel,iMPanel,PIndex,MIndex
 *                                   (iMPanel==-1: half-RWG edge, only the P panel)
 *   edge_length   [num_edges]                                                          */
int scuffcuda_set_num_surfaces(scuffcuda_context *ctx, int num_surfaces);
int scuffcuda_set_surface(scuffcuda_context *ctx, int ns,
                          int num_vertices, const double *vertices,
                          int num_panels, const int *panel_vi,
                          const double *panel_centroid, const double *panel_radius,
                          int num_edges, const int *edge_topo, const double *edge_length,
                          int is_pec);

/* ---- device matrices (the HMatrix the path fills; libs/libhmat/libhmat.h:165-383) ---- */
/* Column-major, ZM[nr + nc*NR] complex128 or DM[...] float64 (LHM_REAL), kept resident
 * in HBM so that the LU consumer runs on the device.                                   */
int scuffcuda_matrix_create(scuffcuda_context *ctx, int num_rows, int num_cols, int is_real, int *handle);
int scuffcuda_matrix_destroy(scuffcuda_context *ctx, int handle);
int scuffcuda_matrix_zero(scuffcuda_context *ctx, int handle);
/* HMatrix::InsertBlock(M, RowOffset, ColOffset, NR, NC, BRowOffset, BColOffset)
 * within one matrix — the Mate[] diagonal-block reuse of AssembleBEMMatrix.cc:731-737  */
int scuffcuda_matrix_copy_block(scuffcuda_context *ctx, int handle, int dst_row, int dst_col,
                                int num_rows, int num_cols, int src_row, int src_col);
/* HMatrix::InsertBlock between two device matrices of the same type (libs/libhmat/HMatrix.cc InsertBlock): the T/U-block stamping
 * of scuff-cas3D's Factorize (CasimirIntegrand.cc:159-196) without a round trip through the host                          */
int scuffcuda_matrix_insert_block(scuffcuda_context *ctx, int handle_dst, int dst_row, int dst_col,
                                  int handle_src, int src_row, int src_col, int num_rows, int num_cols);
/* sparse accumulate M[rows[i], cols[i]] += vals[2i] + i*vals[2i+1] (HMatrix::AddEntry; a real matrix takes the real part):
 * the surface-impedance term of AddSurfaceZetaContributionToBEMMatrix (SurfaceSurfaceInteractions.cc:339-421), whose
 * O(NumEdges) overlap entries the host computes                                           */
int scuffcuda_matrix_add_entries(scuffcuda_context *ctx, int handle, int n, const int *rows, const int *cols, const double *vals);
/* StampInNeighborBlock (AssembleBEMMatrix.cc:72-114), the Bloch-phase accumulation of one neighbour-cell block of a
 * periodic geometry: M[row_off+r, col_off+c] += BPF*B[r,c] (+ conj(BPF)*B[c,r] if use_symmetry), BPF = phase[0]+i*phase[1]
 * = exp(i kBloch.L); `handle_B` is an NBFA x NBFB matrix (complex, or real: the
 * LHM_REAL cache blocks of CreateABMBAccelerator) filled by scuffcuda_assemble_block with use_displacement          */
int scuffcuda_matrix_stamp_neighbor_block(scuffcuda_context *ctx, int handle, int row_off, int col_off,
                                          int handle_B, const double phase[2], int use_symmetry);
/* upper -> lower triangle copy of the square sub-block at (offset,offset) of size n
 * (AssembleBEMMatrix.cc:753-758, SurfaceSurfaceInteractions.cc:573-583)                */
int scuffcuda_matrix_symmetrize(scuffcuda_context *ctx, int handle, int offset, int n);
/* host <-> device; host buffer is column-major with leading dimension ld (elements)    */
int scuffcuda_matrix_download(scuffcuda_context *ctx, int handle, void *host, long ld);
int scuffcuda_matrix_upload(scuffcuda_context *ctx, int handle, const void *host, long ld);
/* copy-back that overlaps with assembly (what HMatrix *AssembleBEMMatrix returns is a HOST matrix,
 * AssembleBEMMatrix.cc:684-765): enqueue the D2H copy of rows [row0,row0+nrows) x columns [col0,col0+ncols)
 * on the context's copy stream, ordered after everything issued on the compute stream so far; `host` is the
 * base of the full column-major host matrix (page-locked memory for a truly asynchronous copy).
 * _wait(ctx,0) blocks until all such copies have landed; _wait(ctx,1) instead makes later compute-stream
 * work wait for them (no host block).                                                   */
int scuffcuda_matrix_download_region_async(scuffcuda_context *ctx, int handle, void *host, long ld,
                                           int row0, int nrows, int col0, int ncols);
int scuffcuda_matrix_download_wait(scuffcuda_context *ctx, int fence_compute);
/* a marker in the copy stream: _mark returns a ticket for "every region copy enqueued so far", _wait_ticket blocks the host until
 * those copies have landed (and releases the ticket).  Lets a caller consume the first regions (e.g. mirror the upper triangle on
 * the host) while later regions are still crossing PCIe.                                  */
int scuffcuda_matrix_download_mark(scuffcuda_context *ctx, int *ticket);
int scuffcuda_matrix_download_wait_ticket(scuffcuda_context *ctx, int ticket);
/* raw device pointer (for zero-copy consumers such as torch / cuSOLVER callers)        */
int scuffcuda_matrix_device_ptr(scuffcuda_context *ctx, int handle, void **dptr);
/* downstream consumer: HMatrix::LUFactorize / LUSolve (libs/libhmat/LBWrappers.cc:150-230)
 * via cuSOLVER getrf/getrs on the resident matrix; rhs is host memory [n*nrhs].        */
int scuffcuda_lu_factor(scuffcuda_context *ctx, int handle);
int scuffcuda_lu_solve(scuffcuda_context *ctx, int handle, void *rhs_host, int nrhs);

/* ---- assembly ----------------------------------------------------------------------- */
/* Mirror of GetSSIArgStruct (libs/libscuff/libscuffInternals.h:217-264) as filled by
 * AssembleBEMMatrixBlock (AssembleBEMMatrix.cc:371-387) and GetSurfaceSurfaceInteractions
 * (SurfaceSurfaceInteractions.cc:462-492): material values are evaluated on the host by
 * MatProp::GetEpsMu and passed by value.                                               */
typedef struct scuffcuda_ssi_args {
  double omega[2];          /* Omega (re, im)                                           */
  double eps_a[2], mu_a[2]; /* first common region; eps_a==0 => skipped                 */
  double sign_a;
  double eps_b[2], mu_b[2]; /* second common region; eps_b==0 => absent                 */
  double sign_b;
  int    symmetric;         /* Args->Symmetric: compute nea<=neb, then mirror           */
  int    accumulate;        /* Args->Accumulate: do not zero the block first            */
  int    row_offset, col_offset;
  int    use_displacement;  /* Args->Displacement != 0                                  */
  double displacement[3];
  int    force_taylor_duffy;/* RWGGeometry::DisableCache (PanelPanelInteractions.cc:589)*/
  int    far_only;          /* debug/bench: evaluate only rRel>4 pairs                  */
  int    col_begin, col_end;/* column-block assembly across GPUs: write only the GLOBAL matrix columns
                             * [col_begin, col_end) (and skip the panel pairs that cannot reach them);
                             * col_end == 0: all columns.  Symmetric blocks still fill the upper triangle
                             * only and are NOT mirrored when a range is given (the caller mirrors the
                             * gathered matrix with scuffcuda_matrix_symmetrize); without `accumulate`
                             * only the owned columns of the block are cleared                  */
  int    gba_a, gba_b;      /* Args->GBA1 / GBA2: handles from scuffcuda_gbar_create.  Non-zero = the outer-cell step of a
                             * periodic geometry (AssembleBEMMatrix.cc:526-557): every panel pair of the block is
                             * integrated with the 6x6 rule and the periodic Green's function of the region
                             * (PanelPanelInteractions.cc:124-153,420-424) instead of exp(ikr)/4 pi r; needs symmetric=0,
                             * no displacement; a region without a table must be passed with eps = 0 (OmitRegion)   */
  int    reserved[3];
} scuffcuda_ssi_args;

/* GetSurfaceSurfaceInteractions(Args) on the device: adds the (nsa,nsb) block into the
 * device matrix `handle` at (row_offset,col_offset).                                   */
int scuffcuda_assemble_block(scuffcuda_context *ctx, int nsa, int nsb,
                             const scuffcuda_ssi_args *args, int handle);

/* Derivative blocks of the same surface pair — the GradM[3] / dMdT[NumTorqueAxes] arguments of AssembleBEMMatrixBlock
 * (AssembleBEMMatrix.cc:349-391 -> GetSurfaceSurfaceInteractions GradB / dBdTheta, SurfaceSurfaceInteractions.cc:205-327;
 * integrands PanelPanelInteractions.cc:176-199, 6x6 rule for far pairs, 78x78 without desingularisation otherwise, zero for
 * Taylor-Duffy pairs :484-486, :504-510).  grad_handles[Mu] / dmdt_handles[n]: matrices of the block's shape or 0 to skip a
 * component; gamma_matrix: 9 doubles per torque axis as CreateGammaMatrix writes them (EdgeEdgeInteractions.cc:241-270).
 * args as for scuffcuda_assemble_block (symmetric must be 0: the reference's callers ask for derivatives of blocks between
 * different surfaces; gba_a/gba_b and column ranges are not available).  The value block itself is NOT computed here.   */
int scuffcuda_assemble_block_gradients(scuffcuda_context *ctx, int nsa, int nsb, const scuffcuda_ssi_args *args,
                                       const int grad_handles[3], int num_torque_axes, const double *gamma_matrix,
                                       const int dmdt_handles[3]);

/* ---- periodic Green's function (libs/libscuff/GBarAccelerator.h:45-92, GBarVDEwald.cc:931-996) ------------------ */
/* CreateGBarAccelerator(LBasis, RhoMin, RhoMax, k, kBloch, RelTol, ExcludeInnerCells) (GBarAccelerator.cc:339-547) on the
 * device: Ewald sums, grid-spacing probes, table fill and tricubic fit are CUDA kernels; the table stays in HBM.
 *   k[2] wavenumber in the region (re,im); lbv = {L1x, L1y, L2x, L2y}; ldim must be 2 (1-D lattices: error)
 *   rho_min > rho_max or force_full_ewald: no table, every lookup is a full Ewald sum (SCUFF_EWALD_FULL, :384-396)
 *   nmax_grid > 0 caps the grid dimensions (SCUFF_GBAR_NMAX, :532-537)                                                 */
int scuffcuda_gbar_create(scuffcuda_context *ctx, const double k[2], const double kbloch[2], const double lbv[4], int ldim,
                          double rho_min, double rho_max, double rel_tol, int exclude_inner_cells, int force_full_ewald,
                          int nmax_grid, int *handle);
int scuffcuda_gbar_destroy(scuffcuda_context *ctx, int handle);
/* table decisions: grid dimensions n[3] (0 if no table), origin xmin[3], spacing dx[3], LMax, full-Ewald flag */
int scuffcuda_gbar_info(scuffcuda_context *ctx, int handle, int n[3], double xmin[3], double dx[3], double *lmax, int *full_ewald);
/* tricubic coefficients, 128 doubles per cell (Interp3D::CTable layout, libs/libMDInterp/Interp3D.cc:63-69), host buffer */
int scuffcuda_gbar_get_table(scuffcuda_context *ctx, int handle, double *coeffs);
/* GetGBar(R, GBA, dGBar) (GBarAccelerator.cc:983-1005) at npts points R[3*npts] (host): out8 = G, dG/dx, dG/dy, dG/dz as (re,im) */
int scuffcuda_gbar_eval(scuffcuda_context *ctx, int handle, int npts, const double *R, double *out8);
/* GBarVDEwald(R, k, kBloch, LBV, LDim, E=-1, ExcludeInnerCells, GBarVD) (GBarVDEwald.cc:931-996) at npts points:
 * out16 = value, d/dx, d/dy, d/dz, d2/dxdy, d2/dxdz, d2/dydz, d3/dxdydz as (re,im)                                      */
int scuffcuda_gbar_ewald(scuffcuda_context *ctx, const double k[2], const double kbloch[2], const double lbv[4], int ldim,
                         int exclude_inner_cells, int npts, const double *R, double *out16);

/* ---- right-hand side, solve and power / force / torque on the device ---------------------------
 * One frequency point of a scattering run (applications/scuff-scatter/scuff-scatter.cc:402-520) without the matrix or the
 * vectors leaving HBM.  Device vectors are N x 1 device matrices (scuffcuda_matrix_create(ctx, N, 1, is_real, &h)).     */
/* PlaneWave (libs/libIncField/libIncField.h:106-125) as UpdateIncFields prepared it (AssembleRHSVector.cc:62-118)      */
typedef struct scuffcuda_planewave {
  double e0[6];             /* PlaneWave::E0, (re,im) x 3                                                             */
  double nhat[3];           /* PlaneWave::nHat                                                                        */
  double eps[2], mu[2];     /* IncField::Eps, Mu of the region that holds the sources                                 */
  double sign;              /* +1: sources in the surface's interior region (PositiveIFs), -1: exterior (NegativeIFs),
                             * AssembleRHSVector.cc:297-306; ignored by scuffcuda_emtpft                              */
} scuffcuda_planewave;
/* cubature points of the RHS inner products of surface ns: for edge ne, panel pm (0 positive, 1 negative), point np of the
 * 78-point rule: X[3*((ne*2+pm)*78+np)...] (host buffer; X may be NULL to query *npts = 2*78*NumEdges).  Incident fields the
 * device does not know (any IncField subclass) are evaluated there by the caller and passed as `tabulated_eh`.            */
int scuffcuda_rhs_points(scuffcuda_context *ctx, int ns, double *X, long *npts);
/* AssembleRHSVector for one surface (AssembleRHSVector.cc:282-345, GetInnerProducts :172-229): writes the entries
 * [offset, offset+NumBFs) of the device vector.  Fields = sum of the plane waves (each with its sign) + tabulated_eh
 * (12 doubles per point of scuffcuda_rhs_points: E then H as (re,im), signs already applied; or NULL).  A real vector takes
 * the real parts (AllocateRHSVector(PureImagFreq), :438-447).                                                            */
int scuffcuda_assemble_rhs(scuffcuda_context *ctx, int ns, const double omega[2], int num_pw, const scuffcuda_planewave *pw,
                           const double *tabulated_eh, int vec_handle, int offset);
/* HMatrix::LUSolve(HVector *) (libs/libhmat/LBWrappers.cc:196-230) with the vector resident on the device                */
int scuffcuda_lu_solve_device(scuffcuda_context *ctx, int handle, int vec_handle);
/* device vector / matrix element-wise copy dst <- src (HVector::Copy): same shape and type                                */
int scuffcuda_matrix_copy(scuffcuda_context *ctx, int handle_dst, int handle_src);

/* GetEMTPFTMatrix (libs/libscuff/EMTPFT.cc:533-905) with the cubature integrals SCUFF_EMTPFTI_EHDERIVATIVES1 (:82-176,
 * :308-334): the energy/momentum-transfer sums over all pairs of basis functions and the extinction sums over all basis
 * functions (GetExtinctionPFTT :466-531), on the device.                                                                 */
typedef struct scuffcuda_pft_args {
  double omega[2];              /* real, positive frequency                                                           */
  int    interior;              /* PFTOptions::Interior                                                               */
  int    num_surfaces;
  const int    *region_indices; /* RWGSurface::RegionIndices, 2 per surface                                           */
  const int    *bf_offsets;     /* RWGGeometry::BFIndexOffset                                                         */
  const double *origins;        /* RWGSurface::Origin (torque centres), 3 per surface                                 */
  int    num_regions;
  const double *eps_mu;         /* per region: eps (re,im), mu (re,im), relative (RegionMPs[nr]->GetEpsMu(Omega))      */
  int    if_region;             /* IncField::RegionIndex of the head of the incident-field chain; -1: no incident field */
  int    num_pw;                /* 1: the incident field is the plane wave *pw; 0: it is tabulated                     */
  const scuffcuda_planewave *pw;
  const double *tabulated_ehd;  /* or NULL; 48 doubles per point of scuffcuda_pft_points, surfaces one after the other:
                                 * EH[6] then dEH[3][6] (IncField::GetFields / GetFieldGradients) as (re,im)           */
} scuffcuda_pft_args;
/* the 21-point rule of every panel of surface ns in the panel's vertex order (PanelCubature.cc:155-236): X[3*(np*21+p)...] */
int scuffcuda_pft_points(scuffcuda_context *ctx, int ns, double *X, long *npts);
/* scattered[(nsa*NS + nsb)*11 + nq] = contribution of surface b to quantity nq on surface a (DeltaPFTT, EMTPFT.cc:604-625;
 * nq: 0 PAbs (unused), 1 PScat, 2-4 force, 5-7 first torque term, 8-10 second torque term); extinction[ns*11 + nq] as
 * GetExtinctionPFTT returns it.  Host buffers.                                                                           */
int scuffcuda_emtpft(scuffcuda_context *ctx, int kn_handle, const scuffcuda_pft_args *args, double *scattered, double *extinction);

/* ---- scattered fields at evaluation points: RWGGeometry::GetRFMatrix / GetFields (libs/libscuff/GetFields.cc:141-395) --
 * Contribution of surface ns to the fields at npts points X[3*npts] (host) that all lie in ONE region bounded by the surface:
 * k, zrel = wavenumber and relative wave impedance of that region, sign = +1 if it is the surface's exterior region, -1 if interior
 * (GetFields.cc:262-270).  kn_handle != 0: F[12*npts] (host; E then H as (re,im)) += sum over the surface's basis functions of
 * (reduced fields) x (solution entries).  rf_handle != 0: the reduced-field matrix entries RFMatrix[bf_offset + .., 6*rf_columns[i] + mu]
 * are stored in that device matrix (TotalBFs x 6 NX complex).  16-point rule beyond 4 basis-function radii, 78-point rule beyond 1
 * (:171-181, :284-295); closer pairs: GetReducedFields_Nearby (libs/libscuff/GetNearFields.cc:60-905: closed-form static panel
 * integrals + 21-point rule for the desingularised remainder), *near_pairs counts them (diagnostic).                      */
int scuffcuda_get_fields(scuffcuda_context *ctx, int ns, const double k[2], const double zrel[2], double sign, int npts, const double *X,
                         int kn_handle, int bf_offset, double *F, int rf_handle, const int *rf_columns, long *near_pairs);

/* RWGGeometry::GetDyadicGFs (libs/libscuff/GetDyadicGFs.cc:58-190), the contraction step: with D = the reduced-field matrix of the
 * destination points and S = the LU-solved reduced-field matrix of the source points (both TotalBFs x 6 NX device matrices from
 * scuffcuda_get_fields / scuffcuda_lu_solve_device), out[36*nx + 2*(3i+j)] = sum_bf D[bf,6nx+i] S[bf,6nx+j] (electric, re/im) and
 * out[36*nx + 18 + 2*(3i+j)] the same with the magnetic columns 3..5; host buffer.  The normalisation factors are the caller's.   */
int scuffcuda_dgf_contract(scuffcuda_context *ctx, int rf_dest_handle, int rf_source_handle, int npts, double *out);

/* PPI algorithm counters of the last assemble_block call, same order as the reference's
 * log line (SurfaceSurfaceInteractions.cc:550-558; libscuffInternals.h PPIALG_*):
 * [0] low-order cubature, [1] high-order cubature, [2] Taylor-Duffy, [3] high-k TD,
 * [4] desingularised; counted per (panel pair, Q pair) PPI the reference would evaluate */
int scuffcuda_get_ppi_counts(scuffcuda_context *ctx, unsigned long long counts[5]);

/* panel pairs of the last assemble_block call: [0] far pairs integrated by the far-field kernel (each ONCE, for
 * all its (Q,Q') combinations), [1] near/singular pairs (rRel <= 4)                                        */
int scuffcuda_get_pair_counts(scuffcuda_context *ctx, unsigned long long counts[2]);
/* bookkeeping for measurement: CUDA kernels launched by this context since creation, and bytes this
 * context copied host->device [0] / device->host [1] since creation                                       */
int scuffcuda_get_launch_count(scuffcuda_context *ctx, unsigned long long *n);
int scuffcuda_get_transfer_bytes(scuffcuda_context *ctx, unsigned long long bytes[2]);
/* pinned (page-locked, zero-filled) host memory for HMatrix::ZM/DM (libs/libhmat/HMatrix.cc:325-360 allocates
 * with mallocEC) so that the copy-back of an assembled matrix runs at full PCIe rate                       */
int scuffcuda_host_alloc(scuffcuda_context *ctx, size_t bytes, void **ptr);
int scuffcuda_host_free(scuffcuda_context *ctx, void *ptr);

/* device time (ms, CUDA events on the context's stream) of the kernels of the last
 * assemble_block call: [0] total, [1] far-field kernel, [2] near/singular kernels,
 * [3] FIPPI table build, [4] classification/plan                                       */
int scuffcuda_get_last_timing(scuffcuda_context *ctx, double ms[5]);
int scuffcuda_synchronize(scuffcuda_context *ctx);

/* ---- FIPPI table (PreloadCache / StoreCache, libs/libscuff/libscuff.h:662-663;
 *      file format FIPPICache.cc:209-400: "FIPPICACHE\0" + 352-byte records)  --------- */
int scuffcuda_fippi_preload(scuffcuda_context *ctx, const char *filename);
int scuffcuda_fippi_store(scuffcuda_context *ctx, const char *filename);
int scuffcuda_fippi_num_records(scuffcuda_context *ctx, long *n);
/* drop every record (and the pair plans that index them); next assembly rebuilds the table on the device */
int scuffcuda_fippi_clear(scuffcuda_context *ctx);
/* FIPPI table built ONCE by all ranks of a sweep instead of once per rank: after _set_partition(rank, world) a pair
 * plan computes only the records with index = rank (mod world); _prepare_block builds the plan of a surface pair (pair
 * classification + its share of the records) without assembling anything; _host_records exposes the table (36 doubles per record,
 * zeros where another rank computes) so that the caller can sum it over the ranks in place — record indices are identical on
 * every rank because the geometry is; _commit sends the completed table to the device.  Reset with _set_partition(0, 1).   */
int scuffcuda_fippi_set_partition(scuffcuda_context *ctx, int rank, int world);
int scuffcuda_prepare_block(scuffcuda_context *ctx, int nsa, int nsb, int symmetric, int use_displacement, const double *displacement);
int scuffcuda_fippi_host_records(scuffcuda_context *ctx, double **records, long *count);
int scuffcuda_fippi_commit(scuffcuda_context *ctx);

/* ---- self-test of the kernels' FP64 elementary functions (csrc/fastmath.cuh): for each x[i] writes
 *      out[6i..] = { rsqrt(|x|), rsqrt one-step variant, sin x, cos x, exp(-|x|), 0 }; host buffers          */
int scuffcuda_selftest_math(scuffcuda_context *ctx, int n, const double *x, double *out);

/* ---- measurement helpers ------------------------------------------------------------ */
/* dependent-chain-free DFMA microkernel: returns achieved FP64 TFLOP/s (2 flop per FMA);
 * used as the in-run FP64 roofline denominator (MEASURED_PEAKS.json has no FP64 entry) */
int scuffcuda_measure_fp64_peak(scuffcuda_context *ctx, double *tflops);

#ifdef __cplusplus
}
#endif
#endif
