// Internal state of a libscuff_cuda context (one per device).  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>
#include <cusolverDn.h>
#include <string>
#include <vector>
#include <map>
#include <unordered_map>
#include <stdint.h>
#include "../../include/libscuff_cuda.h"
#include "ppi_math.cuh"
#include "gbar_params.h"

namespace scuffcuda {

// ---- device panel record: everything the kernels read for one panel (288 B, 16B aligned) ----
struct __align__(16) PanelRec
{ double C[3];     // centroid            (RWGPanel::Centroid)
  double R;        // bounding radius     (RWGPanel::Radius)
  double V[9];     // vertices V0,V1,V2 in the surface's order (Sa->Vertices + 3*P->VI[i])
  double x[18];    // 6 cubature points X = V0 + u(V1-V0) + v(V2-V0), minus C (local coordinates)
  double sl[3];    // signed length of the RWG edge opposite vertex i: +len (panel is the
                   // edge's P panel), -len (M panel), 0 (exterior edge without basis function)
  int ei[3];       // index of that edge in the surface (RWGPanel::EI), -1 if none
  int pad;
};
static_assert(sizeof(PanelRec) == 288, "PanelRec layout");

struct Surface
{ int NV = 0, NP = 0, NE = 0; bool isPEC = true;
  std::vector<double> V;           // host copies (for plans / FIPPI keys)
  std::vector<PanelRec> hpanels;
  PanelRec *dpanels = nullptr;
  uint64_t generation = 0;         // bumped on every set_surface (invalidates plans)
  // edge topology as scuffcuda_set_surface received it (8 ints per edge) + lengths: the per-edge kernels of rhs_pft.cu
  // (RHS vector, PFT near pairs) build their device edge records from these
  std::vector<int> edgeTopo; std::vector<double> edgeLen;
  void *dedges = nullptr; uint64_t edgesGen = 0;
};

struct Matrix
{ int NR = 0, NC = 0; bool isReal = false;
  double *d = nullptr;             // NR*NC*(1|2) doubles, column-major
  int *ipiv = nullptr; bool factored = false;
};

// stamping parameters for one region pass (SurfaceSurfaceInteractions.cc:165-176,203-327)
struct StampParams
{ double *M; long ld; int isReal;
  int rowOff, colOff; int aPEC, bPEC; int symmetric;
  cd c1, c2, c3;   // PreFac1, PreFac2/(ik), PreFac3
  int colLo, colHi; // only matrix columns in [colLo, colHi) are written (column-block assembly across GPUs); 0,INT_MAX = all
};

// near-pair plan for one (nsa, nsb, displacement): list of panel pairs with rRel<=4 and
// what AssessPanelPair / CanonicallyOrderVertices decided for each
struct NearPair
{ int pa, pb;
  int ncv;
  unsigned char ia[3], ib[3];      // AssessPanelPair vertex order (common vertices first)
  unsigned char oa[3], ob[3];      // canonical order; if flipped, oa indexes panel b's vertices
  int flipped;
  int rec;                         // FIPPI record index (-1: none needed)
  // Symmetric blocks visit an unordered pair once, but the reference evaluates PPI(pb,pa) (not
  // PPI(pa,pb)) for the edge pairs whose lower-numbered edge lies on pb.  For ncv==0 the
  // "canonical" ordering is not symmetric under swapping the panels (AssessPanelPair.cc:413-451), so
  // the swapped order has its own canonical vertices, key and FIPPI record:
  unsigned char oa2[3], ob2[3];
  unsigned char flipped2, hasSwap;
  int rec2;
};
static_assert(sizeof(NearPair) == 44, "NearPair layout");

struct Plan
{ uint64_t genA = 0, genB = 0;
  int nsa = -1, nsb = -1; bool symmetric = false; bool useDisp = false; double L[3] = {0, 0, 0};
  int nNear = 0;
  std::vector<NearPair> hnear;
  NearPair *dnear = nullptr;
  bool fippiAssigned = false;
};

struct FIPPIKey { float k[15]; };
struct FIPPIKeyHash
{ size_t operator()(const FIPPIKey &K) const
  { const uint32_t *p = (const uint32_t *)K.k; uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < 15; i++) { h ^= p[i]; h *= 1099511628211ull; }
    return (size_t)h; }
};
struct FIPPIKeyEq
{ bool operator()(const FIPPIKey &a, const FIPPIKey &b) const { return memcmp(a.k, b.k, sizeof(a.k)) == 0; } };

struct FIPPITable
{ std::unordered_map<FIPPIKey, int, FIPPIKeyHash, FIPPIKeyEq> index;
  std::vector<FIPPIKey> keys;
  std::vector<double> hrec;        // 36 doubles per record (QIFIPPIData, libscuffInternals.h:282-293)
  double *drec = nullptr; long dcap = 0; long dcount = 0;  // device copy (resident across the sweep)
};

// periodic Green's function table (GBarAccelerator, libs/libscuff/GBarAccelerator.h:45-61), resident in HBM
struct GBarTable
{ GBarParams P;
  double RhoMin = 0, RhoMax = 0, LMax = 0;
  int forceFull = 0;
  int N[3] = { 0, 0, 0 };
  double XMin[3] = { 0, 0, 0 }, DX[3] = { 0, 0, 0 };
  double minDelta[3] = { 0, 0, 0 };   // the probed optimal spacings (diagnostics)
  double *dC = nullptr;               // 128 doubles per cell
};

}  // namespace scuffcuda

struct scuffcuda_context
{ int device = 0;
  cudaStream_t stream = nullptr;
  cusolverDnHandle_t solver = nullptr;
  std::string err;
  std::vector<scuffcuda::Surface> surfaces;
  std::map<int, scuffcuda::Matrix> matrices; int nextHandle = 1;
  std::vector<scuffcuda::Plan> plans;
  scuffcuda::FIPPITable fippi;
  std::map<int, scuffcuda::GBarTable> gbarTables;   // handles share nextHandle with the matrices
  unsigned long long counts[5] = {0, 0, 0, 0, 0};
  unsigned long long pairCounts[2] = {0, 0};  // far / near panel pairs of the last assemble_block call
  unsigned long long launches = 0;            // kernels launched by this context since creation
  unsigned long long h2dBytes = 0, d2hBytes = 0;
  double timing[5] = {0, 0, 0, 0, 0};
  unsigned long long *dcounters = nullptr;   // device counters (near-list length, algorithm counts)
  void *dscratch = nullptr; size_t scratchBytes = 0;
  int smCount = 148;
  // per-context helper streams, events and mapped host buffers (created with the context, destroyed with it)
  cudaStream_t nearStream = nullptr;          // high priority: near/singular kernels, concurrent with the far kernel
  cudaStream_t tdStream = nullptr;            // high priority: Taylor-Duffy kernels, concurrent with the desingularised-pair kernel
  cudaEvent_t evTd[2] = { nullptr, nullptr };
  cudaStream_t copyStream = nullptr;          // downloads overlapping with assembly
  cudaEvent_t ev[7] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
  struct MappedBuf { void *p = nullptr; size_t n = 0; };
  std::map<int, MappedBuf> mapped;            // slot -> page-locked, device-mapped scratch
  int *dHoIdx = nullptr; size_t hoIdxCap = 0; // index list of the high-order cubature kernel
  void *pftState = nullptr;                   // rhs_pft.cu: cubature tables, near basis-function pair lists (freed by pft_release)
};
