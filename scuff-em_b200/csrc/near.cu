// Near / singular panel pairs (rRel <= 4): the pair plan (AssessPanelPair + CanonicallyOrderVertices
// decisions, cached per surface pair), the FIPPI table resident in HBM (GlobalFIPPICache), and the
// desingularised-cubature kernel (PanelPanelInteractions.cc:504-553).  Taylor-Duffy pairs are handled
// in td_kernel.cu, FIPPI records are built on the device in fippi_kernel.cu.
#include <algorithm>
#include <numeric>
#include <map>
#include <mutex>
#include <cstdio>
#include <cstring>
#include "kernels.h"
#include "classify.cuh"
#include "stamp.cuh"
#include "fippi_math.cuh"

namespace scuffcuda {

#define NEAR_THREADS 128

// ---------------------------------------------------------------------------------------
// pass 1: list all pairs with rRel <= 4
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NEAR_THREADS)
near_list_kernel(const PanelRec *pa, int NPa, const PanelRec *pb, int NPb, int TA, int symmetric,
                 int useDisp, double L0, double L1, double L2,
                 int2 *out, unsigned long long *counter, unsigned long long capacity)
{ extern __shared__ double sm[];   // TA x {C[3], R}
  const int a0 = blockIdx.y * TA, bBase = blockIdx.x * NEAR_THREADS;
  if (symmetric && bBase + NEAR_THREADS - 1 < a0) return;
  const int nA = min(TA, NPa - a0);
  for (int i = threadIdx.x; i < nA * 4; i += NEAR_THREADS) sm[i] = ((const double *)(pa + a0 + i / 4))[i % 4];
  __syncthreads();
  const int b = bBase + threadIdx.x;
  if (b >= NPb) return;
  const double Cb[3] = { pb[b].C[0], pb[b].C[1], pb[b].C[2] }; const double Rb = pb[b].R;
  const double L[3] = { L0, L1, L2 };
  for (int ia = 0; ia < nA; ia++)
   { const int a = a0 + ia;
     if (symmetric && b < a) continue;
     double D[3];
     const double rRel = pair_rrel(sm + 4 * ia, sm[4 * ia + 3], Cb, Rb, useDisp != 0, L, D);
     if (rRel > SCUFF_DESING_RADIUS) continue;
     unsigned long long idx = atomicAdd(counter, 1ull);
     if (idx < capacity) out[idx] = make_int2(a, b);
   }
}

// ---------------------------------------------------------------------------------------
// pass 2: AssessPanelPair + CanonicallyOrderVertices + FIPPI key for every listed pair
// ---------------------------------------------------------------------------------------
__global__ void near_assess_kernel(const PanelRec *pa, const PanelRec *pb, const int2 *list, int n,
                                   int useDisp, double L0, double L1, double L2, int symmetric,
                                   NearPair *out, float *keys, float *keys2)
{ int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int a = list[idx].x, b = list[idx].y;
  const PanelRec &Pa = pa[a], &Pb = pb[b];
  double Va[9], Vb[9];
  const double L[3] = { L0, L1, L2 };
  for (int i = 0; i < 9; i++) { Va[i] = Pa.V[i]; Vb[i] = useDisp ? SC_ADD(Pb.V[i], SC_MUL(1.0, L[i % 3])) : Pb.V[i]; }
  double D[3];
  const double rMax = fmax(Pa.R, Pb.R);
  const double rRel = pair_rrel(Pa.C, Pa.R, Pb.C, Pb.R, useDisp != 0, L, D);
  NearPair N; N.pa = a; N.pb = b; N.rec = -1;
  // AssessPanelPair.cc:167-168: without a displacement, rRel > 2 short-circuits to ncv = 0
  if (!useDisp && rRel > 2.0) { N.ncv = 0; for (int i = 0; i < 3; i++) { N.ia[i] = i; N.ib[i] = i; } }
  else N.ncv = assess_vertices(Va, Vb, rMax, N.ia, N.ib);
  const double *VaO[3] = { Va + 3 * N.ia[0], Va + 3 * N.ia[1], Va + 3 * N.ia[2] };
  const double *VbO[3] = { Vb + 3 * N.ib[0], Vb + 3 * N.ib[1], Vb + 3 * N.ib[2] };
  N.flipped = canonical_order(VaO, VbO, N.ncv, N.oa, N.ob);
  const double *const *A = N.flipped ? VbO : VaO, *const *B = N.flipped ? VaO : VbO;
  const double *OVa[3] = { A[N.oa[0]], A[N.oa[1]], A[N.oa[2]] };
  const double *OVb[3] = { B[N.ob[0]], B[N.ob[1]], B[N.ob[2]] };
  fippi_key(OVa, OVb, keys + 15 * (size_t)idx);
  N.hasSwap = 0; N.flipped2 = 0; N.rec2 = -1;
  for (int i = 0; i < 3; i++) { N.oa2[i] = i; N.ob2[i] = i; }
  if (symmetric && N.ncv == 0 && a != b)
   { // the swapped evaluation PPI(pb,pa): AssessPanelPair leaves both vertex lists in original order
     N.hasSwap = 1;
     N.flipped2 = (unsigned char)canonical_order(VbO, VaO, 0, N.oa2, N.ob2);
     const double *const *A2 = N.flipped2 ? VaO : VbO, *const *B2 = N.flipped2 ? VbO : VaO;
     const double *OVa2[3] = { A2[N.oa2[0]], A2[N.oa2[1]], A2[N.oa2[2]] };
     const double *OVb2[3] = { B2[N.ob2[0]], B2[N.ob2[1]], B2[N.ob2[2]] };
     fippi_key(OVa2, OVb2, keys2 + 15 * (size_t)idx);
   }
  out[idx] = N;
}

// ---------------------------------------------------------------------------------------
// desingularised cubature + FIPPI terms, one thread per near pair
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NEAR_THREADS)
near_desing_kernel(const NearArgs A)
{ int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx == 0) A.counters[6] = (unsigned long long)A.nNear;     // number of near pairs of this block (scuffcuda_get_pair_counts)
  if (idx >= A.nNear) return;
  const NearPair N = A.near[idx];
  const PanelRec &Pa = A.pa[N.pa], &Pb = A.pb[N.pb];
  bool verySW;
  if (near_algorithm(N.ncv, A.kabs, Pa.R, Pb.R, A.forceTD, verySW) != 4) return;
  if (A.S.colHi != 0x7fffffff)
   { // column-block assembly: skip pairs none of whose (<= 9) entries falls into the owned column range (exact per entry,
     // as in far_kernel)
     const int w = A.S.bPEC ? 1 : 2;
     bool any = false;
     for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
       { const int e = Pa.ei[i], f = Pb.ei[j];
         const int c0 = A.S.colOff + w * (A.S.symmetric ? max(e, f) : f);
         any |= (e >= 0 && f >= 0 && c0 + w - 1 >= A.S.colLo && c0 < A.S.colHi); }
     if (!any) return; }

  double D[3];
  pair_rrel(Pa.C, Pa.R, Pb.C, Pb.R, A.useDisp != 0, A.L, D);
  double xa[18], xb[18];
  for (int i = 0; i < 18; i++) { xa[i] = Pa.x[i]; xb[i] = Pb.x[i]; }
  Moments<cd> M;
  panel_pair_moments<KMODE_CPLX, true>(xa, xb, D, A.K, M);

  // global vertex coordinates (b displaced as the reference does: V + 1.0*L)
  double Va[9], Vb[9];
  for (int i = 0; i < 9; i++) { Va[i] = Pa.V[i]; Vb[i] = A.useDisp ? (Pb.V[i] + 1.0 * A.L[i % 3]) : Pb.V[i]; }
  const double *VaO[3] = { Va + 3 * N.ia[0], Va + 3 * N.ia[1], Va + 3 * N.ia[2] };
  const double *VbO[3] = { Vb + 3 * N.ib[0], Vb + 3 * N.ib[1], Vb + 3 * N.ib[2] };
  const double *const *PA = N.flipped ? VbO : VaO, *const *PB = N.flipped ? VaO : VbO;
  const double *OVa[3] = { PA[N.oa[0]], PA[N.oa[1]], PA[N.oa[2]] };
  const double *OVb[3] = { PB[N.ob[0]], PB[N.ob[1]], PB[N.ob[2]] };
  // centroid of the panel that plays "a" in the canonical ordering, in AssessPanelPair order
  const double *const *CV = N.flipped ? VbO : VaO;
  double Cen[3];
  for (int m = 0; m < 3; m++) Cen[m] = (CV[0][m] + CV[1][m] + CV[2][m]) / 3.0;
  const double *rec = A.fippi + (size_t)N.rec * FIPPI_REC_DOUBLES;
  // swapped evaluation PPI(pb,pa) (see NearPair): canonical vertices, centroid and record of that order
  const double *const *PA2 = N.flipped2 ? VaO : VbO, *const *PB2 = N.flipped2 ? VbO : VaO;
  const double *OVa2[3] = { PA2[N.oa2[0]], PA2[N.oa2[1]], PA2[N.oa2[2]] };
  const double *OVb2[3] = { PB2[N.ob2[0]], PB2[N.ob2[1]], PB2[N.ob2[2]] };
  const double *const *CV2 = N.flipped2 ? VaO : VbO;
  double Cen2[3];
  for (int m = 0; m < 3; m++) Cen2[m] = (CV2[0][m] + CV2[1][m] + CV2[2][m]) / 3.0;
  const double *rec2 = N.hasSwap ? A.fippi + (size_t)N.rec2 * FIPPI_REC_DOUBLES : rec;

  unsigned long long nPPI = 0;
  for (int i = 0; i < 3; i++)
   { const int ea = Pa.ei[i]; if (ea < 0) continue;
     const double qa[3] = { Va[3 * i] - Pa.C[0], Va[3 * i + 1] - Pa.C[1], Va[3 * i + 2] - Pa.C[2] };
     for (int j = 0; j < 3; j++)
      { const int eb = Pb.ei[j]; if (eb < 0) continue;
        if (A.S.symmetric && N.pa == N.pb && ea > eb) continue;
        const double qb[3] = { Pb.V[3 * j] - Pb.C[0], Pb.V[3 * j + 1] - Pb.C[1], Pb.V[3 * j + 2] - Pb.C[2] };
        cd H0, H1;
        contract_moments(M, qa, qb, D, A.K, H0, H1);
        const double *Qa = Va + 3 * i, *Qb = Vb + 3 * j;
        QDFIPPI q;
        cd F0 = mk(0, 0), F1 = mk(0, 0), G0 = mk(0, 0), G1 = mk(0, 0);
        const bool needFwd = !N.hasSwap || ea <= eb, needSwap = N.hasSwap && ea >= eb;
        if (needFwd)
         { qdfippi(OVa, N.flipped ? Qb : Qa, OVb, N.flipped ? Qa : Qb, Qa, Qb, Cen, N.flipped, rec, q);
           add_fippi_terms(q, A.K, F0, F1); }
        if (needSwap)   // roles of the two panels exchanged: PPI(pb,Qb ; pa,Qa)
         { qdfippi(OVa2, N.flipped2 ? Qa : Qb, OVb2, N.flipped2 ? Qb : Qa, Qb, Qa, Cen2, N.flipped2, rec2, q);
           add_fippi_terms(q, A.K, G0, G1); }
        if (N.hasSwap && ea == eb)
         { // both orders feed the diagonal edge pair: scatter_ppi doubles, so pass the mean
           H0 += 0.5 * (F0 + G0); H1 += 0.5 * (F1 + G1); }
        else if (needFwd) { H0 += F0; H1 += F1; }
        else { H0 += G0; H1 += G1; }
        scatter_ppi(A.S, A.K, N.pa, N.pb, ea, eb, Pa.sl[i], Pb.sl[j], H0, H1);
        nPPI += (A.S.symmetric && N.pa != N.pb && ea == eb) ? 2 : 1;
      }
   }
  if (nPPI) atomicAdd(A.counters + 4, nPPI);
}

// ---------------------------------------------------------------------------------------
// plan management
// ---------------------------------------------------------------------------------------
static int failc(scuffcuda_context *ctx, const char *msg, cudaError_t e)
{ char buf[512]; snprintf(buf, sizeof buf, "%s: %s", msg, cudaGetErrorString(e)); ctx->err = buf; return 1; }
#define CKN(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return failc(ctx, #call, e_); } while (0)


static int fippi_sync_device(scuffcuda_context *ctx)
{ FIPPITable &T = ctx->fippi;
  long n = (long)T.keys.size();
  if (n > T.dcap)
   { long cap = std::max(n * 2, 1024L);
     double *d; CKN(cudaMalloc(&d, (size_t)cap * FIPPI_REC_DOUBLES * sizeof(double)));
     if (T.drec) { CKN(cudaDeviceSynchronize()); CKN(cudaMemcpy(d, T.drec, (size_t)T.dcount * FIPPI_REC_DOUBLES * sizeof(double), cudaMemcpyDeviceToDevice)); CKN(cudaDeviceSynchronize()); cudaFree(T.drec); }
     T.drec = d; T.dcap = cap;
   }
  if (n > T.dcount)
   { // stream-ordered (and waited for): the consumers run on non-blocking streams, which the legacy stream does not order
     CKN(cudaMemcpyAsync(T.drec + (size_t)T.dcount * FIPPI_REC_DOUBLES, T.hrec.data() + (size_t)T.dcount * FIPPI_REC_DOUBLES,
                         (size_t)(n - T.dcount) * FIPPI_REC_DOUBLES * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
     CKN(cudaDeviceSynchronize());
     T.dcount = n; }
  return 0;
}

// FIPPI build partitioned over the ranks of a sweep: with partition (rank, world) set, a plan computes
// only the records whose index is congruent to rank mod world; the others stay zero until the caller has summed the
// tables of all ranks (identical geometry => identical record indices on every rank) and called scuffcuda_fippi_commit.
static std::map<scuffcuda_context *, std::pair<int, int> > g_fippiPartition;
static std::mutex g_fippiPartitionLock;
static Plan *get_plan(scuffcuda_context *ctx, int nsa, int nsb, bool symmetric, bool useDisp, const double *L, int &status)
{ status = 0;
  Surface &Sa = ctx->surfaces[nsa], &Sb = ctx->surfaces[nsb];
  for (size_t i = 0; i < ctx->plans.size(); i++)
   { Plan &P = ctx->plans[i];
     if (P.nsa == nsa && P.nsb == nsb && P.symmetric == symmetric && P.useDisp == useDisp
         && (!useDisp || (P.L[0] == L[0] && P.L[1] == L[1] && P.L[2] == L[2])))
      { if (P.genA == Sa.generation && P.genB == Sb.generation) return &P;
        if (P.dnear) cudaFree(P.dnear);
        ctx->plans.erase(ctx->plans.begin() + i); break; }
   }
  // drop plans that refer to stale geometry (bounded memory across Transform() sweeps)
  for (size_t i = 0; i < ctx->plans.size();)
   { Plan &P = ctx->plans[i];
     bool stale = P.nsa >= (int)ctx->surfaces.size() || P.nsb >= (int)ctx->surfaces.size()
                  || P.genA != ctx->surfaces[P.nsa].generation || P.genB != ctx->surfaces[P.nsb].generation;
     if (stale) { if (P.dnear) cudaFree(P.dnear); ctx->plans.erase(ctx->plans.begin() + i); } else i++; }

  Plan P; P.nsa = nsa; P.nsb = nsb; P.symmetric = symmetric; P.useDisp = useDisp;
  for (int i = 0; i < 3; i++) P.L[i] = useDisp ? L[i] : 0.0;
  P.genA = Sa.generation; P.genB = Sb.generation;

  // pass 1 (retry with a larger buffer if the guess was too small)
  unsigned long long cap = 64ull * std::max(Sa.NP, Sb.NP) + 1024, count = 0;
  int2 *dlist = nullptr; unsigned long long *dcount = ctx->dcounters + 8;
  int TA = 32; dim3 grid((Sb.NP + NEAR_THREADS - 1) / NEAR_THREADS, (Sa.NP + TA - 1) / TA);
  for (int attempt = 0; attempt < 2; attempt++)
   { if (cudaMalloc(&dlist, cap * sizeof(int2)) != cudaSuccess) { status = failc(ctx, "near list alloc", cudaGetLastError()); return nullptr; }
     cudaMemsetAsync(dcount, 0, sizeof(unsigned long long), ctx->stream);
     ctx->launches++;
     near_list_kernel<<<grid, NEAR_THREADS, TA * 4 * sizeof(double), ctx->stream>>>(Sa.dpanels, Sa.NP, Sb.dpanels, Sb.NP, TA, symmetric ? 1 : 0,
                                                                                    useDisp ? 1 : 0, P.L[0], P.L[1], P.L[2], dlist, dcount, cap);
     cudaMemcpyAsync(&count, dcount, sizeof count, cudaMemcpyDeviceToHost, ctx->stream);
     cudaError_t e = cudaStreamSynchronize(ctx->stream);
     if (e != cudaSuccess) { cudaFree(dlist); status = failc(ctx, "near_list_kernel", e); return nullptr; }
     if (count <= cap) break;
     cudaFree(dlist); dlist = nullptr; cap = count;
   }
  P.nNear = (int)count;
  if (count > 0)
   { // deterministic order: sort the (pa,pb) list on the host, then assess on the device
     std::vector<int2> hl(count);
     cudaMemcpy(hl.data(), dlist, count * sizeof(int2), cudaMemcpyDeviceToHost);
     std::sort(hl.begin(), hl.end(), [](const int2 &u, const int2 &v) { return u.x != v.x ? u.x < v.x : u.y < v.y; });
     cudaMemcpyAsync(dlist, hl.data(), count * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream);   // ordered before near_assess_kernel on the same stream
     float *dkeys = nullptr;
     if (cudaMalloc(&P.dnear, count * sizeof(NearPair)) != cudaSuccess || cudaMalloc(&dkeys, 2 * count * 15 * sizeof(float)) != cudaSuccess)
      { status = failc(ctx, "near plan alloc", cudaGetLastError()); cudaFree(dlist); return nullptr; }
     ctx->launches++;
     near_assess_kernel<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(Sa.dpanels, Sb.dpanels, dlist, (int)count, useDisp ? 1 : 0,
                                                                                  P.L[0], P.L[1], P.L[2], symmetric ? 1 : 0, P.dnear, dkeys, dkeys + count * 15);
     P.hnear.resize(count);
     std::vector<float> hkeys(2 * count * 15);
     cudaMemcpyAsync(P.hnear.data(), P.dnear, count * sizeof(NearPair), cudaMemcpyDeviceToHost, ctx->stream);
     cudaMemcpyAsync(hkeys.data(), dkeys, 2 * count * 15 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
     cudaError_t e = cudaStreamSynchronize(ctx->stream);
     cudaFree(dkeys);
     if (e != cudaSuccess) { cudaFree(dlist); status = failc(ctx, "near_assess_kernel", e); return nullptr; }

     // FIPPI record assignment: same key, same sharing semantics as FIPPICache::GetQIFIPPIData,
     // insertion in (pa,pb) order.  Records are needed for every pair with ncv < 3.
     FIPPITable &T = ctx->fippi;
     std::vector<int> newIds, newNcv; std::vector<double> newVerts;
     // insertion order mirrors a single-threaded reference run over (pa,pb): forward order first
     for (size_t i = 0; i < count; i++)
      { NearPair &N = P.hnear[i];
        if (N.ncv == 3) continue;
        bool need[2] = { true, false };
        if (N.hasSwap)
         { // which evaluation orders does the reference actually perform for this pair?
           const PanelRec &Qa = Sa.hpanels[N.pa], &Qb = Sb.hpanels[N.pb];
           need[0] = need[1] = false;
           for (int ii = 0; ii < 3; ii++) for (int jj = 0; jj < 3; jj++)
            { int ea = Qa.ei[ii], eb = Qb.ei[jj]; if (ea < 0 || eb < 0) continue;
              if (ea <= eb) need[0] = true;
              if (ea >= eb) need[1] = true; }
           if (!need[1]) N.hasSwap = 0;
         }
        for (int pass = 0; pass < 2; pass++)
         { if (!need[pass]) continue;
           FIPPIKey K; memcpy(K.k, &hkeys[15 * (i + (pass ? count : 0))], sizeof K.k);
           int &slot = pass ? N.rec2 : N.rec;
           auto it = T.index.find(K);
           if (it != T.index.end()) { slot = it->second; continue; }
           int id = (int)T.keys.size();
           T.index[K] = id; T.keys.push_back(K); T.hrec.resize((size_t)(id + 1) * FIPPI_REC_DOUBLES, 0.0);
           slot = id;
           // canonically ordered vertices of the (possibly swapped) pair that defines the record
           const PanelRec &Pa = Sa.hpanels[N.pa], &Pb = Sb.hpanels[N.pb];
           double Va[9], Vb[9];
           for (int q = 0; q < 9; q++) { Va[q] = Pa.V[q]; Vb[q] = useDisp ? (Pb.V[q] + 1.0 * P.L[q % 3]) : Pb.V[q]; }
           const double *VaO[3] = { Va + 3 * N.ia[0], Va + 3 * N.ia[1], Va + 3 * N.ia[2] };
           const double *VbO[3] = { Vb + 3 * N.ib[0], Vb + 3 * N.ib[1], Vb + 3 * N.ib[2] };
           const double *const *F = pass ? VbO : VaO, *const *S = pass ? VaO : VbO;   // first / second panel of this evaluation
           int fl = pass ? N.flipped2 : N.flipped;
           const unsigned char *oa = pass ? N.oa2 : N.oa, *ob = pass ? N.ob2 : N.ob;
           const double *const *A = fl ? S : F, *const *B = fl ? F : S;
           for (int v = 0; v < 3; v++) for (int m = 0; m < 3; m++) newVerts.push_back(A[oa[v]][m]);
           for (int v = 0; v < 3; v++) for (int m = 0; m < 3; m++) newVerts.push_back(B[ob[v]][m]);
           newIds.push_back(id); newNcv.push_back(N.ncv);
         }
        if (N.rec < 0) N.rec = N.rec2;   // forward order never evaluated: keep a valid index
      }
     cudaMemcpyAsync(P.dnear, P.hnear.data(), count * sizeof(NearPair), cudaMemcpyHostToDevice, ctx->stream);
     cudaStreamSynchronize(ctx->stream);          // consumers run on the near stream as well as on this one
     if (!newIds.empty())
      { std::pair<int, int> part(0, 1);
        { std::lock_guard<std::mutex> lock(g_fippiPartitionLock); auto it = g_fippiPartition.find(ctx); if (it != g_fippiPartition.end()) part = it->second; }
        if (part.second > 1)
         { std::vector<int> ids, ncv; std::vector<double> verts;
           for (size_t q = 0; q < newIds.size(); q++)
            if (newIds[q] % part.second == part.first)
             { ids.push_back(newIds[q]); ncv.push_back(newNcv[q]); verts.insert(verts.end(), newVerts.begin() + 18 * q, newVerts.begin() + 18 * (q + 1)); }
           newIds.swap(ids); newNcv.swap(ncv); newVerts.swap(verts); }
        if (fippi_compute_records(ctx, newIds, newVerts, newNcv)) { cudaFree(dlist); status = 1; return nullptr; } }
     if (fippi_sync_device(ctx)) { cudaFree(dlist); status = 1; return nullptr; }
   }
  if (dlist) cudaFree(dlist);
  ctx->plans.push_back(P);
  return &ctx->plans.back();
}


// the pair plan of a surface pair for other translation units (grad_kernel.cu)
Plan *near_plan(scuffcuda_context *ctx, int nsa, int nsb, bool symmetric, bool useDisp, const double *L, int &status)
{ return get_plan(ctx, nsa, nsb, symmetric, useDisp, L, status); }

int run_near(scuffcuda_context *ctx, int nsa, int nsb, const scuffcuda_ssi_args *a,
             const KParams &K, const StampParams &SP, int kmode, double kabs)
{ int status = 0;
  double L[3] = { a->displacement[0], a->displacement[1], a->displacement[2] };
  Plan *P = get_plan(ctx, nsa, nsb, a->symmetric != 0, a->use_displacement != 0, L, status);
  if (!P) return status ? status : 1;
  if (P->nNear == 0) return 0;
  NearArgs A;
  A.pa = ctx->surfaces[nsa].dpanels; A.pb = ctx->surfaces[nsb].dpanels; A.near = P->dnear; A.nNear = P->nNear;
  A.useDisp = a->use_displacement ? 1 : 0; for (int i = 0; i < 3; i++) A.L[i] = A.useDisp ? L[i] : 0.0;
  A.forceTD = a->force_taylor_duffy ? 1 : 0; A.kabs = kabs; A.fippi = ctx->fippi.drec;
  A.K = K; A.S = SP; A.counters = ctx->dcounters;
  // The desingularised pairs (near_desing_kernel) and the Taylor-Duffy pairs (run_td) are disjoint and both scatter with atomic
  // adds, so they run side by side: the first on the caller's (near) stream, the second on the Taylor-Duffy stream, which starts
  // where the near stream stands now (block cleared, far pass done) and is joined again before returning.  Nothing here blocks
  // the host before the kernels are queued (the pair count reaches the counters through the kernel), so both start the moment
  // the far pass ends.  SCUFF_CUDA_TD_STREAM=0: one stream, one after the other.
  static const bool twoStreams = !(getenv("SCUFF_CUDA_TD_STREAM") && getenv("SCUFF_CUDA_TD_STREAM")[0] == '0');
  cudaStream_t sNear = ctx->stream;
  if (twoStreams && ctx->tdStream)
   { CKN(cudaEventRecord(ctx->evTd[0], sNear));
     CKN(cudaStreamWaitEvent(ctx->tdStream, ctx->evTd[0], 0)); }
  ctx->launches++;
  near_desing_kernel<<<(P->nNear + NEAR_THREADS - 1) / NEAR_THREADS, NEAR_THREADS, 0, sNear>>>(A);
  CKN(cudaGetLastError());
  if (!(twoStreams && ctx->tdStream)) return run_td(ctx, *P, A, kmode);
  ctx->stream = ctx->tdStream;
  const int rc = run_td(ctx, *P, A, kmode);
  ctx->stream = sNear;
  if (rc) return rc;
  CKN(cudaEventRecord(ctx->evTd[1], ctx->tdStream));
  CKN(cudaStreamWaitEvent(sNear, ctx->evTd[1], 0));
  return 0;
}

}  // namespace scuffcuda

// ---------------------------------------------------------------------------------------
// FIPPI table <-> reference cache files (FIPPICache.cc:209-400): 11-byte signature, then records of
// { float Key[15] (+4 bytes padding), QIFIPPIData (36 doubles) } = 352 bytes
// ---------------------------------------------------------------------------------------
using namespace scuffcuda;
static const char FIPPICF_Signature[] = "FIPPICACHE";
struct FIPPICF_Record { float key[15]; float pad; double data[FIPPI_REC_DOUBLES]; };
static_assert(sizeof(FIPPICF_Record) == 352, "FIPPI cache record size");

extern "C" {

int scuffcuda_fippi_preload(scuffcuda_context *ctx, const char *filename)
{ FILE *f = fopen(filename, "rb");
  if (!f) { ctx->err = std::string("could not open FIPPI cache file ") + filename + " (skipping cache preload)"; return 1; }
  char sig[sizeof FIPPICF_Signature];
  if (fread(sig, sizeof sig, 1, f) != 1 || memcmp(sig, FIPPICF_Signature, sizeof sig))
   { fclose(f); ctx->err = std::string(filename) + ": invalid cache file (skipping cache preload)"; return 1; }
  fseek(f, 0, SEEK_END); long sz = ftell(f) - (long)sizeof sig; fseek(f, sizeof sig, SEEK_SET);
  if (sz % (long)sizeof(FIPPICF_Record)) { fclose(f); ctx->err = std::string(filename) + ": cache file has incorrect size (skipping cache preload)"; return 1; }
  long n = sz / (long)sizeof(FIPPICF_Record);
  FIPPITable &T = ctx->fippi;
  FIPPICF_Record R;
  for (long i = 0; i < n; i++)
   { if (fread(&R, sizeof R, 1, f) != 1) break;
     FIPPIKey K; memcpy(K.k, R.key, sizeof K.k);
     if (T.index.find(K) != T.index.end()) continue;      // unordered_map::insert keeps the first
     int id = (int)T.keys.size();
     T.index[K] = id; T.keys.push_back(K);
     T.hrec.insert(T.hrec.end(), R.data, R.data + FIPPI_REC_DOUBLES);
   }
  fclose(f);
  cudaSetDevice(ctx->device);
  return fippi_sync_device(ctx);
}

int scuffcuda_fippi_store(scuffcuda_context *ctx, const char *filename)
{ FILE *f = fopen(filename, "wb");
  if (!f) { ctx->err = std::string("could not open file ") + filename + " (aborting cache dump)"; return 1; }
  fwrite(FIPPICF_Signature, sizeof FIPPICF_Signature, 1, f);
  FIPPITable &T = ctx->fippi;
  FIPPICF_Record R; memset(&R, 0, sizeof R);
  for (size_t i = 0; i < T.keys.size(); i++)
   { memcpy(R.key, T.keys[i].k, sizeof R.key);
     memcpy(R.data, &T.hrec[i * FIPPI_REC_DOUBLES], sizeof R.data);
     fwrite(&R, sizeof R, 1, f); }
  fclose(f);
  return 0;
}

int scuffcuda_fippi_num_records(scuffcuda_context *ctx, long *n) { *n = (long)ctx->fippi.keys.size(); return 0; }

int scuffcuda_fippi_set_partition(scuffcuda_context *ctx, int rank, int world)
{ if (world < 1 || rank < 0 || rank >= world) { ctx->err = "fippi_set_partition: need 0 <= rank < world"; return 1; }
  std::lock_guard<std::mutex> lock(g_fippiPartitionLock);
  if (world == 1) g_fippiPartition.erase(ctx); else g_fippiPartition[ctx] = std::make_pair(rank, world);
  return 0;
}
int scuffcuda_prepare_block(scuffcuda_context *ctx, int nsa, int nsb, int symmetric, int use_displacement, const double *displacement)
{ const int NS = (int)ctx->surfaces.size();
  if (nsa < 0 || nsa >= NS || nsb < 0 || nsb >= NS || !ctx->surfaces[nsa].dpanels || !ctx->surfaces[nsb].dpanels)
   { ctx->err = "prepare_block: invalid surface index or surface not set"; return 1; }
  if (symmetric && nsa != nsb) { ctx->err = "prepare_block: symmetric requires nsa==nsb"; return 1; }
  cudaSetDevice(ctx->device);
  int status = 0; const double zero[3] = { 0, 0, 0 };
  Plan *P = get_plan(ctx, nsa, nsb, symmetric != 0, use_displacement != 0, (use_displacement && displacement) ? displacement : zero, status);
  return P ? 0 : (status ? status : 1);
}
int scuffcuda_fippi_host_records(scuffcuda_context *ctx, double **records, long *count)
{ if (!records || !count) { ctx->err = "fippi_host_records: null argument"; return 1; }
  *records = ctx->fippi.hrec.data(); *count = (long)ctx->fippi.keys.size();
  return 0;
}
int scuffcuda_fippi_commit(scuffcuda_context *ctx)
{ cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream);
  ctx->fippi.dcount = 0;                       // every record travels to the device again
  return fippi_sync_device(ctx);
}
int scuffcuda_fippi_clear(scuffcuda_context *ctx)
{ cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream);
  for (auto &p : ctx->plans) if (p.dnear) cudaFree(p.dnear);
  ctx->plans.clear();
  FIPPITable &T = ctx->fippi;
  T.index.clear(); T.keys.clear(); T.hrec.clear(); T.dcount = 0;
  return 0;
}

}  // extern "C"
