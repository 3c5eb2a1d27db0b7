// Far-field panel-pair kernel: every panel pair with rRel > 4 (the reference's
// PPIALG_LOCUBATURE branch, PanelPanelInteractions.cc:420-424) is integrated ONCE with the 6x6
// order-4 rule into Q-independent moments (ppi_math.cuh) and scattered, with RWG signs and edge
// lengths, into the up-to-nine edge-pair entries it feeds (stamp.cuh).
//
// Mapping: one thread per panel pair.  A CTA owns a tile of TA a-panels (staged in shared memory,
// broadcast reads) x 128 b-panels (one per lane, kept in registers for the whole tile).
// FP64 DFMA-bound; panel data is 288 B per panel and is read once per tile.
#include "context.h"
#include "classify.cuh"
#include "stamp.cuh"
#include "kernels.h"

namespace scuffcuda {

#define FAR_THREADS 128

template <int KMODE>
__global__ void __launch_bounds__(FAR_THREADS)
far_kernel(const FarArgs A)
{ extern __shared__ double sm[];
  const int TA = A.TA;
  const int a0 = blockIdx.y * TA;
  const int bBase = blockIdx.x * FAR_THREADS;
  // symmetric: skip tiles entirely below the diagonal
  if (A.symmetric && bBase + FAR_THREADS - 1 < a0) return;
  const int nA = min(TA, A.NPa - a0);

  // stage the a-tile (nA * 36 doubles) in shared memory
  { const double *src = (const double *)(A.pa + a0);
    const int n = nA * 36;
    for (int i = threadIdx.x; i < n; i += FAR_THREADS) sm[i] = src[i];
  }
  __syncthreads();

  const int pb = bBase + threadIdx.x;
  if (pb >= A.NPb) return;
  if (A.symmetric && pb < a0) return;

  // this lane's b-panel
  const PanelRec *Pb = A.pb + pb;
  double Cb[3], Rb, xb[18], qb[9], slb[3]; int eib[3];
#pragma unroll
  for (int i = 0; i < 3; i++) Cb[i] = Pb->C[i];
  Rb = Pb->R;
#pragma unroll
  for (int i = 0; i < 18; i++) xb[i] = Pb->x[i];
#pragma unroll
  for (int i = 0; i < 9; i++) qb[i] = Pb->V[i] - Cb[i % 3];
#pragma unroll
  for (int i = 0; i < 3; i++) { slb[i] = Pb->sl[i]; eib[i] = Pb->ei[i]; }

  typedef typename KScalar<KMODE>::type S;
  unsigned long long nPPI = 0; unsigned int nPairs = 0;

  for (int ia = 0; ia < nA; ia++)
   { const int pa = a0 + ia;
     if (A.symmetric && pb < pa) continue;
     const double *Pa = sm + ia * 36;
     double D[3];
     const double rRel = pair_rrel(Pa, Pa[3], Cb, Rb, A.useDisp != 0, A.L, D);
     if (!(rRel > SCUFF_DESING_RADIUS)) continue;

     nPairs++;
     Moments<S> M;
     panel_pair_moments<KMODE, false>(Pa + 13, xb, D, A.K, M);

     const int *eia = (const int *)(Pa + 34);
#pragma unroll
     for (int i = 0; i < 3; i++)
      { const int ea = eia[i];
        if (ea < 0) continue;
        double qa[3] = { Pa[4 + 3 * i] - Pa[0], Pa[5 + 3 * i] - Pa[1], Pa[6 + 3 * i] - Pa[2] };
        const double sla = Pa[31 + i];
#pragma unroll
        for (int j = 0; j < 3; j++)
         { if (eib[j] < 0) continue;
           cd H0, H1;
           contract_moments(M, qa, qb + 3 * j, D, A.K, H0, H1);
           scatter_ppi(A.S, A.K, pa, pb, ea, eib[j], sla, slb[j], H0, H1);
           nPPI += (A.symmetric && pa != pb && ea == eib[j]) ? 2 : 1;
         }
      }
   }
  // algorithm counter (reference log line SurfaceSurfaceInteractions.cc:550-558)
  if (nPPI) atomicAdd(A.counters, nPPI);
  if (nPairs) atomicAdd(A.counters + 5, (unsigned long long)nPairs);
}

void launch_far(scuffcuda_context *ctx, const FarArgs &Ain, int kmode)
{ FarArgs A = Ain;
  // tile height: enough CTAs to fill the machine on small meshes, 32 on large ones
  int bTiles = (A.NPb + FAR_THREADS - 1) / FAR_THREADS;
  int TA = 32;
  while (TA > 2 && (long)bTiles * ((A.NPa + TA - 1) / TA) < 4L * ctx->smCount) TA >>= 1;
  A.TA = TA;
  dim3 grid(bTiles, (A.NPa + TA - 1) / TA);
  size_t smem = (size_t)TA * sizeof(PanelRec);
  ctx->launches++;
  if (kmode == KMODE_REAL) far_kernel<KMODE_REAL><<<grid, FAR_THREADS, smem, ctx->stream>>>(A);
  else if (kmode == KMODE_IMAG) far_kernel<KMODE_IMAG><<<grid, FAR_THREADS, smem, ctx->stream>>>(A);
  else far_kernel<KMODE_CPLX><<<grid, FAR_THREADS, smem, ctx->stream>>>(A);
}

}  // namespace scuffcuda
