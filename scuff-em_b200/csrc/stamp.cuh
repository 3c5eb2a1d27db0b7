// Scatter of one panel-pair result into the BEM matrix (device only).
// Reference: GetEdgeEdgeInteractions EdgeEdgeInteractions.cc:183-188 (l_a l_b (H++ - H+- - H-+ + H--)),
// GSSIThread stamping SurfaceSurfaceInteractions.cc:205-327, HMatrix::AddEntry HMatrix.cc:419-447
// (a complex value added to a real matrix keeps its real part).
#pragma once
#include "context.h"

namespace scuffcuda {

__device__ __forceinline__ void add_entry(const StampParams &P, int row, int col, cd v)
{ if (col < P.colLo || col >= P.colHi) return;
  size_t idx = (size_t)row + (size_t)col * (size_t)P.ld;
  if (P.isReal) atomicAdd(P.M + idx, v.x);
  else { atomicAdd(P.M + 2 * idx, v.x); atomicAdd(P.M + 2 * idx + 1, v.y); }
}

// g0 = l_a l_b (+-) H[0] ; g1 = l_a l_b (+-) H[1]/(ik)  for edge pair (ea,eb)
__device__ __forceinline__ void stamp_edge_pair(const StampParams &P, int ea, int eb, cd g0, cd g1, bool diag)
{ if (P.aPEC && P.bPEC)
   add_entry(P, P.rowOff + ea, P.colOff + eb, P.c1 * g0);
  else if (P.aPEC && !P.bPEC)
   { int X = P.rowOff + ea, Y = P.colOff + 2 * eb;
     add_entry(P, X, Y, P.c1 * g0); add_entry(P, X, Y + 1, P.c2 * g1); }
  else if (!P.aPEC && P.bPEC)
   { int X = P.rowOff + 2 * ea, Y = P.colOff + eb;
     add_entry(P, X, Y, P.c1 * g0); add_entry(P, X + 1, Y, P.c2 * g1); }
  else
   { const int X = P.rowOff + 2 * ea, Y = P.colOff + 2 * eb;
     const cd v11 = P.c1 * g0, v12 = P.c2 * g1, v22 = P.c3 * g0;
     // the 2x2 PMCHWT stamp as a rolled loop (keeps the nine inlined copies of this function small: the dielectric far
     // kernel was instruction-fetch bound): e = 0..3 -> (X,Y), (X,Y+1), (X+1,Y), (X+1,Y+1)
#pragma unroll 1
     for (int e = 0; e < 4; e++)
      { if (e == 2 && diag) continue;
        const cd v = (e == 0) ? v11 : ((e == 3) ? v22 : v12);
        add_entry(P, X + (e >> 1), Y + (e & 1), v); }
   }
}

// Scatter the contribution of the PPI (pa,iQa ; pb,iQb) = (H0,H1) to its edge pair, handling
// the upper-triangle convention of Symmetric blocks (only nea<=neb is stored; an unordered
// panel pair p!=q visited once feeds (e,f) or (f,e), twice the value when e==f).
template <bool NEED_G1 = true>
__device__ __forceinline__ void scatter_ppi(const StampParams &P, const KParams &K,
                                            int pa, int pb, int ea, int eb, double sla, double slb,
                                            cd H0, cd H1)
{ if (ea < 0 || eb < 0) return;
  double s = sla * slb;
  bool diag = false;
  if (P.symmetric)
   { if (pa == pb) { if (ea > eb) return; }
     else if (ea == eb) s *= 2.0;
     else if (ea > eb) { int t = ea; ea = eb; eb = t; }
     diag = (ea == eb);
   }
  cd g0 = s * H0;
  cd g1 = NEED_G1 ? s * (K.ooik * H1) : mk(0.0, 0.0);   // PEC/PEC blocks stamp GC[0] only
  stamp_edge_pair(P, ea, eb, g0, g1, diag);
}

}  // namespace scuffcuda
