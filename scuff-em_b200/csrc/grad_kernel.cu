// Derivative blocks of the BEM matrix: GradM[Mu] = dM/dx_Mu for a rigid displacement of surface a against surface b and
// dMdT[n] = dM/dTheta for a rotation of surface a about torque axis n (AssembleBEMMatrixBlock's GradM / dMdT arguments,
// AssembleBEMMatrix.cc:349-391; scuff-cas3D asks for them on the U blocks of different objects, CasimirIntegrand.cc:443-447).
//
// What the reference computes (PanelPanelInteractions.cc:108-205, 420-561), restated:
//   * the derivative integrands of one cubature point pair are
//       GradH0_Mu = R_Mu hPlus Psi ,              GradH1_Mu = R_Mu hTimes Zeta + (F x F')_Mu Psi ,
//       dHdT0     = hPlus (R.dX) Psi + (dF.F') Phi ,   dHdT1 = hTimes (R.dX) Zeta + ((dF x F').R + (F x F').dX) Psi ,
//     with hPlus = F.F' + 4/(ik)^2, hTimes = (F x F').R, dX = Gamma X, dF = Gamma F, Psi = Phi (ik - 1/r)/r,
//     Zeta = Phi ((ik)^2 - 3ik/r + 3/r^2)/r^2;
//   * panel pairs with rRel > 4 use the 6x6 rule; every other pair that is not a Taylor-Duffy pair uses the 78x78 rule
//     WITHOUT desingularisation for the derivatives (":504-510: desingularization of derivative integrals is not implemented
//     and probably never will be"), also when the panels share vertices; Taylor-Duffy pairs contribute zero (:484-486);
//   * edge-edge sums, l_a l_b, 1/(ik) on the second component and the stamps with the region prefactors are those of the
//     matrix itself (EdgeEdgeInteractions.cc:190-198, SurfaceSurfaceInteractions.cc:205-327).
//
// These blocks are small (objects of a Casimir geometry) and not on the benchmarked path: the kernels keep the reference's
// per-(Q_a,Q_b) formulation and stamp with atomics.  far pairs: one thread per (panel pair, Q_a, Q_b) combination; near pairs
// (from the pair plan of near.cu): one warp per combination, lanes split the 78 outer points.
#include <vector>
#include <complex>
#include <mutex>
#include "kernels.h"
#include "classify.cuh"
#include "stamp.cuh"
#include "generated_tcr20.h"

namespace scuffcuda {

Plan *near_plan(scuffcuda_context *ctx, int nsa, int nsb, bool symmetric, bool useDisp, const double *L, int &status);   // near.cu

struct GradArgs
{ const PanelRec *pa, *pb; int NPa, NPb;
  const NearPair *near; int nNear;
  int useDisp; double L[3];
  KParams K; double kabs; int forceTD;
  StampParams S;                 // S.M is replaced by the block being stamped
  double *GM[3]; int nTorque; double *TM[3]; double gamma[27];
  const double *tcr20;           // 78 x (u, v, w)
};

__device__ __forceinline__ cd cmulr(cd a, double s) { return mk(a.x * s, a.y * s); }

// derivative integrals of one (Q_a, Q_b) combination over the outer points np = first, first + stride, ...
__device__ void grad_ppi(const double *tcr, int npts, int first, int stride, const double *Va, const double *Qa, const double *Vb, const double *Qb,
                         const KParams &K, int nTorque, const double *gamma, cd GradH[6], cd dHdT[6])
{ double A[3], B[3], AP[3], BP[3];
  for (int m = 0; m < 3; m++) { A[m] = Va[3 + m] - Va[m]; B[m] = Va[6 + m] - Va[m]; AP[m] = Vb[3 + m] - Vb[m]; BP[m] = Vb[6 + m] - Vb[m]; }
  const cd ik = K.ik, ik2 = ik * ik, c4 = K.c4;
  for (int q = 0; q < 6; q++) { GradH[q] = mk(0.0, 0.0); dHdT[q] = mk(0.0, 0.0); }
  for (int np = first; np < npts; np += stride)
   { const double u = tcr[3 * np], v = tcr[3 * np + 1], w = tcr[3 * np + 2];
     double X[3], F[3];
     for (int m = 0; m < 3; m++) { X[m] = Va[m] + u * A[m] + v * B[m]; F[m] = X[m] - Qa[m]; }
     cd gI[6], tI[6];
     for (int q = 0; q < 6; q++) { gI[q] = mk(0.0, 0.0); tI[q] = mk(0.0, 0.0); }
     for (int npp = 0; npp < npts; npp++)
      { const double up = tcr[3 * npp], vp = tcr[3 * npp + 1], wp = tcr[3 * npp + 2];
        double FP[3], R[3];
        for (int m = 0; m < 3; m++) { const double XP = Vb[m] + up * AP[m] + vp * BP[m]; FP[m] = XP - Qb[m]; R[m] = X[m] - XP; }
        const double r2 = R[0] * R[0] + R[1] * R[1] + R[2] * R[2], r = sqrt(r2);
        // Phi = w' exp(ikr)/(4 pi r)
        double s, c; sincos(K.kr * r, &s, &c);
        const double e = (K.ki == 0.0 ? 1.0 : exp(-K.ki * r)) * wp / (4.0 * M_PI * r);
        cd Phi = mk(e * c, e * s);
        if (!isfinite(Phi.x)) Phi = mk(0.0, 0.0);
        const double rinv = 1.0 / r;
        const cd Psi = cmulr(Phi * mk(ik.x - rinv, ik.y), rinv);
        const cd Zeta = cmulr(Phi * (ik2 - cmulr(ik, 3.0 * rinv) + mk(3.0 / r2, 0.0)), 1.0 / r2);
        const cd hPlus = mk(F[0] * FP[0] + F[1] * FP[1] + F[2] * FP[2], 0.0) + c4;
        double FxFP[3]; FxFP[0] = F[1] * FP[2] - F[2] * FP[1]; FxFP[1] = F[2] * FP[0] - F[0] * FP[2]; FxFP[2] = F[0] * FP[1] - F[1] * FP[0];
        const double hTimes = FxFP[0] * R[0] + FxFP[1] * R[1] + FxFP[2] * R[2];
        const cd hPsi = hPlus * Psi, tZeta = cmulr(Zeta, hTimes);
        for (int m = 0; m < 3; m++)
         { gI[2 * m] += cmulr(hPsi, R[m]);
           gI[2 * m + 1] += cmulr(tZeta, R[m]) + cmulr(Psi, FxFP[m]); }
        for (int nta = 0; nta < nTorque; nta++)
         { double dX[3] = { 0, 0, 0 }, dF[3] = { 0, 0, 0 };
           for (int m = 0; m < 3; m++) for (int n = 0; n < 3; n++) { dX[m] += gamma[9 * nta + m + 3 * n] * X[n]; dF[m] += gamma[9 * nta + m + 3 * n] * F[n]; }
           const double Puv = R[0] * dX[0] + R[1] * dX[1] + R[2] * dX[2];
           double dFxFP[3]; dFxFP[0] = dF[1] * FP[2] - dF[2] * FP[1]; dFxFP[1] = dF[2] * FP[0] - dF[0] * FP[2]; dFxFP[2] = dF[0] * FP[1] - dF[1] * FP[0];
           tI[2 * nta] += cmulr(hPsi, Puv) + cmulr(Phi, dF[0] * FP[0] + dF[1] * FP[1] + dF[2] * FP[2]);
           tI[2 * nta + 1] += cmulr(tZeta, Puv)
                              + cmulr(Psi, (dFxFP[0] * R[0] + dFxFP[1] * R[1] + dFxFP[2] * R[2]) + (FxFP[0] * dX[0] + FxFP[1] * dX[1] + FxFP[2] * dX[2])); } }
     for (int q = 0; q < 6; q++) { GradH[q] += cmulr(gI[q], w); dHdT[q] += cmulr(tI[q], w); } }
}

__device__ void grad_stamp(const GradArgs &A, int pa, int pb, int ea, int eb, double sla, double slb, const cd GradH[6], const cd dHdT[6])
{ StampParams SP = A.S;
  for (int m = 0; m < 3; m++)
   if (A.GM[m]) { SP.M = A.GM[m]; scatter_ppi(SP, A.K, pa, pb, ea, eb, sla, slb, GradH[2 * m], GradH[2 * m + 1]); }
  for (int n = 0; n < A.nTorque; n++)
   if (A.TM[n]) { SP.M = A.TM[n]; scatter_ppi(SP, A.K, pa, pb, ea, eb, sla, slb, dHdT[2 * n], dHdT[2 * n + 1]); }
}

// w_p, u_p, v_p of the 6-point rule in (u, v, w) triples (TCR.cc:83-91 literals)
__constant__ double c_grad_tcr4[18] = {
  0.0915762135098, 0.0915762135098, 0.0549758718276,  0.8168475729805, 0.0915762135098, 0.0549758718276,
  0.0915762135098, 0.8168475729805, 0.0549758718276,  0.1081030181681, 0.4459484909160, 0.1116907948390,
  0.4459484909160, 0.1081030181681, 0.1116907948390,  0.4459484909160, 0.4459484909160, 0.1116907948390 };

// far pairs: thread <-> (pa, pb, i, j)
__global__ void __launch_bounds__(128)
grad_far_kernel(const GradArgs A)
{ const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long nItems = (long)A.NPa * A.NPb * 9;
  if (t >= nItems) return;
  const int ij = (int)(t % 9), i = ij / 3, j = ij % 3;
  const long pr = t / 9;
  const int pa = (int)(pr % A.NPa), pb = (int)(pr / A.NPa);
  const PanelRec &Pa = A.pa[pa], &Pb = A.pb[pb];
  const int ea = Pa.ei[i], eb = Pb.ei[j];
  if (ea < 0 || eb < 0) return;
  double D[3];
  if (pair_rrel(Pa.C, Pa.R, Pb.C, Pb.R, A.useDisp != 0, A.L, D) <= SCUFF_DESING_RADIUS) return;
  double Va[9], Vb[9];
  for (int q = 0; q < 9; q++) { Va[q] = Pa.V[q]; Vb[q] = A.useDisp ? (Pb.V[q] + 1.0 * A.L[q % 3]) : Pb.V[q]; }
  cd GradH[6], dHdT[6];
  grad_ppi(c_grad_tcr4, 6, 0, 1, Va, Va + 3 * i, Vb, Vb + 3 * j, A.K, A.nTorque, A.gamma, GradH, dHdT);
  grad_stamp(A, pa, pb, ea, eb, Pa.sl[i], Pb.sl[j], GradH, dHdT);
}

// near pairs: warp <-> (near pair, i, j); 78 x 78 rule, lanes split the outer points
__global__ void __launch_bounds__(128)
grad_near_kernel(const GradArgs A)
{ const long w = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (long)A.nNear * 9) return;
  const int ij = (int)(w % 9), i = ij / 3, j = ij % 3;
  const NearPair N = A.near[w / 9];
  const PanelRec &Pa = A.pa[N.pa], &Pb = A.pb[N.pb];
  const int ea = Pa.ei[i], eb = Pb.ei[j];
  if (ea < 0 || eb < 0) return;
  bool verySW;
  if (near_algorithm(N.ncv, A.kabs, Pa.R, Pb.R, A.forceTD, verySW) == 2) return;      // Taylor-Duffy pairs: derivative integrals are zero
  double Va[9], Vb[9];
  for (int q = 0; q < 9; q++) { Va[q] = Pa.V[q]; Vb[q] = A.useDisp ? (Pb.V[q] + 1.0 * A.L[q % 3]) : Pb.V[q]; }
  cd GradH[6], dHdT[6];
  grad_ppi(A.tcr20, SC_TCR20_NPTS, lane, 32, Va, Va + 3 * i, Vb, Vb + 3 * j, A.K, A.nTorque, A.gamma, GradH, dHdT);
  for (int q = 0; q < 6; q++)
   for (int o = 16; o > 0; o >>= 1)
    { GradH[q].x += __shfl_xor_sync(0xffffffffu, GradH[q].x, o); GradH[q].y += __shfl_xor_sync(0xffffffffu, GradH[q].y, o);
      dHdT[q].x += __shfl_xor_sync(0xffffffffu, dHdT[q].x, o); dHdT[q].y += __shfl_xor_sync(0xffffffffu, dHdT[q].y, o); }
  if (lane == 0) grad_stamp(A, N.pa, N.pb, ea, eb, Pa.sl[i], Pb.sl[j], GradH, dHdT);
}

// one region pass: all panel pairs of the two surfaces
int run_gradients(scuffcuda_context *ctx, int nsa, int nsb, const scuffcuda_ssi_args *a, const KParams &K, const StampParams &SP, double kabs,
                  double *const GM[3], int nTorque, double *const TM[3], const double *gamma)
{ int status = 0;
  double L[3] = { a->displacement[0], a->displacement[1], a->displacement[2] };
  Plan *P = near_plan(ctx, nsa, nsb, false, a->use_displacement != 0, L, status);
  if (!P) return status ? status : 1;
  static double *dtcr[64] = { nullptr };
  static std::mutex dtcrLock;
  std::lock_guard<std::mutex> tlock(dtcrLock);
  double *&T = dtcr[ctx->device & 63];
  if (!T)
   { if (cudaMalloc(&T, sizeof SC_TCR20_H) != cudaSuccess) { ctx->err = "derivative blocks: cubature table allocation failed"; return 1; }
     cudaMemcpy(T, SC_TCR20_H, sizeof SC_TCR20_H, cudaMemcpyHostToDevice); cudaDeviceSynchronize(); }
  GradArgs A;
  A.pa = ctx->surfaces[nsa].dpanels; A.NPa = ctx->surfaces[nsa].NP; A.pb = ctx->surfaces[nsb].dpanels; A.NPb = ctx->surfaces[nsb].NP;
  A.near = P->dnear; A.nNear = P->nNear;
  A.useDisp = a->use_displacement ? 1 : 0; for (int i = 0; i < 3; i++) A.L[i] = A.useDisp ? L[i] : 0.0;
  A.K = K; A.kabs = kabs; A.forceTD = a->force_taylor_duffy ? 1 : 0; A.S = SP;
  for (int m = 0; m < 3; m++) { A.GM[m] = GM[m]; A.TM[m] = (m < nTorque) ? TM[m] : nullptr; }
  A.nTorque = nTorque;
  for (int q = 0; q < 27; q++) A.gamma[q] = (gamma && q < 9 * nTorque) ? gamma[q] : 0.0;
  A.tcr20 = T;
  const long nFar = (long)A.NPa * A.NPb * 9;
  ctx->launches++;
  grad_far_kernel<<<(unsigned)((nFar + 127) / 128), 128, 0, ctx->stream>>>(A);
  if (A.nNear > 0)
   { const long nW = (long)A.nNear * 9;
     ctx->launches++;
     grad_near_kernel<<<(unsigned)((nW * 32 + 127) / 128), 128, 0, ctx->stream>>>(A); }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { ctx->err = std::string("derivative-block kernels: ") + cudaGetErrorString(e); return 1; }
  return 0;
}

}  // namespace scuffcuda
