// C++ test of scuff::SweepDriver (include/libscuff_b200.h): the Mie sweep of the reference's tests (tests/Mie/OmegaFile, 15
// frequencies, SiO2Sphere_501 PMCHWT) sharded over W workers (one host thread + device context each), against the same sweep on
// ONE worker.  Per point: Frobenius norm, trace and one LU-solved unknown of the assembled matrix.  Built and run by
// tests/test_sweep_driver_gpu.py:   sweep_driver_test <geo> <OmegaFile> <W> [dev0 dev1 ...]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "libscuff_b200.h"
using namespace scuff;

struct Result { double fro, trRe, trIm, x0Re, x0Im; double pabs, pscat, fz; int worker; };

static void Point(RWGGeometry *G, HMatrix *M, cdouble Omega, double *kBloch, int j, int w, void *ud)
{ (void)G; (void)Omega; (void)kBloch;
  Result *R = (Result *)ud + j;
  M->SyncToHost();
  double fro = 0; cdouble tr = 0;
  for (int c = 0; c < M->NC; c++) for (int r = 0; r < M->NR; r++) { cdouble v = M->GetEntry(r, c); fro += std::norm(v); if (r == c) tr += v; }
  R->fro = sqrt(fro); R->trRe = real(tr); R->trIm = imag(tr); R->worker = w;
  // the downstream consumer on the same device: LU + one solve
  HMatrix X(M->NR, 1, LHM_COMPLEX);
  for (int r = 0; r < M->NR; r++) X.SetEntry(r, 0, cdouble(1.0 + 0.001 * r, 0.5));
  M->LUFactorize(); M->LUSolve(&X);
  R->x0Re = real(X.GetEntry(0, 0)); R->x0Im = imag(X.GetEntry(0, 0));
  // the rest of a scuff-scatter frequency point on the worker's device: plane-wave right-hand side, solve on the
  // device-resident vector, energy/momentum-transfer power and force — the loop body of scuff-scatter.cc:466-520
  cdouble E0[3] = { 1.0, 0.0, 0.0 }; double nHat[3] = { 0.0, 0.0, 1.0 };
  PlaneWave PW(E0, nHat);
  HVector *KN = G->AssembleRHSVector(Omega, &PW, 0);
  M->LUSolve(KN);
  PFTOptions Opts; InitPFTOptions(&Opts); Opts.IF = &PW; Opts.EMTPFTIMethod = SCUFF_EMTPFTI_EHDERIVATIVES1;
  HMatrix *PFT = G->GetPFTMatrix(KN, Omega, &Opts, 0);
  R->pabs = PFT->GetEntryD(0, PFT_PABS); R->pscat = PFT->GetEntryD(0, PFT_PSCAT); R->fz = PFT->GetEntryD(0, PFT_ZFORCE);
  delete PFT; delete KN;
}

int main(int argc, char **argv)
{ if (argc < 4) { fprintf(stderr, "usage: sweep_driver_test geo OmegaFile W [devices]\n"); return 2; }
  std::vector<cdouble> Omegas;
  { FILE *f = fopen(argv[2], "r"); if (!f) return 2; double w; while (fscanf(f, "%lf", &w) == 1) Omegas.push_back(cdouble(w, 0.0)); fclose(f); }
  const int NP = (int)Omegas.size(), W = atoi(argv[3]);
  std::vector<int> dev; for (int i = 4; i < argc; i++) dev.push_back(atoi(argv[i]));
  while ((int)dev.size() < W) dev.push_back(dev.empty() ? 0 : dev.back());
  std::vector<Result> one(NP), many(NP);
  double cold1, coldW0, coldW1;
  { SweepDriver D(argv[1], 1, dev.data()); D.Run(NP, Omegas.data(), 0, Point, one.data()); cold1 = D.ColdSeconds[0]; }
  { SweepDriver D(argv[1], W, dev.data());
    D.Run(NP, Omegas.data(), 0, Point, many.data());
    coldW0 = D.ColdSeconds[0]; coldW1 = D.ColdSeconds[W - 1];
    D.Run(3, Omegas.data(), 0, Point, many.data()); }          // a second sweep on the same driver reuses contexts, plans and table
  double worst = 0; int counts[16] = { 0 };
  for (int j = 0; j < NP; j++)
   { worst = fmax(worst, fabs(many[j].fro - one[j].fro) / one[j].fro);
     worst = fmax(worst, hypot(many[j].trRe - one[j].trRe, many[j].trIm - one[j].trIm) / hypot(one[j].trRe, one[j].trIm));
     worst = fmax(worst, hypot(many[j].x0Re - one[j].x0Re, many[j].x0Im - one[j].x0Im) / hypot(one[j].x0Re, one[j].x0Im));
     worst = fmax(worst, fabs(many[j].pscat - one[j].pscat) / fabs(one[j].pscat));
     worst = fmax(worst, fabs(many[j].pabs - one[j].pabs) / fabs(one[j].pscat));
     worst = fmax(worst, fabs(many[j].fz - one[j].fz) / fabs(one[j].fz));
     if (many[j].worker != j % W) { printf("{\"error\": \"point %d ran on worker %d\"}\n", j, many[j].worker); return 1; }
     counts[many[j].worker & 15]++; }
  printf("{\"pscat_first\": %.6e, \"pscat_last\": %.6e, ", many[0].pscat, many[NP - 1].pscat);
  printf("\"points\": %d, \"workers\": %d, \"worst_rel_diff\": %.3e, \"cold_s_one_worker\": %.3f, \"cold_s_first\": %.3f, \"cold_s_last\": %.3f, \"points_worker0\": %d}\n",
         NP, W, worst, cold1, coldW0, coldW1, counts[0]);
  return worst <= 1e-9 ? 0 : 1;
}
