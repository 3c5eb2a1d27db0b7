// Tricubic interpolation table of the periodic Green's function and its evaluation (device side of the reference's
// GBarAccelerator: GetGBar_2D GBarAccelerator.cc:823-978, Interp3D::EvaluatePlus libs/libMDInterp/Interp3D.cc:698-795,
// FindInterval BinSearch.cc:93-150, coefficient fit Interp3D.cc:385-505).
#pragma once
#include "gbar.cuh"

namespace scuffcuda {

// device view of one table.  Coefficients: 64 doubles per (cell, function); function 0 = Re GBar, 1 = Im GBar; the
// coefficient of t1^p t2^q t3^r (cell-normalised coordinates) sits at 16p + 4q + r, cells ordered n3 + (N3-1)(n2 + (N2-1) n1)
struct GBarTableView
{ GBarParams P;
  double RhoMin, RhoMax, LMax;   // range test of GetGBar_2D (the ORIGINAL RhoMin/RhoMax, even when the grid was widened)
  int forceFull;                 // no table: every evaluation is a full Ewald sum
  int N[3];
  double XMin[3], DX[3];
  const double *C;
};

// tricubic Hermite fit of one cell: D[corner][8] holds {f, fx, fy, fz, fxy, fxz, fyz, fxyz} at the corners ordered
// (dn1, dn2, dn3) with dn3 fastest; the interpolant that matches these 64 numbers is unique, so the tensor product of
// 1-D cubic Hermite bases is the solution of the reference's 64x64 system (Interp3D.cc:397-488)
SCUFF_HD void tricubic_fit(const double D[8][8], const double DX[3], double C[64])
{ // 1-D Hermite basis polynomials, coefficients of t^0..t^3: h[corner][derivative][power]
  const double h[2][2][4] = { { { 1.0, 0.0, -3.0, 2.0 }, { 0.0, 1.0, -2.0, 1.0 } },
                              { { 0.0, 0.0, 3.0, -2.0 }, { 0.0, 0.0, -1.0, 1.0 } } };
  // derivative data index -> (d1,d2,d3)
  const int dd[8][3] = { { 0, 0, 0 }, { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 }, { 1, 1, 0 }, { 1, 0, 1 }, { 0, 1, 1 }, { 1, 1, 1 } };
  for (int n = 0; n < 64; n++) C[n] = 0.0;
  for (int c = 0; c < 8; c++)
   { const int c1 = c >> 2, c2 = (c >> 1) & 1, c3 = c & 1;
     for (int d = 0; d < 8; d++)
      { const int d1 = dd[d][0], d2 = dd[d][1], d3 = dd[d][2];
        const double v = D[c][d] * (d1 ? DX[0] : 1.0) * (d2 ? DX[1] : 1.0) * (d3 ? DX[2] : 1.0);
        if (v == 0.0) continue;
        for (int p = 0; p < 4; p++)
         { const double a = h[c1][d1][p]; if (a == 0.0) continue;
           for (int q = 0; q < 4; q++)
            { const double b = h[c2][d2][q]; if (b == 0.0) continue;
              for (int r = 0; r < 4; r++) C[16 * p + 4 * q + r] += v * a * b * h[c3][d3][r]; } } } }
}

// FindInterval on a uniform grid (BinSearch.cc:117-141)
SCUFF_HD void find_interval(double X, int NumPoints, double XMin, double DX, int &nCell, double &XBar)
{ if (X <= XMin) { nCell = 0; XBar = 0.0; }
  else if (X >= XMin + ((double)(NumPoints - 1)) * DX) { nCell = NumPoints - 2; XBar = 1.0; }
  else
   { XBar = (X - XMin) / DX;
     nCell = (int)trunc(XBar);
     XBar -= (double)nCell;
     if (nCell == NumPoints - 1) { nCell = NumPoints - 2; XBar = 1.0; } }
}

// value and gradient of both tabulated functions at (X1,X2,X3): out = { f, fx, fy, fz } for f = Re GBar, then Im GBar
SCUFF_HD void tricubic_eval_plus(const int N[3], const double XMin[3], const double DX[3], const double *CTable,
                                 double X1, double X2, double X3, double out[8])
{ int n1, n2, n3; double t1, t2, t3;
  find_interval(X1, N[0], XMin[0], DX[0], n1, t1);
  find_interval(X2, N[1], XMin[1], DX[1], n2, t2);
  find_interval(X3, N[2], XMin[2], DX[2], n3, t3);
  const double a[4] = { 1.0, t1, t1 * t1, t1 * t1 * t1 }, b[4] = { 1.0, t2, t2 * t2, t2 * t2 * t2 }, c[4] = { 1.0, t3, t3 * t3, t3 * t3 * t3 };
  const double da[4] = { 0.0, 1.0 / DX[0], 2.0 * t1 / DX[0], 3.0 * t1 * t1 / DX[0] };
  const double db[4] = { 0.0, 1.0 / DX[1], 2.0 * t2 / DX[1], 3.0 * t2 * t2 / DX[1] };
  const double dc[4] = { 0.0, 1.0 / DX[2], 2.0 * t3 / DX[2], 3.0 * t3 * t3 / DX[2] };
  const double *C0 = CTable + (size_t)64 * 2 * (n3 + (N[2] - 1) * (n2 + (N[1] - 1) * n1));
  for (int nf = 0; nf < 2; nf++)
   { const double *C = C0 + 64 * nf;
     double P = 0.0, PX = 0.0, PY = 0.0, PZ = 0.0;
     for (int p = 0; p < 4; p++)
      { double s = 0.0, sy = 0.0, sz = 0.0;
        for (int q = 0; q < 4; q++)
         { const double *Cq = C + 16 * p + 4 * q;
           const double u = Cq[0] * c[0] + Cq[1] * c[1] + Cq[2] * c[2] + Cq[3] * c[3];
           const double uz = Cq[1] * dc[1] + Cq[2] * dc[2] + Cq[3] * dc[3];
           s += u * b[q]; sy += u * db[q]; sz += uz * b[q]; }
        P += s * a[p]; PX += s * da[p]; PY += sy * a[p]; PZ += sz * a[p]; }
     out[4 * nf + 0] = P; out[4 * nf + 1] = PX; out[4 * nf + 2] = PY; out[4 * nf + 3] = PZ; }
}

// GFull term with first derivatives only (AddGFullTerm, GBarAccelerator.cc:572-607)
SCUFF_HD void gbar_add_full_term(const double R[3], cd k, const double kBloch[2], double Lx, double Ly, cd T[4])
{ double s, c; sincos(Lx * kBloch[0] + Ly * kBloch[1], &s, &c);
  const cd Phase = mk(c, s);
  const double X = R[0] - Lx, Y = R[1] - Ly, Z = R[2];
  const double r2 = X * X + Y * Y + Z * Z, r = sqrt(r2);
  if (r < 1.0e-8) return;
  const cd IKR = mk(-k.y * r, k.x * r);
  const cd Phi = gb_cexp(IKR) * (1.0 / (4.0 * M_PI * r));
  const cd Psi = (mk(IKR.x - 1.0, IKR.y) * Phi) * (1.0 / r2);
  const cd PPsi = Phase * Psi;
  T[0] += Phase * Phi; T[1] += PPsi * X; T[2] += PPsi * Y; T[3] += PPsi * Z;
}


// ---- grid-spacing probe (GetOptimalGridSpacing2D, GBarAccelerator.cc:257-333) ----
// sample points of one probe at (x,y,z): the 8 corners of the one-cell trial grid [x,x+D0]x[y,y+D1]x[z,z+D2] (corner
// c = 4 dn1 + 2 dn2 + dn3) followed by the three mid-edge points; Delta as the reference chooses it
SCUFF_HD void spacing_probe_setup(const GBarParams &P, double RhoMin, double RhoMax, const double X[3], double Delta[3], double DX[3], double Pts[11][3])
{ Delta[0] = P.LBV[0][0] / 10.0; Delta[1] = P.LBV[1][1] / 10.0; Delta[2] = (RhoMax - RhoMin) / 10.0;
  if (Delta[2] <= 0.0) Delta[2] = Delta[0];
  for (int m = 0; m < 3; m++) DX[m] = ((X[m] + Delta[m]) - X[m]) / 1.0;          // Interp3D: (XMax - XMin)/(N-1)
  for (int c = 0; c < 8; c++)
   { const int dn[3] = { c >> 2, (c >> 1) & 1, c & 1 };
     for (int m = 0; m < 3; m++) Pts[c][m] = X[m] + dn[m] * DX[m]; }
  for (int Mu = 0; Mu < 3; Mu++)
   { for (int m = 0; m < 3; m++) Pts[8 + Mu][m] = X[m];
     Pts[8 + Mu][Mu] += 0.5 * Delta[Mu]; }
}
// V[11][8]: GBarVD at the 11 sample points.  Returns the optimal spacings.
SCUFF_HD void spacing_probe_finish(const double X[3], const double Delta[3], const double DX[3], const cd V[11][SC_GB_NSUM], double RelTol, double Opt[3])
{ double C[2][64], D[8][8];
  for (int nf = 0; nf < 2; nf++)
   { for (int c = 0; c < 8; c++) for (int d = 0; d < 8; d++) D[c][d] = nf ? V[c][d].y : V[c][d].x;
     tricubic_fit(D, DX, C[nf]); }
  const int N[3] = { 2, 2, 2 };
  for (int Mu = 0; Mu < 3; Mu++)
   { double R[3] = { X[0], X[1], X[2] }; R[Mu] += 0.5 * Delta[Mu];
     double Phi[8];
     tricubic_eval_plus(N, X, DX, &C[0][0], R[0], R[1], R[2], Phi);
     const cd GE = V[8 + Mu][0], GI = mk(Phi[0], Phi[4]);
     const double aGE = gb_cabs(GE);
     double RelError = gb_cabs(GI - GE) / aGE;
     Opt[Mu] = Delta[Mu] * pow(RelTol / RelError, 0.25);
     for (int Nu = 0; Nu < 3; Nu++)
      { const cd dE = V[8 + Mu][1 + Nu], dI = mk(Phi[1 + Nu], Phi[5 + Nu]);
        if (gb_cabs(dE) < 1.0e-6 * aGE) continue;
        RelError = gb_cabs(dI - dE) / gb_cabs(dE);
        Opt[Mu] = fmin(Opt[Mu], Delta[Mu] * pow(RelTol / RelError, 0.33)); } }
}


// ---- table layout decisions (CreateGBarAccelerator, GBarAccelerator.cc:375-382,492-542), shared by gbar.cu and the
//      host check in tests/host_harness.cc ----
struct GBarGrid { double LMax; double Lx, Ly; int N[3]; double XMin[3], DX[3]; };
inline void gbar_grid_lmax(const GBarParams &P, double RhoMin, GBarGrid &G)
{ G.LMax = 1.0e87;
  const double Kappa = P.k.y;
  if (Kappa != 0.0) { const double GoK = 14.0 / Kappa; G.LMax = sqrt(GoK * GoK + 2.0 * RhoMin * GoK); }
  G.Lx = P.LBV[0][0]; G.Ly = P.LBV[1][1];
  if (G.LMax < G.Lx) G.Lx = G.LMax;
  if (G.LMax < G.Ly) G.Ly = G.LMax;
}
// probe points (x,y,z) of the spacing estimate: 4 at RhoMin, 4 more at RhoMax if the range is not empty
inline int gbar_grid_probes(const GBarGrid &G, double RhoMin, double RhoMax, double probes[8][3])
{ const double px[4] = { 0.0 * G.Lx, 0.0 * G.Lx, -0.5 * G.Lx, -0.5 * G.Lx }, py[4] = { 0.0 * G.Ly, -0.5 * G.Ly, 0.0 * G.Ly, -0.5 * G.Ly };
  int n = 0;
  for (int i = 0; i < 4; i++, n++) { probes[n][0] = px[i]; probes[n][1] = py[i]; probes[n][2] = RhoMin; }
  if (RhoMax != RhoMin) for (int i = 0; i < 4; i++, n++) { probes[n][0] = px[i]; probes[n][1] = py[i]; probes[n][2] = RhoMax; }
  return n;
}
inline void gbar_grid_dims(GBarGrid &G, double RhoMin, double RhoMax, const double MinDelta[3], int nMaxGrid)
{ int nx = (int)ceil(G.Lx / MinDelta[0]); if (nx < 2) nx = 2;
  int ny = (int)ceil(G.Ly / MinDelta[1]); if (ny < 2) ny = 2;
  int nRho; double GridRhoMax = RhoMax;
  if (RhoMax <= RhoMin) { GridRhoMax = RhoMin + MinDelta[0]; nRho = 2; }
  else { nRho = (int)ceil((RhoMax - RhoMin) / MinDelta[1]); if (nRho < 2) nRho = 2; }       // [1], as the reference (:528)
  if (nMaxGrid > 0) { if (nx > nMaxGrid) nx = nMaxGrid; if (ny > nMaxGrid) ny = nMaxGrid; if (nRho > nMaxGrid) nRho = nMaxGrid; }
  G.N[0] = nx; G.N[1] = ny; G.N[2] = nRho;
  G.XMin[0] = -0.5 * G.Lx; G.XMin[1] = -0.5 * G.Ly; G.XMin[2] = RhoMin;
  G.DX[0] = (0.5 * G.Lx - G.XMin[0]) / ((double)(nx - 1));
  G.DX[1] = (0.5 * G.Ly - G.XMin[1]) / ((double)(ny - 1));
  G.DX[2] = (GridRhoMax - RhoMin) / ((double)(nRho - 1));
}

SCUFF_HD bool in_m101(int x) { return -1 <= x && x <= 1; }

// GetGBar (value + gradient): table lookup with Wigner-Seitz wrap, Bloch phase and add-back of the cells that the wrap
// moves in or out of the excluded inner 3x3 block; full Ewald sum outside the table's Rho range
SCUFF_HD void gbar_get(const GBarTableView &T, const double R[3], cd &G, cd dG[3])
{ const double Rho = fabs(R[2]);
  if (T.forceFull || Rho < T.RhoMin || Rho > T.RhoMax)
   { cd V[SC_GB_NSUM]; gbar_vd_ewald(R, T.P, V);
     G = V[0]; dG[0] = V[1]; dG[1] = V[2]; dG[2] = V[3]; return; }
  const double L0x = T.P.LBV[0][0], L0y = T.P.LBV[1][1];
  const int mx = (int)lround(R[0] / L0x), my = (int)lround(R[1] / L0y);
  const double xBar = R[0] - mx * L0x, yBar = R[1] - my * L0y;
  if (fabs(xBar) > T.LMax || fabs(yBar) > T.LMax)
   { G = mk(0.0, 0.0); dG[0] = dG[1] = dG[2] = G; return; }
  double PhiVD[8];
  tricubic_eval_plus(T.N, T.XMin, T.DX, T.C, xBar, yBar, Rho, PhiVD);
  G = mk(PhiVD[0], PhiVD[4]); dG[0] = mk(PhiVD[1], PhiVD[5]); dG[1] = mk(PhiVD[2], PhiVD[6]); dG[2] = mk(PhiVD[3], PhiVD[7]);
  if (mx != 0 || my != 0)
   { const double ExpArg = T.P.kBloch[0] * (mx * T.P.LBV[0][0] + my * T.P.LBV[1][0]) + T.P.kBloch[1] * (mx * T.P.LBV[0][1] + my * T.P.LBV[1][1]);
     double s, c; sincos(ExpArg, &s, &c);
     const cd Phase = mk(c, s);
     if (T.P.excludeInner)
      { cd TP[4], TM[4];
        for (int i = 0; i < 4; i++) { TP[i] = mk(0.0, 0.0); TM[i] = mk(0.0, 0.0); }
        const double RBar[3] = { xBar, yBar, Rho };
        for (int nx = -1; nx <= 1; nx++)
         for (int ny = -1; ny <= 1; ny++)
          { if (!(in_m101(nx + mx) && in_m101(ny + my))) gbar_add_full_term(RBar, T.P.k, T.P.kBloch, nx * L0x, ny * L0y, TP);
            if (!(in_m101(nx - mx) && in_m101(ny - my))) gbar_add_full_term(RBar, T.P.k, T.P.kBloch, (nx - mx) * L0x, (ny - my) * L0y, TM); }
        G += TP[0] - TM[0];
        for (int Mu = 0; Mu < 3; Mu++) dG[Mu] += TP[Mu + 1] - TM[Mu + 1]; }
     G = G * Phase;
     for (int Mu = 0; Mu < 3; Mu++) dG[Mu] = dG[Mu] * Phase; }
  if (!(R[2] > 0.0)) dG[2] = -dG[2];
}

}  // namespace scuffcuda
