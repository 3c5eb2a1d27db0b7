// Parameters of one Bloch-periodic Green's function (shared by the device code and the context bookkeeping)
#pragma once
#include "ppi_math.cuh"

namespace scuffcuda {

struct GBarParams
{ cd k;                 // wavenumber in the region
  double kBloch[2];
  double LBV[2][2];     // lattice basis vectors (x,y components)
  int LDim;             // 2 (1-D lattices are refused by the launcher)
  int excludeInner;     // omit the 3x3 innermost cells (they are assembled directly, AssembleBEMMatrix.cc:465-519)
};

}  // namespace scuffcuda
