// FIPPI builder, integration case 2 (see fippi_build.cuh)
#include "fippi_build.cuh"
namespace scuffcuda { FIPPI_DEFINE_LAUNCH(2) }
