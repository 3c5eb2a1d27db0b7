// FIPPI builder, integration case 1 (see fippi_build.cuh)
#include "fippi_build.cuh"
namespace scuffcuda { FIPPI_DEFINE_LAUNCH(1) }
