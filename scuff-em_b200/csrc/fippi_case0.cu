// FIPPI builder, integration case 0 (see fippi_build.cuh)
#include "fippi_build.cuh"
namespace scuffcuda { FIPPI_DEFINE_LAUNCH(0) }
