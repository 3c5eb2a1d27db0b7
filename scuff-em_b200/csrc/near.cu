// Near / singular panel pairs (rRel <= 4): plan construction, FIPPI table, desingularised
// cubature, Taylor-Duffy.  (stub: filled in below)
#include "kernels.h"
namespace scuffcuda {
int run_near(scuffcuda_context *ctx, int nsa, int nsb, const scuffcuda_ssi_args *args,
             const KParams &K, const StampParams &SP, int kmode, double kabs)
{ return 0; }
}
extern "C" {
int scuffcuda_fippi_preload(scuffcuda_context *ctx, const char *f) { return 0; }
int scuffcuda_fippi_store(scuffcuda_context *ctx, const char *f) { return 0; }
int scuffcuda_fippi_num_records(scuffcuda_context *ctx, long *n) { *n = (long)ctx->fippi.keys.size(); return 0; }
}
