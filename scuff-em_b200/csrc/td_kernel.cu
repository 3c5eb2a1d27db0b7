// Taylor-Duffy evaluation of panel pairs with common vertices (stub; filled in below)
#include "kernels.h"
namespace scuffcuda {
int run_td(scuffcuda_context *ctx, const Plan &P, const NearArgs &A, int kmode) { return 0; }
}
