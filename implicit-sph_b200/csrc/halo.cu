// Multi-GPU plumbing (row partition along LAMMPS subdomains): halo plan, NCCL send/recv of ghost columns, allreduce.
// Placeholder in this revision: single-GPU only, fails loudly for nranks > 1.
#include "isph_internal.h"
namespace isph {
void halo_setup(Ctx *) { ISPH_REQUIRE(false, "nranks > 1: halo exchange not built yet in this revision"); }
void halo_exchange(Ctx *, double *, int, int) { ISPH_REQUIRE(false, "nranks > 1 not built yet"); }
void halo_allreduce(Ctx *, double *, int) { ISPH_REQUIRE(false, "nranks > 1 not built yet"); }
void halo_forward_field(Ctx *, int, int) { ISPH_REQUIRE(false, "nranks > 1 not built yet"); }
void halo_destroy(Ctx *) {}
int halo_ncols(Ctx *c) { return c->nlocal; }
}  // namespace isph
extern "C" int isph_nccl_unique_id(void *) { return ISPH_FAILURE; }
