// ILU(0) block-Jacobi preconditioner (Ifpack_ILU level-of-fill 0, overlap 0) — see DESIGN.md.  Placeholder until the
// level-scheduled factorisation lands; fails loudly instead of falling back to anything else.
#include "isph_internal.h"
namespace isph {
void ilu_create(Ctx *) { ISPH_REQUIRE(false, "Precond Type ILU: not built yet in this revision"); }
void ilu_free(Ctx *) {}
void ilu_apply(Ctx *, const double *, double *) { ISPH_REQUIRE(false, "Precond Type ILU: not built yet in this revision"); }
}  // namespace isph
