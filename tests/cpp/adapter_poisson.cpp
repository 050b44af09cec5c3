// Drives the drop-in boundary from C++ exactly in the order of PairISPH::compute (pair_isph.cpp:1241-1380) and the
// Poisson block of PairISPH::computeIncompressibleNavierStokes (pair_isph.cpp:986-1026), through the adapter classes
// of include/solver_lin_b200.h.  Reads one particle set from a flat binary file written by tests/test_gpu_adapter.py,
// hands LAMMPS-layout arrays (numneigh / firstneigh pages) to the ABI, writes dp.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <string>
#include "solver_lin_b200.h"

using namespace isph_b200;

template <class T> static std::vector<T> rd(FILE *f, size_t n) { std::vector<T> v(n); if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } return v; }
#define CK(e) do { if ((e) != ISPH_SUCCESS) { fprintf(stderr, "%s failed: %s\n", #e, isph_last_error(ctx)); return 1; } } while (0)

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  FILE *f = fopen(argv[1], "rb"); if (!f) return 2;
  int hdr[6]; if (fread(hdr, sizeof(int), 6, f) != 6) return 2;
  const int dim = hdr[0], nlocal = hdr[1], nghost = hdr[2], nall = nlocal + nghost; const long long nneigh = ((long long)hdr[4] << 31) | hdr[3];
  double hdt[2]; if (fread(hdt, sizeof(double), 2, f) != 2) return 2;          // h, dt
  std::vector<double> x = rd<double>(f, 3 * (size_t)nall), vstar = rd<double>(f, 3 * (size_t)nall), rho = rd<double>(f, nall);
  std::vector<int> type = rd<int>(f, nall), tag = rd<int>(f, nall), numneigh = rd<int>(f, nlocal), neigh = rd<int>(f, (size_t)nneigh);
  fclose(f);
  // LAMMPS NeighList: ilist, numneigh[i], firstneigh[i] pointing into pages
  std::vector<int> ilist(nlocal); std::vector<int *> firstneigh(nall, nullptr); std::vector<int> nn(nall, 0);
  { size_t off = 0; for (int i = 0; i < nlocal; ++i) { ilist[i] = i; nn[i] = numneigh[i]; firstneigh[i] = neigh.data() + off; off += numneigh[i]; } }

  isph_ctx *ctx = NULL;
  if (isph_ctx_create(&ctx, 0, 1, 0, NULL) != ISPH_SUCCESS) { fprintf(stderr, "no GPU context\n"); return 3; }
  const int kinds[2] = {0, ISPH_KIND_FLUID};
  CK(isph_pair_coeff(ctx, dim, 1, kinds, hdt[0], hdt[0], 2.0, ISPH_KERNEL_WENDLAND, 0.43301));       // pair_coeff * * file.xml h
  CK(isph_atoms_set(ctx, nlocal, nghost, x.data(), type.data(), tag.data()));                        // grow(); atom arrays
  CK(isph_neighbors_set(ctx, nlocal, ilist.data(), nn.data(), firstneigh.data()));                   // list->...
  CK(isph_field_set(ctx, ISPH_F_VSTAR, vstar.data())); CK(isph_field_set(ctx, ISPH_F_DENSITY, rho.data()));
  CK(isph_compute_volumes(ctx)); CK(isph_compute_gradient_correction(ctx)); CK(isph_compute_laplacian_correction(ctx));   // computePre()
  CK(isph_graph_build(ctx));                                                                         // nodal map + computeGraph + new CrsMatrix

  // pair_isph.cpp:325-329: the wrapper class follows "Precond Package" (default "ML")
  const std::string prec_package = argc > 3 ? argv[3] : "Ifpack";
  PrecondWrapper_B200 *precp = NULL;
  if (prec_package == "Ifpack") { precp = new PrecondWrapper_B200(ctx); precp->set("Precond Type", "point relaxation"); precp->set("Overlap Level", 0); }
  else if (prec_package == "ML") { precp = new PrecondWrapper_ML_B200(ctx); precp->setParameters(); precp->setNullVector(NULL); }
  else { fprintf(stderr, "Preconditioner is not in supported list: Ifpack, ML\n"); return 4; }
  PrecondWrapper_B200 &prec = *precp;
  SolverLin_B200 li_solver(ctx); li_solver.setParameters();
  li_solver.setNodalMap(ctx); li_solver.setMatrix(ctx); prec.setMatrix(ctx);                         // pair_isph.cpp:924-926
  std::vector<double> dp(nlocal, 0.0);
  li_solver.createSolutionMultiVector(dp.data(), nlocal, 1);                                         // :988
  li_solver.createLoadMultiVector(NULL, nlocal, 1);                                                  // :989
  CK(isph_ns_poisson(ctx, hdt[1], 1, ISPH_NULLSPACE, 0));                                            // computePoisson(b->Values()), :993
  std::vector<int> null_mask(nlocal);
  for (int i = 0; i < nlocal; ++i) null_mask[i] = !(type[i] & ISPH_KIND_SOLID);                      // :998-1001 (tests the raw type number)
  li_solver.setNullVectorMask(null_mask.data()); li_solver.setMatrixIsSingular(true);                // :1003-1004
  li_solver.setInitialSolution(SolverLin_B200::Zero);                                                // :1010
  if (li_solver.solveProblem(&prec, "Poisson") != LAMMPS_SUCCESS) return 1;                          // :1011
  li_solver.setMatrixIsSingular(false);                                                              // :1015
  CK(isph_matrix_invalidate(ctx));                                                                   // A.is_filled = 0, :1026
  printf("iterations %d\n", li_solver.iterations());
  FILE *o = fopen(argv[2], "wb"); fwrite(dp.data(), sizeof(double), nlocal, o); fclose(o);
  delete precp;
  isph_ctx_destroy(ctx);
  return 0;
}
