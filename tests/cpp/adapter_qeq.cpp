// The SECOND client of the SolverLin API (USER-REAXC-T/fix_qeq_reax.cpp:509-694): the caller assembles its own Epetra_CrsMatrix
// (here: a symmetric positive definite charge-equilibration-like matrix H + diag(eta) on a random point set, built with the same
// Epetra call sequence — graph, SumIntoGlobalValues, FillComplete) and only SOLVES through SolverLin / PrecondWrapper, twice, with
// the caller's arrays as Views for x (initial guess in, solution out) and b.  Compiled against the Epetra stand-ins of
// oracle/ref_shim (test infrastructure) and the Epetra-typed adapter include/solver_lin_b200_epetra.h; the solver block itself is
// the call sequence of fix_qeq_reax.cpp:671-693.  Prints the residuals of both solves.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "mpi.h"
#include "Epetra_CrsMatrix.h"
#define ISPH_B200_REPLACE_TRILINOS_SOLVERS
#include "solver_lin_b200_epetra.h"

using namespace LAMMPS_NS;

int main() {
  MPI_Comm world = MPI_COMM_WORLD;
  const int n = 1500; const double swb = 0.16;
  std::vector<double> x(3 * n); std::vector<int> tag(n);
  unsigned long long s_ = 12345;
  for (int i = 0; i < 3 * n; ++i) { s_ = s_ * 6364136223846793005ull + 1442695040888963407ull; x[i] = (double)(s_ >> 11) / 9007199254740992.0; }
  for (int i = 0; i < n; ++i) tag[i] = i + 1;
  Epetra_Map nodalmap(-1, n, tag.data(), 1, Epetra_MpiComm());
  // neighbor pairs within swb
  std::vector<std::vector<int> > nb(n);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) if (i != j) {
    const double dx = x[3 * j] - x[3 * i], dy = x[3 * j + 1] - x[3 * i + 1], dz = x[3 * j + 2] - x[3 * i + 2];
    if (dx * dx + dy * dy + dz * dz <= swb * swb) nb[i].push_back(j);
  }
  Epetra_IntSerialDenseVector row(n); int mx = 0;
  for (int i = 0; i < n; ++i) { row[i] = (int)nb[i].size() + 1; mx = row[i] > mx ? row[i] : mx; }
  Epetra_CrsGraph graph(Copy, nodalmap, row.Values(), true);
  std::vector<int> idx(mx); std::vector<double> val(mx);
  for (int i = 0; i < n; ++i) { int cnt = 0; for (size_t k = 0; k < nb[i].size(); ++k) idx[cnt++] = tag[nb[i][k]]; idx[cnt++] = tag[i]; graph.InsertGlobalIndices(tag[i], cnt, idx.data()); }
  graph.FillComplete(); graph.OptimizeStorage();
  Epetra_CrsMatrix AA(Copy, graph);
  for (int i = 0; i < n; ++i) {
    int cnt = 0; double off = 0.0;
    for (size_t k = 0; k < nb[i].size(); ++k) {
      const int j = nb[i][k]; const double dx = x[3 * j] - x[3 * i], dy = x[3 * j + 1] - x[3 * i + 1], dz = x[3 * j + 2] - x[3 * i + 2];
      const double r = std::sqrt(dx * dx + dy * dy + dz * dz), h = 0.3 * (1.0 - r / swb) * (1.0 - r / swb);      // a short-ranged, symmetric interaction
      idx[cnt] = tag[j]; val[cnt] = h; off += h; ++cnt;
    }
    idx[cnt] = tag[i]; val[cnt] = 1.0 + off; ++cnt;                                                          // eta + row sum: diagonally dominant, SPD
    AA.SumIntoGlobalValues(tag[i], cnt, val.data(), idx.data());
  }
  AA.FillComplete(); AA.OptimizeStorage();

  std::vector<double> s(n, 0.1), t(n, -0.2), b_s(n), b_t(n, -1.0);
  for (int i = 0; i < n; ++i) b_s[i] = -(0.5 + 0.1 * std::sin(7.0 * x[3 * i]));

  PrecondWrapper_ML prec(world);
  prec.setParameters();

  // ---- the solver block, call for call as in fix_qeq_reax.cpp:671-693
  SolverLin_Belos li_solver(world);
  li_solver.setParameters();
  li_solver.setNodalMap(&nodalmap);
  li_solver.setMatrix(&AA);
  prec.setMatrix(&AA);
  {
    li_solver.createSolutionMultiVector(s.data(), n, 1);
    li_solver.createLoadMultiVector(b_s.data(), n, 1);
    if (li_solver.solveProblem(&prec, "fix_qeq_reax:: b_s, s") != LAMMPS_SUCCESS) return 1;
  }
  const int it_s = li_solver.iterations();
  {
    li_solver.createSolutionMultiVector(t.data(), n, 1);
    li_solver.createLoadMultiVector(b_t.data(), n, 1);
    if (li_solver.solveProblem(&prec, "fix_qeq_reax:: b_t, t") != LAMMPS_SUCCESS) return 1;
  }
  // ---- check both solutions against the Epetra (stand-in) matrix on the host
  Epetra_Vector xs(View, nodalmap, s.data()), xt(View, nodalmap, t.data()), ys(nodalmap), yt(nodalmap);
  AA.Multiply(false, xs, ys); AA.Multiply(false, xt, yt);
  double rs = 0, rt = 0, ns = 0, nt = 0;
  for (int i = 0; i < n; ++i) { rs += (ys[i] - b_s[i]) * (ys[i] - b_s[i]); ns += b_s[i] * b_s[i]; rt += (yt[i] - b_t[i]) * (yt[i] - b_t[i]); nt += b_t[i] * b_t[i]; }
  std::printf("residuals %.3e %.3e iterations %d %d\n", std::sqrt(rs / ns), std::sqrt(rt / nt), it_s, li_solver.iterations());
  return (std::sqrt(rs / ns) <= 1e-7 && std::sqrt(rt / nt) <= 1e-7) ? 0 : 4;
}
