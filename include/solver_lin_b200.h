// solver_lin_b200.h — header-only C++ adapter: the reference's SolverLin / PrecondWrapper method names on top of the
// C ABI (isph_b200.h), so that the call sites in pair_isph.cpp (computeIncompressibleNavierStokes :910-1034,
// computeAppliedElectricField :635-657, computeSoluteTransport :811-835) and USER-REAXC-T/fix_qeq_reax.cpp:671-693
// read the same after the switch.  Same names, argument meaning, ownership (everything borrowed) and error behaviour
// (int LAMMPS_SUCCESS / LAMMPS_FAILURE returns; non-convergence is reported, not an error) as
//   IMPLICIT-SPH/solver_lin.h:23-98, solver_lin.cpp:30-160, solver_lin_belos.h:35-49,130-264,
//   IMPLICIT-SPH/precond.h:17-46, precond_ifpack.h:14-85.
// Differences that the type system forces (no Trilinos types cross the boundary):
//   * Epetra_Map* / Epetra_CrsMatrix* arguments become the isph_ctx handle that owns the device-resident map+matrix
//     (built by isph_graph_build / the assembly entry points, or uploaded with isph_matrix_set_csr);
//   * Teuchos::ParameterList* becomes set(name, value) calls with the same Belos / Ifpack key names;
//   * getLoadMultiVector()->Values() becomes loadSet()/loadGet() because b lives in HBM.
#pragma once
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include "isph_b200.h"

namespace isph_b200 {

#ifndef LAMMPS_SUCCESS
#define LAMMPS_SUCCESS 0
#endif
#ifndef LAMMPS_FAILURE
#define LAMMPS_FAILURE (-1)
#endif

// PrecondWrapper / PrecondWrapper_Ifpack (precond.h:17-46, precond_ifpack.h:14-85)
class PrecondWrapper_B200 {
 public:
  explicit PrecondWrapper_B200(isph_ctx *ctx) : _ctx(ctx) {}
  virtual ~PrecondWrapper_B200() {}
  // setMatrix(Epetra_CrsMatrix*) : the matrix is the context's device matrix; kept for call-site compatibility
  virtual void setMatrix(isph_ctx *ctx) { if (ctx) _ctx = ctx; }
  // setParameters(NULL) loads the wrapper defaults of precond_ifpack.h:30-44 restricted to what BASELINE names (fill 0, overlap 0).
  // The reference's own values (level-of-fill 1, overlap 1) can be set explicitly: level-of-fill k > 0 is supported, overlap 1 across
  // ranks too (ILU, one block per rank: additive Schwarz with combine mode Add); see isph_b200.h.
  virtual void setParameters() { set("Precond Type", "ILU"); set("Overlap Level", 0); set("fact: level-of-fill", 0); }
  int set(const char *name, int v) { return isph_precond_set_param_int(_ctx, name, v); }
  int set(const char *name, double v) { return isph_precond_set_param_double(_ctx, name, v); }
  int set(const char *name, const char *v) { return isph_precond_set_param_str(_ctx, name, v); }
  virtual void setNullVector(double *) {}                       // ML only in the reference (precond_ml.h); no-op like the base class
  virtual void create() { check(isph_precond_create(_ctx), "PrecondWrapper::create"); }
  virtual void free() { isph_precond_free(_ctx); }
  isph_ctx *getPrecondOperator() { return _ctx; }
  // Ifpack factors one block per MPI rank; name the rank-equivalent brick of every local row to reproduce a CPU run
  int setBlocks(const int *block_of_row) { return isph_precond_set_blocks(_ctx, block_of_row); }
 protected:
  void check(int rc, const char *what) { if (rc != ISPH_SUCCESS) throw std::runtime_error(std::string(what) + ": " + isph_last_error(_ctx)); }
  isph_ctx *_ctx;
};

// PrecondWrapper_ML (precond_ml.h:17-172) — the reference's default package (pair_isph.cpp:325-329,359-361).  The multilevel
// preconditioner of csrc/amg.cu stands in for ML ("parity unpinned": ML is un-vendored third-party code; see isph_b200.h for what is
// and is not ML's algorithm).  Same method set as the reference's class; ML's parameter names are kept.
class PrecondWrapper_ML_B200 : public PrecondWrapper_B200 {
 public:
  explicit PrecondWrapper_ML_B200(isph_ctx *ctx) : PrecondWrapper_B200(ctx) {}
  // setParameters(NULL), precond_ml.h:44-58: "max levels" 5, "aggregation: type" Uncoupled, smoother 1 sweep pre and post, coarse solve.
  // One deviation, forced by the device: "smoother: type" Chebyshev — the alternative the reference's own benchmark file keeps
  // commented next to its Gauss-Seidel line (bench-script/hopper/tgv/4096/ml.xml:15-18) — instead of symmetric Gauss-Seidel (sequential
  // within a rank).  "coarse: type" Amesos-KLU = a direct solve of the coarsest operator; for singular problems it is replaced by the
  // smoother, as PrecondWrapper_ML::setNullVector does (:118-120).
  virtual void setParameters() {
    set("Precond Package", "ML"); set("max levels", 5); set("increasing or decreasing", "increasing"); set("aggregation: type", "Uncoupled");
    set("smoother: type", "Chebyshev"); set("smoother: pre or post", "both"); set("coarse: type", "Amesos-KLU");
  }
  // for zoltan re-partition (precond_ml.h:65-99): one GPU per rank, the coarse levels are replicated — nothing to repartition
  void setCoordinates(const int, double *, double *, double *) {}
  // precond_ml.h:101-126: ML is told the null vector and switches the coarse solver to the smoother.  Here the coarse solver already is
  // the smoother and the tentative prolongator is piecewise constant (the null vector of every singular system of the reference is the
  // normalised 0/1 mask of solver_lin.cpp:59-77, which piecewise constants span inside the masked rows)
  virtual void setNullVector(double *) {}
  virtual void create(const int /*dim*/) { create(); }          // block variant: one hierarchy, applied to every diagonal block (precond_ml.h:137-154)
  virtual void create() { set("Precond Package", "ML"); PrecondWrapper_B200::create(); }
};

// SolverLin / SolverLin_Belos (solver_lin.h:23-98, solver_lin_belos.h:35-49)
class SolverLin_B200 {
 public:
  enum SolutionInitType { Random = ISPH_INIT_RANDOM, Zero = ISPH_INIT_ZERO, Value = ISPH_INIT_VALUE };   // solver_lin.h:25

  explicit SolverLin_B200(isph_ctx *ctx) : _ctx(ctx), _is_singular(false) {}
  virtual ~SolverLin_B200() {}

  // solver_lin.h:31-39
  int createLoadMultiVector(double *b, int lda, int num_vectors) { return rc(isph_solver_create_load_multivector(_ctx, b, lda, num_vectors)); }
  int createSolutionMultiVector(double *x, int lda, int num_vectors) { return rc(isph_solver_create_solution_multivector(_ctx, x, lda, num_vectors)); }
  // solver_lin.h:46-52
  void setNullVectorMask(const int *mask) { isph_solver_set_null_vector_mask(_ctx, mask); }
  void setMatrixIsSingular(const bool is_singular) { _is_singular = is_singular; isph_solver_set_matrix_is_singular(_ctx, is_singular ? 1 : 0); }
  void setNodalMap(isph_ctx *ctx) { if (ctx) _ctx = ctx; }        // the nodal map is part of the context (isph_atoms_set)
  void setMatrix(isph_ctx *ctx) { if (ctx) _ctx = ctx; }
  // solver_lin.h:58
  void setInitialSolution(SolutionInitType init, double val = 0.0) { isph_solver_set_initial_solution(_ctx, (int)init, val); }
  // getLoadMultiVector()->Values() replacement (b is device resident)
  int loadSet(const double *b, int lda) { return rc(isph_solver_load_set(_ctx, b, lda)); }
  int loadGet(double *b, int lda) { return rc(isph_solver_load_get(_ctx, b, lda)); }

  // setParameters(NULL): the hard-coded Belos list of solver_lin_belos.h:224-264
  virtual void setParameters() { isph_solver_set_default_params(_ctx); }
  int set(const char *name, int v) { return isph_solver_set_param_int(_ctx, name, v); }
  int set(const char *name, double v) { return isph_solver_set_param_double(_ctx, name, v); }
  int set(const char *name, const char *v) { return isph_solver_set_param_str(_ctx, name, v); }

  // SolverLin_Belos::solveProblem(prec, name), solver_lin_belos.h:130-222 — always returns LAMMPS_SUCCESS unless the
  // device path itself failed; prints the Belos-style status lines on rank 0 (the ABI prints the residual on failure)
  virtual int solveProblem(PrecondWrapper_B200 *prec = NULL, const char *name = NULL) {
    if (name != NULL) std::printf(">> isph_b200::Label - %s\n", name);
    const int r = isph_solver_solve(_ctx, prec != NULL ? 1 : 0, name);
    if (r != ISPH_SUCCESS) { std::fprintf(stderr, ">> isph_b200 error: %s\n", isph_last_error(_ctx)); return LAMMPS_FAILURE; }
    int iters = 0, conv = 0; double relres = 0.0;
    isph_solver_stats(_ctx, &iters, &relres, &conv, NULL);
    if (conv) std::printf(">> isph_b200::Status - Passed! %s (%d iterations, %.3e)\n", name ? name : " ", iters, relres);
    return LAMMPS_SUCCESS;
  }
  int iterations() const { int it = 0; isph_solver_stats(_ctx, &it, NULL, NULL, NULL); return it; }

  // block interface, solver_lin.h:43-56 / solver_lin.cpp:78-138; a block = CSR over the nodal map instead of an Epetra_CrsMatrix*
  int createBlockMatrix(const int dim, const char *name) { return rc(isph_solver_create_block_matrix(_ctx, dim, name)); }
  int freeBlockMatrix() { return rc(isph_solver_free_block_matrix(_ctx)); }
  void setMatrixIsBlocked(const bool) {}
  void setBlockBegin() {}
  void setBlock(const int i, const int j, int n, const int *rowptr, const int *col, const double *val) { isph_solver_set_block_csr(_ctx, i, j, n, rowptr, col, val); }
  void setBlockEnd() { isph_solver_set_block_end(_ctx); }
  // SolverLin_Belos::solveBlockProblem(prec, name), solver_lin_belos.h:53-128
  virtual int solveBlockProblem(PrecondWrapper_B200 *prec = NULL, const char *name = NULL) {
    if (name != NULL) std::printf(">> isph_b200(Block)::Label - %s\n", name);
    const int r = isph_solver_solve_block(_ctx, prec != NULL ? 1 : 0, name);
    if (r != ISPH_SUCCESS) { std::fprintf(stderr, ">> isph_b200 error: %s\n", isph_last_error(_ctx)); return LAMMPS_FAILURE; }
    int iters = 0, conv = 0; double relres = 0.0; isph_solver_stats(_ctx, &iters, &relres, &conv, NULL);
    if (conv) std::printf(">> isph_b200::Status - Passed! %s (%d iterations, %.3e)\n", name ? name : " ", iters, relres);
    return LAMMPS_SUCCESS;
  }

 protected:
  static int rc(int r) { return r == ISPH_SUCCESS ? LAMMPS_SUCCESS : LAMMPS_FAILURE; }
  isph_ctx *_ctx;
  bool _is_singular;
};

// SolverNOX<PairISPH> as PairISPH::computePoissonBoltzmann uses it (pair_isph.cpp:572-600, solver_nox.h, solver_nox_impl.h:26-200):
// the Newton iteration with the reference's default lists (line search "Full Step", NormF 1e-8 AND NormUpdate 1e-5, 100
// iterations) runs on the device (isph_pb_newton); computeF / computeJacobian are no longer host callbacks.
class SolverNonlinear_B200 {
 public:
  enum SolutionInitType { Random, Zero, Value };                // solver_nox.h SolutionInitType
  enum JacobianType { Analytic, MatrixFree };                   // only Analytic (pair_isph.cpp:373) is provided

  explicit SolverNonlinear_B200(isph_ctx *ctx) : _ctx(ctx), _psi(NULL), _nlocal(0), _nall(0), _mh(0), _lin(0), _ezcb(0.5), _psiref(1.0), _gamma(0.0),
                                                 _extra(NULL), _max_newton(100), _tol_f(1.0e-8), _tol_update(1.0e-5), _newton(0), _linear(0), _normf(0.0) {}
  void setNodalMap(isph_ctx *ctx) { if (ctx) _ctx = ctx; }
  void setJacobianMatrix(isph_ctx *ctx) { if (ctx) _ctx = ctx; }
  // createSolutionVector(atom->psi): a View like the reference's; the ghost part (nall - nlocal values) is refreshed on return
  int createSolutionVector(double *psi, int nlocal, int nall) { _psi = psi; _nlocal = nlocal; _nall = nall; return LAMMPS_SUCCESS; }
  void setInitialSolution(SolutionInitType init, double val = 0.0) {
    if (!_psi) return;
    unsigned long long s = 37482ull;                            // seed of pair_isph.cpp:1454 ; Epetra's generator itself is not restated
    for (int i = 0; i < _nall; ++i) {
      if (init == Zero) _psi[i] = 0.0; else if (init == Value) _psi[i] = val;
      else { s = s * 6364136223846793005ull + 1442695040888963407ull; _psi[i] = 2.0 * ((double)(s >> 11) * (1.0 / 9007199254740992.0)) - 1.0; }
    }
  }
  void setParameters() {}                                       // "Full Step" Newton: what isph_pb_newton does
  void setConvergenceTests(int max_iters = 100, double normf = 1.0e-8, double normupdate = 1.0e-5) { _max_newton = max_iters; _tol_f = normf; _tol_update = normupdate; }
  // pair->pb.* (pair_isph.h) and the evaluated pb.extra_f expression (functor_poisson_boltzmann_extra_f.h), per owned particle or NULL
  void setPoissonBoltzmann(bool morris_holmes, bool linearized, double ezcb, double psiref, double gamma, const double *extra_f) {
    _mh = morris_holmes; _lin = linearized; _ezcb = ezcb; _psiref = psiref; _gamma = gamma; _extra = extra_f;
  }
  int solveProblem(PrecondWrapper_B200 *prec = NULL, const char *name = NULL) {
    if (!_psi) return LAMMPS_FAILURE;
    if (name != NULL) std::printf(">> isph_b200::Label - %s\n", name);
    int conv = 0;
    if (isph_field_set(_ctx, ISPH_F_PSI, _psi) != ISPH_SUCCESS ||
        isph_solver_create_solution_multivector(_ctx, NULL, _nlocal, 1) != ISPH_SUCCESS || isph_solver_create_load_multivector(_ctx, NULL, _nlocal, 1) != ISPH_SUCCESS ||
        isph_pb_newton(_ctx, _mh, _lin, _ezcb, _psiref, _gamma, _extra, _max_newton, _tol_f, _tol_update, prec != NULL ? 1 : 0, &_newton, &_linear, &_normf, &conv) != ISPH_SUCCESS ||
        isph_field_get(_ctx, ISPH_F_PSI, _psi) != ISPH_SUCCESS) { std::fprintf(stderr, ">> isph_b200 error: %s\n", isph_last_error(_ctx)); return LAMMPS_FAILURE; }
    std::printf(">> isph_b200::Status - %s %s (%d Newton / %d linear iterations, ||F|| = %.3e)\n", conv ? "Passed!" : "Failed to converge!", name ? name : " ", _newton, _linear, _normf);
    return LAMMPS_SUCCESS;                                       // non-convergence is reported, not an error (as for the linear solver)
  }
  int newtonIterations() const { return _newton; }
  int linearIterations() const { return _linear; }
 protected:
  isph_ctx *_ctx; double *_psi; int _nlocal, _nall; int _mh, _lin; double _ezcb, _psiref, _gamma; const double *_extra;
  int _max_newton; double _tol_f, _tol_update; int _newton, _linear; double _normf;
};

}  // namespace isph_b200
