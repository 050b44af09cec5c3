// solver_lin_b200_epetra.h — the reference's OWN call-site signatures on top of the C ABI.
//
// include/solver_lin_b200.h takes an isph_ctx where the reference takes MPI_Comm / Epetra_Map* / Epetra_CrsMatrix*.  This header
// closes that gap for callers that hold Epetra objects: the classes below have the constructor and method signatures of
//   LAMMPS_NS::SolverLin / SolverLin_Belos   (IMPLICIT-SPH/solver_lin.h:23-98, solver_lin_belos.h:35-49,130-264)
//   LAMMPS_NS::PrecondWrapper[_Ifpack|_ML]   (IMPLICIT-SPH/precond.h:17-46, precond_ifpack.h:14-85)
// so that a call site such as USER-REAXC-T/fix_qeq_reax.cpp:671-693 (the second client of the API: it assembles its own
// Epetra_CrsMatrix and only SOLVES through SolverLin) compiles unchanged once the two typedefs at the bottom are in effect
// (tests/test_boundary_cpu.py compiles exactly those reference lines against this header).
//
//   setMatrix(Epetra_CrsMatrix*)      local CSR extracted with NumMyRows / ExtractMyRowCopy (local column ids) -> isph_matrix_set_csr
//   createLoadMultiVector(NULL, ..)   b is owned by the solver as in the reference (solver_lin.cpp:45-51): a host mirror that
//                                     getLoadMultiVector()->Values() exposes; it is a borrowed View for the ABI, so what the
//                                     caller writes there is what solveProblem uploads (isph_b200.h, borrowed load vector)
//   createSolutionMultiVector(x, ..)  a View of caller memory: the initial guess on entry, the solution on return
//   PrecondWrapper::setMatrix / create / free  the preconditioner object only carries its parameter list; SolverLin::solveProblem
//                                     applies it to the solver's context and creates / frees the preconditioner around the solve
//                                     (solver_lin_belos.h:153,190)
// One process = one GPU = one context (device = $ISPH_DEVICE or 0).  This variant is single-rank (a fix_qeq_reax-style external
// matrix, isph_matrix_set_csr); the multi-rank path is the functor-driven one (isph_graph_build).
//
// Include AFTER the Epetra headers (real Trilinos, or any stand-in providing the few methods used here).
#pragma once
#include <cstdlib>
#include <string>
#include <utility>
#include <vector>
#include "solver_lin_b200.h"

namespace Teuchos { class ParameterList; }

namespace LAMMPS_NS {

class PrecondWrapper_B200E {
 public:
  explicit PrecondWrapper_B200E(MPI_Comm) : _A(NULL) { setParameters(); }
  virtual ~PrecondWrapper_B200E() {}
  virtual void setMatrix(Epetra_CrsMatrix *A) { if (A != NULL) _A = A; }                       // precond.h:29-32 (borrowed)
  // setParameters(NULL): wrapper defaults (precond_ifpack.h:30-44 restricted to what BASELINE.json names: ILU, fill 0, overlap 0)
  virtual Teuchos::ParameterList *setParameters(Teuchos::ParameterList *param = NULL) {
    (void)param; _ints.clear(); _dbls.clear(); _strs.clear();
    set("Precond Type", "ILU"); set("Overlap Level", 0); set("fact: level-of-fill", 0);
    return NULL;
  }
  void set(const char *name, int v) { _ints.push_back(std::make_pair(std::string(name), v)); }
  void set(const char *name, double v) { _dbls.push_back(std::make_pair(std::string(name), v)); }
  void set(const char *name, const char *v) { _strs.push_back(std::make_pair(std::string(name), std::string(v))); }
  virtual void setNullVector(double *) {}
  virtual void create() {}                                       // created by solveProblem on the solver's context
  virtual void free() {}
  int apply(isph_ctx *ctx) const {                               // push the list into a context
    int rc = ISPH_SUCCESS;
    for (size_t k = 0; k < _strs.size(); ++k) rc |= isph_precond_set_param_str(ctx, _strs[k].first.c_str(), _strs[k].second.c_str());
    for (size_t k = 0; k < _ints.size(); ++k) rc |= isph_precond_set_param_int(ctx, _ints[k].first.c_str(), _ints[k].second);
    for (size_t k = 0; k < _dbls.size(); ++k) rc |= isph_precond_set_param_double(ctx, _dbls[k].first.c_str(), _dbls[k].second);
    return rc;
  }
 protected:
  Epetra_CrsMatrix *_A;
  std::vector<std::pair<std::string, int> > _ints; std::vector<std::pair<std::string, double> > _dbls; std::vector<std::pair<std::string, std::string> > _strs;
};

// PrecondWrapper_ML(MPI_Comm) as USER-REAXC-T/fix_qeq_reax.cpp:551-552 constructs it: the multilevel stand-in of csrc/amg.cu with the default
// list of precond_ml.h:44-58 except for the smoother (Chebyshev instead of symmetric Gauss-Seidel, see solver_lin_b200.h / isph_b200.h)
class PrecondWrapper_ML_B200E : public PrecondWrapper_B200E {
 public:
  explicit PrecondWrapper_ML_B200E(MPI_Comm comm) : PrecondWrapper_B200E(comm) { setParameters(); }
  virtual Teuchos::ParameterList *setParameters(Teuchos::ParameterList *param = NULL) {
    (void)param; _ints.clear(); _dbls.clear(); _strs.clear();
    set("Precond Package", "ML"); set("ML output", 10); set("max levels", 5); set("increasing or decreasing", "increasing"); set("aggregation: type", "Uncoupled");
    set("smoother: type", "Chebyshev"); set("smoother: sweeps", 1); set("smoother: pre or post", "both"); set("coarse: type", "Amesos-KLU");
    return NULL;
  }
  void setCoordinates(const int, double *, double *, double *) {}      // Zoltan repartitioning (precond_ml.h:65-99): nothing to repartition
};

class SolverLin_B200E {
 public:
  enum SolutionInitType { Random = ISPH_INIT_RANDOM, Zero = ISPH_INIT_ZERO, Value = ISPH_INIT_VALUE };     // solver_lin.h:25

  // what getLoadMultiVector() / getSolutionMultiVector() hand out: ->Values() is the column-major host array
  struct HostMultiVector {
    double *p; int lda, nvec;
    double *Values() const { return p; }
    int Stride() const { return lda; }
    int NumVectors() const { return nvec; }
    HostMultiVector *operator->() { return this; }
  };

  explicit SolverLin_B200E(MPI_Comm &) : _ctx(NULL), _map(NULL), _A(NULL), _dirty(true), _is_singular(false), _init(-1), _init_val(0.0) {
    const char *dev = std::getenv("ISPH_DEVICE");
    if (isph_ctx_create(&_ctx, dev ? std::atoi(dev) : 0, 1, 0, NULL) != ISPH_SUCCESS) _ctx = NULL;    // no CUDA device: every call below fails loudly, there is no CPU path
    _x.p = _b.p = NULL; _x.lda = _b.lda = 0; _x.nvec = _b.nvec = 0;
  }
  virtual ~SolverLin_B200E() { if (_ctx) isph_ctx_destroy(_ctx); }

  void setNodalMap(Epetra_Map *map) { _map = map; }                                              // solver_lin.cpp:109-113 (borrowed)
  void setMatrix(Epetra_CrsMatrix *A) { _A = A; _dirty = true; }                                 // solver_lin.cpp:114-118 (borrowed; uploaded at solveProblem)
  void setMatrixIsSingular(const bool s) { _is_singular = s; }
  void setNullVectorMask(Epetra_IntSerialDenseVector *mask) { _mask = mask ? std::vector<int>(mask->Values(), mask->Values() + mask->Length()) : std::vector<int>(); }
  void setInitialSolution(SolutionInitType init, double val = 0.0) { _init = (int)init; _init_val = val; }
  virtual void setParameters(Teuchos::ParameterList *param = NULL) { (void)param; if (_ctx) isph_solver_set_default_params(_ctx); }   // solver_lin_belos.h:224-264
  int set(const char *name, int v) { return _ctx ? isph_solver_set_param_int(_ctx, name, v) : ISPH_FAILURE; }
  int set(const char *name, double v) { return _ctx ? isph_solver_set_param_double(_ctx, name, v) : ISPH_FAILURE; }
  int set(const char *name, const char *v) { return _ctx ? isph_solver_set_param_str(_ctx, name, v) : ISPH_FAILURE; }

  int createSolutionMultiVector(double *x, int lda, int num_vectors) { _x.p = x; _x.lda = lda; _x.nvec = num_vectors; return LAMMPS_SUCCESS; }    // View, solver_lin.cpp:52-58
  int createLoadMultiVector(double *b, int lda, int num_vectors) {                               // b == NULL: owned by the solver, solver_lin.cpp:45-51
    if (b == NULL) { _b_own.assign((size_t)lda * num_vectors, 0.0); b = _b_own.data(); }
    _b.p = b; _b.lda = lda; _b.nvec = num_vectors; return LAMMPS_SUCCESS;
  }
  HostMultiVector getLoadMultiVector() { return _b; }
  HostMultiVector getSolutionMultiVector() { return _x; }

  // SolverLin_Belos::solveProblem, solver_lin_belos.h:130-222
  virtual int solveProblem(PrecondWrapper_B200E *prec = NULL, const char *name = NULL) {
    if (!_ctx) { std::fprintf(stderr, ">> isph_b200 error: no CUDA device (there is no CPU fallback)\n"); return LAMMPS_FAILURE; }
    if (!_A || !_x.p || !_b.p) { std::fprintf(stderr, ">> isph_b200 error: setMatrix / createSolutionMultiVector / createLoadMultiVector first\n"); return LAMMPS_FAILURE; }
    if (name != NULL) std::printf(">> isph_b200::Label - %s\n", name);
    if (_dirty && upload() != ISPH_SUCCESS) return fail();
    if (isph_solver_create_solution_multivector(_ctx, _x.p, _x.lda, _x.nvec) != ISPH_SUCCESS) return fail();
    if (isph_solver_create_load_multivector(_ctx, _b.p, _b.lda, _b.nvec) != ISPH_SUCCESS) return fail();
    isph_solver_set_matrix_is_singular(_ctx, _is_singular ? 1 : 0);
    isph_solver_set_null_vector_mask(_ctx, _mask.empty() ? NULL : _mask.data());
    if (_init >= 0) isph_solver_set_initial_solution(_ctx, _init, _init_val);                   // otherwise x is the initial guess, as in the reference
    if (prec != NULL && prec->apply(_ctx) != ISPH_SUCCESS) return fail();
    if (isph_solver_solve(_ctx, prec != NULL ? 1 : 0, name) != ISPH_SUCCESS) return fail();
    int iters = 0, conv = 0; double relres = 0.0; isph_solver_stats(_ctx, &iters, &relres, &conv, NULL);
    if (conv) std::printf(">> isph_b200::Status - Passed! %s (%d iterations, %.3e)\n", name ? name : " ", iters, relres);
    return LAMMPS_SUCCESS;                                       // non-convergence is reported by the ABI, not an error (solver_lin_belos.h:197-213)
  }
  int iterations() const { int it = 0; if (_ctx) isph_solver_stats(_ctx, &it, NULL, NULL, NULL); return it; }
  isph_ctx *context() { return _ctx; }

 protected:
  int fail() { std::fprintf(stderr, ">> isph_b200 error: %s\n", isph_last_error(_ctx)); return LAMMPS_FAILURE; }
  // Epetra_CrsMatrix -> local CSR (rows in row-map order, local column ids, ascending within a row is NOT required by the ABI)
  int upload() {
    const int n = _A->NumMyRows(), mx = _A->MaxNumEntries();
    std::vector<int> rp(n + 1, 0), ci, idx(mx > 0 ? mx : 1); std::vector<double> va, val(mx > 0 ? mx : 1);
    for (int i = 0; i < n; ++i) {
      int cnt = 0;
      if (_A->ExtractMyRowCopy(i, mx, cnt, val.data(), idx.data()) != 0) return ISPH_FAILURE;
      for (int k = 0; k < cnt; ++k) { if (idx[k] < 0 || idx[k] >= n) return ISPH_FAILURE; ci.push_back(idx[k]); va.push_back(val[k]); }      // single rank: every column is a local row
      rp[i + 1] = (int)ci.size();
    }
    const int rc = isph_matrix_set_csr(_ctx, n, rp.data(), ci.data(), va.data());
    if (rc == ISPH_SUCCESS) _dirty = false;
    return rc;
  }
  isph_ctx *_ctx; Epetra_Map *_map; Epetra_CrsMatrix *_A; bool _dirty, _is_singular; int _init; double _init_val;
  HostMultiVector _x, _b; std::vector<double> _b_own; std::vector<int> _mask;
};

#ifdef ISPH_B200_REPLACE_TRILINOS_SOLVERS          // the switch a maintainer adds to USER-REAXC-T/lammps-trilinos.h
typedef SolverLin_B200E SolverLin_Belos;
typedef PrecondWrapper_B200E PrecondWrapper_Ifpack;
typedef PrecondWrapper_ML_B200E PrecondWrapper_ML;
#endif

}  // namespace LAMMPS_NS
