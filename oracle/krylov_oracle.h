/* TEST INFRASTRUCTURE ONLY — CPU oracle for the Krylov/preconditioner half (see krylov_oracle.cpp header:
 * "parity unpinned": Belos/Ifpack are un-vendored, un-pinned third-party code). */
#ifndef ISPH_KRYLOV_ORACLE_H
#define ISPH_KRYLOV_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif
enum { ORC_SOLVER_GMRES = 0, ORC_SOLVER_CG = 1 };
enum { ORC_PREC_NONE = 0, ORC_PREC_JACOBI = 1, ORC_PREC_CHEBYSHEV = 2, ORC_PREC_ILU0 = 3, ORC_PREC_AMG = 4 };
typedef struct {
  int solver;            /* "Solver Type": Block GMRES | Block CG   (solver_lin_belos.h:173-182) */
  int flexible;          /* "Flexible Gmres" */
  int num_blocks;        /* "Num Blocks" (restart length) */
  int max_iters;         /* "Maximum Iterations" */
  int max_restarts;      /* "Maximum Restarts" */
  double tol;            /* "Convergence Tolerance" */
  int precond;           /* Ifpack "Precond Type": none | point relaxation (Jacobi) | Chebyshev | ILU (fill 0, overlap 0) */
  int jacobi_sweeps;     /* "relaxation: sweeps" */
  double jacobi_damping; /* "relaxation: damping factor" */
  double min_diag;       /* "relaxation: min diagonal value" / "chebyshev: min diagonal value" */
  int cheb_degree;       /* "chebyshev: degree" */
  double cheb_ratio;     /* "chebyshev: ratio eigenvalue" */
  double cheb_lambda_max;/* "chebyshev: max eigenvalue" (<=0: power method) */
  int cheb_eig_iters;    /* "chebyshev: eigenvalue max iterations" */
  const int *row_gid;    /* global id per row for the power-method start vector (NULL: row+1) */
  int ilu_fill;          /* "fact: level-of-fill" (Ifpack_IlukGraph level rule) */
  int overlap;           /* "Overlap Level" 0 | 1 with block_of_row = the rank of every row (Ifpack_AdditiveSchwarz over an
                            Ifpack_OverlappingRowMatrix, "schwarz: combine mode" = Add: precond_ifpack.h:35-43).  Every block is extended
                            by the rows of its off-block columns (one level), ordered behind its own rows by (block, row_gid) */
  /* ORC_PREC_AMG: the multilevel stand-in for ML (amg_oracle.h; names = ML's parameter list, precond_ml.h:44-58) */
  int amg_max_levels;        /* "max levels" */
  double amg_threshold;      /* "aggregation: threshold" */
  int amg_smoother;          /* "smoother: type": 0 Chebyshev | 1 Jacobi */
  int amg_pre, amg_post;     /* degree / sweeps before and after the coarse correction on the finest level ("smoother: sweeps", "smoother: pre or post") */
  int amg_level_sweeps;      /* the same on the intermediate levels */
  int amg_coarse_sweeps;     /* "coarse: sweeps" */
  double amg_alpha;          /* "smoother: Chebyshev alpha" (eigenvalue ratio) */
  double amg_coarse_alpha;   /* the same on the coarsest level */
  int amg_eig_iters;         /* "eigen-analysis: iterations" (power method) */
  int amg_max_coarse;        /* "coarse: max size" */
  double amg_scale;          /* scaling of the coarse-grid correction */
  double amg_damping;        /* "smoother: damping factor" (Jacobi) */
  double amg_level_alpha;    /* eigenvalue ratio on the levels between the finest and the coarsest (amg_alpha: finest level) */
  double amg_level_scale;    /* coarse-correction scaling on those levels (amg_scale: finest level) */
  int amg_coarse_direct;     /* "coarse: type" = Amesos-KLU: dense inverse of the coarsest operator (used for non-singular problems; the caller clears it for singular ones as PrecondWrapper_ML::setNullVector does, precond_ml.h:118-120) */
} orc_krylov_params;

int orc_set_num_threads(int n);   /* OpenMP threads of the port's row loops; returns the count in effect */
void orc_krylov_default_params(orc_krylov_params *p);
/* col = local row index of each column, -1 for a column that is not in this process' rows.  b is modified in place
 * when use_null (as solver_lin_belos.h:141-143 does).  returns 0 converged, 1 not converged. */
int orc_krylov_solve(int n, const int *rowptr, const int *col, const double *val, const orc_krylov_params *prm,
                     const int *block_of_row, const int *null_mask, int use_null, double *b, double *x,
                     int *iters_out, double *relres_out, double *history, int history_cap);
int orc_krylov_solve_block(int nb, int dim, const int *rowptr, const int *col, const double *val, const int *prec_rowptr, const int *prec_col, const double *prec_val,
                           const orc_krylov_params *prm, double *b, double *x, int *iters_out, double *relres_out);
int orc_precond_apply(int n, const int *rowptr, const int *col, const double *val, const orc_krylov_params *prm,
                      const int *block_of_row, const double *r, double *z, double *lambda_max_out);
/* hierarchy of ORC_PREC_AMG for the tests: returns the number of levels; rows[l], nnz[l], lmax[l] per level; agg0[n] = aggregate of every
 * finest-level row (-1: none); the level-1 operator in CSR when it fits the given capacities (c_rowptr: cap_rows + 1, c_col/c_val: cap_nnz) */
int orc_amg_hierarchy(int n, const int *rowptr, const int *col, const double *val, const orc_krylov_params *prm, const int *block_of_row,
                      int *rows, long long *nnz, double *lmax, int *agg0, int cap_rows, long long cap_nnz, int *c_rowptr, int *c_col, double *c_val);
#ifdef __cplusplus
}
#endif
#endif
