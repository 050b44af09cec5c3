/* isph_b200.h — C ABI of the B200-native linear-solve hot path of implicit-sph.
 *
 * Drop-in boundary: everything the reference does between "LAMMPS hands over atoms + full
 * neighbor list" and "x holds the Krylov solution" — graph build, operator assembly, preconditioner, Krylov —
 * runs on one B200 per process behind these entry points.  Plain pointers and sizes only; all pointers are HOST
 * pointers unless a function says otherwise; nothing passed in is retained after the call returns except where
 * stated ("borrowed", mirroring Teuchos::rcp(p,false) in the reference).  Every function returns
 * ISPH_SUCCESS (0, = LAMMPS_SUCCESS, macrodef.h:23) or ISPH_FAILURE (-1, = LAMMPS_FAILURE, macrodef.h:20);
 * isph_last_error() gives the message.  Non-convergence of a solve is NOT an error (solver_lin_belos.h:194-213).
 *
 * Each group cites the reference interface it replaces (paths relative to IMPLICIT-SPH/).
 * The header-only C++ adapter include/solver_lin_b200.h re-creates the SolverLin / PrecondWrapper method names
 * on top of this ABI; INTEGRATION.md shows the binding a reference maintainer would add.
 */
#ifndef ISPH_B200_H
#define ISPH_B200_H
#ifdef __cplusplus
extern "C" {
#endif

#define ISPH_SUCCESS 0
#define ISPH_FAILURE (-1)

typedef struct isph_ctx isph_ctx;

/* ParticleKind bit masks, pair_isph.h:113-123 */
enum { ISPH_KIND_FLUID = 99, ISPH_KIND_SOLID = 12, ISPH_KIND_BOUNDARY = 16, ISPH_KIND_BUFFER_DIRICHLET = 32,
       ISPH_KIND_BUFFER_NEUMANN = 64, ISPH_KIND_ALL = 127 };
/* SingularPoisson, pair_isph.h:133-137 */
enum { ISPH_NOT_SINGULAR = 0, ISPH_NULLSPACE = 1, ISPH_PINZERO = 2, ISPH_DOUBLEDIAG = 3 };
/* kernel functions, pair_isph_corrected.cpp:1295-1302 */
enum { ISPH_KERNEL_WENDLAND = 0, ISPH_KERNEL_CUBIC = 1, ISPH_KERNEL_QUINTIC = 2 };
/* SolverLin::SolutionInitType, solver_lin.h:25 */
enum { ISPH_INIT_RANDOM = 0, ISPH_INIT_ZERO = 1, ISPH_INIT_VALUE = 2 };
/* per-particle fields, row-major [nlocal+nghost][ncomp]  (atom_vec_isph.h:56-89 / pair_isph.h work arrays) */
enum { ISPH_F_VFRAC = 0,   /* 1 */ ISPH_F_GC = 1,       /* 9: dim x dim column-major (VIEW2, macrodef.h:61) in the first dim*dim */
       ISPH_F_LC = 2,      /* 6: packed symmetric, first dim(dim+1)/2 */
       ISPH_F_NORMAL = 3,  /* 3 */ ISPH_F_PND = 4,      /* 1 */
       ISPH_F_DENSITY = 5, ISPH_F_VISCOSITY = 6, ISPH_F_PRESSURE = 7,
       ISPH_F_VELOCITY = 8,/* 3 */ ISPH_F_VSTAR = 9,    /* 3 */ ISPH_F_FORCE = 10, /* 3 */
       ISPH_F_EPS = 11, ISPH_F_PSI = 12, ISPH_F_DP = 13 /* 1: pressure increment dp, owned + ghost */,
       ISPH_F_PSI0 = 14 /* 1: prescribed (normalised) potential of solid / boundary particles, atom->psi0 */, 
       ISPH_F_SIGMA = 15 /* 1: electric conductivity, atom->sigma */, ISPH_F_PHI = 16 /* 1: applied electric potential, atom->phi */, ISPH_F_COUNT = 17 };

/* ---- context ------------------------------------------------------------------------------------------------
 * replaces: SolverLin(MPI_Comm&) solver_lin.h:28, PrecondWrapper(MPI_Comm) precond.h:26 (one communicator per
 * process; here one NCCL communicator per GPU).  nccl_unique_id: 128 bytes obtained from isph_nccl_unique_id() on
 * rank 0 and broadcast by the host program (MPI in LAMMPS, torch.distributed in bench.py); NULL when nranks == 1. */
int isph_ctx_create(isph_ctx **ctx, int device, int nranks, int rank, const void *nccl_unique_id);
int isph_ctx_destroy(isph_ctx *ctx);
int isph_nccl_unique_id(void *id128);
/* Host-only (no CUDA, no context) halo planner used by the NCCL path: the role of Epetra's column map / Epetra_Import
 * (SURVEY.md §2.2).  For every ghost atom: its tag, the rank that owns the tag and the owner-local index.  Out: the
 * column id of every ghost (owner-local index for ghosts owned here, nlocal + halo slot otherwise; slots number the
 * distinct remote tags in (owner, tag) order), how many values come from each peer, and the owner-local indices to
 * request, in slot order. */
int isph_halo_plan_host(int nranks, int rank, int nlocal, int nghost, const int *ghost_tag, const int *ghost_owner,
                        const int *ghost_owner_idx, int *ghost_col, int *recv_count, int *request_idx, int *nhalo_out);
/* The same plan computed by the device planner that multi-GPU runs use (sort + numbering on the GPU); host arrays in and out.
 * Exists so that a one-GPU parity test can hold the device planner against isph_halo_plan_host. */
int isph_halo_plan_device(isph_ctx *ctx, int nranks, int rank, int nlocal, int nghost, const int *ghost_tag, const int *ghost_owner,
                          const int *ghost_owner_idx, int *ghost_col, int *recv_count, int *request_idx, int *nhalo_out);
const char *isph_last_error(const isph_ctx *ctx);
int isph_set_stream(isph_ctx *ctx, void *cuda_stream);   /* run on the caller's CUDA stream (borrowed) */
int isph_synchronize(isph_ctx *ctx);
const char *isph_version(void);

/* ---- pair / atom / neighbor data ------------------------------------------------------------------------------
 * replaces: PairISPH_Corrected::coeff pair_isph_corrected.cpp:1273-1347 (kernel, cutsq, h tables, MorrisSafeCoeff)
 *           and the FunctorOuter<PairIsph> constructor capture functor.h:64-91 (atom->{x,type,tag,vfrac}, list->...) */
/* Limits: 1 <= ntypes <= 7 (particle types 1..7; the reference's scripts use at most 3); ilist must enumerate the owned atoms
 * in order, ilist[ii] == ii, and inum == nlocal (what a LAMMPS full list built for pair_style isph gives, pair_isph.cpp:1887-1894:
 * the row map IS the owned-atom order, :1258); at most 511 neighbors per row; atom tags fit a 32-bit int (functor.h:62). */
int isph_pair_coeff(isph_ctx *ctx, int dim, int ntypes, const int *kind_of_type /*[ntypes+1]*/, double h, double h_min,
                    double cut_over_h, int kernel, double morris_safe);
int isph_atoms_set(isph_ctx *ctx, int nlocal, int nghost, const double *x /*[nall][3]*/, const int *type, const int *tag);
/* LAMMPS NeighList layout: numneigh[i], firstneigh[i] indexed by atom i = ilist[ii]; entries are masked with NEIGHMASK */
int isph_neighbors_set(isph_ctx *ctx, int inum, const int *ilist, const int *numneigh, int *const *firstneigh);
/* packed layout: row ii owns neigh[noff[ii] .. noff[ii+1]) */
int isph_neighbors_set_packed(isph_ctx *ctx, int inum, const int *ilist, const long long *noff, const int *neigh);
/* The full neighbor list built ON THE DEVICE from the atoms already set (replaces the host-built LAMMPS list the
 * pair style requests in init_style, pair_isph.cpp:1887-1894): every owned atom's row holds all atoms j != i (owned or ghost) with
 * |x_i - x_j|^2 <= cutneigh^2, cutneigh <= 0 meaning the pair cutoff (skin 0, as the reference's scripts run).  Cell binning; row
 * order = cell stencil order, then atom index.  The list is a superset for the functors' own rsq < cutsq test, so the graph does not
 * depend on it; removes the per-step upload of the list (the bulk of the host-to-device traffic of a step). */
int isph_neighbors_build(isph_ctx *ctx, double cutneigh);
long long isph_neighbors_count(isph_ctx *ctx);                     /* entries of the current list */
int isph_neighbors_get(isph_ctx *ctx, long long *noff /*[inum+1]*/, int *neigh /*[count]*/);   /* copy the current list out (packed layout) */
int isph_field_set(isph_ctx *ctx, int field, const double *data);
int isph_field_get(isph_ctx *ctx, int field, double *data);
/* owner -> ghost copy of a field: comm->forward_comm_pair(this), pair_isph.cpp:1924-2074 */
int isph_forward_comm(isph_ctx *ctx, int field);

/* ---- pre-computation (PairISPH_Corrected::computePre, pair_isph_corrected.cpp:302-313) ------------------------ */
int isph_compute_volumes(isph_ctx *ctx);               /* functor_volume.h:42-81 */
int isph_compute_gradient_correction(isph_ctx *ctx);   /* functor_gradient_correction.h:24-71 */
int isph_compute_laplacian_correction(isph_ctx *ctx);  /* functor_laplacian_correction.h:25-153 */
int isph_compute_normals(isph_ctx *ctx);               /* functor_normal.h:56-125, pair_isph_corrected.cpp:404-427 */

/* ---- graph + matrix (the Epetra_CrsGraph / Epetra_CrsMatrix method set the functors call) -------
 * isph_graph_build = nodal map (pair_isph.cpp:1258-1259) + FunctorOuterGraph (functor_graph.h:38-99)
 *                    + new Epetra_CrsMatrix(Copy, graph) + zero diagonal vectors (pair_isph.cpp:1266-1270) */
int isph_graph_build(isph_ctx *ctx);
long long isph_graph_nnz(isph_ctx *ctx);                /* after duplicate merging (Epetra FillComplete) */
int isph_graph_max_row(isph_ctx *ctx);                  /* Epetra_CrsGraph::MaxNumIndices */
/* canonical form: rows in nodal-map order, columns = ascending global tag, duplicates merged */
int isph_graph_get(isph_ctx *ctx, int *rowptr /*[nlocal+1]*/, int *col_tags /*[nnz]*/);
int isph_matrix_get(isph_ctx *ctx, double *val /*[nnz], aligned with isph_graph_get*/);
/* matrix supplied by the caller (second API client, USER-REAXC-T/fix_qeq_reax.cpp:509-694): local column ids */
int isph_matrix_set_csr(isph_ctx *ctx, int n, const int *rowptr, const int *col, const double *val);
int isph_matrix_put_scalar(isph_ctx *ctx, double a);                       /* PutScalar */
int isph_matrix_scale(isph_ctx *ctx, double a);                            /* Scale */
int isph_matrix_left_scale(isph_ctx *ctx, const double *s /*[nlocal]*/);   /* LeftScale */
int isph_matrix_extract_diagonal(isph_ctx *ctx, double *d /*[nlocal]*/);   /* ExtractDiagonalCopy */
int isph_matrix_replace_diagonal(isph_ctx *ctx, const double *d);          /* ReplaceDiagonalValues */
int isph_matrix_multiply(isph_ctx *ctx, const double *x, double *y, int lda, int nvec);   /* Multiply(false, X, Y) */
int isph_matrix_invalidate(isph_ctx *ctx);                                 /* A.is_filled = 0, pair_isph.cpp:982,1026 */
/* end of step: delete A.crs, tags_in_cut, nodalmap (pair_isph.cpp:1351-1372); the next isph_graph_build rebuilds the pattern */
int isph_graph_invalidate(isph_ctx *ctx);
/* Corrected::FunctorOuterLaplacianMatrix<Pair,Anti>[_MorrisHolmes], functor_laplacian_matrix.h:56-328, including the
 * PutScalar(0.0) that precedes it at every call site.  material_field < 0: material == 1. */
#define ISPH_FILTER_MATCH 0x100   /* OR into filter_i: FilterMatchBinary (filter.h:84-108): i must EQUAL the kind, j & mask */
int isph_assemble_laplacian(isph_ctx *ctx, double alpha, int material_field, int anti, int morris_holmes,
                            int filter_i, int filter_j);
/* FunctorOuterGradientDotOperatorMatrix, functor_gradient_dot_operator_matrix.h:36-79 (SumInto) */
int isph_assemble_gradient_dot(isph_ctx *ctx, double alpha, int vector_field, int filter_i, int filter_j);

/* ---- system functors ----------------------------------------------------------------------------------------
 * They fill A and the solver's load vector b on the device (b = li_solver->getLoadMultiVector()->Values()). */
/* FunctorOuterIncompNavierStokesPoisson, functor_incomp_navier_stokes_poisson.h:47-181 (via computePoisson) */
int isph_ns_poisson(isph_ctx *ctx, double dt, int anti, int singular, int morris_holmes);
/* FunctorOuterIncompNavierStokesHelmholtz, functor_incomp_navier_stokes_helmholtz.h:48-159; b must hold v^n */
int isph_ns_helmholtz(isph_ctx *ctx, double dt, double theta, int anti, int morris_holmes, int incremental_pressure,
                      const double *g /*[3]*/);
/* FunctorOuterPoissonBoltzmannJacobian, functor_poisson_boltzmann_jacobian.h:35-107 */
int isph_pb_jacobian(isph_ctx *ctx, int morris_holmes, int linearized, double ezcb, double psiref, double gamma);
/* the block right after the Poisson solve, pair_isph.cpp:1017-1031: forward_comm(DeltaP),
 * computeZeroMeanPressure(dp) (:422-464, only with incremental pressure), correctVelocity (functor_correct_velocity.h:52-78:
 * vstar -= dt/rho grad(dp), then forward_comm(Vstar)), correctPressure (functor_correct_pressure.h:29-43).  dp = the device
 * solution of the last solve, or dp_owned[nlocal] when given.  Results: fields ISPH_F_DP, ISPH_F_VSTAR, ISPH_F_PRESSURE. */
int isph_ns_correct(isph_ctx *ctx, double dt, int anti, int incremental_pressure, const double *dp_owned);
/* "fixed" particle types, pinfo[1][type] (pair_isph.cpp:165-167; XML `type:N = "solid:fixed"`), fixed_of_type[ntypes+1]; default none */
int isph_pair_fixed(isph_ctx *ctx, const int *fixed_of_type);
/* PairISPH_Corrected::advanceTime (pair_isph_corrected.cpp:1183-1194): FunctorOuterAdvanceTimeBegin
 * (functor_advance_time_begin.h:52-81: dp_i = grad(p)_i . 0.5 dt (v^{n+1} + v^n) on fluid rows, matrix-free corrected gradient with
 * FilterBinary(Fluid, Fluid); forward_comm(DeltaP)) and FunctorOuterAdvanceTimeEnd over owned + ghost atoms
 * (functor_advance_time_end.h:48-66: fixed types v = v^{n+1}; others p += dp, x += 0.5 dt (v^{n+1} + v^n), v = v^{n+1}).
 * Fields: ISPH_F_VELOCITY = v^n in / v^{n+1} out, ISPH_F_VSTAR = v^{n+1}, ISPH_F_PRESSURE, ISPH_F_DP; the device positions move
 * (isph_atoms_get_x reads them back) and the graph is invalidated.  Needs the graph of the current positions. */
int isph_advance_time(isph_ctx *ctx, double dt, int anti);
int isph_atoms_get_x(isph_ctx *ctx, double *x /*[nall][3]*/);
/* Row modifiers applied after the Helmholtz functor for ns.boundary = NavierSlip / Dirichlet (pair_isph_corrected.cpp:918-934):
 * Corrected::FunctorOuterBoundaryNavierSlip (functor_boundary_navier_slip.h:54-174, iblock < 0, add_neumann_term; Robin rows summed
 * into A; normals = ISPH_F_NORMAL, rho = ISPH_F_DENSITY) and Corrected::FunctorOuterBoundaryDirichlet (functor_boundary_dirichlet.h:
 * 47-150; fluid rows with a solid within h REPLACED by the extrapolation stencil, their load-vector rows zeroed). */
int isph_boundary_navier_slip(isph_ctx *ctx, double beta);
int isph_boundary_dirichlet(isph_ctx *ctx);
/* FunctorOuterAppliedElectricPotential (functor_applied_electric_potential.h:34-96; call site pair_isph_corrected.cpp:598-617,
 * pair_isph.cpp:628-663): A = Laplacian(alpha = -1, material = sigma, FilterMatchBinary(Fluid, Fluid)); A.diagonal = diag(A);
 * b = 0; Solid rows: diagonal 1; buffer rows: diagonal 1, b = phi; ReplaceDiagonalValues.  Fields ISPH_F_SIGMA, ISPH_F_PHI;
 * b = the load vector (one column). */
int isph_applied_electric_potential(isph_ctx *ctx);
/* FunctorOuterSoluteTransport (functor_solute_transport.h:47-134; call site pair_isph_corrected.cpp:843-861, pair_isph.cpp:811-835):
 * A = Laplacian(alpha = dt * dcoeff, FilterMatchBinary(Fluid, Fluid - BufferNeumann)); w = (1 - theta) A c^n; A *= -theta;
 * scaled_laplace_diagonal = diag(A); Fluid rows: diagonal 1 + sld, b += w; Solid / buffer rows: diagonal 1.  The load vector
 * (one column) holds the concentration c^n on entry and the right-hand side on return. */
int isph_solute_transport(isph_ctx *ctx, double dt, double theta, double dcoeff);
/* PairISPH_Corrected::computeF (pair_isph_corrected.cpp:438-485): forward_comm(psi), then FunctorOuterPoissonBoltzmannF
 * (functor_poisson_boltzmann_f.h:58-88; fields ISPH_F_PSI, ISPH_F_PSI0, ISPH_F_EPS) on the matrix-free corrected Laplacian
 * (functor_laplacian.h:67-277), plus the caller-evaluated source of FunctorOuterPoissonBoltzmannExtraF
 * (functor_poisson_boltzmann_extra_f.h:76-90; extra_f[nlocal] or NULL).  F is left in the load vector when one exists and
 * copied to f_out[nlocal] when given. */
int isph_pb_residual(isph_ctx *ctx, int morris_holmes, int linearized, double ezcb, double psiref, double gamma,
                     const double *extra_f, double *f_out);
/* The Newton iteration NOX runs for PairISPH::computePoissonBoltzmann (pair_isph.cpp:572-600) with the reference's default
 * lists (solver_nox_impl.h:76-160: full steps; converged when ||F||_2/sqrt(N) <= tol_f AND ||dpsi||_2/sqrt(N) <= tol_update;
 * reference values 1e-8, 1e-5, 100 iterations): computeF, computeJacobian and the Jacobian solve of every iteration stay on
 * the device; the solve uses this context's Krylov / preconditioner parameter lists.  ISPH_F_PSI holds the initial guess on
 * entry and the solution (owned + ghost) on return.  One solution and one load vector must exist. */
int isph_pb_newton(isph_ctx *ctx, int morris_holmes, int linearized, double ezcb, double psiref, double gamma, const double *extra_f,
                   int max_newton, double tol_f, double tol_update, int use_prec,
                   int *newton_iters, int *linear_iters, double *normf, int *converged);
int isph_diagonals_get(isph_ctx *ctx, double *diagonal, double *scaled_laplace_diagonal);   /* A.diagonal, A.scaled_laplace_diagonal */

/* ---- SolverLin / SolverLin_Belos mirror (solver_lin.h:23-98, solver_lin.cpp, solver_lin_belos.h:130-264) --------- */
int isph_solver_create_solution_multivector(isph_ctx *ctx, double *x /*borrowed; NULL: owned*/, int lda, int nvec);
/* b borrowed = a View of caller memory as in the reference (solver_lin.cpp:45-58): the system functors (isph_ns_poisson, ...) copy
 * the right-hand side they form on the device back into b, and whatever b holds when a functor that reads it or the solve starts
 * is uploaded, unless a device functor / isph_solver_load_set wrote the load vector after the previous solve. */
int isph_solver_create_load_multivector(isph_ctx *ctx, double *b /*borrowed; NULL: owned, device-resident*/, int lda, int nvec);
int isph_solver_load_set(isph_ctx *ctx, const double *b, int lda);        /* write getLoadMultiVector()->Values() */
int isph_solver_load_get(isph_ctx *ctx, double *b, int lda);
int isph_solver_solution_set(isph_ctx *ctx, const double *x, int lda);    /* write getSolutionMultiVector()->Values() (initial guess of an owned x) */
int isph_solver_solution_get(isph_ctx *ctx, double *x, int lda);
int isph_solver_set_null_vector_mask(isph_ctx *ctx, const int *mask /*[nlocal] or NULL = all ones*/);
int isph_solver_set_matrix_is_singular(isph_ctx *ctx, int is_singular);
int isph_solver_set_initial_solution(isph_ctx *ctx, int init_type, double val);
/* Belos parameter names are kept: "Solver Type" ("Block GMRES"|"Block CG"), "Flexible Gmres", "Num Blocks",
 * "Maximum Iterations", "Maximum Restarts", "Convergence Tolerance", "Orthogonalization" ("DGKS") */
int isph_solver_set_param_int(isph_ctx *ctx, const char *name, int v);
int isph_solver_set_param_double(isph_ctx *ctx, const char *name, double v);
int isph_solver_set_param_str(isph_ctx *ctx, const char *name, const char *v);
int isph_solver_set_default_params(isph_ctx *ctx);                        /* setParameters(NULL), solver_lin_belos.h:224-264 */
/* PrecondWrapper_Ifpack parameter names are kept (precond_ifpack.h:28-48): "Precond Type" ("point relaxation" |
 * "Chebyshev" | "ILU" | "none"), "Overlap Level" (0; 1 = the reference's default, precond_ifpack.h:37: across ranks for "ILU" with one block per rank — halo rows imported,
 * additive Schwarz with "schwarz: combine mode" Add; a no-op on one rank), "fact: level-of-fill" (0; k > 0 supported), "relaxation: type" ("Jacobi"),
 * "relaxation: sweeps", "relaxation: damping factor", "chebyshev: degree", "chebyshev: ratio eigenvalue",
 * "chebyshev: max eigenvalue", "chebyshev: eigenvalue max iterations".  "b200: ilu blocks" = {bx,by,bz} split of the
 * local rows into Ifpack-rank-equivalent bricks is set with isph_precond_set_blocks(). */
/* PrecondWrapper_ML (precond_ml.h:17-172; "Precond Package" = "ML" is the reference's default, pair_isph.cpp:325-329,359-361): a multilevel
 * preconditioner built on the device stands in for ML — PARITY UNPINNED (ML is un-vendored, un-pinned third-party code) and not ML's default
 * algorithm but the data-parallel member of its option space: distance-2 independent-set aggregation inside each rank, non-smoothed
 * aggregation ("aggregation: damping factor" 0), Chebyshev or Jacobi smoothers on every level incl. the coarsest, V-cycle.  ML's names are
 * kept: "Precond Package" ("ML" | "Ifpack"), "max levels" (5), "aggregation: type" ("Uncoupled" | "MIS" | "Uncoupled-MIS"),
 * "aggregation: threshold" (0.02; ML's criterion a_ij^2 > eps^2 |a_ii a_jj|), "aggregation: damping factor" (must be 0), "smoother: type"
 * ("Chebyshev" | "Jacobi"; "symmetric Gauss-Seidel", the value precond_ml.h:53 sets, is sequential within a rank and is refused),
 * "smoother: sweeps" (Chebyshev degree / Jacobi sweeps before and after the coarse correction), "smoother: pre or post", "smoother: Chebyshev
 * alpha" (eigenvalue ratio on the finest level, 2), "smoother: damping factor" (Jacobi, 0.67), "coarse: type" (= smoother: type, or "Amesos-KLU" = direct solve of the coarsest operator; replaced by the smoother for singular problems as
 * PrecondWrapper_ML::setNullVector does, precond_ml.h:118-120), "coarse: sweeps" (8),
 * "coarse: Chebyshev alpha" (30), "coarse: max size" (128), "eigen-analysis: iterations" (10).  Extensions (not ML keys): "smoother: pre
 * sweeps" (1), "smoother: post sweeps" (1), "smoother: sweeps (coarse levels)" (3), "smoother: Chebyshev alpha (coarse levels)" (10),
 * "coarse correction scale" (2.5 on the finest level), "coarse correction scale (coarse levels)" (2.0). */
int isph_precond_set_param_int(isph_ctx *ctx, const char *name, int v);
int isph_precond_set_param_double(isph_ctx *ctx, const char *name, double v);
int isph_precond_set_param_str(isph_ctx *ctx, const char *name, const char *v);
int isph_precond_set_blocks(isph_ctx *ctx, const int *block_of_row /*[nlocal] or NULL = one block*/);
int isph_precond_create(isph_ctx *ctx);     /* PrecondWrapper::create(), precond_ifpack.h:50-75 */
/* Pure host code (no CUDA): level-of-fill pattern of Ifpack_ILU ("fact: level-of-fill" k, precond_ifpack.h:38) for a CSR
 * pattern with ascending columns — the host pass behind ILU(k), k > 0; exported so that CPU-only tests can drive it.
 * ISPH_FAILURE when cap is too small (*nnz_out holds the size needed). */
int isph_iluk_symbolic_host(int n, const int *rowptr, const int *col, int fill, int *rowptr_out, int *col_out, long long cap, long long *nnz_out);
int isph_precond_free(isph_ctx *ctx);       /* PrecondWrapper::free() */
int isph_precond_apply(isph_ctx *ctx, const double *r, double *z);        /* ApplyInverse, host vectors (tests) */
/* SolverLin block interface (solver_lin.h:43-56, solver_lin.cpp:78-138) and SolverLin_Belos::solveBlockProblem (solver_lin_belos.h:53-128): a
 * dim x dim block operator whose blocks are CSR matrices over the nodal map (col = local row index; a NULL block is ignored like a NULL
 * Epetra_CrsMatrix*, solver_lin.cpp:133).  x and b are the n x dim multivectors of createSolution/LoadMultiVector, one column per block row.
 * The preconditioner is ONE operator built from the scalar matrix of the context (prec->setMatrix(A.crs), pair_isph.cpp:925) and applied to
 * every diagonal block (getBlockPrecondOperator).  Errors as the reference: rhs columns != dim, singular problems.  One rank only. */
int isph_solver_create_block_matrix(isph_ctx *ctx, int dim, const char *name);      /* createBlockMatrix + setBlockBegin */
int isph_solver_set_block_csr(isph_ctx *ctx, int i, int j, int n, const int *rowptr, const int *col, const double *val);   /* setBlock(i, j, A) */
int isph_solver_set_block_end(isph_ctx *ctx);                                       /* setBlockEnd */
int isph_solver_free_block_matrix(isph_ctx *ctx);                                   /* freeBlockMatrix */
int isph_solver_solve_block(isph_ctx *ctx, int use_prec, const char *label);        /* solveBlockProblem(prec, name) */
/* SolverLin_Belos::solveProblem(prec, name): use_prec != 0 creates and frees the preconditioner around the solve */
int isph_solver_solve(isph_ctx *ctx, int use_prec, const char *label);
int isph_solver_stats(isph_ctx *ctx, int *iters, double *relres, int *converged, double *lambda_max);
/* Arnoldi steps of the last GMRES solve that took the second (DGKS) Gram-Schmidt pass — bookkeeping for the traffic model */
long long isph_solver_second_passes(isph_ctx *ctx);

/* ---- timers (the reference's Teuchos::Time scopes, utils.cpp:16-43): "computePoisson", "solvePoisson", ... ----- */
double isph_timer_ms(isph_ctx *ctx, const char *name);    /* accumulated device time (CUDA events) */
int isph_timer_reset(isph_ctx *ctx);
/* halo plan of this rank (after isph_atoms_set with nranks > 1): values received / sent per vector import and the number of
 * peers sent to; 8 * nsend bytes leave this GPU over NVLink per SpMV */
int isph_halo_counts(isph_ctx *ctx, int *nhalo, int *nsend, int *npeers);
long long isph_kernel_launches(isph_ctx *ctx);            /* number of kernels this context has launched */
/* per-launch CUDA-event timing of the SpMV kernel inside whatever runs next (solve, assembly): enable, run, read */
int isph_profile_spmv(isph_ctx *ctx, int enable);
int isph_profile_spmv_get(isph_ctx *ctx, double *total_ms, long long *launches);   /* also resets the counters */
/* the same for the ILU triangular-solve kernel (enabled by isph_profile_spmv) */
int isph_profile_precond_get(isph_ctx *ctx, double *total_ms, long long *launches);
/* the ILU factors of the last isph_precond_create / solve: stored entries of L + D + U (all blocks), dependency levels of the
 * forward and backward sweeps (the critical path of the level-scheduled solves), longest factor row */
/* hierarchy of the last multilevel preconditioner: number of levels, rows / stored entries / lambda_max(D^-1 A) per level (cap entries);
 * the aggregate (coarse index) of every local row of the finest level (-1: row without strong connections); device time of a setup
 * phase ("aggregate" | "galerkin" | "eigen") in ms */
int isph_precond_ml_info(isph_ctx *ctx, int *levels, int *rows, long long *nnz, double *lambda_max, int cap);
int isph_precond_ml_aggregates(isph_ctx *ctx, int *agg /*[nlocal]*/);
double isph_precond_ml_setup_ms(isph_ctx *ctx, const char *phase);
int isph_precond_info(isph_ctx *ctx, long long *factor_nnz, int *levels_lower, int *levels_upper, int *max_row);
/* FP64 FMA-loop peak of this GPU in TFLOP/s (16 independent DFMA chains per thread, no memory traffic): the denominator the
 * assembly kernels' flop rates are quoted against ("not in MEASURED_PEAKS.json; measure") */
int isph_measure_fp64_peak(isph_ctx *ctx, double *tflops);
/* last SpMV-only micro benchmark: runs `reps` SpMVs on the current matrix, returns average ms (device events) */
int isph_bench_spmv(isph_ctx *ctx, int reps, double *avg_ms);

#ifdef __cplusplus
}
#endif
#endif
