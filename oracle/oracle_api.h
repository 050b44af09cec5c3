/* TEST INFRASTRUCTURE ONLY — CPU oracle C interface.
 *
 * Two shared libraries export exactly this symbol set:
 *   oracle/_ref/libisph_ref.so  — the reference's OWN functor headers (read from /root/reference at
 *                                 build time, never copied) compiled against oracle/ref_shim stand-ins;
 *                                 covers graph / pre-computation / assembly / RHS (not Krylov: Trilinos absent)
 *   oracle/libisph_oracle.so    — our plain C++ restatement ("port") of the same algorithms plus the
 *                                 Belos/Ifpack-semantics Krylov + preconditioner restatement.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load them.
 */
#ifndef ISPH_ORACLE_API_H
#define ISPH_ORACLE_API_H
#ifdef __cplusplus
extern "C" {
#endif

/* particle kinds, pair_isph.h:113-123 */
#define ORC_FILTER_MATCH 0x100   /* FilterMatchBinary, filter.h:84-108 */
enum { ORC_FLUID = 99, ORC_SOLID = 12, ORC_BOUNDARY = 16, ORC_BUFFER_DIRICHLET = 32, ORC_BUFFER_NEUMANN = 64, ORC_ALL = 127 };
/* SingularPoisson, pair_isph.h:133-137 */
enum { ORC_NOT_SINGULAR = 0, ORC_NULLSPACE = 1, ORC_PINZERO = 2, ORC_DOUBLEDIAG = 3 };
enum { ORC_KERNEL_WENDLAND = 0, ORC_KERNEL_CUBIC = 1, ORC_KERNEL_QUINTIC = 2 };
/* per-particle fields (length nall = nlocal+nghost, row-major [nall][ncomp]) */
enum { ORC_F_VFRAC = 0,   /* 1  */  ORC_F_GC = 1,       /* 9: dim x dim column-major in the first dim*dim */
       ORC_F_LC = 2,      /* 6: packed upper, first dim(dim+1)/2 */
       ORC_F_NORMAL = 3,  /* 3  */  ORC_F_PND = 4,      /* 1 */
       ORC_F_DENSITY = 5, ORC_F_VISCOSITY = 6, ORC_F_PRESSURE = 7,
       ORC_F_VELOCITY = 8,/* 3  */  ORC_F_VSTAR = 9,    /* 3 */  ORC_F_FORCE = 10, /* 3 */
       ORC_F_EPS = 11,    ORC_F_PSI = 12, ORC_F_DP = 13 /* 1: pressure increment, owned + ghost */,
       ORC_F_PSI0 = 14, /* 1: atom->psi0 */ ORC_F_SIGMA = 15, ORC_F_PHI = 16, ORC_F_COUNT = 17 };

typedef struct orc_problem orc_problem;

/* neighbor list is "packed": row ii (atom ilist[ii]) owns neigh[noff[ii] .. noff[ii+1]) */
orc_problem *orc_create(int dim, int nlocal, int nghost, const double *x, const int *type, const int *tag,
                        int inum, const int *ilist, const long long *noff, const int *neigh,
                        int ntypes, const int *kind_of_type, double h_one, double h_min, double cut_over_h,
                        int kernel_id, double morris_safe);
void orc_destroy(orc_problem *p);
const char *orc_name(void);
int orc_field_ncomp(int field);
int orc_set_field(orc_problem *p, int field, const double *data);
int orc_get_field(orc_problem *p, int field, double *data);

int orc_compute_volumes(orc_problem *p);                /* functor_volume.h:42-81 (+ owner->ghost copy) */
int orc_compute_gradient_correction(orc_problem *p);    /* functor_gradient_correction.h:24-71 */
int orc_compute_laplacian_correction(orc_problem *p);   /* functor_laplacian_correction.h:25-153 */
int orc_compute_normals(orc_problem *p);                /* functor_normal.h:56-125, pair_isph_corrected.cpp:404-427 */

long long orc_graph(orc_problem *p);                    /* functor_graph.h:38-99; returns nnz */
int orc_graph_get(orc_problem *p, int *rowptr, int *col_tags);
int orc_graph_max_row(orc_problem *p);

/* functor_incomp_navier_stokes_poisson.h:47-181; b[nlocal] out */
int orc_ns_poisson(orc_problem *p, double dt, int anti, int singular, int morris_holmes, double *b);
/* functor_incomp_navier_stokes_helmholtz.h:48-159; b[nlocal*dim] column-major lda=nlocal: in v^n, out rhs */
int orc_ns_helmholtz(orc_problem *p, double dt, double theta, int anti, int morris_holmes,
                     int incremental_pressure, const double *g, double *b);
/* functor_poisson_boltzmann_jacobian.h:35-107 (A.is_filled kept between calls) */
int orc_pb_jacobian(orc_problem *p, int morris_holmes, int linearized, double ezcb, double psiref, double gamma);
/* Corrected::FunctorOuterGradient<Pair, anti>(field, alpha = 1, grad) with FilterBinary(filter_i, filter_j) — the matrix-free
 * corrected gradient of a scalar per-particle field (functor_gradient.h:80-169), as PairISPH_Corrected::computePsiGradient
 * calls it (pair_isph_corrected.cpp:528-553; psi, (Fluid, All)); the field is forward-communicated first; grad[nlocal][3] out */
int orc_scalar_gradient(orc_problem *p, int field, int anti, int morris_holmes, int filter_i, int filter_j, double *grad);
/* functor_applied_electric_potential.h:34-96 (sigma, phi from the fields); b[nlocal] out */
int orc_applied_electric_potential(orc_problem *p, double *b);
/* functor_solute_transport.h:47-134; b[nlocal]: in c^n, out rhs */
int orc_solute_transport(orc_problem *p, double dt, double theta, double dcoeff, double *b);
/* functor_poisson_boltzmann_f.h:58-88 (psi, psi0, eps from the fields) + functor_poisson_boltzmann_extra_f.h:76-90 with the
 * caller's precomputed source extra_f[nlocal] (NULL: none); psi is forward-communicated first (pair_isph_corrected.cpp:446-450);
 * f[nlocal] out */
int orc_pb_residual(orc_problem *p, int morris_holmes, int linearized, double ezcb, double psiref, double gamma,
                    const double *extra_f, double *f);
/* the block right after the Poisson solve, pair_isph.cpp:1017-1031: forward_comm(DeltaP), computeZeroMeanPressure(dp)
 * (:422-464, when incremental pressure is used), correctVelocity (functor_correct_velocity.h:52-78, pair_isph_corrected.cpp
 * :1019-1034) incl. forward_comm(Vstar), correctPressure (functor_correct_pressure.h:29-43).  dp_owned[nlocal] = the solution. */
int orc_ns_correct(orc_problem *p, double dt, int anti, int incremental_pressure, const double *dp_owned);
/* pinfo[1][type] ("fixed" particles, pair_isph.cpp:165-167), [ntypes+1]; default all 0 */
int orc_set_fixed(orc_problem *p, const int *fixed_of_type);
int orc_get_x(orc_problem *p, double *x);               /* positions [nall][3] (moved by orc_advance_time) */
/* PairISPH_Corrected::advanceTime (pair_isph_corrected.cpp:1183-1194): FunctorOuterAdvanceTimeBegin (functor_advance_time_begin.h:52-81:
 * dp_i = grad(p)_i . 0.5 dt (v^{n+1} + v^n) on fluid rows, gradient with FilterBinary(Fluid, Fluid), then forward_comm(DeltaP)) and
 * FunctorOuterAdvanceTimeEnd over owned + ghost atoms (functor_advance_time_end.h:48-66: fixed: v = v^{n+1}; else p += dp,
 * x += 0.5 dt (v^{n+1} + v^n), v = v^{n+1}).  Fields: VELOCITY (v^n in, v^{n+1} out), VSTAR (v^{n+1}), PRESSURE, DP; positions. */
int orc_advance_time(orc_problem *p, double dt, int anti);
/* Corrected::FunctorOuterBoundaryNavierSlip (functor_boundary_navier_slip.h:54-174, iblock < 0, add_neumann_term): Robin rows summed
 * into A (call site pair_isph_corrected.cpp:921-926, after the Helmholtz functor); normals from the NORMAL field */
int orc_boundary_navier_slip(orc_problem *p, double beta);
/* Corrected::FunctorOuterBoundaryDirichlet (functor_boundary_dirichlet.h:47-150): rows of fluid particles with a solid within h are
 * REPLACED by the least-squares extrapolation stencil, their b entries (dim columns, leading dimension lda) zeroed */
int orc_boundary_dirichlet(orc_problem *p, double *b, int lda);
int orc_invalidate_matrix(orc_problem *p);              /* A.is_filled = 0, pair_isph.cpp:982,1026 */

int orc_matrix_get(orc_problem *p, double *val);        /* aligned with orc_graph_get order */
int orc_diag_get(orc_problem *p, double *diagonal, double *scaled_laplace_diagonal);
int orc_spmv(orc_problem *p, const double *x, double *y, int nvec); /* column-major lda=nlocal */

#ifdef __cplusplus
}
#endif
#endif
