// Internal state of the B200 linear-solve path (host C++ side).  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/isph_b200.h"
#include "p2p_device.cuh"

#define ISPH_NEIGHMASK 0x3FFFFFFF        /* LAMMPS NEIGHMASK, functor_graph.h:72 */
#define ISPH_EPS_R 1.0e-24               /* ISPH_EPSILON, macrodef.h:6 */
#define ISPH_MAXT 8                      /* particle types 1..7 */
#define ISPH_SLICE 32                    /* SELL-C slice height = warp size */

#define CUDA_CHECK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) \
  throw std::runtime_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " @" + __FILE__ + ":" + std::to_string(__LINE__)); } while (0)
#define ISPH_REQUIRE(cond, msg) do { if (!(cond)) throw std::runtime_error(std::string(msg)); } while (0)

namespace isph {

// grow-only device buffer (the matrix is rebuilt every step, pair_isph.cpp:1351-1372: never free/realloc on the hot path)
template <class T> struct DevBuf {
  T *p = nullptr; size_t cap = 0;
  void ensure(size_t n, bool keep = false) {
    if (n <= cap) return;
    size_t ncap = n + n / 8 + 64; T *q = nullptr;
    CUDA_CHECK(cudaMalloc(&q, ncap * sizeof(T)));
    if (keep && p && cap) CUDA_CHECK(cudaMemcpy(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice));
    if (p) cudaFree(p);
    p = q; cap = ncap;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
template <class T> struct PinBuf {
  T *p = nullptr; size_t cap = 0;
  void ensure(size_t n) { if (n <= cap) return; if (p) cudaFreeHost(p); cap = n + n / 8 + 64; CUDA_CHECK(cudaMallocHost(&p, cap * sizeof(T))); }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// pair tables (PairISPH_Corrected::coeff, pair_isph_corrected.cpp:1302-1337); kC/kCh are the kernel's cached
// normalisation `_C` and `_C/_h` per type pair, computed on the host with the reference's expressions
struct PairTab {
  double cutsq[ISPH_MAXT][ISPH_MAXT], h[ISPH_MAXT][ISPH_MAXT], kC[ISPH_MAXT][ISPH_MAXT], kCh[ISPH_MAXT][ISPH_MAXT];
  double cut[ISPH_MAXT][ISPH_MAXT];      // sqrt(cutsq), the argument the functors pass to computeMirrorCoefficient
  int kind[ISPH_MAXT];
  int kernel, ntypes, dim;
  double morris_safe;
};

// SELL-32 matrix with slack: slice s holds rows 32s..32s+31 column-major; entry k of row r at slice_off[s] + 32k + (r&31).
// Columns ascend within a row (local column id: owned rows first, then halo slots); `atom` = the neighbor atom that
// produced the entry (-1 for padding / external matrices); row r's self entry has atom == ilist[r].
struct Matrix {
  int n = 0, ncols = 0, nslices = 0; long long total = 0, nnz = 0; int max_row = 0, ndup = 0;
  DevBuf<long long> slice_off; DevBuf<int> slice_len, row_len, diag_k, col, atom; DevBuf<double> val;
  DevBuf<double> diagonal, sld;          // A.diagonal, A.scaled_laplace_diagonal (pair_isph.h:385-392)
  std::vector<long long> h_slice_off;
  int is_filled = 0; bool built = false; bool external = false;
};

struct Timer { cudaEvent_t a = nullptr, b = nullptr; double ms = 0.0; bool open = false, pending = false; };

struct SolverParams {              // SolverLin_Belos::setParameters defaults, solver_lin_belos.h:224-264
  std::string solver_type = "Block GMRES", ortho = "DGKS";
  bool flexible = true; int num_blocks = 50, block_size = 1, max_iters = 500, max_restarts = 15; double tol = 1.0e-8;
};
struct PrecondParams {             // names of precond_ifpack.h:28-48 + Ifpack's own lists; defaults differ where BASELINE says so
  std::string type = "ILU", relax_type = "Jacobi";
  int overlap = 0, fill = 0, sweeps = 1, cheb_degree = 1, cheb_eig_iters = 10;
  double damping = 1.0, min_diag = 0.0, cheb_ratio = 30.0, cheb_lmax = -1.0;
  // "Precond Package" = ML: the multilevel stand-in of amg.cu; names = ML's parameter list (precond_ml.h:44-58), defaults chosen for the GPU
  // (Chebyshev instead of the sequential symmetric Gauss-Seidel, non-smoothed MIS aggregation)
  std::string ml_smoother = "Chebyshev", ml_coarse = "Chebyshev", ml_agg_type = "MIS";
  int ml_max_levels = 5, ml_pre = 1, ml_post = 1, ml_level_sweeps = 3, ml_coarse_sweeps = 8, ml_eig_iters = 10, ml_max_coarse = 128;
  double ml_threshold = 0.02, ml_agg_damping = 0.0, ml_alpha = 2.0, ml_level_alpha = 10.0, ml_coarse_alpha = 30.0, ml_scale = 2.5, ml_level_scale = 2.0, ml_damping = 0.67;   // alpha / scale: finest level; level_*: between the finest and the coarsest
};

struct Halo;      // halo.cu
struct NeighWork; // neighbor.cu
struct IluData;   // precond.cu
struct AmgData;   // amg.cu
struct BlockSys;  // capi.cu: the dim x dim block operator of solveBlockProblem

struct Ctx {
  int device = 0, nranks = 1, rank = 0;
  cudaStream_t stream = nullptr; bool own_stream = false;
  std::string err;
  long long launches = 0;
  // pair
  bool have_pair = false; PairTab tab; DevBuf<PairTab> d_tab; int fixed_of_type[ISPH_MAXT] = {};   // pinfo[1] ("fixed" particle types), pair_isph.cpp:165-167
  // atoms
  int nlocal = 0, nghost = 0, nall = 0, first_fluid_row = -1, max_tag = 0; bool have_atoms = false; unsigned long long tag_hash = 0;
  DevBuf<double> x; DevBuf<int> type, tag, kind, col_of_atom, tag2own; std::vector<int> h_type, h_tag;
  DevBuf<double> field[ISPH_F_COUNT];
  // neighbors
  int inum = 0, max_jnum = 0; long long nneigh = 0; bool have_neigh = false;
  DevBuf<int> ilist, neigh; DevBuf<long long> noff; std::vector<long long> h_noff;
  PinBuf<int> pin_neigh; bool neigh_on_device = false; NeighWork *nwork = nullptr;    // list built by isph_neighbors_build: no host copy of the offsets
  // matrix
  Matrix A;
  // solver state (SolverLin members, solver_lin.h:70-97)
  SolverParams sp; PrecondParams pp;
  double *x_host = nullptr, *b_host = nullptr; int x_lda = 0, b_lda = 0, x_nvec = 0, b_nvec = 0; bool x_owned = true, b_owned = true;
  bool b_dev_fresh = false;        // a device functor / isph_solver_load_set wrote the load vector after the last solve: a borrowed host b is not re-uploaded over it
  DevBuf<double> xs, bs;           // device solution / load multivectors, column-major, leading dimension ld
  int ld = 0;                      // padded length of every Krylov vector (>= ncols)
  bool is_singular = false, have_mask = false; DevBuf<double> nullvec; DevBuf<int> mask;
  int init_type = -1; double init_val = 0.0;
  DevBuf<double> V, Z, wk, red, hbuf; DevBuf<int> flag;      // Krylov workspace
  PinBuf<double> h_scal;
  long long last_second_passes = 0; int last_iters = 0, last_converged = 0; double last_relres = 0.0, last_lmax = 0.0;
  // preconditioner
  bool prec_ready = false; int prec_kind = 0; DevBuf<double> invdiag, cw, cv; DevBuf<int> block_of_row; bool have_blocks = false;
  IluData *ilu = nullptr; AmgData *amg = nullptr;
  BlockSys *blk = nullptr;          // createBlockMatrix / setBlock (solver_lin.cpp:78-138)
  Ctx *prec_parent = nullptr; int prec_dim = 0;   // block-diagonal preconditioner of a block solve: the parent's preconditioner applied to every diagonal block (precond_ifpack.h:77-81, precond_ml.h:137-154)
  DevBuf<double> pb_extra;         // Poisson-Boltzmann extra source term staged for the Newton loop
  // multi-GPU
  Halo *halo = nullptr;
  const double *prepush_x = nullptr; unsigned long long prepush_seq = 0;   // SpMV input whose halo was pushed by its producer kernel
  unsigned char nccl_id[128]; bool have_nccl_id = false;
  // timers
  std::map<std::string, Timer> timers;
  bool prof_spmv = false; std::vector<cudaEvent_t> prof_ev; size_t prof_used = 0; double prof_ms = 0.0; long long prof_cnt = 0;
  std::vector<cudaEvent_t> pprof_ev; size_t pprof_used = 0; double pprof_ms = 0.0; long long pprof_cnt = 0;      // the same for the ILU apply kernel
  // optional per-phase device timing of the Krylov loop (ISPH_PROFILE=1): name -> event pairs
  bool prof_phases = false; std::map<std::string, std::vector<cudaEvent_t>> phase_ev; std::map<std::string, size_t> phase_used;

  void tic(const char *name);
  void toc(const char *name);
};

inline int ceil_div(long long a, int b) { return (int)((a + b - 1) / b); }

struct ProfScope {          // CUDA-event bracket around a phase of the solver when ISPH_PROFILE is set (off: zero cost)
  Ctx *c; cudaEvent_t e1 = nullptr;
  ProfScope(Ctx *c_, const char *name) : c(c_) {
    if (!c->prof_phases) return;
    auto &v = c->phase_ev[name]; size_t &u = c->phase_used[name];
    if (u + 2 > v.size()) { size_t o = v.size(); v.resize(o + 256, nullptr); for (size_t q = o; q < v.size(); ++q) cudaEventCreate(&v[q]); }
    cudaEventRecord(v[u], c->stream); e1 = v[u + 1]; u += 2;
  }
  ~ProfScope() { if (e1) cudaEventRecord(e1, c->stream); }
};

// ---- kernels / drivers implemented across the .cu files ---------------------------------------------------------
void build_column_map(Ctx *c);                               // graph.cu
void graph_build(Ctx *c);
void graph_export(Ctx *c, int *rowptr, int *col_tags, double *val);
void matrix_from_csr(Ctx *c, int n, const int *rowptr, const int *col, const double *val);
void matrix_put_scalar(Ctx *c, double a);
void matrix_scale(Ctx *c, double a);
void matrix_left_scale_dev(Ctx *c, const double *d_s, bool reciprocal);
void matrix_extract_diag_dev(Ctx *c, double *d_out);
void matrix_replace_diag_dev(Ctx *c, const double *d_in);
void matrix_merge_duplicates(Ctx *c);

void compute_volumes(Ctx *c);                                // assemble.cu
void compute_gradient_correction(Ctx *c);
void compute_laplacian_correction(Ctx *c);
void compute_normals(Ctx *c);
void forward_comm(Ctx *c, int field);
void assemble_laplacian(Ctx *c, double alpha, const double *d_material, bool anti, bool mh, int f0, int f1);
void assemble_gradient_dot(Ctx *c, double alpha, const double *d_vec, int f0, int f1);
void ns_poisson(Ctx *c, double dt, bool anti, int singular, bool mh);
void ns_helmholtz(Ctx *c, double dt, double theta, bool anti, bool mh, bool incp, const double *g);
void pb_jacobian(Ctx *c, bool mh, bool linearized, double ezcb, double psiref, double gamma);
void ns_correct(Ctx *c, double dt, bool anti, bool incp, const double *dp_owned_host);
void applied_electric_potential(Ctx *c);
void solute_transport(Ctx *c, double dt, double theta, double dcoeff);
void advance_time(Ctx *c, double dt, bool anti);
void boundary_navier_slip(Ctx *c, double beta);
void boundary_dirichlet(Ctx *c);
void pb_residual(Ctx *c, bool mh, bool linearized, double ezcb, double psiref, double gamma, const double *d_extra, double *d_f);
void forward_comm(Ctx *c, int field);

void spmv(Ctx *c, const double *d_x, double *d_y, int nvec, int ldx, int ldy, const double *dot_vec = nullptr, double *dot_out = nullptr);   // spmv.cu (does the halo exchange when nranks > 1)

void precond_create(Ctx *c);                                 // precond.cu
void precond_free(Ctx *c);
void precond_apply(Ctx *c, const double *d_r, double *d_z);  // z = M^-1 r
void amg_create(Ctx *c);                                     // amg.cu: the multilevel stand-in for PrecondWrapper_ML
void amg_free(Ctx *c);
void amg_destroy(Ctx *c);
void amg_apply(Ctx *c, const double *d_r, double *d_z);
int amg_info(Ctx *c, int *rows, long long *nnz, double *lmax, int cap);
void amg_aggregates(Ctx *c, int *agg_host);
double amg_setup_ms(Ctx *c, const char *phase);
void ilu_destroy(Ctx *c);                                    // ilu.cu
bool ilu_fault(Ctx *c);
void ilu_info(Ctx *c, long long *nnz, int *nlev_l, int *nlev_u, int *maxlen);

void load_from_host(Ctx *c);                                  // capi.cu: borrowed (View) load vector, see there
void load_written(Ctx *c);
void solver_prepare_vectors(Ctx *c);                         // krylov.cu
void solver_solve(Ctx *c, bool use_prec, const char *label);
void pb_newton(Ctx *c, bool mh, bool linearized, double ezcb, double psiref, double gamma, const double *d_extra, int max_newton, double tol_f, double tol_update,
               bool use_prec, int *newton_iters, int *linear_iters, double *normf, int *converged);

// device-resident exchange plan of the halo kernels (written once per halo_setup; indexed dynamically on the device,
// which is why it is not a by-value kernel parameter).  Staging buffer of a rank: [MB_SLOTS][3 vectors][cap] doubles.
struct HaloDev {
  double *peer[ISPH_MAX_RANKS]; double *mine;                 // halo staging buffers (peer-mapped) and this rank's own
  int send_off[ISPH_MAX_RANKS + 1], dst_off[ISPH_MAX_RANKS], recv_cnt[ISPH_MAX_RANKS];
  int nranks, rank; long long cap; int *fault;
};
// "push from the producer": the kernel that WRITES the next SpMV input (k_finish: z_{j+1} = D^-1 v_{j+1}) also stores the
// rows on the send list straight into the peers' staging buffers, so the import of that SpMV is already under way (usually
// complete) when the SpMV is launched.  sp/sd = per-row send list (CSR over owned rows; entry = peer << 28 | offset).
struct PrePush { const HaloDev *plan; const int *sp, *sd; unsigned long long seq; };

void neighbors_build(Ctx *c, double cutneigh);                // neighbor.cu
void neighbors_destroy(Ctx *c);
long long slice_offsets_device(Ctx *c, int n, int nslices, long long *d_slice_off);
// matrix rows of the halo particles, imported from their owners for "Overlap Level" 1 (halo.cu): row of halo slot s = entries
// off_r[s] .. off_r[s+1] of (col_r2, val_r2), sorted by local column id, columns outside the extended set = 0x7fffffff at the end
struct OverlapRows {
  int nhalo = 0; long long total = 0;
  DevBuf<int> coltag, tag2col, len_s, len_r, tag_s, tag_r, col_r, col_r2; DevBuf<long long> off_s, off_r; DevBuf<double> val_s, val_r, val_r2;
  void release() { coltag.release(); tag2col.release(); len_s.release(); len_r.release(); tag_s.release(); tag_r.release(); col_r.release(); col_r2.release(); off_s.release(); off_r.release(); val_s.release(); val_r.release(); val_r2.release(); }
};
void halo_import_rows(Ctx *c, OverlapRows *out, const int *slot_ref /* per halo slot: part of the extended set? (device) */);
void halo_export_add(Ctx *c, const double *zext_halo, double *z);
void halo_setup(Ctx *c);                                     // halo.cu
bool halo_prepush_begin(Ctx *c, const double *x_next, PrePush *pp);   // reserves the exchange of the SpMV that will read x_next
void halo_wait_unstage(Ctx *c, double *d_x, unsigned long long seq);  // completes a pre-pushed exchange: halo lands behind x's owned rows
void halo_exchange(Ctx *c, double *d_x, int nvec, int ldx);   // import: off-rank entries of x land behind its owned rows
P2PRed halo_p2p_ticket(Ctx *c);          // sequence ticket for an in-kernel peer all-reduce (nranks <= 1 in it: not available)
bool halo_fault(Ctx *c);
void halo_recover(Ctx *c);                                   // collective: re-arm the peer slots after a timed-out wait
void halo_allreduce(Ctx *c, double *d_buf, int count);
void halo_forward_field(Ctx *c, int field, int ncomp);
void halo_destroy(Ctx *c);
int halo_ncols(Ctx *c);
void halo_counts(Ctx *c, int *nhalo, int *nsend, int *npeers);

}  // namespace isph
