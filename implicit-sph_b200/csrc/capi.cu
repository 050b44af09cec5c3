// extern "C" boundary (include/isph_b200.h).  Thin: argument checks, host<->device copies, exception -> status code.
#include "isph_internal.h"
#include <algorithm>

namespace isph {

static void timer_flush(Timer &t) {
  if (!t.pending) return;
  cudaEventSynchronize(t.b); float ms = 0.f; cudaEventElapsedTime(&ms, t.a, t.b); t.ms += ms; t.pending = false;
}
void Ctx::tic(const char *name) {
  Timer &t = timers[name];
  if (!t.a) { cudaEventCreate(&t.a); cudaEventCreate(&t.b); }
  timer_flush(t);
  cudaEventRecord(t.a, stream); t.open = true;
}
void Ctx::toc(const char *name) {      // no host synchronisation here: the elapsed time is collected when it is asked for
  Timer &t = timers[name]; if (!t.open) return;
  cudaEventRecord(t.b, stream); t.open = false; t.pending = true;
}

static const int FIELD_NC[ISPH_F_COUNT] = {1, 9, 6, 3, 1, 1, 1, 1, 3, 3, 3, 1, 1, 1, 1, 1, 1};

static void ensure_fields(Ctx *c) {
  for (int f = 0; f < ISPH_F_COUNT; ++f) {
    const size_t need = (size_t)c->nall * FIELD_NC[f];
    if (need > c->field[f].cap) {
      c->field[f].ensure(need);
      CUDA_CHECK(cudaMemsetAsync(c->field[f].p, 0, sizeof(double) * c->field[f].cap, c->stream));
      if (f == ISPH_F_DENSITY || f == ISPH_F_EPS) {   // same defaults as the oracle: rho = 1, eps = 1
        std::vector<double> ones(c->field[f].cap, 1.0);
        CUDA_CHECK(cudaMemcpyAsync(c->field[f].p, ones.data(), sizeof(double) * ones.size(), cudaMemcpyHostToDevice, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
      }
    }
  }
}

static void compute_first_fluid_row(Ctx *c) {
  c->first_fluid_row = -1;
  if (!c->have_pair || !c->have_atoms) return;
  for (int i = 0; i < c->nlocal; ++i) { const int k = c->tab.kind[c->h_type[i]];
    if (k == ISPH_KIND_FLUID || k == ISPH_KIND_BUFFER_DIRICHLET || k == ISPH_KIND_BUFFER_NEUMANN) { c->first_fluid_row = i; break; } }
}

static double kernel_C(int kernel, int dim, double h) {   // kernel_{wendland,cubic,quintic}.h setSmoothingLength
  switch (kernel) {
  case ISPH_KERNEL_CUBIC: return dim == 3 ? 1.0 / (pow(h, 3) * M_PI) : 10.0 / (pow(h, 2) * 7.0 * M_PI);
  case ISPH_KERNEL_QUINTIC: return dim == 3 ? 14.0 / (pow(h, 3) * 1745.0 * M_PI) : 7.0 / (pow(h, 2) * 478.0 * M_PI);
  default: return dim == 3 ? 21.0 / (16 * M_PI * pow(h, 3)) : 7.0 / (4 * M_PI * pow(h, 2));
  }
}

// A load multivector created over caller memory is a View in the reference (solver_lin.cpp:45-58): the functors write the
// right-hand side INTO that memory and the caller may write it too, at any time before solveProblem.  Here the functors write
// the device copy, so (1) what a device functor wrote is copied back into the borrowed host array, and the solve does not
// upload the (older) host content over it; (2) whatever else is in the host array when a functor that reads b (Helmholtz,
// solute transport) or the solve starts is uploaded first.
void load_from_host(Ctx *c) {
  if (!c->b_host || c->b_owned || c->b_dev_fresh || c->b_nvec < 1) return;
  for (int q = 0; q < c->b_nvec; ++q) CUDA_CHECK(cudaMemcpyAsync(c->bs.p + (size_t)q * c->ld, c->b_host + (size_t)q * c->b_lda, sizeof(double) * c->A.n, cudaMemcpyHostToDevice, c->stream));
}
void load_written(Ctx *c) {
  c->b_dev_fresh = true;
  if (!c->b_host || c->b_owned || c->b_nvec < 1) return;
  for (int q = 0; q < c->b_nvec; ++q) CUDA_CHECK(cudaMemcpyAsync(c->b_host + (size_t)q * c->b_lda, c->bs.p + (size_t)q * c->ld, sizeof(double) * c->A.n, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

}  // namespace isph

using namespace isph;

#define API_BEGIN(ctx) if (!(ctx)) return ISPH_FAILURE; Ctx *c = reinterpret_cast<Ctx *>(ctx); try { CUDA_CHECK(cudaSetDevice(c->device));
#define API_END } catch (const std::exception &e) { c->err = e.what(); cudaGetLastError(); return ISPH_FAILURE; } return ISPH_SUCCESS;

namespace isph { struct BlockSys { int dim = 0, n = 0; std::string name; std::vector<std::vector<int>> rp, ci; std::vector<std::vector<double>> va; isph_ctx *child = nullptr; bool filled = false; }; }

extern "C" {

const char *isph_version(void) { return "isph_b200 0.1 (sm_100a)"; }

int isph_ctx_create(isph_ctx **out, int device, int nranks, int rank, const void *nccl_unique_id) {
  if (!out) return ISPH_FAILURE;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); fprintf(stderr, "isph_b200: no CUDA device (there is no CPU fallback)\n"); return ISPH_FAILURE; }
  if (device < 0 || device >= ndev || nranks < 1 || rank < 0 || rank >= nranks) return ISPH_FAILURE;
  Ctx *c = new Ctx();
  try {
    c->device = device; c->nranks = nranks; c->rank = rank;
    CUDA_CHECK(cudaSetDevice(device));
    CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true;
    if (nranks > 1) { ISPH_REQUIRE(nccl_unique_id, "nranks > 1 needs the NCCL unique id"); memcpy(c->nccl_id, nccl_unique_id, 128); c->have_nccl_id = true; }
    c->h_scal.ensure(1024); c->flag.ensure(16); c->hbuf.ensure(8192);
    CUDA_CHECK(cudaMemset(c->flag.p, 0, 16 * sizeof(int)));
  } catch (const std::exception &e) { fprintf(stderr, "isph_ctx_create: %s\n", e.what()); delete c; return ISPH_FAILURE; }
  *out = reinterpret_cast<isph_ctx *>(c);
  return ISPH_SUCCESS;
}

int isph_ctx_destroy(isph_ctx *ctx) {
  if (!ctx) return ISPH_FAILURE; Ctx *c = reinterpret_cast<Ctx *>(ctx);
  cudaSetDevice(c->device); cudaStreamSynchronize(c->stream);
  if (c->prec_ready) { try { precond_free(c); } catch (...) {} }
  if (c->blk) { if (c->blk->child) isph_ctx_destroy(c->blk->child); delete c->blk; c->blk = nullptr; }
  halo_destroy(c); ilu_destroy(c); amg_destroy(c); neighbors_destroy(c);
  c->d_tab.release(); c->x.release(); c->type.release(); c->tag.release(); c->kind.release(); c->col_of_atom.release(); c->tag2own.release();
  for (auto &f : c->field) f.release();
  c->ilist.release(); c->neigh.release(); c->noff.release(); c->pin_neigh.release();
  Matrix &A = c->A; A.slice_off.release(); A.slice_len.release(); A.row_len.release(); A.diag_k.release(); A.col.release(); A.atom.release(); A.val.release(); A.diagonal.release(); A.sld.release();
  c->xs.release(); c->bs.release(); c->nullvec.release(); c->mask.release(); c->V.release(); c->Z.release(); c->wk.release(); c->red.release(); c->hbuf.release(); c->flag.release(); c->h_scal.release();
  c->invdiag.release(); c->cw.release(); c->cv.release(); c->block_of_row.release(); c->pb_extra.release();
  for (auto &kv : c->timers) { if (kv.second.a) cudaEventDestroy(kv.second.a); if (kv.second.b) cudaEventDestroy(kv.second.b); }
  for (auto e : c->prof_ev) if (e) cudaEventDestroy(e);
  for (auto e : c->pprof_ev) if (e) cudaEventDestroy(e);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c; return ISPH_SUCCESS;
}

const char *isph_last_error(const isph_ctx *ctx) { return ctx ? reinterpret_cast<const Ctx *>(ctx)->err.c_str() : "null context"; }

int isph_set_stream(isph_ctx *ctx, void *s) { API_BEGIN(ctx) CUDA_CHECK(cudaStreamSynchronize(c->stream)); if (c->own_stream) cudaStreamDestroy(c->stream); c->stream = (cudaStream_t)s; c->own_stream = false; API_END }
int isph_synchronize(isph_ctx *ctx) { API_BEGIN(ctx) CUDA_CHECK(cudaStreamSynchronize(c->stream)); API_END }

// ---- pair / atoms / neighbors -------------------------------------------------------------------------------------
int isph_pair_coeff(isph_ctx *ctx, int dim, int ntypes, const int *kind_of_type, double h, double h_min, double cut_over_h, int kernel, double morris_safe) {
  API_BEGIN(ctx)
  ISPH_REQUIRE(dim == 2 || dim == 3, "dimension must be 2 or 3"); ISPH_REQUIRE(ntypes >= 1 && ntypes < ISPH_MAXT, "1 <= ntypes <= 7");
  ISPH_REQUIRE(kernel >= 0 && kernel <= 2, "Kernel is not in supported list: Wendland, Quintic, and Cubic");      // pair_isph_corrected.cpp:1301
  PairTab &T = c->tab; memset(&T, 0, sizeof(T));
  T.dim = dim; T.ntypes = ntypes; T.kernel = kernel; T.morris_safe = morris_safe;
  for (int i = 0; i <= ntypes; ++i) T.kind[i] = kind_of_type[i];
  const double cut_one = h * cut_over_h, cut_one_sq = cut_one * cut_one;                                           // :1303-1310
  for (int i = 1; i <= ntypes; ++i) for (int j = 1; j <= ntypes; ++j) {
    T.cutsq[i][j] = cut_one_sq; T.cut[i][j] = sqrt(cut_one_sq);
    const double hij = (T.kind[i] == T.kind[j]) ? h : h_min;                                                       // :1325-1328
    T.h[i][j] = hij; const double C = kernel_C(kernel, dim, hij); T.kC[i][j] = C; T.kCh[i][j] = C / hij;
  }
  c->d_tab.ensure(1);
  CUDA_CHECK(cudaMemcpyAsync(c->d_tab.p, &T, sizeof(T), cudaMemcpyHostToDevice, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream));
  c->have_pair = true; c->A.built = false; compute_first_fluid_row(c);
  API_END
}

int isph_atoms_set(isph_ctx *ctx, int nlocal, int nghost, const double *x, const int *type, const int *tag) {
  API_BEGIN(ctx)
  ISPH_REQUIRE(c->have_pair, "isph_pair_coeff must be called before isph_atoms_set");
  ISPH_REQUIRE(nlocal >= 0 && nghost >= 0 && x && type && tag, "bad atom arrays");
  c->tic("h2dAtoms");
  c->nlocal = nlocal; c->nghost = nghost; c->nall = nlocal + nghost; const int nall = c->nall;
  c->x.ensure((size_t)3 * nall); c->type.ensure(nall); c->tag.ensure(nall);
  CUDA_CHECK(cudaMemcpyAsync(c->x.p, x, sizeof(double) * 3 * nall, cudaMemcpyHostToDevice, c->stream));
  CUDA_CHECK(cudaMemcpyAsync(c->type.p, type, sizeof(int) * nall, cudaMemcpyHostToDevice, c->stream));
  CUDA_CHECK(cudaMemcpyAsync(c->tag.p, tag, sizeof(int) * nall, cudaMemcpyHostToDevice, c->stream));
  c->h_type.assign(type, type + nall); c->h_tag.assign(tag, tag + nall);
  int mt = 0; unsigned long long hsh = 0x9E3779B97F4A7C15ull ^ ((unsigned long long)nlocal << 32 | (unsigned)nghost);      // hash of the tag set: key of the halo-plan cache
  for (int i = 0; i < nall; ++i) { ISPH_REQUIRE(tag[i] >= 0, "negative atom tag"); mt = std::max(mt, tag[i]); ISPH_REQUIRE(type[i] >= 1 && type[i] <= c->tab.ntypes, "atom type out of range");
    hsh = (hsh ^ (unsigned long long)(unsigned)tag[i]) * 0x100000001B3ull; hsh ^= hsh >> 29; }
  c->max_tag = mt; c->tag_hash = hsh; c->have_atoms = true; c->A.built = false;
  ensure_fields(c);
  compute_first_fluid_row(c);
  c->toc("h2dAtoms");
  build_column_map(c);
  API_END
}

static void finish_neighbors(Ctx *c, int inum, const int *ilist) {
  for (int ii = 0; ii < inum; ++ii) ISPH_REQUIRE(ilist[ii] == ii, "ilist must enumerate the owned atoms in order (LAMMPS full neighbor lists do)");
  c->inum = inum; c->ilist.ensure(inum); c->noff.ensure(inum + 1);
  CUDA_CHECK(cudaMemcpyAsync(c->ilist.p, ilist, sizeof(int) * inum, cudaMemcpyHostToDevice, c->stream));
  CUDA_CHECK(cudaMemcpyAsync(c->noff.p, c->h_noff.data(), sizeof(long long) * (inum + 1), cudaMemcpyHostToDevice, c->stream));
  long long mj = 0; for (int ii = 0; ii < inum; ++ii) mj = std::max(mj, c->h_noff[ii + 1] - c->h_noff[ii]);
  c->max_jnum = (int)mj; c->nneigh = c->h_noff[inum]; c->have_neigh = true; c->neigh_on_device = false; c->A.built = false;
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

int isph_neighbors_set(isph_ctx *ctx, int inum, const int *ilist, const int *numneigh, int *const *firstneigh) {
  API_BEGIN(ctx)
  ISPH_REQUIRE(inum >= 0 && ilist && numneigh && firstneigh, "bad neighbor list");
  c->tic("h2dNeighbors");
  c->h_noff.assign(inum + 1, 0);
  for (int ii = 0; ii < inum; ++ii) c->h_noff[ii + 1] = c->h_noff[ii] + numneigh[ilist[ii]];
  const long long tot = c->h_noff[inum];
  c->pin_neigh.ensure(tot); c->neigh.ensure(tot);
  for (int ii = 0; ii < inum; ++ii) memcpy(c->pin_neigh.p + c->h_noff[ii], firstneigh[ilist[ii]], sizeof(int) * numneigh[ilist[ii]]);   // LAMMPS pages -> one pinned block
  CUDA_CHECK(cudaMemcpyAsync(c->neigh.p, c->pin_neigh.p, sizeof(int) * tot, cudaMemcpyHostToDevice, c->stream));
  finish_neighbors(c, inum, ilist);
  c->toc("h2dNeighbors");
  API_END
}

int isph_neighbors_set_packed(isph_ctx *ctx, int inum, const int *ilist, const long long *noff, const int *neigh) {
  API_BEGIN(ctx)
  ISPH_REQUIRE(inum >= 0 && ilist && noff && neigh, "bad neighbor list");
  c->tic("h2dNeighbors");
  c->h_noff.assign(noff, noff + inum + 1);
  const long long tot = noff[inum]; c->neigh.ensure(tot);
  CUDA_CHECK(cudaMemcpyAsync(c->neigh.p, neigh, sizeof(int) * tot, cudaMemcpyHostToDevice, c->stream));
  finish_neighbors(c, inum, ilist);
  c->toc("h2dNeighbors");
  API_END
}

int isph_neighbors_build(isph_ctx *ctx, double cutneigh) { API_BEGIN(ctx) neighbors_build(c, cutneigh); API_END }
long long isph_neighbors_count(isph_ctx *ctx) { if (!ctx) return -1; Ctx *c = reinterpret_cast<Ctx *>(ctx); return c->have_neigh ? c->nneigh : -1; }
int isph_neighbors_get(isph_ctx *ctx, long long *noff, int *neigh) {
  API_BEGIN(ctx) ISPH_REQUIRE(c->have_neigh && noff && neigh, "no neighbor list");
  CUDA_CHECK(cudaMemcpyAsync(noff, c->noff.p, sizeof(long long) * (c->inum + 1), cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaMemcpyAsync(neigh, c->neigh.p, sizeof(int) * c->nneigh, cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream)); API_END
}
int isph_field_set(isph_ctx *ctx, int f, const double *data) {
  API_BEGIN(ctx) ISPH_REQUIRE(f >= 0 && f < ISPH_F_COUNT && data && c->have_atoms, "bad field / atoms not set");
  CUDA_CHECK(cudaMemcpyAsync(c->field[f].p, data, sizeof(double) * c->nall * FIELD_NC[f], cudaMemcpyHostToDevice, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream)); API_END
}
int isph_field_get(isph_ctx *ctx, int f, double *data) {
  API_BEGIN(ctx) ISPH_REQUIRE(f >= 0 && f < ISPH_F_COUNT && data && c->have_atoms, "bad field / atoms not set");
  CUDA_CHECK(cudaMemcpyAsync(data, c->field[f].p, sizeof(double) * c->nall * FIELD_NC[f], cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream)); API_END
}
int isph_forward_comm(isph_ctx *ctx, int f) { API_BEGIN(ctx) ISPH_REQUIRE(f >= 0 && f < ISPH_F_COUNT, "bad field"); forward_comm(c, f); API_END }

int isph_compute_volumes(isph_ctx *ctx) { API_BEGIN(ctx) compute_volumes(c); API_END }
int isph_compute_gradient_correction(isph_ctx *ctx) { API_BEGIN(ctx) compute_gradient_correction(c); API_END }
int isph_compute_laplacian_correction(isph_ctx *ctx) { API_BEGIN(ctx) compute_laplacian_correction(c); API_END }
int isph_compute_normals(isph_ctx *ctx) { API_BEGIN(ctx) compute_normals(c); API_END }

// ---- graph + matrix ---------------------------------------------------------------------------------------------
int isph_graph_build(isph_ctx *ctx) {
  API_BEGIN(ctx)
  c->tic("computeGraph");
  if (c->A.built && !c->A.external) {   // same atoms + list as the pattern already on the device: a fresh zero matrix on it (pair_isph.cpp:1266-1270)
    Matrix &A = c->A;
    CUDA_CHECK(cudaMemsetAsync(A.val.p, 0, sizeof(double) * A.total, c->stream));
    CUDA_CHECK(cudaMemsetAsync(A.diagonal.p, 0, sizeof(double) * A.n, c->stream)); CUDA_CHECK(cudaMemsetAsync(A.sld.p, 0, sizeof(double) * A.n, c->stream));
    A.is_filled = 0;
  } else graph_build(c);
  c->toc("computeGraph");
  API_END
}
long long isph_graph_nnz(isph_ctx *ctx) { if (!ctx) return -1; return reinterpret_cast<Ctx *>(ctx)->A.built ? reinterpret_cast<Ctx *>(ctx)->A.nnz : -1; }
int isph_graph_max_row(isph_ctx *ctx) {
  if (!ctx) return -1; Ctx *c = reinterpret_cast<Ctx *>(ctx); if (!c->A.built) return -1;
  try { if (c->A.max_row < 0) graph_export(c, nullptr, nullptr, nullptr); } catch (const std::exception &e) { c->err = e.what(); return -1; }
  return c->A.max_row;
}
int isph_graph_get(isph_ctx *ctx, int *rowptr, int *col_tags) { API_BEGIN(ctx) graph_export(c, rowptr, col_tags, nullptr); API_END }
int isph_matrix_get(isph_ctx *ctx, double *val) { API_BEGIN(ctx) graph_export(c, nullptr, nullptr, val); API_END }
int isph_matrix_set_csr(isph_ctx *ctx, int n, const int *rowptr, const int *col, const double *val) {
  API_BEGIN(ctx) ISPH_REQUIRE(n > 0 && rowptr && col && val, "bad csr"); ISPH_REQUIRE(c->nranks == 1, "isph_matrix_set_csr is single-rank"); matrix_from_csr(c, n, rowptr, col, val); API_END
}
int isph_matrix_put_scalar(isph_ctx *ctx, double a) { API_BEGIN(ctx) ISPH_REQUIRE(c->A.built, "no matrix"); matrix_put_scalar(c, a); API_END }
int isph_matrix_scale(isph_ctx *ctx, double a) { API_BEGIN(ctx) ISPH_REQUIRE(c->A.built, "no matrix"); matrix_scale(c, a); API_END }
int isph_matrix_left_scale(isph_ctx *ctx, const double *s) {
  API_BEGIN(ctx) ISPH_REQUIRE(c->A.built && s, "no matrix"); c->cv.ensure(c->ld);
  CUDA_CHECK(cudaMemcpyAsync(c->cv.p, s, sizeof(double) * c->A.n, cudaMemcpyHostToDevice, c->stream)); matrix_left_scale_dev(c, c->cv.p, false); CUDA_CHECK(cudaStreamSynchronize(c->stream)); API_END
}
int isph_matrix_extract_diagonal(isph_ctx *ctx, double *d) {
  API_BEGIN(ctx) ISPH_REQUIRE(c->A.built && d, "no matrix"); c->cv.ensure(c->ld); matrix_extract_diag_dev(c, c->cv.p);
  CUDA_CHECK(cudaMemcpyAsync(d, c->cv.p, sizeof(double) * c->A.n, cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream)); API_END
}
int isph_matrix_replace_diagonal(isph_ctx *ctx, const double *d) {
  API_BEGIN(ctx) ISPH_REQUIRE(c->A.built && d, "no matrix"); c->cv.ensure(c->ld);
  CUDA_CHECK(cudaMemcpyAsync(c->cv.p, d, sizeof(double) * c->A.n, cudaMemcpyHostToDevice, c->stream)); matrix_replace_diag_dev(c, c->cv.p); CUDA_CHECK(cudaStreamSynchronize(c->stream)); API_END
}
int isph_matrix_multiply(isph_ctx *ctx, const double *x, double *y, int lda, int nvec) {
  API_BEGIN(ctx) ISPH_REQUIRE(c->A.built && x && y && nvec >= 1 && lda >= c->A.n, "bad multiply arguments");
  const int ld = c->ld, n = c->A.n; c->V.ensure((size_t)2 * nvec * ld);
  double *dx = c->V.p, *dy = c->V.p + (size_t)nvec * ld;
  for (int q = 0; q < nvec; ++q) CUDA_CHECK(cudaMemcpyAsync(dx + (size_t)q * ld, x + (size_t)q * lda, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
  spmv(c, dx, dy, nvec, ld, ld);
  for (int q = 0; q < nvec; ++q) CUDA_CHECK(cudaMemcpyAsync(y + (size_t)q * lda, dy + (size_t)q * ld, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream)); API_END
}
int isph_matrix_invalidate(isph_ctx *ctx) { API_BEGIN(ctx) c->A.is_filled = 0; API_END }
int isph_graph_invalidate(isph_ctx *ctx) { API_BEGIN(ctx) c->A.built = false; c->A.is_filled = 0; API_END }
int isph_assemble_laplacian(isph_ctx *ctx, double alpha, int mf, int anti, int mh, int f0, int f1) {
  API_BEGIN(ctx) ISPH_REQUIRE(mf < ISPH_F_COUNT, "bad material field"); assemble_laplacian(c, alpha, mf < 0 ? nullptr : c->field[mf].p, anti != 0, mh != 0, f0, f1); API_END
}
int isph_assemble_gradient_dot(isph_ctx *ctx, double alpha, int vf, int f0, int f1) {
  API_BEGIN(ctx) ISPH_REQUIRE(vf >= 0 && vf < ISPH_F_COUNT && FIELD_NC[vf] == 3, "vector field expected"); assemble_gradient_dot(c, alpha, c->field[vf].p, f0, f1); API_END
}

int isph_ns_poisson(isph_ctx *ctx, double dt, int anti, int singular, int mh) { API_BEGIN(ctx) ns_poisson(c, dt, anti != 0, singular, mh != 0); load_written(c); API_END }
int isph_ns_helmholtz(isph_ctx *ctx, double dt, double theta, int anti, int mh, int incp, const double *g) { API_BEGIN(ctx) load_from_host(c); ns_helmholtz(c, dt, theta, anti != 0, mh != 0, incp != 0, g); load_written(c); API_END }
int isph_pb_jacobian(isph_ctx *ctx, int mh, int lin, double ezcb, double psiref, double gamma) { API_BEGIN(ctx) pb_jacobian(c, mh != 0, lin != 0, ezcb, psiref, gamma); API_END }
int isph_applied_electric_potential(isph_ctx *ctx) { API_BEGIN(ctx) applied_electric_potential(c); load_written(c); API_END }
int isph_solute_transport(isph_ctx *ctx, double dt, double theta, double dcoeff) { API_BEGIN(ctx) load_from_host(c); solute_transport(c, dt, theta, dcoeff); load_written(c); API_END }
// extra source (functor_poisson_boltzmann_extra_f.h) staged on the device, indexed by owned atom
static const double *stage_extra(Ctx *c, const double *extra_f) {
  if (!extra_f) return nullptr;
  c->cv.ensure(c->ld > c->nlocal ? c->ld : c->nlocal);
  CUDA_CHECK(cudaMemcpyAsync(c->cv.p, extra_f, sizeof(double) * c->nlocal, cudaMemcpyHostToDevice, c->stream));
  return c->cv.p;
}
int isph_pb_residual(isph_ctx *ctx, int mh, int lin, double ezcb, double psiref, double gamma, const double *extra_f, double *f_out) {
  API_BEGIN(ctx) ISPH_REQUIRE(c->have_atoms && c->have_neigh, "atoms and neighbors must be set first");
  double *df; if (c->b_nvec >= 1 && c->bs.p) df = c->bs.p; else { c->wk.ensure((size_t)c->nall + (size_t)c->nlocal); df = c->wk.p + c->nall; }
  pb_residual(c, mh != 0, lin != 0, ezcb, psiref, gamma, stage_extra(c, extra_f), df);
  if (df == c->bs.p) load_written(c);
  if (f_out) CUDA_CHECK(cudaMemcpyAsync(f_out, df, sizeof(double) * c->nlocal, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream)); API_END
}
int isph_pb_newton(isph_ctx *ctx, int mh, int lin, double ezcb, double psiref, double gamma, const double *extra_f, int max_newton, double tol_f, double tol_update,
                   int use_prec, int *newton_iters, int *linear_iters, double *normf, int *converged) {
  API_BEGIN(ctx) int a = 0, b = 0, cv = 0; double nf = 0.0;
  // the extra source must not share a buffer with the preconditioner work vectors used inside the solves
  const double *dex = nullptr;
  if (extra_f) { c->pb_extra.ensure(c->nlocal); CUDA_CHECK(cudaMemcpyAsync(c->pb_extra.p, extra_f, sizeof(double) * c->nlocal, cudaMemcpyHostToDevice, c->stream)); dex = c->pb_extra.p; }
  pb_newton(c, mh != 0, lin != 0, ezcb, psiref, gamma, dex, max_newton, tol_f, tol_update, use_prec != 0, &a, &b, &nf, &cv);
  load_written(c); c->b_dev_fresh = false;                        // F(psi) of the last iteration is what the load vector holds
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  if (newton_iters) *newton_iters = a; if (linear_iters) *linear_iters = b; if (normf) *normf = nf; if (converged) *converged = cv; API_END
}
int isph_pair_fixed(isph_ctx *ctx, const int *fixed_of_type) {
  API_BEGIN(ctx) ISPH_REQUIRE(c->have_pair && fixed_of_type, "isph_pair_coeff first"); for (int t = 0; t < ISPH_MAXT; ++t) c->fixed_of_type[t] = (t <= c->tab.ntypes) ? (fixed_of_type[t] != 0) : 0; API_END
}
int isph_advance_time(isph_ctx *ctx, double dt, int anti) { API_BEGIN(ctx) advance_time(c, dt, anti != 0); CUDA_CHECK(cudaStreamSynchronize(c->stream)); API_END }
int isph_atoms_get_x(isph_ctx *ctx, double *x) {
  API_BEGIN(ctx) ISPH_REQUIRE(c->have_atoms && x, "atoms not set");
  CUDA_CHECK(cudaMemcpyAsync(x, c->x.p, sizeof(double) * 3 * c->nall, cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream)); API_END
}
int isph_boundary_navier_slip(isph_ctx *ctx, double beta) { API_BEGIN(ctx) boundary_navier_slip(c, beta); API_END }
int isph_boundary_dirichlet(isph_ctx *ctx) { API_BEGIN(ctx) boundary_dirichlet(c); load_written(c); API_END }
int isph_ns_correct(isph_ctx *ctx, double dt, int anti, int incp, const double *dp) { API_BEGIN(ctx) ns_correct(c, dt, anti != 0, incp != 0, dp); CUDA_CHECK(cudaStreamSynchronize(c->stream)); API_END }
int isph_diagonals_get(isph_ctx *ctx, double *d, double *s) {
  API_BEGIN(ctx) ISPH_REQUIRE(c->A.built, "no matrix");
  if (d) CUDA_CHECK(cudaMemcpyAsync(d, c->A.diagonal.p, sizeof(double) * c->A.n, cudaMemcpyDeviceToHost, c->stream));
  if (s) CUDA_CHECK(cudaMemcpyAsync(s, c->A.sld.p, sizeof(double) * c->A.n, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream)); API_END
}

// ---- SolverLin mirror ---------------------------------------------------------------------------------------------
static void alloc_mv(Ctx *c, DevBuf<double> &buf, int nvec) {
  ISPH_REQUIRE(c->A.built, "set the matrix / build the graph before creating multivectors (the row map comes from it)");
  const size_t need = (size_t)c->ld * nvec;
  buf.ensure(need); CUDA_CHECK(cudaMemsetAsync(buf.p, 0, sizeof(double) * need, c->stream));
}
int isph_solver_create_solution_multivector(isph_ctx *ctx, double *x, int lda, int nvec) {
  API_BEGIN(ctx) ISPH_REQUIRE(nvec >= 1 && nvec <= 3 && (x == nullptr || lda >= c->A.n), "bad solution multivector");
  alloc_mv(c, c->xs, nvec); c->x_host = x; c->x_lda = lda; c->x_nvec = nvec; c->x_owned = (x == nullptr); c->init_type = -1; API_END
}
int isph_solver_create_load_multivector(isph_ctx *ctx, double *b, int lda, int nvec) {
  API_BEGIN(ctx) ISPH_REQUIRE(nvec >= 1 && nvec <= 3 && (b == nullptr || lda >= c->A.n), "bad load multivector");
  alloc_mv(c, c->bs, nvec); c->b_host = b; c->b_lda = lda; c->b_nvec = nvec; c->b_owned = (b == nullptr); c->b_dev_fresh = false; API_END
}
int isph_solver_load_set(isph_ctx *ctx, const double *b, int lda) {
  API_BEGIN(ctx) ISPH_REQUIRE(c->b_nvec >= 1 && b && lda >= c->A.n, "no load multivector");
  for (int q = 0; q < c->b_nvec; ++q) CUDA_CHECK(cudaMemcpyAsync(c->bs.p + (size_t)q * c->ld, b + (size_t)q * lda, sizeof(double) * c->A.n, cudaMemcpyHostToDevice, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream)); load_written(c); API_END
}
int isph_solver_load_get(isph_ctx *ctx, double *b, int lda) {
  API_BEGIN(ctx) ISPH_REQUIRE(c->b_nvec >= 1 && b && lda >= c->A.n, "no load multivector");
  for (int q = 0; q < c->b_nvec; ++q) CUDA_CHECK(cudaMemcpyAsync(b + (size_t)q * lda, c->bs.p + (size_t)q * c->ld, sizeof(double) * c->A.n, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream)); API_END
}
int isph_solver_solution_set(isph_ctx *ctx, const double *x, int lda) {      // write getSolutionMultiVector()->Values(): the initial guess of an owned x
  API_BEGIN(ctx) ISPH_REQUIRE(c->x_nvec >= 1 && x && lda >= c->A.n, "no solution multivector");
  for (int q = 0; q < c->x_nvec; ++q) CUDA_CHECK(cudaMemcpyAsync(c->xs.p + (size_t)q * c->ld, x + (size_t)q * lda, sizeof(double) * c->A.n, cudaMemcpyHostToDevice, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream)); API_END
}
int isph_solver_solution_get(isph_ctx *ctx, double *x, int lda) {
  API_BEGIN(ctx) ISPH_REQUIRE(c->x_nvec >= 1 && x && lda >= c->A.n, "no solution multivector");
  for (int q = 0; q < c->x_nvec; ++q) CUDA_CHECK(cudaMemcpyAsync(x + (size_t)q * lda, c->xs.p + (size_t)q * c->ld, sizeof(double) * c->A.n, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream)); API_END
}
int isph_solver_set_null_vector_mask(isph_ctx *ctx, const int *mask) {
  API_BEGIN(ctx)
  if (!mask) { c->have_mask = false; }
  else { ISPH_REQUIRE(c->A.built, "no matrix"); c->mask.ensure(c->A.n); CUDA_CHECK(cudaMemcpyAsync(c->mask.p, mask, sizeof(int) * c->A.n, cudaMemcpyHostToDevice, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream)); c->have_mask = true; }
  API_END
}
int isph_solver_set_matrix_is_singular(isph_ctx *ctx, int s) { API_BEGIN(ctx) c->is_singular = s != 0; API_END }
int isph_solver_set_initial_solution(isph_ctx *ctx, int t, double v) { API_BEGIN(ctx) ISPH_REQUIRE(t >= 0 && t <= 2, "bad SolutionInitType"); c->init_type = t; c->init_val = v; API_END }

int isph_solver_set_default_params(isph_ctx *ctx) { API_BEGIN(ctx) c->sp = SolverParams(); API_END }
int isph_solver_set_param_int(isph_ctx *ctx, const char *name, int v) {
  API_BEGIN(ctx) const std::string k(name ? name : "");
  if (k == "Num Blocks") c->sp.num_blocks = v; else if (k == "Block Size") c->sp.block_size = v; else if (k == "Maximum Iterations") c->sp.max_iters = v;
  else if (k == "Maximum Restarts") c->sp.max_restarts = v; else if (k == "Flexible Gmres") c->sp.flexible = v != 0;
  else if (k == "Output Frequency" || k == "Output Style" || k == "Verbosity" || k == "Num Recycled Blocks") { /* accepted, no effect */ }
  else ISPH_REQUIRE(false, "unknown integer solver parameter: " + k);
  API_END
}
int isph_solver_set_param_double(isph_ctx *ctx, const char *name, double v) {
  API_BEGIN(ctx) const std::string k(name ? name : "");
  if (k == "Convergence Tolerance") c->sp.tol = v; else ISPH_REQUIRE(false, "unknown double solver parameter: " + k);
  API_END
}
int isph_solver_set_param_str(isph_ctx *ctx, const char *name, const char *v) {
  API_BEGIN(ctx) const std::string k(name ? name : ""), s(v ? v : "");
  if (k == "Solver Type") { ISPH_REQUIRE(s == "Block GMRES" || s == "Block CG", "Solver Type: Block GMRES | Block CG"); c->sp.solver_type = s; }
  else if (k == "Orthogonalization") { ISPH_REQUIRE(s == "DGKS", "Orthogonalization: only DGKS (the reference's choice, solver_lin_belos.h:238)"); c->sp.ortho = s; }
  else ISPH_REQUIRE(false, "unknown string solver parameter: " + k);
  API_END
}
int isph_precond_set_param_int(isph_ctx *ctx, const char *name, int v) {
  API_BEGIN(ctx) const std::string k(name ? name : "");
  if (k == "Overlap Level") c->pp.overlap = v; else if (k == "fact: level-of-fill") c->pp.fill = v; else if (k == "relaxation: sweeps") c->pp.sweeps = v;
  else if (k == "chebyshev: degree") c->pp.cheb_degree = v; else if (k == "chebyshev: eigenvalue max iterations") c->pp.cheb_eig_iters = v;
  // ML's parameter list (precond_ml.h:44-58 and the keys PrecondWrapper_ML::setNullVector adds, :108-120)
  else if (k == "max levels") c->pp.ml_max_levels = v; else if (k == "smoother: sweeps") { ISPH_REQUIRE(v >= 0, "smoother: sweeps >= 0"); c->pp.ml_pre = c->pp.ml_post = v; }
  else if (k == "smoother: pre sweeps") c->pp.ml_pre = v; else if (k == "smoother: post sweeps") c->pp.ml_post = v; else if (k == "smoother: sweeps (coarse levels)") c->pp.ml_level_sweeps = v;
  else if (k == "coarse: sweeps") c->pp.ml_coarse_sweeps = v; else if (k == "coarse: max size") c->pp.ml_max_coarse = v; else if (k == "eigen-analysis: iterations") c->pp.ml_eig_iters = v;
  else if (k == "ML output" || k == "repartition: enable") { /* verbosity / Zoltan repartitioning: nothing to do on one GPU per rank */ }
  else if (k == "null space: dimension") ISPH_REQUIRE(v == 1, "null space: dimension must be 1");
  else ISPH_REQUIRE(false, "unknown integer preconditioner parameter: " + k);
  API_END
}
int isph_precond_set_param_double(isph_ctx *ctx, const char *name, double v) {
  API_BEGIN(ctx) const std::string k(name ? name : "");
  if (k == "relaxation: damping factor") c->pp.damping = v; else if (k == "relaxation: min diagonal value" || k == "chebyshev: min diagonal value") c->pp.min_diag = v;
  else if (k == "chebyshev: ratio eigenvalue") c->pp.cheb_ratio = v; else if (k == "chebyshev: max eigenvalue") c->pp.cheb_lmax = v;
  else if (k == "fact: drop tolerance" || k == "fact: relax value" || k == "fact: absolute threshold" || k == "fact: relative threshold") {
    // Ifpack_ILU defaults (drop 0, relax 0, absolute 0, relative 1): anything else changes the factors and is not implemented — refuse, do not ignore
    const double dflt = (k == "fact: relative threshold") ? 1.0 : 0.0;
    ISPH_REQUIRE(v == dflt, "preconditioner parameter '" + k + "' is only supported at its Ifpack default value");
  }
  else if (k == "aggregation: threshold") { ISPH_REQUIRE(v >= 0.0, "aggregation: threshold >= 0"); c->pp.ml_threshold = v; }
  else if (k == "aggregation: damping factor") c->pp.ml_agg_damping = v;          // only 0 (non-smoothed aggregation) is provided: checked at create
  else if (k == "smoother: Chebyshev alpha") c->pp.ml_alpha = v; else if (k == "coarse: Chebyshev alpha") c->pp.ml_coarse_alpha = v;
  else if (k == "smoother: Chebyshev alpha (coarse levels)") c->pp.ml_level_alpha = v; else if (k == "coarse correction scale (coarse levels)") c->pp.ml_level_scale = v;
  else if (k == "smoother: damping factor") c->pp.ml_damping = v; else if (k == "coarse correction scale") c->pp.ml_scale = v;
  else ISPH_REQUIRE(false, "unknown double preconditioner parameter: " + k);
  API_END
}
int isph_precond_set_param_str(isph_ctx *ctx, const char *name, const char *v) {
  API_BEGIN(ctx) const std::string k(name ? name : ""), s(v ? v : "");
  if (k == "Precond Type") c->pp.type = s; else if (k == "relaxation: type") c->pp.relax_type = s; else if (k == "schwarz: combine mode") { /* overlap 0: no combine */ }
  // pair_isph.cpp:325-329: "Precond Package" picks the wrapper class; ML = the multilevel stand-in (amg.cu), Ifpack = "Precond Type" as set
  else if (k == "Precond Package") { ISPH_REQUIRE(s == "ML" || s == "Ifpack", "Preconditioner is not in supported list: Ifpack, ML"); if (s == "ML") c->pp.type = "ML"; else if (c->pp.type == "ML") c->pp.type = "ILU"; }
  else if (k == "aggregation: type") { ISPH_REQUIRE(s == "Uncoupled" || s == "MIS" || s == "Uncoupled-MIS", "aggregation: type: Uncoupled | MIS | Uncoupled-MIS (aggregates are always formed inside a rank by a distance-2 independent set)"); c->pp.ml_agg_type = s; }
  else if (k == "smoother: type") c->pp.ml_smoother = s;                             // checked at create: Chebyshev | Jacobi
  else if (k == "coarse: type") c->pp.ml_coarse = s;
  else if (k == "smoother: pre or post") { ISPH_REQUIRE(s == "both" || s == "pre" || s == "post", "smoother: pre or post: both | pre | post"); if (s == "pre") c->pp.ml_post = 0; if (s == "post") c->pp.ml_pre = 0; }
  else if (k == "increasing or decreasing" || k == "null space: type" || k == "eigen-analysis: type") { /* level numbering / pre-computed null space / power method: the only behaviour here */ }
  else ISPH_REQUIRE(false, "unknown string preconditioner parameter: " + k);
  API_END
}
int isph_precond_set_blocks(isph_ctx *ctx, const int *blk) {
  API_BEGIN(ctx)
  if (!blk) c->have_blocks = false;
  else { ISPH_REQUIRE(c->A.built, "no matrix"); c->block_of_row.ensure(c->A.n); CUDA_CHECK(cudaMemcpyAsync(c->block_of_row.p, blk, sizeof(int) * c->A.n, cudaMemcpyHostToDevice, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream)); c->have_blocks = true; }
  API_END
}
int isph_precond_create(isph_ctx *ctx) { API_BEGIN(ctx) precond_create(c); CUDA_CHECK(cudaStreamSynchronize(c->stream)); API_END }
int isph_precond_free(isph_ctx *ctx) { API_BEGIN(ctx) precond_free(c); API_END }
int isph_precond_apply(isph_ctx *ctx, const double *r, double *z) {
  API_BEGIN(ctx) ISPH_REQUIRE(c->prec_ready && r && z, "preconditioner not created"); const int n = c->A.n, ld = c->ld; c->V.ensure((size_t)2 * ld);
  CUDA_CHECK(cudaMemcpyAsync(c->V.p, r, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream)); precond_apply(c, c->V.p, c->V.p + ld);
  CUDA_CHECK(cudaMemcpyAsync(z, c->V.p + ld, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream)); API_END
}
// ---- SolverLin block interface (solver_lin.h:43-56, solver_lin.cpp:78-138) and SolverLin_Belos::solveBlockProblem (solver_lin_belos.h:53-128) ----
// The dim x dim block operator (Thyra::PhysicallyBlockedLinearOp of Epetra matrices that share the nodal map) becomes ONE stacked SELL matrix
// with dim * n rows held by a child context; x and b (n x dim multivectors, one column per block row: createBlockVector, solver_lin.cpp:93-107)
// are the stacked vectors.  The preconditioner is the reference's: ONE operator built from the scalar matrix the wrapper holds, applied to
// every diagonal block (getBlockPrecondOperator).  One rank only (the reference's own block path is experimental, pair_isph.cpp:1750).
int isph_solver_create_block_matrix(isph_ctx *ctx, int dim, const char *name) {
  API_BEGIN(ctx) ISPH_REQUIRE(dim >= 1 && dim <= 3, "createBlockMatrix: dim must be 1..3"); ISPH_REQUIRE(c->nranks == 1, "createBlockMatrix: the block system is provided on one rank only");
  if (!c->blk) c->blk = new BlockSys();
  BlockSys &B = *c->blk; B.dim = dim; B.n = 0; B.name = name ? name : ""; B.filled = false;
  B.rp.assign((size_t)dim * dim, {}); B.ci.assign((size_t)dim * dim, {}); B.va.assign((size_t)dim * dim, {});
  API_END
}
int isph_solver_free_block_matrix(isph_ctx *ctx) {
  API_BEGIN(ctx) if (c->blk) { if (c->blk->child) isph_ctx_destroy(c->blk->child); delete c->blk; c->blk = nullptr; } API_END
}
int isph_solver_set_block_csr(isph_ctx *ctx, int i, int j, int n, const int *rowptr, const int *col, const double *val) {
  API_BEGIN(ctx) ISPH_REQUIRE(c->blk && c->blk->dim > 0, "setBlock: createBlockMatrix first");
  BlockSys &B = *c->blk;
  if (!rowptr || i < 0 || j < 0 || i >= B.dim || j >= B.dim) return ISPH_SUCCESS;          // solver_lin.cpp:133: silently ignored, like the reference
  ISPH_REQUIRE(n > 0 && (B.n == 0 || B.n == n), "setBlock: all blocks share the nodal map (same number of rows)");
  B.n = n; const size_t q = (size_t)i * B.dim + j;
  B.rp[q].assign(rowptr, rowptr + n + 1); B.ci[q].assign(col, col + rowptr[n]); B.va[q].assign(val, val + rowptr[n]); B.filled = false;
  API_END
}
int isph_solver_set_block_end(isph_ctx *ctx) {
  API_BEGIN(ctx) ISPH_REQUIRE(c->blk && c->blk->n > 0, "setBlockEnd: no block was set");
  BlockSys &B = *c->blk; const int n = B.n, d = B.dim; const long long N = (long long)n * d; ISPH_REQUIRE(N < (1ll << 31) - 64, "block system too large");
  std::vector<int> rp(N + 1, 0), ci; std::vector<double> va;
  for (int ib = 0; ib < d; ++ib) for (int r = 0; r < n; ++r) {
    for (int jb = 0; jb < d; ++jb) { const size_t q = (size_t)ib * d + jb; if (B.rp[q].empty()) continue;
      for (int e = B.rp[q][r]; e < B.rp[q][r + 1]; ++e) { ISPH_REQUIRE(B.ci[q][e] >= 0 && B.ci[q][e] < n, "setBlock: column outside the nodal map"); ci.push_back(jb * n + B.ci[q][e]); va.push_back(B.va[q][e]); } }
    rp[(size_t)ib * n + r + 1] = (int)ci.size();
  }
  if (!B.child) { ISPH_REQUIRE(isph_ctx_create(&B.child, c->device, 1, 0, nullptr) == ISPH_SUCCESS, "block system: no child context"); }
  Ctx *k = reinterpret_cast<Ctx *>(B.child);
  CUDA_CHECK(cudaStreamSynchronize(k->stream)); if (k->own_stream) { cudaStreamDestroy(k->stream); k->own_stream = false; } k->stream = c->stream;
  matrix_from_csr(k, (int)N, rp.data(), ci.data(), va.data()); B.filled = true;
  API_END
}
int isph_solver_solve_block(isph_ctx *ctx, int use_prec, const char *label) {
  API_BEGIN(ctx) ISPH_REQUIRE(c->blk && c->blk->filled, "solveBlockProblem: createBlockMatrix / setBlock / setBlockEnd first");
  BlockSys &B = *c->blk; Ctx *k = reinterpret_cast<Ctx *>(B.child); const int n = B.n, d = B.dim;
  k->stream = c->stream; k->own_stream = false;                    // the parent's stream may have been replaced since setBlockEnd (isph_set_stream)
  ISPH_REQUIRE(c->b_nvec == d && c->x_nvec == d && c->xs.p && c->bs.p, ">> SolverLin_Belos::solveBlockProblem, dimension of rhs does not match to the block matrix");   // solver_lin_belos.h:58-59
  ISPH_REQUIRE(!c->is_singular, ">> SolverLin_Belos::solveBlockProblem does not support singular problems");                                                    // :60-61
  ISPH_REQUIRE(!use_prec || (c->A.built && c->A.n == n), "solveBlockProblem: the preconditioner is built from the scalar matrix (prec->setMatrix): it must have the blocks' row map");
  const int ld = c->ld; const size_t xl = (size_t)ld * d;
  // initial solution and load vector exactly as solveProblem prepares them, then stacked
  if (c->init_type == ISPH_INIT_ZERO) CUDA_CHECK(cudaMemsetAsync(c->xs.p, 0, sizeof(double) * xl, c->stream));
  else if (c->init_type == ISPH_INIT_VALUE) { std::vector<double> f(xl, c->init_val); CUDA_CHECK(cudaMemcpyAsync(c->xs.p, f.data(), sizeof(double) * xl, cudaMemcpyHostToDevice, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream)); }
  else if (c->init_type == ISPH_INIT_RANDOM) ISPH_REQUIRE(false, "solveBlockProblem: setInitialSolution(Random) is not provided for the block system");
  else if (c->x_host) for (int q = 0; q < d; ++q) CUDA_CHECK(cudaMemcpyAsync(c->xs.p + (size_t)q * ld, c->x_host + (size_t)q * c->x_lda, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
  load_from_host(c); c->b_dev_fresh = false;
  k->sp = c->sp; k->xs.ensure(k->ld); k->bs.ensure(k->ld); k->x_nvec = k->b_nvec = 1; k->x_host = k->b_host = nullptr; k->x_owned = k->b_owned = true; k->init_type = -1; k->is_singular = false;
  for (int q = 0; q < d; ++q) { CUDA_CHECK(cudaMemcpyAsync(k->xs.p + (size_t)q * n, c->xs.p + (size_t)q * ld, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
                                CUDA_CHECK(cudaMemcpyAsync(k->bs.p + (size_t)q * n, c->bs.p + (size_t)q * ld, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream)); }
  k->prec_parent = use_prec ? c : nullptr; k->prec_dim = d; k->launches = 0;
  solver_solve(k, use_prec != 0, label);                          // Belos on the product space: the same GMRES / CG on the stacked vectors
  if (use_prec && c->prec_kind == 3) ISPH_REQUIRE(!ilu_fault(c), "ILU(0): a dependency wait timed out or a factor entry is not a number (zero pivot?)");
  for (int q = 0; q < d; ++q) CUDA_CHECK(cudaMemcpyAsync(c->xs.p + (size_t)q * ld, k->xs.p + (size_t)q * n, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
  if (c->x_host) for (int q = 0; q < d; ++q) CUDA_CHECK(cudaMemcpyAsync(c->x_host + (size_t)q * c->x_lda, c->xs.p + (size_t)q * ld, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  c->last_iters = k->last_iters; c->last_converged = k->last_converged; c->last_relres = k->last_relres; c->launches += k->launches; c->init_type = -1;
  API_END
}
int isph_solver_solve(isph_ctx *ctx, int use_prec, const char *label) { API_BEGIN(ctx) solver_solve(c, use_prec != 0, label); API_END }
int isph_solver_stats(isph_ctx *ctx, int *iters, double *relres, int *converged, double *lmax) {
  API_BEGIN(ctx) if (iters) *iters = c->last_iters; if (relres) *relres = c->last_relres; if (converged) *converged = c->last_converged; if (lmax) *lmax = c->last_lmax; API_END
}

int isph_halo_counts(isph_ctx *ctx, int *nhalo, int *nsend, int *npeers) {
  API_BEGIN(ctx) int a = 0, b = 0, p = 0; halo_counts(c, &a, &b, &p); if (nhalo) *nhalo = a; if (nsend) *nsend = b; if (npeers) *npeers = p; API_END
}
long long isph_solver_second_passes(isph_ctx *ctx) { return ctx ? reinterpret_cast<Ctx *>(ctx)->last_second_passes : -1; }

double isph_timer_ms(isph_ctx *ctx, const char *name) { if (!ctx || !name) return -1.0; Ctx *c = reinterpret_cast<Ctx *>(ctx); auto it = c->timers.find(name); if (it == c->timers.end()) return 0.0; timer_flush(it->second); return it->second.ms; }
int isph_timer_reset(isph_ctx *ctx) { API_BEGIN(ctx) for (auto &kv : c->timers) { timer_flush(kv.second); kv.second.ms = 0.0; } API_END }
long long isph_kernel_launches(isph_ctx *ctx) { return ctx ? reinterpret_cast<Ctx *>(ctx)->launches : -1; }

int isph_profile_spmv(isph_ctx *ctx, int enable) { API_BEGIN(ctx) c->prof_spmv = enable != 0; API_END }
int isph_profile_spmv_get(isph_ctx *ctx, double *total_ms, long long *launches) {
  API_BEGIN(ctx)
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  for (size_t q = 0; q + 1 < c->prof_used; q += 2) { float ms = 0.f; CUDA_CHECK(cudaEventElapsedTime(&ms, c->prof_ev[q], c->prof_ev[q + 1])); c->prof_ms += ms; ++c->prof_cnt; }
  c->prof_used = 0;
  if (total_ms) *total_ms = c->prof_ms; if (launches) *launches = c->prof_cnt;
  c->prof_ms = 0.0; c->prof_cnt = 0;
  API_END
}
int isph_profile_precond_get(isph_ctx *ctx, double *total_ms, long long *launches) {
  API_BEGIN(ctx)
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  for (size_t q = 0; q + 1 < c->pprof_used; q += 2) { float ms = 0.f; CUDA_CHECK(cudaEventElapsedTime(&ms, c->pprof_ev[q], c->pprof_ev[q + 1])); c->pprof_ms += ms; ++c->pprof_cnt; }
  c->pprof_used = 0;
  if (total_ms) *total_ms = c->pprof_ms; if (launches) *launches = c->pprof_cnt;
  c->pprof_ms = 0.0; c->pprof_cnt = 0;
  API_END
}
int isph_precond_ml_info(isph_ctx *ctx, int *levels, int *rows, long long *nnz, double *lambda_max, int cap) {
  API_BEGIN(ctx) const int nl = amg_info(c, rows, nnz, lambda_max, cap); if (levels) *levels = nl; API_END
}
int isph_precond_ml_aggregates(isph_ctx *ctx, int *agg) { API_BEGIN(ctx) amg_aggregates(c, agg); API_END }
double isph_precond_ml_setup_ms(isph_ctx *ctx, const char *phase) { if (!ctx) return 0.0; return amg_setup_ms((Ctx *)ctx, phase); }
int isph_precond_info(isph_ctx *ctx, long long *factor_nnz, int *levels_lower, int *levels_upper, int *max_row) {
  API_BEGIN(ctx) long long z = 0; int a = 0, b = 0, m = 0; ilu_info(c, &z, &a, &b, &m);
  if (factor_nnz) *factor_nnz = z; if (levels_lower) *levels_lower = a; if (levels_upper) *levels_upper = b; if (max_row) *max_row = m; API_END
}
int isph_bench_spmv(isph_ctx *ctx, int reps, double *avg_ms) {
  API_BEGIN(ctx) ISPH_REQUIRE(c->A.built && reps > 0 && avg_ms, "no matrix"); const int ld = c->ld; c->V.ensure((size_t)2 * ld);
  CUDA_CHECK(cudaMemsetAsync(c->V.p, 0, sizeof(double) * 2 * ld, c->stream));
  cudaEvent_t a, b; CUDA_CHECK(cudaEventCreate(&a)); CUDA_CHECK(cudaEventCreate(&b));
  for (int w = 0; w < 3; ++w) spmv(c, c->V.p, c->V.p + ld, 1, ld, ld);
  CUDA_CHECK(cudaEventRecord(a, c->stream));
  for (int r = 0; r < reps; ++r) spmv(c, c->V.p, c->V.p + ld, 1, ld, ld);
  CUDA_CHECK(cudaEventRecord(b, c->stream)); CUDA_CHECK(cudaEventSynchronize(b));
  float ms = 0.f; CUDA_CHECK(cudaEventElapsedTime(&ms, a, b)); *avg_ms = ms / reps; cudaEventDestroy(a); cudaEventDestroy(b);
  API_END
}

}  // extern "C"

// isph_nccl_unique_id lives in halo.cu (it needs NCCL)
