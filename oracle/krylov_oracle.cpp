// TEST INFRASTRUCTURE ONLY — CPU oracle ("port") for the Krylov + preconditioner half of the path.
//
// PARITY UNPINNED: the arithmetic restated here lives in Trilinos (Belos / Ifpack / Epetra), which is neither
// vendored in /root/reference nor pinned to a version (README:11-13 "git clone ... Trilinos.git").  What is
// restated is the published behaviour of those packages for the parameter set the reference hard-codes
// (solver_lin_belos.h:224-264, precond_ifpack.h:28-48), anchored on the reference's own call sites:
//   solver_lin_belos.h:130-222  (solveProblem: null-space projection, right preconditioning, solver choice)
//   solver_lin.h:130-140        (PoissonProjection::Apply  y = A x ; y -= (y.n) n)
//   solver_lin.cpp:59-77        (createNullVector: mask / ||mask||_2)
//   precond_ifpack.h:50-75      (Ifpack::Create(type, A, overlap) -> SetParameters -> Initialize -> Compute; ApplyInverse)
// Belos semantics restated: BlockGmresSolMgr/Block(F)GmresIter with block size 1 — right-preconditioned (flexible)
// Arnoldi, DGKS orthogonalisation (classical Gram-Schmidt, a 2nd pass when ||w||^2 drops below dep_tol=1/sqrt(2) of
// its pre-projection value), Givens least squares, implicit residual |g_{j+1}| scaled by the norm of the FIRST
// initial residual, test `<= tol`, restart every `Num Blocks`, stop at Maximum Iterations / Maximum Restarts.
// BlockCGSolMgr with block size 1 — standard PCG, test ||r||_2 / ||r_0||_2 before every iteration.
// Ifpack semantics restated: point relaxation (Jacobi, zero start), Ifpack_Chebyshev (ApplyInverse recurrence with
// alpha = lmax/ratio, beta = 1.1 lmax; lmax from a 10-step power method on D^-1 A), Ifpack_ILU with level-of-fill k on
// the local block (overlap 0: off-block columns dropped): pattern from Ifpack_IlukGraph's level rule
// level(i,j) = min_k level(i,k) + level(k,j) + 1 <= k, then row-wise IKJ, L unit-lower, D stored inverted, U unit-upper scaled.
// Deviation (documented in DESIGN.md): Epetra's Random() start vector of the power method is replaced by a
// deterministic per-row hash, identically here and in the CUDA path.
#ifdef _OPENMP
#include <omp.h>
#endif
#include <vector>
#include <cmath>
#include <cstring>
#include <cstdio>
#include <algorithm>
#include <map>
#include "krylov_oracle.h"
#include "amg_oracle.h"

namespace {

struct Csr { int n; const int *rp, *ci; const double *v; };

double dot(int n, const double *a, const double *b) {
  double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}
void axpy(int n, double a, const double *x, double *y) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) y[i] += a * x[i];
}
void spmv(const Csr &A, const double *x, double *y) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < A.n; ++i) { double s = 0.0; for (int p = A.rp[i]; p < A.rp[i + 1]; ++p) if (A.ci[p] >= 0) s += A.v[p] * x[A.ci[p]]; y[i] = s; }
}

double hash01(unsigned long long t, int salt) {   // same generator as implicit-sph_b200/lattice.py::_hash01 and the CUDA path
  unsigned long long z = t + 0x9E3779B97F4A7C15ULL * (unsigned long long)(salt + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; z = z ^ (z >> 31);
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

struct Precond {
  int type; const Csr *A; const orc_krylov_params *prm;
  std::vector<double> invdiag;
  double lmax;
  // ILU(0): factors stored on A's pattern restricted to the row's block
  std::vector<double> fv, dinv; std::vector<int> diagpos, frp, fci; const int *blk;
  std::vector<double> V, W;
  // "Overlap Level" 1: one extended local problem per block (Ifpack_OverlappingRowMatrix + Ifpack_AdditiveSchwarz, combine mode Add)
  struct Sub { std::vector<int> idx, rp, ci; std::vector<double> v, r, z; Csr A; orc_krylov_params prm; Precond *M; };
  std::vector<Sub> subs;
  amg_oracle::Hierarchy *amg = nullptr;
  ~Precond() { for (auto &s : subs) delete s.M; delete amg; }

  void setup_overlap(const Csr &A_, const orc_krylov_params *p, const int *block_of_row) {
    const int n = A_.n; int nb = 0; for (int i = 0; i < n; ++i) nb = std::max(nb, block_of_row[i] + 1);
    subs.resize(nb);
    std::vector<int> pos(n, -1);
    for (int b = 0; b < nb; ++b) {
      Sub &S = subs[b];
      for (int i = 0; i < n; ++i) if (block_of_row[i] == b) S.idx.push_back(i);                     // own rows, local order
      std::vector<int> ghost; std::vector<char> seen(n, 0);
      for (int i : S.idx) for (int q = A_.rp[i]; q < A_.rp[i + 1]; ++q) { const int c = A_.ci[q]; if (c >= 0 && block_of_row[c] != b && !seen[c]) { seen[c] = 1; ghost.push_back(c); } }
      std::sort(ghost.begin(), ghost.end(), [&](int a, int c) {                                       // behind the own rows, by (owner block, global id)
        if (block_of_row[a] != block_of_row[c]) return block_of_row[a] < block_of_row[c];
        const int ga = p->row_gid ? p->row_gid[a] : a, gc = p->row_gid ? p->row_gid[c] : c; return ga < gc; });
      S.idx.insert(S.idx.end(), ghost.begin(), ghost.end());
      for (size_t k = 0; k < S.idx.size(); ++k) pos[S.idx[k]] = (int)k;
      S.rp.assign(1, 0);
      for (int i : S.idx) {                                                                           // rows restricted to the extended set (Ifpack_LocalFilter)
        std::vector<std::pair<int, double>> row;
        for (int q = A_.rp[i]; q < A_.rp[i + 1]; ++q) { const int c = A_.ci[q]; if (c >= 0 && pos[c] >= 0) row.emplace_back(pos[c], A_.v[q]); }
        std::stable_sort(row.begin(), row.end(), [](const std::pair<int, double> &a, const std::pair<int, double> &c) { return a.first < c.first; });
        for (auto &e : row) { S.ci.push_back(e.first); S.v.push_back(e.second); }
        S.rp.push_back((int)S.ci.size());
      }
      for (int i : S.idx) pos[i] = -1;
      S.A = Csr{(int)S.idx.size(), S.rp.data(), S.ci.data(), S.v.data()};
      S.prm = *p; S.prm.overlap = 0; S.prm.row_gid = nullptr;
      S.M = new Precond(); S.M->setup(S.A, &S.prm, nullptr);
      S.r.assign(S.idx.size(), 0.0); S.z.assign(S.idx.size(), 0.0);
    }
  }

  void setup(const Csr &A_, const orc_krylov_params *p, const int *block_of_row) {
    A = &A_; prm = p; type = p->precond; blk = block_of_row; const int n = A_.n;
    if (type == ORC_PREC_ILU0 && p->overlap >= 1 && block_of_row) { setup_overlap(A_, p, block_of_row); return; }
    if (type == ORC_PREC_AMG) {                             // PrecondWrapper_ML::create, precond_ml.h:128-135 (stand-in: amg_oracle.h)
      amg_oracle::Params q; q.max_levels = p->amg_max_levels; q.theta = p->amg_threshold; q.smoother = p->amg_smoother; q.pre = p->amg_pre; q.post = p->amg_post;
      q.level_sweeps = p->amg_level_sweeps; q.coarse_sweeps = p->amg_coarse_sweeps; q.alpha = p->amg_alpha; q.coarse_alpha = p->amg_coarse_alpha;
      q.eig_iters = p->amg_eig_iters; q.max_coarse = p->amg_max_coarse; q.oc = p->amg_scale; q.damping = p->amg_damping; q.coarse_direct = p->amg_coarse_direct; q.level_alpha = p->amg_level_alpha; q.level_oc = p->amg_level_scale;
      amg = new amg_oracle::Hierarchy(); amg->setup(n, A_.rp, A_.ci, A_.v, p->row_gid, block_of_row, q); lmax = amg->L[0].lmax; return;
    }
    if (type == ORC_PREC_JACOBI || type == ORC_PREC_CHEBYSHEV) {
      invdiag.assign(n, 0.0);
      for (int i = 0; i < n; ++i) { double d = 0.0; for (int q = A_.rp[i]; q < A_.rp[i + 1]; ++q) if (A_.ci[q] == i) d += A_.v[q];
        if (fabs(d) < p->min_diag) d = p->min_diag; if (d != 0.0) invdiag[i] = 1.0 / d; }
      V.assign(n, 0.0);
    }
    if (type == ORC_PREC_CHEBYSHEV) {
      lmax = p->cheb_lambda_max;
      if (lmax <= 0.0) {                                  // Ifpack_Chebyshev::PowerMethod
        std::vector<double> x(n), y(n);
        for (int i = 0; i < n; ++i) x[i] = 2.0 * hash01((unsigned long long)(p->row_gid ? p->row_gid[i] : i + 1), 7) - 1.0;
        double nrm = sqrt(dot(n, x.data(), x.data()));
        for (int i = 0; i < n; ++i) x[i] *= 1.0 / nrm;
        for (int it = 0; it < p->cheb_eig_iters; ++it) {
          spmv(A_, x.data(), y.data());
          for (int i = 0; i < n; ++i) y[i] *= invdiag[i];
          const double top = dot(n, y.data(), x.data()), bot = dot(n, x.data(), x.data());
          lmax = top / bot;
          nrm = sqrt(dot(n, y.data(), y.data()));
          for (int i = 0; i < n; ++i) x[i] = y[i] * (1.0 / nrm);
        }
      }
      V.assign(n, 0.0); W.assign(n, 0.0);
    }
    if (type == ORC_PREC_ILU0) {                           // Ifpack_ILU::Compute, level-of-fill p->ilu_fill, relax 0, athresh 0, rthresh 1
      auto inblk = [&](int i, int c) { return c >= 0 && (!blk || blk[c] == blk[i]); };
      // pattern: Ifpack_IlukGraph::ConstructFilledGraph.  Level 0 = entries of A inside the row's block; a fill entry (i,j)
      // created through pivot k gets level(i,k) + level(k,j) + 1 and is kept when that is <= fill.
      const int fill = p->ilu_fill;
      frp.assign(1, 0); fci.clear(); fv.clear(); std::vector<int> flev;
      std::vector<int> ubeg(n, 0);                            // first strictly-upper entry of every finished row
      for (int i = 0; i < n; ++i) {
        std::map<int, std::pair<int, double>> row;            // col -> (level, value)
        for (int q = A_.rp[i]; q < A_.rp[i + 1]; ++q) { const int c = A_.ci[q]; if (!inblk(i, c)) continue; auto it = row.find(c); if (it == row.end()) row[c] = {0, A_.v[q]}; else it->second.second += A_.v[q]; }
        if (fill > 0) for (auto it = row.begin(); it != row.end() && it->first < i; ++it) {
          const int k = it->first, lk = it->second.first;
          for (int u = ubeg[k]; u < frp[k + 1]; ++u) { const int nl = lk + flev[u] + 1; if (nl > fill) continue;
            auto jt = row.find(fci[u]); if (jt == row.end()) row[fci[u]] = {nl, 0.0}; else if (nl < jt->second.first) jt->second.first = nl; }
        }
        ubeg[i] = (int)fci.size();
        for (auto &e : row) { if (e.first <= i) ++ubeg[i]; fci.push_back(e.first); flev.push_back(e.second.first); fv.push_back(e.second.second); }
        frp.push_back((int)fci.size());
      }
      dinv.assign(n, 0.0); diagpos.assign(n, -1);
      std::vector<int> colflag(n, -1);
      for (int i = 0; i < n; ++i) for (int q = frp[i]; q < frp[i + 1]; ++q) if (fci[q] == i) diagpos[i] = q;
      for (int i = 0; i < n; ++i) {                           // numeric phase on the filled pattern
        for (int q = frp[i]; q < frp[i + 1]; ++q) colflag[fci[q]] = q;
        for (int q = frp[i]; q < frp[i + 1]; ++q) {           // columns ascending => strictly-lower part first
          const int j = fci[q]; if (j >= i) continue;
          const double multiplier = fv[q];
          fv[q] *= dinv[j];
          for (int u = frp[j]; u < frp[j + 1]; ++u) { const int k = fci[u]; if (k <= j) continue;
            const int kk = colflag[k]; if (kk > -1) fv[kk] -= multiplier * fv[u]; }
        }
        double d = fv[diagpos[i]];
        d = 1.0 / d;
        dinv[i] = d;
        for (int q = frp[i]; q < frp[i + 1]; ++q) if (fci[q] > i) fv[q] *= d;
        for (int q = frp[i]; q < frp[i + 1]; ++q) colflag[fci[q]] = -1;
      }
    }
  }

  // z = M^-1 r   (Epetra_Operator::ApplyInverse of the Ifpack object, as wrapped by Belos::EpetraPrecOp)
  void apply(const double *r, double *z) {
    const int n = A->n;
    switch (type) {
    case ORC_PREC_NONE: memcpy(z, r, sizeof(double) * n); break;
    case ORC_PREC_AMG: amg->apply(r, z); break;
    case ORC_PREC_JACOBI:                                   // Ifpack_PointRelaxation, Jacobi, zero starting solution
      for (int i = 0; i < n; ++i) z[i] = 0.0;
      for (int s = 0; s < prm->jacobi_sweeps; ++s) {
        if (s == 0) { for (int i = 0; i < n; ++i) z[i] = prm->jacobi_damping * invdiag[i] * r[i]; }
        else { spmv(*A, z, V.data()); for (int i = 0; i < n; ++i) z[i] += prm->jacobi_damping * invdiag[i] * (r[i] - V[i]); }
      }
      break;
    case ORC_PREC_CHEBYSHEV: {                              // Ifpack_Chebyshev::ApplyInverse
      const double alpha = lmax / prm->cheb_ratio, beta = 1.1 * lmax, delta = 2.0 / (beta - alpha), theta = 0.5 * (beta + alpha), s1 = theta * delta;
      const double oneOverTheta = 1.0 / theta;
      for (int i = 0; i < n; ++i) { W[i] = invdiag[i] * r[i] * oneOverTheta; z[i] = W[i]; }
      double rhok = 1.0 / s1;
      for (int deg = 0; deg < prm->cheb_degree - 1; ++deg) {
        spmv(*A, z, V.data());
        const double rhokp1 = 1.0 / (2.0 * s1 - rhok), dtemp1 = rhokp1 * rhok, dtemp2 = 2.0 * rhokp1 * delta; rhok = rhokp1;
        for (int i = 0; i < n; ++i) { W[i] *= dtemp1; W[i] += dtemp2 * invdiag[i] * (r[i] - V[i]); z[i] += W[i]; }
      }
      break; }
    case ORC_PREC_ILU0: {                                   // Ifpack_ILU::ApplyInverse: L (unit) solve, D^-1 scale, U (unit) solve
      if (!subs.empty()) {                                   // additive Schwarz, combine mode Add: every block's extended solution is added
        for (int i = 0; i < n; ++i) z[i] = 0.0;
        for (auto &S : subs) {
          for (size_t k = 0; k < S.idx.size(); ++k) S.r[k] = r[S.idx[k]];
          S.M->apply(S.r.data(), S.z.data());
          for (size_t k = 0; k < S.idx.size(); ++k) z[S.idx[k]] += S.z[k];
        }
        break;
      }
      for (int i = 0; i < n; ++i) { double s = r[i]; for (int q = frp[i]; q < frp[i + 1]; ++q) { const int j = fci[q]; if (j < i) s -= fv[q] * z[j]; } z[i] = s; }
      for (int i = 0; i < n; ++i) z[i] *= dinv[i];
      for (int i = n - 1; i >= 0; --i) { double s = z[i]; for (int q = frp[i]; q < frp[i + 1]; ++q) { const int j = fci[q]; if (j > i) s -= fv[q] * z[j]; } z[i] = s; }
      break; }
    }
  }
};

struct Op {   // A, or PoissonProjection (I - n n^T) A   (solver_lin.h:130-140)
  const Csr *A; const double *nv;
  void apply(const double *x, double *y) const {
    spmv(*A, x, y);
    if (nv) { const double val = dot(A->n, y, nv); axpy(A->n, -val, nv, y); }
  }
};

}  // namespace

extern "C" {

// thread count of the OpenMP row loops of this library (bench.py's CPU arm: torchrun exports OMP_NUM_THREADS=1)
int orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n; return 1;
#endif
}

void orc_krylov_default_params(orc_krylov_params *p) {
  memset(p, 0, sizeof(*p));
  p->solver = ORC_SOLVER_GMRES; p->flexible = 1; p->num_blocks = 50; p->max_iters = 500; p->max_restarts = 15; p->tol = 1.0e-8;   // solver_lin_belos.h:231-240
  p->overlap = 0; p->precond = ORC_PREC_NONE; p->jacobi_sweeps = 1; p->jacobi_damping = 1.0; p->min_diag = 0.0;
  p->cheb_degree = 1; p->cheb_ratio = 30.0; p->cheb_lambda_max = -1.0; p->cheb_eig_iters = 10; p->row_gid = 0; p->ilu_fill = 0;
  amg_oracle::Params q;                                    // defaults of the multilevel stand-in (implicit-sph_b200/csrc/amg.cu)
  p->amg_max_levels = q.max_levels; p->amg_threshold = q.theta; p->amg_smoother = q.smoother; p->amg_pre = q.pre; p->amg_post = q.post; p->amg_level_sweeps = q.level_sweeps;
  p->amg_coarse_sweeps = q.coarse_sweeps; p->amg_alpha = q.alpha; p->amg_coarse_alpha = q.coarse_alpha; p->amg_eig_iters = q.eig_iters; p->amg_max_coarse = q.max_coarse;
  p->amg_scale = q.oc; p->amg_damping = q.damping; p->amg_coarse_direct = 0; p->amg_level_alpha = q.level_alpha; p->amg_level_scale = q.level_oc;
}

}  // extern "C"
namespace {
// the preconditioner of SolverLin_Belos::solveBlockProblem: ONE operator built from the scalar matrix, applied to every diagonal block
// (PrecondWrapper_Ifpack::getBlockPrecondOperator, precond_ifpack.h:77-81; PrecondWrapper_ML::create(dim), precond_ml.h:137-154)
struct BlockDiagPrecond { Precond *P; int dim, nb; double lmax = 0.0; void apply(const double *r, double *z) { for (int k = 0; k < dim; ++k) P->apply(r + (size_t)k * nb, z + (size_t)k * nb); } };
template <class MT> int krylov_impl(const Csr &A, const orc_krylov_params *prm, MT &M, const int *null_mask, int use_null, double *b, double *x,
                                    int *iters_out, double *relres_out, double *history, int history_cap) {
  const int n = A.n;
  std::vector<double> nvec;
  if (use_null) {                                           // solver_lin.cpp:59-77 ; solver_lin_belos.h:138-144
    nvec.assign(n, 1.0);
    if (null_mask) for (int i = 0; i < n; ++i) nvec[i] = null_mask[i];
    const double nrm = sqrt(dot(n, nvec.data(), nvec.data()));
    for (auto &v : nvec) v *= 1.0 / nrm;
    const double bn = dot(n, b, nvec.data()); axpy(n, -bn, nvec.data(), b);
  }
  Op op{&A, use_null ? nvec.data() : nullptr};
  int iters = 0, nhist = 0; bool converged = false; double scale = 0.0, res = 0.0;
  std::vector<double> r(n), w(n);

  if (prm->solver == ORC_SOLVER_CG) {                       // Belos CGIter (block size 1)
    std::vector<double> z(n), p(n), Ap(n);
    op.apply(x, w.data()); for (int i = 0; i < n; ++i) r[i] = b[i] - w[i];
    M.apply(r.data(), z.data()); p = z;
    double rHz = dot(n, r.data(), z.data());
    scale = sqrt(dot(n, r.data(), r.data())); res = scale;
    if (history && nhist < history_cap) history[nhist++] = res;
    while (true) {
      if (scale == 0.0 || res / scale <= prm->tol) { converged = true; break; }
      if (iters >= prm->max_iters) break;
      ++iters;
      op.apply(p.data(), Ap.data());
      const double pAp = dot(n, p.data(), Ap.data()), alpha = rHz / pAp;
      axpy(n, alpha, p.data(), x); axpy(n, -alpha, Ap.data(), r.data());
      res = sqrt(dot(n, r.data(), r.data()));
      if (history && nhist < history_cap) history[nhist++] = res;
      if (res / scale <= prm->tol) { converged = true; break; }
      if (iters >= prm->max_iters) break;
      M.apply(r.data(), z.data());
      const double rHz_old = rHz; rHz = dot(n, r.data(), z.data());
      const double beta = rHz / rHz_old;
      for (int i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
    }
  } else {                                                  // Belos Block(F)GmresIter, block size 1
    const int m = prm->num_blocks;
    std::vector<double> V((size_t)(m + 1) * n), Z(prm->flexible ? (size_t)m * n : (size_t)n);
    std::vector<double> H((size_t)(m + 1) * m, 0.0), cs(m), sn(m), g(m + 1), h(m + 2), h2(m + 2), y(m);
    int restarts = 0; bool first = true;
    while (true) {
      op.apply(x, w.data()); for (int i = 0; i < n; ++i) r[i] = b[i] - w[i];
      const double beta = sqrt(dot(n, r.data(), r.data()));
      if (first) { scale = beta; first = false; if (history && nhist < history_cap) history[nhist++] = beta; }
      res = beta;
      if (scale == 0.0 || res / scale <= prm->tol) { converged = true; break; }
      for (int i = 0; i < n; ++i) V[i] = r[i] / beta;
      std::fill(g.begin(), g.end(), 0.0); g[0] = beta;
      int j = 0; bool stop = false;
      for (; j < m; ++j) {
        double *vj = &V[(size_t)j * n], *zj = prm->flexible ? &Z[(size_t)j * n] : Z.data(), *vn = &V[(size_t)(j + 1) * n];
        M.apply(vj, zj); op.apply(zj, vn);
        // DGKS (DGKSOrthoManager::blkOrtho1 + normalisation)
        const double oldDot = dot(n, vn, vn);
        for (int k = 0; k <= j; ++k) h[k] = dot(n, &V[(size_t)k * n], vn);
        for (int k = 0; k <= j; ++k) axpy(n, -h[k], &V[(size_t)k * n], vn);
        double newDot = dot(n, vn, vn);
        if (newDot < 0.70710678118654752440 * oldDot) {
          for (int k = 0; k <= j; ++k) h2[k] = dot(n, &V[(size_t)k * n], vn);
          for (int k = 0; k <= j; ++k) axpy(n, -h2[k], &V[(size_t)k * n], vn);
          for (int k = 0; k <= j; ++k) h[k] += h2[k];
          newDot = dot(n, vn, vn);
        }
        const double hn = sqrt(newDot); h[j + 1] = hn;
        if (hn > 0.0) for (int i = 0; i < n; ++i) vn[i] *= 1.0 / hn;
        // Givens update of column j (BlockGmresIter::updateLSQR)
        for (int k = 0; k < j; ++k) { const double t = cs[k] * h[k] + sn[k] * h[k + 1]; h[k + 1] = -sn[k] * h[k] + cs[k] * h[k + 1]; h[k] = t; }
        { const double a = h[j], bb = h[j + 1], rr = hypot(a, bb); cs[j] = rr == 0.0 ? 1.0 : a / rr; sn[j] = rr == 0.0 ? 0.0 : bb / rr; h[j] = rr; h[j + 1] = 0.0;
          g[j + 1] = -sn[j] * g[j]; g[j] = cs[j] * g[j]; }
        for (int k = 0; k <= j; ++k) H[(size_t)k * m + j] = h[k];
        ++iters; res = fabs(g[j + 1]);
        if (history && nhist < history_cap) history[nhist++] = res;
        if (res / scale <= prm->tol) { converged = true; stop = true; ++j; break; }
        if (iters >= prm->max_iters) { stop = true; ++j; break; }
      }
      // y = H^-1 g ; x += Z y  (flexible)  or  x += M^-1 (V y)
      for (int k = j - 1; k >= 0; --k) { double s = g[k]; for (int l = k + 1; l < j; ++l) s -= H[(size_t)k * m + l] * y[l]; y[k] = s / H[(size_t)k * m + k]; }
      if (prm->flexible) { for (int k = 0; k < j; ++k) axpy(n, y[k], &Z[(size_t)k * n], x); }
      else { std::fill(w.begin(), w.end(), 0.0); for (int k = 0; k < j; ++k) axpy(n, y[k], &V[(size_t)k * n], w.data()); M.apply(w.data(), r.data()); axpy(n, 1.0, r.data(), x); }
      if (stop) break;
      if (restarts >= prm->max_restarts) break;
      ++restarts;
    }
  }
  if (use_null) { const double xn = dot(n, x, nvec.data()); axpy(n, -xn, nvec.data(), x); }   // solver_lin_belos.h:215-219
  if (iters_out) *iters_out = iters;
  if (relres_out) *relres_out = scale > 0.0 ? res / scale : 0.0;
  return converged ? 0 : 1;
}
}  // namespace
extern "C" {
int orc_krylov_solve(int n, const int *rowptr, const int *col, const double *val, const orc_krylov_params *prm,
                     const int *block_of_row, const int *null_mask, int use_null, double *b, double *x,
                     int *iters_out, double *relres_out, double *history, int history_cap) {
  Csr A{n, rowptr, col, val};
  Precond M; M.setup(A, prm, block_of_row);                 // prec->create(), solver_lin_belos.h:153
  return krylov_impl(A, prm, M, null_mask, use_null, b, x, iters_out, relres_out, history, history_cap);
}
// SolverLin_Belos::solveBlockProblem (solver_lin_belos.h:53-128): the dim x dim block operator as ONE stacked CSR matrix of dim * nb rows
// (block row ib = rows ib*nb .. ib*nb+nb-1, block column jb = columns jb*nb ..), the preconditioner built from the scalar nb x nb matrix
// (prec_*) and applied to every diagonal block; no null space (the reference refuses singular block problems, :60-61)
int orc_krylov_solve_block(int nb, int dim, const int *rowptr, const int *col, const double *val, const int *prec_rowptr, const int *prec_col, const double *prec_val,
                           const orc_krylov_params *prm, double *b, double *x, int *iters_out, double *relres_out) {
  Csr A{nb * dim, rowptr, col, val}, Ap{nb, prec_rowptr, prec_col, prec_val};
  Precond P; P.setup(Ap, prm, nullptr);
  BlockDiagPrecond M{&P, dim, nb};
  return krylov_impl(A, prm, M, nullptr, 0, b, x, iters_out, relres_out, nullptr, 0);
}

int orc_precond_apply(int n, const int *rowptr, const int *col, const double *val, const orc_krylov_params *prm,
                      const int *block_of_row, const double *r, double *z, double *lambda_max_out) {
  Csr A{n, rowptr, col, val}; Precond M; M.setup(A, prm, block_of_row); M.apply(r, z);
  if (lambda_max_out) *lambda_max_out = M.lmax;
  return 0;
}

int orc_amg_hierarchy(int n, const int *rowptr, const int *col, const double *val, const orc_krylov_params *prm, const int *block_of_row,
                      int *rows, long long *nnz, double *lmax, int *agg0, int cap_rows, long long cap_nnz, int *c_rowptr, int *c_col, double *c_val) {
  Csr A{n, rowptr, col, val}; orc_krylov_params p = *prm; p.precond = ORC_PREC_AMG; Precond M; M.setup(A, &p, block_of_row);
  const auto &L = M.amg->L;
  for (size_t l = 0; l < L.size(); ++l) { rows[l] = L[l].A.n; nnz[l] = (long long)L[l].A.ci.size(); lmax[l] = L[l].lmax; }
  if (agg0) for (int i = 0; i < n; ++i) agg0[i] = L[0].agg.empty() ? -1 : L[0].agg[i];
  if (L.size() > 1 && c_rowptr && L[1].A.n <= cap_rows && (long long)L[1].A.ci.size() <= cap_nnz) {
    std::copy(L[1].A.rp.begin(), L[1].A.rp.end(), c_rowptr); std::copy(L[1].A.ci.begin(), L[1].A.ci.end(), c_col); std::copy(L[1].A.v.begin(), L[1].A.v.end(), c_val);
  }
  return (int)L.size();
}

}  // extern "C"
