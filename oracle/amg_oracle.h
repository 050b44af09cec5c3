// TEST INFRASTRUCTURE ONLY — CPU oracle for the multilevel preconditioner that stands in for ML (precond_ml.h:17-172, the reference's
// default `Precond Package`, pair_isph.cpp:325-329,359-361).
//
// PARITY UNPINNED BY CONSTRUCTION: ML (Trilinos) is neither vendored nor pinned, and what is built here is NOT ML's default algorithm
// but the member of ML's own option space that is data-parallel:
//   "aggregation: type" = MIS (distance-2 maximal independent set, hashed priorities; aggregates never cross a rank = "Uncoupled"),
//   "aggregation: damping factor" = 0 (non-smoothed / plain aggregation, tentative prolongator of piecewise constants),
//   "aggregation: threshold" eps with ML's criterion a_ij^2 > eps^2 |a_ii a_jj|,
//   "smoother: type" = Chebyshev (or Jacobi) on every level, "coarse: type" = the same smoother with "coarse: sweeps" (what
//   PrecondWrapper_ML::setNullVector selects for singular problems, precond_ml.h:118-120), Galerkin coarse operators P^T A P,
//   V-cycle with a scaled coarse-grid correction.
// The reference's own default smoother (symmetric Gauss-Seidel, precond_ml.h:53) is sequential per rank and is not provided.
// This file is the sequential restatement of implicit-sph_b200/csrc/amg.cu: same decisions (strength, MIS rounds, joins, numbering),
// same summation orders up to the column numbering, so iteration counts agree and solutions agree to rounding.
#pragma once
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cmath>

namespace amg_oracle {

struct Params {
  int max_levels = 5, pre = 1, post = 1, level_sweeps = 3, coarse_sweeps = 8, eig_iters = 10, max_coarse = 128, smoother = 0;
  double theta = 0.02, alpha = 2.0, level_alpha = 10.0, coarse_alpha = 30.0, oc = 2.5, level_oc = 2.0, damping = 0.67;    // alpha / oc: finest level; level_*: the levels between the finest and the coarsest
  int coarse_direct = 0;      // "coarse: type" = Amesos-KLU: dense inverse of the coarsest operator (non-singular problems only)
};

struct CsrM { int n = 0; std::vector<int> rp, ci; std::vector<double> v; };

inline uint64_t mix64(uint64_t t, int salt) {
  uint64_t z = t + 0x9E3779B97F4A7C15ULL * (uint64_t)(salt + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; z = z ^ (z >> 31);
  return z;
}
inline double hash01(uint64_t t, int salt) { return (double)(mix64(t, salt) >> 11) * (1.0 / 9007199254740992.0); }
// MIS priority: hashed global id in the high word, the id itself in the low word (unique, independent of the local numbering)
inline uint64_t prio(int gid) { return ((mix64((uint64_t)gid, 3) >> 32) << 32) | (uint32_t)gid; }

struct Level {
  CsrM A; std::vector<int> gid, agg, root; int nc = 0;        // root[I] = the row that founded aggregate I (its MIS root, or the leftover node itself)
  std::vector<double> invdiag; double lmax = 0.0;
  std::vector<std::vector<int>> members;
  std::vector<double> x, b, w, t;
};

template <class SpmvF> double power_method(int n, const std::vector<int> &gid, const std::vector<double> &invdiag, int iters, SpmvF spmv) {
  std::vector<double> x(n), y(n); double nrm = 0.0, lmax = 0.0;
  for (int i = 0; i < n; ++i) { x[i] = 2.0 * hash01((uint64_t)gid[i], 7) - 1.0; nrm += x[i] * x[i]; }
  nrm = std::sqrt(nrm); for (int i = 0; i < n; ++i) x[i] *= 1.0 / nrm;
  for (int it = 0; it < iters; ++it) {
    spmv(x.data(), y.data());
    double top = 0, bot = 0, yy = 0;
    for (int i = 0; i < n; ++i) { y[i] *= invdiag[i]; top += y[i] * x[i]; bot += x[i] * x[i]; yy += y[i] * y[i]; }
    lmax = top / bot; const double s = 1.0 / std::sqrt(yy);
    for (int i = 0; i < n; ++i) x[i] = y[i] * s;
  }
  return lmax;
}

inline void spmv(const CsrM &A, const double *x, double *y) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < A.n; ++i) { double s = 0.0; for (int q = A.rp[i]; q < A.rp[i + 1]; ++q) if (A.ci[q] >= 0) s += A.v[q] * x[A.ci[q]]; y[i] = s; }
}

// aggregates of one level.  blk (level 0 only) = rank of every row: strong connections never cross it.
inline void aggregate(Level &L, double theta, const int *blk) {
  const CsrM &A = L.A; const int n = A.n;
  std::vector<double> d(n, 0.0);
  for (int i = 0; i < n; ++i) for (int q = A.rp[i]; q < A.rp[i + 1]; ++q) if (A.ci[q] == i) d[i] += A.v[q];
  std::vector<std::vector<int>> sg(n); std::vector<std::vector<float>> sw(n);
  for (int i = 0; i < n; ++i) for (int q = A.rp[i]; q < A.rp[i + 1]; ++q) {
    const int c = A.ci[q]; if (c < 0 || c == i || (blk && blk[c] != blk[i])) continue;
    const double a = A.v[q];
    if (a * a > theta * theta * std::fabs(d[i] * d[c])) { sg[i].push_back(c); sw[i].push_back((float)std::fabs(a)); }
  }
  const uint64_t MAXKEY = ~0ULL;
  std::vector<uint64_t> key(n), m1(n); std::vector<int> state(n);       // 0 undecided, 1 root, -1 out, -2 no strong connection ("Dirichlet")
  for (int i = 0; i < n; ++i) { key[i] = prio(L.gid[i]); state[i] = sg[i].empty() ? -2 : 0; }
  while (true) {
    long long und = 0; for (int i = 0; i < n; ++i) und += state[i] == 0;
    if (!und) break;
    auto T = [&](int j) -> uint64_t { return state[j] == 1 ? MAXKEY : (state[j] == 0 ? key[j] : 0ULL); };
    for (int i = 0; i < n; ++i) { uint64_t m = T(i); for (int j : sg[i]) m = std::max(m, T(j)); m1[i] = m; }
    std::vector<int> ns(state);
    for (int i = 0; i < n; ++i) if (state[i] == 0) {
      uint64_t m = m1[i]; for (int j : sg[i]) m = std::max(m, m1[j]);
      if (m == MAXKEY) ns[i] = -1; else if (m == key[i]) ns[i] = 1;
    }
    state.swap(ns);
  }
  std::vector<int> agg(n, -1); int nroot = 0; L.root.clear();
  for (int i = 0; i < n; ++i) if (state[i] == 1) { agg[i] = nroot++; L.root.push_back(i); }
  for (int pass = 0; pass < 3; ++pass) {                                    // join the aggregate of the strongest aggregated neighbour (snapshot per pass)
    std::vector<int> na(agg);
    for (int i = 0; i < n; ++i) if (agg[i] < 0 && state[i] != -2) {
      int best = -1; float bw = -1.0f; uint64_t bk = 0;
      for (size_t k = 0; k < sg[i].size(); ++k) { const int j = sg[i][k]; if (agg[j] < 0) continue;
        if (best < 0 || sw[i][k] > bw || (sw[i][k] == bw && key[j] > bk)) { best = j; bw = sw[i][k]; bk = key[j]; } }
      if (best >= 0) na[i] = agg[best];
    }
    agg.swap(na);
  }
  int nc = nroot;
  for (int i = 0; i < n; ++i) if (agg[i] < 0 && state[i] != -2) { agg[i] = nc++; L.root.push_back(i); }      // leftovers: singletons, numbered behind the roots
  L.agg = agg; L.nc = nc;
  L.members.assign(nc, {});
  for (int i = 0; i < n; ++i) if (agg[i] >= 0) L.members[agg[i]].push_back(i);
}

// Galerkin operator of plain aggregation: Ac[I,J] = sum_{i in I (ascending)} sum_{entries of row i with agg(col) = J (stored order)} a
inline void galerkin(const Level &L, Level &C) {
  const CsrM &A = L.A; C.A.n = L.nc; C.A.rp.assign(1, 0); C.A.ci.clear(); C.A.v.clear(); C.gid.assign(L.nc, 0);
  std::vector<int> pos(L.nc, -1);
  for (int I = 0; I < L.nc; ++I) C.gid[I] = L.gid[L.root[I]];                 // a coarse node inherits the global id of its root: priorities stay independent of the numbering
  for (int I = 0; I < L.nc; ++I) {
    std::vector<int> cols; std::vector<double> sums;
    for (int i : L.members[I]) {
      // per fine row first (the device compresses every row to its aggregate columns before the rows of an aggregate are merged)
      std::vector<int> rj; std::vector<double> rs;
      for (int q = A.rp[i]; q < A.rp[i + 1]; ++q) { const int c = A.ci[q]; if (c < 0) continue; const int J = L.agg[c]; if (J < 0) continue;
        size_t k = 0; for (; k < rj.size(); ++k) if (rj[k] == J) break;
        if (k == rj.size()) { rj.push_back(J); rs.push_back(A.v[q]); } else rs[k] += A.v[q]; }
      for (size_t k = 0; k < rj.size(); ++k) { if (pos[rj[k]] < 0) { pos[rj[k]] = (int)cols.size(); cols.push_back(rj[k]); sums.push_back(rs[k]); } else sums[pos[rj[k]]] += rs[k]; }
    }
    std::vector<int> ord(cols.size()); for (size_t k = 0; k < ord.size(); ++k) ord[k] = (int)k;
    std::sort(ord.begin(), ord.end(), [&](int a, int b) { return cols[a] < cols[b]; });
    for (int k : ord) { C.A.ci.push_back(cols[k]); C.A.v.push_back(sums[k]); }
    for (int cidx : cols) pos[cidx] = -1;
    C.A.rp.push_back((int)C.A.ci.size());
  }
}

// dense inverse by Gauss-Jordan elimination with partial pivoting (row-major n x n); the SAME routine runs on the host side of csrc/amg.cu.
// returns false on a zero pivot (singular coarsest operator)
inline bool dense_inverse(int n, std::vector<double> &a, std::vector<double> &inv) {
  inv.assign((size_t)n * n, 0.0); for (int i = 0; i < n; ++i) inv[(size_t)i * n + i] = 1.0;
  for (int c = 0; c < n; ++c) {
    int piv = c; double best = std::fabs(a[(size_t)c * n + c]);
    for (int r = c + 1; r < n; ++r) { const double v = std::fabs(a[(size_t)r * n + c]); if (v > best) { best = v; piv = r; } }
    if (best == 0.0) return false;
    if (piv != c) for (int k = 0; k < n; ++k) { std::swap(a[(size_t)c * n + k], a[(size_t)piv * n + k]); std::swap(inv[(size_t)c * n + k], inv[(size_t)piv * n + k]); }
    const double d = 1.0 / a[(size_t)c * n + c];
    for (int k = 0; k < n; ++k) { a[(size_t)c * n + k] *= d; inv[(size_t)c * n + k] *= d; }
    for (int r = 0; r < n; ++r) if (r != c) { const double f = a[(size_t)r * n + c]; if (f == 0.0) continue;
      for (int k = 0; k < n; ++k) { a[(size_t)r * n + k] -= f * a[(size_t)c * n + k]; inv[(size_t)r * n + k] -= f * inv[(size_t)c * n + k]; } }
  }
  return true;
}

struct Hierarchy {
  Params P; std::vector<Level> L; std::vector<double> coarse_inv; bool direct_ok = false;

  void smoother_setup(Level &Lv) {
    const CsrM &A = Lv.A; const int n = A.n;
    Lv.invdiag.assign(n, 0.0);
    for (int i = 0; i < n; ++i) { double d = 0.0; for (int q = A.rp[i]; q < A.rp[i + 1]; ++q) if (A.ci[q] == i) d += A.v[q]; if (d != 0.0) Lv.invdiag[i] = 1.0 / d; }
    Lv.lmax = n > 0 ? power_method(n, Lv.gid, Lv.invdiag, P.eig_iters, [&](const double *x, double *y) { spmv(A, x, y); }) : 1.0;
    Lv.x.assign(n, 0.0); Lv.b.assign(n, 0.0); Lv.w.assign(n, 0.0); Lv.t.assign(n, 0.0);
  }

  void setup(int n, const int *rp, const int *ci, const double *v, const int *gid, const int *blk, const Params &prm) {
    P = prm; L.clear(); L.emplace_back();
    Level &F = L[0]; F.A.n = n; F.A.rp.assign(rp, rp + n + 1); F.A.ci.assign(ci, ci + rp[n]); F.A.v.assign(v, v + rp[n]);
    F.gid.resize(n); for (int i = 0; i < n; ++i) F.gid[i] = gid ? gid[i] : i + 1;
    while (true) {
      smoother_setup(L.back());
      const int nl = L.back().A.n;
      if ((int)L.size() >= P.max_levels || nl <= P.max_coarse) break;
      aggregate(L.back(), P.theta, L.size() == 1 ? blk : nullptr);
      const int nc = L.back().nc;
      if (nc == 0 || nc >= (long long)nl * 9 / 10) { L.back().agg.clear(); L.back().nc = 0; break; }       // no coarsening left: this level is the coarsest
      Level C; galerkin(L.back(), C);
      L.push_back(std::move(C));
    }
    if (P.coarse_direct) {                                                        // Amesos-KLU stand-in: explicit inverse of the (small) coarsest operator
      const CsrM &A = L.back().A; const int nc = A.n; std::vector<double> a((size_t)nc * nc, 0.0);
      for (int i = 0; i < nc; ++i) for (int q = A.rp[i]; q < A.rp[i + 1]; ++q) if (A.ci[q] >= 0) a[(size_t)i * nc + A.ci[q]] += A.v[q];
      direct_ok = dense_inverse(nc, a, coarse_inv);
    }
  }

  // Ifpack_Chebyshev recurrence on D^-1 A; zero = start from x = 0 (no product needed for the first term)
  void cheb(Level &Lv, const double *r, double *x, bool zero, int degree, double ratio) {
    const int n = Lv.A.n; if (degree <= 0) { if (zero) for (int i = 0; i < n; ++i) x[i] = 0.0; return; }
    if (P.smoother == 1) {                                                    // damped Jacobi sweeps
      for (int s = 0; s < degree; ++s) {
        if (s == 0 && zero) { for (int i = 0; i < n; ++i) x[i] = P.damping * Lv.invdiag[i] * r[i]; continue; }
        spmv(Lv.A, x, Lv.t.data()); for (int i = 0; i < n; ++i) x[i] += P.damping * Lv.invdiag[i] * (r[i] - Lv.t[i]);
      }
      return;
    }
    const double lmax = Lv.lmax, alpha = lmax / ratio, beta = 1.1 * lmax, delta = 2.0 / (beta - alpha), theta = 0.5 * (beta + alpha), s1 = theta * delta;
    const double oneOverTheta = 1.0 / theta; double *W = Lv.w.data(), *V = Lv.t.data();
    if (zero) { for (int i = 0; i < n; ++i) { W[i] = Lv.invdiag[i] * r[i] * oneOverTheta; x[i] = W[i]; } }
    else { spmv(Lv.A, x, V); for (int i = 0; i < n; ++i) { W[i] = Lv.invdiag[i] * (r[i] - V[i]) * oneOverTheta; x[i] += W[i]; } }
    double rhok = 1.0 / s1;
    for (int deg = 0; deg < degree - 1; ++deg) {
      spmv(Lv.A, x, V);
      const double rhokp1 = 1.0 / (2.0 * s1 - rhok), dtemp1 = rhokp1 * rhok, dtemp2 = 2.0 * rhokp1 * delta; rhok = rhokp1;
      for (int i = 0; i < n; ++i) { W[i] *= dtemp1; W[i] += dtemp2 * Lv.invdiag[i] * (r[i] - V[i]); x[i] += W[i]; }
    }
  }

  void vcycle(size_t l, const double *r, double *x) {
    Level &Lv = L[l]; const int n = Lv.A.n;
    if (l + 1 == L.size()) {
      if (P.coarse_direct && direct_ok) { for (int i = 0; i < n; ++i) { double s = 0.0; for (int k = 0; k < n; ++k) s += coarse_inv[(size_t)i * n + k] * r[k]; x[i] = s; } return; }
      cheb(Lv, r, x, true, P.coarse_sweeps, P.coarse_alpha); return;
    }
    const int pre = l == 0 ? P.pre : P.level_sweeps, post = l == 0 ? P.post : P.level_sweeps;
    Level &C = L[l + 1];
    const double ratio = l == 0 ? P.alpha : P.level_alpha, oc = l == 0 ? P.oc : P.level_oc;
    cheb(Lv, r, x, true, pre, ratio);
    if (pre > 0) spmv(Lv.A, x, Lv.t.data()); else for (int i = 0; i < n; ++i) Lv.t[i] = 0.0;
    for (int I = 0; I < Lv.nc; ++I) { double s = 0.0; for (int i : Lv.members[I]) s += r[i] - Lv.t[i]; C.b[I] = s; }
    vcycle(l + 1, C.b.data(), C.x.data());
    for (int i = 0; i < n; ++i) if (Lv.agg[i] >= 0) x[i] += oc * C.x[Lv.agg[i]];
    cheb(Lv, r, x, false, post, ratio);
  }
  void apply(const double *r, double *z) { vcycle(0, r, z); }
};

}  // namespace amg_oracle
