// TEST INFRASTRUCTURE ONLY — CPU oracle ("port"): a plain C++ restatement of the reference's
// graph / pre-computation / operator-assembly algorithms for the linear-solve hot path.
// Every function cites the reference file:line it follows (paths relative to /root/reference/IMPLICIT-SPH).
// PARITY PINNED (this half of the oracle): (1) bit-identical to oracle/_ref — the reference's own functor headers compiled
// here — on every case of tests/test_oracle_cpu.py and on the fixtures under tests/golden/ that were generated from _ref;
// (2) against the reference's own recorded outputs (sph-script/conv-poisson-boltzmann-harmonic-2d-rev390.txt): total volume
// to 13 digits and, end to end through volumes / corrections / Poisson-Boltzmann residual / Jacobian / Newton iteration, the
// recorded err.psi.norm2 at N = 16, 32, 64, 128 to <= 1e-12 relative (observed 3e-15 .. 7e-15), the recorded gradient errors, and —
// for solid walls, normals, the Morris-Holmes mirror and the linearized equation — the six recorded errors and the volumes of
// sph-script/conv-channel-edl-potential-2d-morrisholmes-rev722.txt (2e-15 .. 2e-12).  (The Krylov half, krylov_oracle.cpp, restates un-vendored
// Trilinos code and stays "parity unpinned".)
// Nothing in the product path (implicit-sph_b200/) may call into this file.
//
// Arithmetic notes: built with -ffp-contract=off; expressions keep the reference's operation order because
// (a) the `rsq < cutsq` test is decided by rounding on lattices and (b) values are compared at 1e-12.
// Rows are independent, so the row loops carry `#pragma omp parallel for` (the reference's PairFor is a
// serial loop, pair_for.h:8-14; one MPI rank per core is how it uses a node) — results do not depend on
// the thread count because every row is computed by exactly one thread with a fixed neighbour order.
#include <vector>
#include <algorithm>
#include <unordered_map>
#include <cmath>
#include <cstring>
#include <cstdio>
#include <string>
#include "oracle_api.h"

#define NEIGHMASK 0x3FFFFFFF              /* LAMMPS neigh_list.h; functor_graph.h:72 */
static const double EPS_R = 1.0e-24;      /* ISPH_EPSILON, macrodef.h:6 */

namespace {

struct KernelFn {                          // kernel.h:9-24 ; kernel_{wendland,cubic,quintic}.h
  int id, dim;
  double C(double h) const {
    switch (id) {
    case ORC_KERNEL_CUBIC:   return dim == 3 ? 1.0 / (pow(h, 3) * M_PI) : 10.0 / (pow(h, 2) * 7.0 * M_PI);            // kernel_cubic.h:36-39
    case ORC_KERNEL_QUINTIC: return dim == 3 ? 14.0 / (pow(h, 3) * 1745.0 * M_PI) : 7.0 / (pow(h, 2) * 478.0 * M_PI); // kernel_quintic.h:36-39
    default:                 return dim == 3 ? 21.0 / (16 * M_PI * pow(h, 3)) : 7.0 / (4 * M_PI * pow(h, 2));         // kernel_wendland.h:36-39
    }
  }
  double val(double r, double h) const {
    const double s = fabs(r / h); double v = 0.0;
    switch (id) {
    case ORC_KERNEL_CUBIC:                                                  // kernel_cubic.h:43-55
      switch ((int)floor(s)) { case 0: v = 1.0 - 0.75 * (2 - s) * s * s; break; case 1: v = 0.25 * pow(2.0 - s, 3); }
      break;
    case ORC_KERNEL_QUINTIC:                                                // kernel_quintic.h:43-65 (fall-through is intended)
      switch ((int)floor(s)) { case 0: v += (15.0 * pow(1.0 - s, 5)); case 1: v -= (6.0 * pow(2.0 - s, 5)); case 2: v += (pow(3.0 - s, 5)); }
      break;
    default: v = pow(1 - 0.5 * s, 4) * (2 * s + 1.) * (s < 2);              // kernel_wendland.h:44-52
    }
    return v * C(h);
  }
  double dval(double r, double h) const {
    const double s = fabs(r / h); double v = 0.0;
    switch (id) {
    case ORC_KERNEL_CUBIC:                                                  // kernel_cubic.h:58-69
      switch ((int)floor(s)) { case 0: v = (2.25 * s - 3) * s; break; case 1: v = -0.75 * pow(2 - s, 2); }
      break;
    case ORC_KERNEL_QUINTIC:                                                // kernel_quintic.h:68-80
      switch ((int)floor(s)) { case 0: v -= (75.0 * pow((1 - s), 4)); case 1: v += (30.0 * pow((2 - s), 4)); case 2: v -= (5 * pow((3 - s), 4)); }
      break;
    default: v = -5.0 * s * pow(1 - 0.5 * s, 3) * (s < 2);                  // kernel_wendland.h:55-63
    }
    return v * (C(h) / h);
  }
};

inline double sph_op(bool anti, double fi, double fj) { return anti ? (fi + fj) : (fj - fi); }   // functor.h:9-20
// FilterBinary, filter.h:49-55 ; with ORC_FILTER_MATCH in m0: FilterMatchBinary, filter.h:101-107 (i == kind, j & mask)
inline bool fyes1(int m0, int ik) { return (m0 & ORC_FILTER_MATCH) ? ik == (m0 & 0xff) : (ik & m0) != 0; }
inline bool fyes2(int m0, int m1, int ik, int jk) { return fyes1(m0, ik) && (jk & m1); }

}  // namespace

struct orc_problem {
  int dim, dimL, nlocal, nghost, nall, ntypes, inum;
  std::vector<double> x, f[ORC_F_COUNT];
  std::vector<int> type, tag, ilist, neigh, kind_of_type, fixed_of_type, owner_of_ghost;
  std::vector<long long> noff;
  std::vector<double> cutsq, h;         // (ntypes+1)^2
  KernelFn kern; double morris_safe;
  double Gi[9], Li[6];
  std::unordered_map<int, int> lid;     // tag -> owned atom index (the Epetra nodal map, pair_isph.cpp:1258-1259)
  // CSR in canonical form: rows = owned atoms, columns = ascending global tag, duplicates merged
  std::vector<int> rowptr, col, collid;
  std::vector<double> val, diagonal, sld;
  int is_filled; bool have_graph;

  int kind(int t) const { return kind_of_type[t]; }
  int row_of(int i) const { return lid.find(tag[i])->second; }
  double cut2(int ti, int tj) const { return cutsq[(size_t)ti * (ntypes + 1) + tj]; }
  double hh(int ti, int tj) const { return h[(size_t)ti * (ntypes + 1) + tj]; }
  const double *X(int i) const { return &x[(size_t)3 * i]; }
  int find(int row, int gcol) const {
    const int *b = col.data() + rowptr[row], *e = col.data() + rowptr[row + 1];
    const int *it = std::lower_bound(b, e, gcol);
    return (it != e && *it == gcol) ? (int)(it - col.data()) : -1;
  }
  // Morris-Holmes mirror coefficient, mirror_morris_holmes.h:39-53
  double mirror(bool mh, int i, int j, double r) const {
    if (!mh) return 1.0;                                       // MirrorNothing, mirror.h:17-19
    const double xi_i = f[ORC_F_PND][i] * f[ORC_F_VFRAC][i], xi_j = f[ORC_F_PND][j] * f[ORC_F_VFRAC][j];
    const double d_i = 2.0 * r * (xi_i - 0.5) + EPS_R, d_j = 2.0 * r * (xi_j - 0.5) + EPS_R;
    return (1.0 + d_j / std::max(d_i, morris_safe * hh(type[i], type[j])));
  }
  // owner -> ghost copy (stand-in for comm->forward_comm_pair, pair_isph.cpp:1924-2074)
  void forward(int field) {
    const int nc = orc_field_ncomp(field); double *a = f[field].data();
    for (int g = 0; g < nghost; ++g) { const int o = owner_of_ghost[g]; if (o >= 0) memcpy(a + (size_t)(nlocal + g) * nc, a + (size_t)o * nc, sizeof(double) * nc); }
  }
};

typedef orc_problem Q;

extern "C" {

const char *orc_name(void) { return "C++ restatement (oracle port)"; }

int orc_field_ncomp(int fl) {
  static const int nc[ORC_F_COUNT] = {1, 9, 6, 3, 1, 1, 1, 1, 3, 3, 3, 1, 1, 1, 1, 1, 1};
  return (fl >= 0 && fl < ORC_F_COUNT) ? nc[fl] : -1;
}

orc_problem *orc_create(int dim, int nlocal, int nghost, const double *x, const int *type, const int *tag,
                        int inum, const int *ilist, const long long *noff, const int *neigh,
                        int ntypes, const int *kind_of_type, double h_one, double h_min, double cut_over_h,
                        int kernel_id, double morris_safe) {
  Q *q = new Q();
  q->dim = dim; q->dimL = dim * (dim + 1) / 2; q->nlocal = nlocal; q->nghost = nghost; q->nall = nlocal + nghost; q->ntypes = ntypes; q->inum = inum;
  const int nall = q->nall;
  q->x.assign(x, x + (size_t)3 * nall); q->type.assign(type, type + nall); q->tag.assign(tag, tag + nall);
  q->ilist.assign(ilist, ilist + inum); q->noff.assign(noff, noff + inum + 1); q->neigh.assign(neigh, neigh + noff[inum]);
  for (int fl = 0; fl < ORC_F_COUNT; ++fl) q->f[fl].assign((size_t)nall * orc_field_ncomp(fl), 0.0);
  std::fill(q->f[ORC_F_DENSITY].begin(), q->f[ORC_F_DENSITY].end(), 1.0);
  std::fill(q->f[ORC_F_EPS].begin(), q->f[ORC_F_EPS].end(), 1.0);
  q->kind_of_type.assign(kind_of_type, kind_of_type + ntypes + 1);
  q->kern.id = kernel_id; q->kern.dim = dim; q->morris_safe = morris_safe;
  // pair_isph_corrected.cpp:1302-1337 (coeff): one cutoff; h = h_one for equal kinds, h_min otherwise
  const double cut_one = h_one * cut_over_h, cut_one_sq = cut_one * cut_one;
  q->cutsq.assign((size_t)(ntypes + 1) * (ntypes + 1), 0.0); q->h.assign((size_t)(ntypes + 1) * (ntypes + 1), 0.0);
  for (int i = 1; i <= ntypes; ++i) for (int j = 1; j <= ntypes; ++j) {
    q->cutsq[(size_t)i * (ntypes + 1) + j] = cut_one_sq;
    q->h[(size_t)i * (ntypes + 1) + j] = (q->kind(i) == q->kind(j)) ? h_one : h_min;
  }
  // identity correction operators, pair_isph_corrected.cpp:342-346,363-366
  memset(q->Gi, 0, sizeof(q->Gi)); memset(q->Li, 0, sizeof(q->Li));
  for (int k2 = 0; k2 < dim; ++k2) for (int k1 = 0; k1 < dim; ++k1) q->Gi[k2 * dim + k1] = (k1 == k2);
  for (int k2 = 0, op = 0; k2 < dim; ++k2) for (int k1 = 0; k1 < (k2 + 1); ++k1, ++op) q->Li[op] = (k1 == k2);
  for (int i = 0; i < nlocal; ++i) q->lid[tag[i]] = i;
  q->owner_of_ghost.assign(nghost, -1);
  for (int g = 0; g < nghost; ++g) { auto it = q->lid.find(tag[nlocal + g]); if (it != q->lid.end()) q->owner_of_ghost[g] = it->second; }
  q->is_filled = 0; q->have_graph = false;
  return q;
}
void orc_destroy(orc_problem *q) { delete q; }
int orc_set_field(orc_problem *q, int fl, const double *d) { if (orc_field_ncomp(fl) < 0) return -1; memcpy(q->f[fl].data(), d, sizeof(double) * q->f[fl].size()); return 0; }
int orc_get_field(orc_problem *q, int fl, double *d) { if (orc_field_ncomp(fl) < 0) return -1; memcpy(d, q->f[fl].data(), sizeof(double) * q->f[fl].size()); return 0; }

// ---- pre-computation ---------------------------------------------------------------------------------------

// functor_volume.h:42-81: vfrac_i = 1 / (W(0,h_ii) + sum_{rsq<cutsq} W(sqrt(rsq), h_ij)), then owner->ghost copy
int orc_compute_volumes(orc_problem *q) {
  double *vfrac = q->f[ORC_F_VFRAC].data(); const int dim = q->dim;
#pragma omp parallel for schedule(static)
  for (int ii = 0; ii < q->inum; ++ii) {
    const int i = q->ilist[ii], itype = q->type[i];
    double wtmp = q->kern.val(0.0, q->hh(itype, itype));
    for (long long p = q->noff[ii]; p < q->noff[ii + 1]; ++p) {
      const int j = q->neigh[p] & NEIGHMASK, jtype = q->type[j];
      double rsq = 0.0;
      for (int k = 0; k < dim; ++k) { const double r = q->X(i)[k] - q->X(j)[k]; rsq += (r * r); }
      if (rsq < q->cut2(itype, jtype)) wtmp += q->kern.val(sqrt(rsq), q->hh(itype, jtype));
    }
    vfrac[i] = 1.0 / wtmp;
  }
  q->forward(ORC_F_VFRAC);
  return 0;
}

// closed-form inverses, utils_reference.cpp:251-313 (invertDenseMatrix dim 1..3) ; A,B column-major dim x dim
static void invert_small(int dim, const double *A, double *B) {
#define A_(i, j) A[(j) * dim + (i)]
#define B_(i, j) B[(j) * dim + (i)]
  if (dim == 1) { B[0] = 1.0 / A[0]; return; }
  if (dim == 2) {
    const double val = A_(0, 0) * A_(1, 1) - A_(0, 1) * A_(1, 0);       // computeDetDenseMatrix, utils_reference.cpp (2x2)
    B_(0, 0) = A_(1, 1) / val; B_(1, 1) = A_(0, 0) / val; B_(1, 0) = -A_(1, 0) / val; B_(0, 1) = -A_(0, 1) / val; return;
  }
  const double val = (A_(0, 0) * A_(1, 1) * A_(2, 2) + A_(1, 0) * A_(2, 1) * A_(0, 2) + A_(2, 0) * A_(0, 1) * A_(1, 2)      // computeDetDenseMatrix, utils_reference.cpp:160-165
                      - A_(2, 0) * A_(1, 1) * A_(0, 2) - A_(0, 0) * A_(2, 1) * A_(1, 2) - A_(1, 0) * A_(0, 1) * A_(2, 2));
  double v0, v1, v2;
  v0 = A_(1, 1) * A_(2, 2) - A_(2, 1) * A_(1, 2); v1 = -A_(1, 0) * A_(2, 2) + A_(2, 0) * A_(1, 2); v2 = A_(1, 0) * A_(2, 1) - A_(2, 0) * A_(1, 1);
  B_(0, 0) = v0 / val; B_(1, 0) = v1 / val; B_(2, 0) = v2 / val;
  v0 = A_(2, 1) * A_(0, 2) - A_(0, 1) * A_(2, 2); v1 = A_(0, 0) * A_(2, 2) - A_(2, 0) * A_(0, 2); v2 = -A_(0, 0) * A_(2, 1) + A_(2, 0) * A_(0, 1);
  B_(0, 1) = v0 / val; B_(1, 1) = v1 / val; B_(2, 1) = v2 / val;
  v0 = A_(0, 1) * A_(1, 2) - A_(1, 1) * A_(0, 2); v1 = -A_(0, 0) * A_(1, 2) + A_(1, 0) * A_(0, 2); v2 = A_(0, 0) * A_(1, 1) - A_(1, 0) * A_(0, 1);
  B_(0, 2) = v0 / val; B_(1, 2) = v1 / val; B_(2, 2) = v2 / val;
#undef A_
#undef B_
}

// functor_gradient_correction.h:24-71: Gc_i = ( - sum r_ij r_ij^T dW/dr / r V_j )^{-1}   (no filter is ever set)
int orc_compute_gradient_correction(orc_problem *q) {
  const int dim = q->dim; const double *vfrac = q->f[ORC_F_VFRAC].data(); double *Gc = q->f[ORC_F_GC].data();
#pragma omp parallel for schedule(static)
  for (int ii = 0; ii < q->inum; ++ii) {
    const int i = q->ilist[ii], itype = q->type[i];
    double G[9] = {};
    for (long long p = q->noff[ii]; p < q->noff[ii + 1]; ++p) {
      const int j = q->neigh[p] & NEIGHMASK, jtype = q->type[j];
      double rsq = 0.0, rij[3];
      for (int k = 0; k < dim; ++k) { rij[k] = q->X(i)[k] - q->X(j)[k]; rsq += (rij[k] * rij[k]); }
      if (rsq < q->cut2(itype, jtype)) {
        const double r = sqrt(rsq) + EPS_R, dwdr = q->kern.dval(r, q->hh(itype, jtype));
        for (int k2 = 0; k2 < dim; ++k2) for (int k1 = 0; k1 < dim; ++k1) G[k2 * dim + k1] -= rij[k1] * rij[k2] * dwdr / r * vfrac[j];
      }
    }
    double *out = Gc + (size_t)9 * i; memset(out, 0, sizeof(double) * 9);
    invert_small(dim, G, out);
  }
  return 0;
}

// LU with partial pivoting + solve, one right-hand side (the role of DGESV at utils_reference.cpp:403)
static void gesv_small(int n, double *A, double *b) {
  int piv[6];
  for (int k = 0; k < n; ++k) {
    int p = k; double mx = fabs(A[k + k * n]);
    for (int i = k + 1; i < n; ++i) if (fabs(A[i + k * n]) > mx) { mx = fabs(A[i + k * n]); p = i; }
    piv[k] = p;
    if (p != k) for (int j = 0; j < n; ++j) std::swap(A[k + j * n], A[p + j * n]);
    const double rinv = 1.0 / A[k + k * n];
    for (int i = k + 1; i < n; ++i) A[i + k * n] *= rinv;
    for (int j = k + 1; j < n; ++j) { const double t = A[k + j * n]; for (int i = k + 1; i < n; ++i) A[i + j * n] -= A[i + k * n] * t; }
  }
  for (int k = 0; k < n; ++k) if (piv[k] != k) std::swap(b[k], b[piv[k]]);
  for (int k = 0; k < n; ++k) for (int i = k + 1; i < n; ++i) b[i] -= A[i + k * n] * b[k];
  for (int k = n - 1; k >= 0; --k) { b[k] /= A[k + k * n]; for (int i = 0; i < k; ++i) b[i] -= A[i + k * n] * b[k]; }
}

// functor_laplacian_correction.h:25-153
int orc_compute_laplacian_correction(orc_problem *q) {
  const int dim = q->dim, dimsq = dim * dim, dimL = q->dimL; const double *vfrac = q->f[ORC_F_VFRAC].data();
#pragma omp parallel for schedule(static)
  for (int ii = 0; ii < q->inum; ++ii) {
    const int i = q->ilist[ii], itype = q->type[i];
    double A[27] = {}, L[36] = {}; const double *G = &q->f[ORC_F_GC][(size_t)9 * i];
    for (long long p = q->noff[ii]; p < q->noff[ii + 1]; ++p) {          // :44-84  third-order tensor A^{kmn}
      const int j = q->neigh[p] & NEIGHMASK, jtype = q->type[j];
      double rsq = 0.0, rij[3] = {};
      for (int k = 0; k < dim; ++k) { rij[k] = q->X(i)[k] - q->X(j)[k]; rsq += (rij[k] * rij[k]); }
      if (rsq < q->cut2(itype, jtype)) {
        const double r = sqrt(rsq) + EPS_R, dwdr = q->kern.dval(r, q->hh(itype, jtype));
        double aij[3] = {};
        for (int k2 = 0; k2 < dim; ++k2) { for (int k1 = 0; k1 < dim; ++k1) aij[k2] += G[k2 * dim + k1] * rij[k1]; aij[k2] *= dwdr / r * vfrac[j]; }
        for (int k3 = 0; k3 < dim; ++k3) { double *slice = &A[k3 * dimsq];
          for (int k2 = 0; k2 < dim; ++k2) for (int k1 = 0; k1 < (k2 + 1); ++k1) slice[k2 * dim + k1] += aij[k3] * rij[k1] * rij[k2]; }
      }
    }
    for (long long p = q->noff[ii]; p < q->noff[ii + 1]; ++p) {          // :86-138  dimL x dimL system
      const int j = q->neigh[p] & NEIGHMASK, jtype = q->type[j];
      double rsq = 0.0, rij[3] = {};
      for (int k = 0; k < dim; ++k) { rij[k] = q->X(i)[k] - q->X(j)[k]; rsq += (rij[k] * rij[k]); }
      if (rsq < q->cut2(itype, jtype)) {
        const double r = sqrt(rsq) + EPS_R, dwdr = q->kern.dval(r, q->hh(itype, jtype));
        double eij[3]; for (int k = 0; k < dim; ++k) eij[k] = (rij[k] / r);
        double Cm[9] = {};
        for (int k3 = 0; k3 < dim; ++k3) { const double *slice = &A[k3 * dimsq];
          for (int k2 = 0; k2 < dim; ++k2) for (int k1 = 0; k1 < (k2 + 1); ++k1) Cm[k2 * dim + k1] += slice[k2 * dim + k1] * eij[k3]; }
        for (int k2 = 0; k2 < dim; ++k2) for (int k1 = 0; k1 < (k2 + 1); ++k1) { Cm[k2 * dim + k1] += rij[k1] * eij[k2]; Cm[k2 * dim + k1] *= dwdr * vfrac[j]; }
        const double scale[2] = {2.0, 1.0};
        for (int k4 = 0, op = 0; k4 < dim; ++k4) for (int k3 = 0; k3 < (k4 + 1); ++k3, ++op)
          for (int k2 = 0, mn = 0; k2 < dim; ++k2) for (int k1 = 0; k1 < (k2 + 1); ++k1, ++mn)
            L[op * dimL + mn] += Cm[k2 * dim + k1] * eij[k3] * eij[k4] * scale[k3 == k4];
      }
    }
    double *Lc = &q->f[ORC_F_LC][(size_t)6 * i]; memset(Lc, 0, sizeof(double) * 6);
    for (int k2 = 0, op = 0; k2 < dim; ++k2) for (int k1 = 0; k1 < (k2 + 1); ++k1, ++op) Lc[op] = -double(k1 == k2);   // :141-143
    gesv_small(dimL, L, Lc);                                                                                            // :146-151
  }
  return 0;
}

// functor_normal.h:56-125 run twice (fluid rows vs solid neighbours, then solid rows vs fluid neighbours) and
// ghost-exchanged, pair_isph_corrected.cpp:404-427
int orc_compute_normals(orc_problem *q) {
  const int dim = q->dim; const double *vfrac = q->f[ORC_F_VFRAC].data();
  double *normal = q->f[ORC_F_NORMAL].data(), *pnd = q->f[ORC_F_PND].data();
  const int pass_masks[2][2] = {{ORC_FLUID, ORC_SOLID}, {ORC_SOLID, ORC_FLUID}};
  for (int pass = 0; pass < 2; ++pass) {
    const int m0 = pass_masks[pass][0], m1 = pass_masks[pass][1];
#pragma omp parallel for schedule(static)
    for (int ii = 0; ii < q->inum; ++ii) {
      const int i = q->ilist[ii], itype = q->type[i], ikind = q->kind(itype);
      if (!fyes1(m0, ikind)) continue;                                    // returns before storing, :70-71
      double n_i[3] = {}, pnd_i = 0.0; const double *G = &q->f[ORC_F_GC][(size_t)9 * i];
      // orientation table, pair_isph_corrected.cpp:381-386
      const double orient = (ikind == ORC_FLUID || ikind == ORC_BUFFER_DIRICHLET || ikind == ORC_BUFFER_NEUMANN) ? -1.0 : ((ikind == ORC_SOLID || ikind == ORC_BOUNDARY) ? 1.0 : 0.0);
      for (long long p = q->noff[ii]; p < q->noff[ii + 1]; ++p) {
        const int j = q->neigh[p] & NEIGHMASK, jtype = q->type[j], jkind = q->kind(jtype);
        double rsq = 0.0, rij[3] = {};
        for (int k = 0; k < dim; ++k) { rij[k] = q->X(i)[k] - q->X(j)[k]; rsq += (rij[k] * rij[k]); }
        const double r = sqrt(rsq) + EPS_R;
        if (rsq < q->cut2(itype, jtype)) {
          if (fyes2(m0, m1, ikind, jkind)) {
            const double dwdr = q->kern.dval(r, q->hh(itype, jtype));
            for (int k2 = 0; k2 < dim; ++k2) { double gitmp = 0.0; for (int k1 = 0; k1 < dim; ++k1) gitmp += G[k2 * dim + k1] * rij[k1];
              n_i[k2] += gitmp * orient * dwdr / r * vfrac[j]; }
          } else {
            pnd_i += q->kern.val(r, q->hh(itype, jtype));
          }
        }
      }
      pnd_i += q->kern.val(0.0, q->hh(itype, itype));
      double alpha = 0.0; for (int k = 0; k < dim; ++k) alpha += n_i[k] * n_i[k];
      alpha = sqrt(alpha);
      if (alpha != 0.0) for (int k = 0; k < dim; ++k) n_i[k] /= alpha;
      memcpy(normal + (size_t)3 * i, n_i, sizeof(double) * dim); pnd[i] = pnd_i;
    }
  }
  q->forward(ORC_F_NORMAL); q->forward(ORC_F_PND);
  return 0;
}

// ---- graph ---------------------------------------------------------------------------------------------------

// functor_graph.h:38-99: row i = { tag[j] : rsq < cutsq } U { tag[i] } ; FillComplete sorts and removes duplicates.
// pair_isph.cpp:1266-1270: zero-valued matrix on that graph + two zero diagonal vectors.
long long orc_graph(orc_problem *q) {
  const int n = q->nlocal, dim = q->dim;
  std::vector<std::vector<int>> rows(n);
#pragma omp parallel for schedule(static)
  for (int ii = 0; ii < q->inum; ++ii) {
    const int i = q->ilist[ii], itype = q->type[i];
    std::vector<int> &r = rows[q->row_of(i)];
    for (long long p = q->noff[ii]; p < q->noff[ii + 1]; ++p) {
      const int j = q->neigh[p] & NEIGHMASK, jtype = q->type[j];
      double rsq = 0.0;
      for (int k = 0; k < dim; ++k) { const double d = q->X(i)[k] - q->X(j)[k]; rsq += (d * d); }
      if (rsq < q->cut2(itype, jtype)) r.push_back(q->tag[j]);
    }
    r.push_back(q->tag[i]);
    std::sort(r.begin(), r.end()); r.erase(std::unique(r.begin(), r.end()), r.end());
  }
  q->rowptr.assign(n + 1, 0);
  for (int i = 0; i < n; ++i) q->rowptr[i + 1] = q->rowptr[i] + (int)rows[i].size();
  q->col.resize(q->rowptr[n]); q->collid.resize(q->rowptr[n]);
  for (int i = 0; i < n; ++i) std::copy(rows[i].begin(), rows[i].end(), q->col.begin() + q->rowptr[i]);
  for (size_t p = 0; p < q->col.size(); ++p) { auto it = q->lid.find(q->col[p]); q->collid[p] = it == q->lid.end() ? -1 : it->second; }
  q->val.assign(q->col.size(), 0.0); q->diagonal.assign(n, 0.0); q->sld.assign(n, 0.0); q->is_filled = 0; q->have_graph = true;
  return (long long)q->col.size();
}
int orc_graph_get(orc_problem *q, int *rowptr, int *col) {
  if (!q->have_graph) return -1;
  memcpy(rowptr, q->rowptr.data(), sizeof(int) * q->rowptr.size()); memcpy(col, q->col.data(), sizeof(int) * q->col.size()); return 0;
}
int orc_graph_max_row(orc_problem *q) { int m = 0; for (int i = 0; i < q->nlocal; ++i) m = std::max(m, q->rowptr[i + 1] - q->rowptr[i]); return m; }

// ---- operator rows ---------------------------------------------------------------------------------------------

// Corrected::FunctorOuterLaplacianMatrix<Pair,Anti>::operator(), functor_laplacian_matrix.h:72-316 (iblock < 0, normal == NULL)
static void laplacian_matrix(Q *q, double alpha, const double *material, bool anti, bool mh, int f0, int f1) {
  const int dim = q->dim; const double *vfrac = q->f[ORC_F_VFRAC].data();
#pragma omp parallel
  {
    std::vector<double> val; std::vector<int> idx;
#pragma omp for schedule(static)
    for (int ii = 0; ii < q->inum; ++ii) {
      const int i = q->ilist[ii], itype = q->type[i], ikind = q->kind(itype), row = q->row_of(i);
      const double m_i = material ? material[i] : 1.0;
      if (!fyes1(f0, ikind)) { const int p = q->find(row, q->tag[i]); q->val[p] = 0.0; continue; }    // :88-96 ReplaceGlobalValues(diag, 0)
      const long long nb = q->noff[ii], ne = q->noff[ii + 1];
      val.assign(ne - nb + 1, 0.0); idx.assign(ne - nb + 1, 0);
      double gm[3] = {}, ci[3] = {};
      const double *L = anti ? q->Li : &q->f[ORC_F_LC][(size_t)6 * i];
      const double *G = anti ? q->Gi : &q->f[ORC_F_GC][(size_t)9 * i];
      int cnt = 0; double diag = 0.0;
      for (long long p = nb; p < ne; ++p) {                                                          // pass 1, :127-195
        const int j = q->neigh[p] & NEIGHMASK, jtype = q->type[j], jkind = q->kind(jtype);
        const double m_j = material ? material[j] : 1.0;
        double rsq = 0.0, rij[3] = {};
        for (int k = 0; k < dim; ++k) { rij[k] = q->X(i)[k] - q->X(j)[k]; rsq += (rij[k] * rij[k]); }
        const double cutsq = q->cut2(itype, jtype);
        if (rsq < cutsq) {
          double coeff = fyes2(f0, f1, ikind, ikind);
          if (!(ikind & ORC_SOLID) && (jkind & ORC_SOLID)) coeff = (fyes2(f0, f1, ikind, jkind) ? q->mirror(mh, i, j, sqrt(cutsq)) : 0.0);
          const double r = sqrt(rsq) + EPS_R, dwdr = q->kern.dval(r, q->hh(itype, jtype));
          double eij[3]; for (int k = 0; k < dim; ++k) eij[k] = rij[k] / r;
          const double vf = (anti ? sqrt(vfrac[i] * vfrac[j]) : vfrac[j]), vjtmp = dwdr * vf;
          for (int k2 = 0; k2 < dim; ++k2) {
            double gitmp = 0.0; for (int k1 = 0; k1 < dim; ++k1) gitmp += G[k2 * dim + k1] * eij[k1];
            const double ijtmp = gitmp * vjtmp;
            if (ikind & jkind) gm[k2] += ijtmp * (sph_op(anti, m_i, m_j));
          }
          double aij = 0.0; const double scale_a[2] = {2.0, 1.0};
          for (int k2 = 0, op = 0; k2 < dim; ++k2) for (int k1 = 0; k1 < (k2 + 1); ++k1, ++op) aij += L[op] * eij[k1] * eij[k2] * scale_a[k1 == k2];
          aij *= 2.0 * dwdr * vf;
          if (!anti) for (int k = 0; k < dim; ++k) ci[k] += aij * eij[k];
          aij *= m_i * coeff / r;
          val[cnt] = -aij; diag += aij; idx[cnt] = q->tag[j]; ++cnt;
        }
      }
      val[cnt] = diag; idx[cnt] = q->tag[i]; ++cnt;
      cnt = 0; diag = 0.0;
      for (long long p = nb; p < ne; ++p) {                                                          // pass 2, :211-259
        const int j = q->neigh[p] & NEIGHMASK, jtype = q->type[j], jkind = q->kind(jtype);
        double rsq = 0.0, rij[3] = {};
        for (int k = 0; k < dim; ++k) { rij[k] = q->X(i)[k] - q->X(j)[k]; rsq += (rij[k] * rij[k]); }
        const double cutsq = q->cut2(itype, jtype);
        if (rsq < cutsq) {
          double coeff = fyes2(f0, f1, ikind, ikind);
          if (!(ikind & ORC_SOLID) && (jkind & ORC_SOLID)) coeff = fyes2(f0, f1, ikind, jkind);
          const double r = sqrt(rsq) + EPS_R, dwdr = q->kern.dval(r, q->hh(itype, jtype));
          const double vf = (anti ? sqrt(vfrac[i] * vfrac[j]) : vfrac[j]), vjtmp = dwdr * vf;
          double eij[3]; for (int k = 0; k < dim; ++k) eij[k] = rij[k] / r;
          double bij[3] = {};
          for (int k2 = 0; k2 < dim; ++k2) for (int k1 = 0; k1 < dim; ++k1) bij[k2] += G[k2 * dim + k1] * eij[k1];
          double dot_c = 0.0, dot_g = 0.0;
          for (int k = 0; k < dim; ++k) dot_c += bij[k] * ci[k];                                      // util.dotVectors
          for (int k = 0; k < dim; ++k) dot_g += bij[k] * gm[k];
          const double tmp = coeff * (m_i * dot_c * vjtmp - dot_g * vjtmp);
          val[cnt] -= tmp; diag += tmp; ++cnt;
        }
      }
      val[cnt] += diag; ++cnt;
      for (int k = 0; k < cnt; ++k) val[k] *= alpha;                                                 // :267
      for (int k = 0; k < cnt; ++k) { const int p = q->find(row, idx[k]); q->val[p] += val[k]; }     // :270 SumIntoGlobalValues
    }
  }
}

// mirror/neighbour loop shared by the matrix-free gradient and divergence (functor_gradient.h:80-169, functor_divergence.h:55-124)
extern "C++" {
template <class Body> static void grad_like_loop(const Q *q, int ii, bool anti, bool mh, int f0, int f1, Body body) {
  const int dim = q->dim, i = q->ilist[ii], itype = q->type[i], ikind = q->kind(itype); const double *vfrac = q->f[ORC_F_VFRAC].data();
  if (!fyes1(f0, ikind)) return;
  const double *G = anti ? q->Gi : &q->f[ORC_F_GC][(size_t)9 * i];
  for (long long p = q->noff[ii]; p < q->noff[ii + 1]; ++p) {
    const int j = q->neigh[p] & NEIGHMASK, jtype = q->type[j], jkind = q->kind(jtype);
    if (!fyes2(f0, f1, ikind, jkind)) continue;
    double rsq = 0.0, rij[3] = {};
    for (int k = 0; k < dim; ++k) { rij[k] = q->X(i)[k] - q->X(j)[k]; rsq += (rij[k] * rij[k]); }
    const double cutsq = q->cut2(itype, jtype);
    if (rsq < cutsq) {
      double coeff = 1.0;
      if (!(ikind & ORC_SOLID) && (jkind & ORC_SOLID)) coeff = q->mirror(mh, i, j, sqrt(cutsq));
      const double r = sqrt(rsq) + EPS_R, dwdr = q->kern.dval(r, q->hh(itype, jtype));
      const double vf = (anti ? sqrt(vfrac[i] * vfrac[j]) : vfrac[j]), vjtmp = dwdr / r * vf * coeff;
      for (int k2 = 0; k2 < dim; ++k2) { double gitmp = 0.0; for (int k1 = 0; k1 < dim; ++k1) gitmp += G[k2 * dim + k1] * rij[k1]; body(j, k2, gitmp, vjtmp); }
    }
  }
}

}  // extern "C++"

// Corrected::FunctorOuterGradientOperator + FunctorOuterGradientDotOperatorMatrix,
// functor_gradient_operator.h:89-170, functor_gradient_dot_operator_matrix.h:36-79 (mirror = MirrorNothing: plain operator is bound)
static void gradient_dot_rows(Q *q, double alpha, const double *vec, int f0, int f1) {
  const int dim = q->dim; const double *vfrac = q->f[ORC_F_VFRAC].data();
#pragma omp parallel for schedule(static)
  for (int ii = 0; ii < q->inum; ++ii) {
    const int i = q->ilist[ii], itype = q->type[i], ikind = q->kind(itype), row = q->row_of(i);
    if (!fyes1(f0, ikind)) continue;
    const double *G = &q->f[ORC_F_GC][(size_t)9 * i];
    std::vector<double> v0, v1, v2; std::vector<int> at;
    at.push_back(i); v0.push_back(0.0); v1.push_back(0.0); v2.push_back(0.0);
    for (long long p = q->noff[ii]; p < q->noff[ii + 1]; ++p) {
      const int j = q->neigh[p] & NEIGHMASK, jtype = q->type[j], jkind = q->kind(jtype);
      if (!fyes2(f0, f1, ikind, jkind)) continue;
      double rsq = 0.0, rij[3] = {};
      for (int k = 0; k < dim; ++k) { rij[k] = q->X(i)[k] - q->X(j)[k]; rsq += (rij[k] * rij[k]); }
      if (rsq < q->cut2(itype, jtype)) {
        const double coeff = 1.0;                                     // MirrorNothing; (kind != Solid && kind == Solid) test only selects the mirror
        const double r = sqrt(rsq) + EPS_R, dwdr = q->kern.dval(r, q->hh(itype, jtype));
        const double vjtmp = dwdr / r * vfrac[j] * coeff;
        double t[3] = {};
        for (int k2 = 0; k2 < dim; ++k2) { double gitmp = 0.0; for (int k1 = 0; k1 < dim; ++k1) gitmp += G[k2 * dim + k1] * rij[k1]; t[k2] = gitmp * vjtmp; }
        at.push_back(j); v0.push_back(t[0]); v1.push_back(t[1]); v2.push_back(t[2]);
        v0[0] -= t[0]; v1[0] -= t[1]; v2[0] -= t[2];
      }
    }
    for (size_t k = 0; k < at.size(); ++k) {
      const double comp[3] = {v0[k] * alpha, v1[k] * alpha, v2[k] * alpha};                           // _val.Scale(_alpha), :168
      double s = 0.0; for (int d = 0; d < dim; ++d) s += comp[d] * vec[(size_t)3 * i + d];            // functor_gradient_dot_operator_matrix.h:69-73
      const int p = q->find(row, q->tag[at[k]]); q->val[p] += s;                                      // :77 SumIntoGlobalValues
    }
  }
}

static void extract_diag(Q *q, std::vector<double> &d) { for (int i = 0; i < q->nlocal; ++i) { const int p = q->find(i, q->tag[i]); d[i] = p < 0 ? 0.0 : q->val[p]; } }
static void replace_diag(Q *q, const std::vector<double> &d) { for (int i = 0; i < q->nlocal; ++i) { const int p = q->find(i, q->tag[i]); if (p >= 0) q->val[p] = d[i]; } }

// FunctorOuterIncompNavierStokesPoisson, functor_incomp_navier_stokes_poisson.h:47-181
int orc_ns_poisson(orc_problem *q, double dt, int anti, int singular, int mh, double *b) {
  if (!q->have_graph || q->is_filled == 1) return -1;                                                // :58-59
  const int dim = q->dim;
  std::fill(q->val.begin(), q->val.end(), 0.0);                                                      // :61 PutScalar(0)
  int f1; bool neumann;
  if (singular == ORC_NOT_SINGULAR) { f1 = ORC_ALL; neumann = false; } else { f1 = ORC_FLUID; neumann = true; }   // :72-87
  std::vector<double> inv_rho(q->nall);
  for (int i = 0; i < q->nall; ++i) inv_rho[i] = 1.0 / q->f[ORC_F_DENSITY][i];                        // :89-92
  laplacian_matrix(q, -dt, inv_rho.data(), anti != 0, false, ORC_FLUID, f1);                         // :94-97 (plain LaplacianMatrix is bound even in the MorrisHolmes variant, pair_isph_corrected.cpp:174-178)
  q->is_filled = 1;                                                                                   // exitFor, functor_laplacian_matrix.h:322-326
  if (neumann) gradient_dot_rows(q, -dt, q->f[ORC_F_NORMAL].data(), ORC_SOLID, ORC_ALL);              // :100-109
  extract_diag(q, q->sld);                                                                            // :111
  const double *vstar = q->f[ORC_F_VSTAR].data(), *normal = q->f[ORC_F_NORMAL].data();
  bool is_once = false;
  std::vector<double> bb(q->nall, 0.0);
  for (int ii = 0; ii < q->inum; ++ii) {                                                              // :127-170 (serial: the "first fluid row" rule)
    const int i = q->ilist[ii], ikind = q->kind(q->type[i]);
    double &diag = q->diagonal[i];
    if (ikind == ORC_SOLID) {
      if (neumann) { double nn = 0.0; for (int k = 0; k < dim; ++k) nn += normal[3 * (size_t)i + k] * normal[3 * (size_t)i + k]; if (nn < 0.5) diag = 1.0; }
      else diag = 1.0;
      bb[i] = 0.0;
    } else if (ikind == ORC_BUFFER_NEUMANN || ikind == ORC_BUFFER_DIRICHLET || ikind == ORC_FLUID) {
      diag = q->sld[i];
      double div = 0.0;                                                                               // divergence of vstar, filter (Fluid, All), alpha = 1
      grad_like_loop(q, ii, anti != 0, mh != 0, ORC_FLUID, ORC_ALL, [&](int j, int k2, double gitmp, double vjtmp) {
        div += gitmp * (sph_op(anti != 0, vstar[3 * (size_t)i + k2], vstar[3 * (size_t)j + k2])) * vjtmp; });
      div *= 1.0;
      bb[i] = -div;
      if (!is_once) {                                                                                 // modifySingularMatrix, pair_isph.cpp:493-520
        if (singular == ORC_PINZERO) { const int row = q->row_of(i); for (int p = q->rowptr[row]; p < q->rowptr[row + 1]; ++p) q->val[p] = 0.0; diag = -1.0; bb[i] = 0.0; }
        else if (singular == ORC_DOUBLEDIAG) diag *= 1.5;
        is_once = true;
      }
    } else return -2;
  }
  replace_diag(q, q->diagonal);                                                                       // :180
  memcpy(b, bb.data(), sizeof(double) * q->nlocal);
  return 0;
}

static void spmv_rows(const Q *q, const double *x, double *y, int nvec) {
  const int n = q->nlocal;
  for (int c = 0; c < nvec; ++c) {
    const double *xc = x + (size_t)c * n; double *yc = y + (size_t)c * n;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) { double s = 0.0; for (int p = q->rowptr[i]; p < q->rowptr[i + 1]; ++p) s += q->val[p] * xc[q->collid[p]]; yc[i] = s; }
  }
}

// FunctorOuterIncompNavierStokesHelmholtz, functor_incomp_navier_stokes_helmholtz.h:48-159
int orc_ns_helmholtz(orc_problem *q, double dt, double theta, int anti, int mh, int incp, const double *g, double *b) {
  if (!q->have_graph || q->is_filled == 1) return -1;
  const int dim = q->dim, n = q->nlocal;
  std::fill(q->val.begin(), q->val.end(), 0.0);                                                       // :58
  std::vector<double> mu(q->nall);
  for (int i = 0; i < q->nall; ++i) mu[i] = q->f[ORC_F_VISCOSITY][i] * q->f[ORC_F_DENSITY][i];         // :69-72
  laplacian_matrix(q, dt, mu.data(), anti != 0, mh != 0, ORC_FLUID, ORC_ALL);                         // :74-77
  q->is_filled = 1;
  for (int i = 0; i < n; ++i) { const double s = 1.0 / q->f[ORC_F_DENSITY][i]; for (int p = q->rowptr[i]; p < q->rowptr[i + 1]; ++p) q->val[p] *= s; }   // :80-83 LeftScale
  std::vector<double> w((size_t)n * dim);
  spmv_rows(q, b, w.data(), dim);                                                                     // :90
  for (auto &v : w) v *= (1.0 - theta);                                                               // :91
  for (auto &v : q->val) v *= (-theta);                                                               // :94
  extract_diag(q, q->sld);                                                                            // :95
  const double *rho = q->f[ORC_F_DENSITY].data(), *f = q->f[ORC_F_FORCE].data(), *pr = q->f[ORC_F_PRESSURE].data();
  for (int ii = 0; ii < q->inum; ++ii) {                                                              // :108-150
    const int i = q->ilist[ii], ikind = q->kind(q->type[i]);
    double &diag = q->diagonal[i];
    if (ikind == ORC_SOLID) { diag = 1.0; }
    else if (ikind == ORC_BUFFER_DIRICHLET || ikind == ORC_BUFFER_NEUMANN || ikind == ORC_FLUID) {
      diag = 1.0 + q->sld[i];
      for (int k = 0; k < dim; ++k) { b[(size_t)k * n + i] += w[(size_t)k * n + i]; b[(size_t)k * n + i] += dt * (f[3 * (size_t)i + k] / rho[i] + (g ? g[k] : 0.0)); }
      if (incp) {
        double gp[3] = {};                                                                            // gradient of p, filter (Fluid, Fluid), plain mirror (functor_gradient.h)
        grad_like_loop(q, ii, anti != 0, false, ORC_FLUID, ORC_FLUID, [&](int j, int k2, double gitmp, double vjtmp) {
          const double ijtmp = gitmp * vjtmp; gp[k2] += ijtmp * (sph_op(anti != 0, pr[i], pr[j])); });
        for (int k = 0; k < dim; ++k) gp[k] *= 1.0;
        for (int k = 0; k < dim; ++k) b[(size_t)k * n + i] += dt * (-1.0 / rho[i] * gp[k]);
      }
    } else return -2;
  }
  replace_diag(q, q->diagonal);                                                                       // :158
  return 0;
}

// FunctorOuterPoissonBoltzmannJacobian, functor_poisson_boltzmann_jacobian.h:35-107
int orc_pb_jacobian(orc_problem *q, int mh, int linearized, double ezcb, double psiref, double gamma) {
  if (!q->have_graph) return -1;
  const double kappasq = 2.0 * ezcb / psiref;
  if (q->is_filled == 0) {                                                                            // :50-65
    std::fill(q->val.begin(), q->val.end(), 0.0);
    laplacian_matrix(q, -1.0, q->f[ORC_F_EPS].data(), false, mh != 0, ORC_FLUID, ORC_ALL);
    extract_diag(q, q->sld); q->is_filled = 1;
  }
  const double *psi_ = q->f[ORC_F_PSI].data();
  for (int ii = 0; ii < q->inum; ++ii) {                                                              // :68-99
    const int i = q->ilist[ii], ikind = q->kind(q->type[i]);
    if (ikind == ORC_SOLID || ikind == ORC_BOUNDARY) q->diagonal[i] = -1.0;
    else if (ikind == ORC_BUFFER_DIRICHLET || ikind == ORC_BUFFER_NEUMANN || ikind == ORC_FLUID) {
      q->diagonal[i] = q->sld[i]; const double psi = psi_[i];
      if (linearized) {
        const double numerator = 4.0 - 2.0 * gamma * pow(psi, 2), denominator = pow(gamma, 2) * pow(psi, 4) + 4.0 * gamma * pow(psi, 2) + 4.0;
        q->diagonal[i] += kappasq * (numerator / denominator);
      } else {
        const double numerator = 2.0 * gamma * cosh(0.5 * psi) * sinh(0.5 * psi) * sinh(psi), denominator = 2.0 * gamma * pow(sinh(0.5 * psi), 2) + 1.0;
        q->diagonal[i] += kappasq * (cosh(psi) / denominator - numerator / pow(denominator, 2));
      }
    }
  }
  replace_diag(q, q->diagonal);                                                                       // :105
  return 0;
}

// Corrected::FunctorOuterGradient (functor_gradient.h:80-169) on a scalar field, alpha = 1: grad_i = sum_j (G_i r_ij) dW/dr / r V_j coeff op(f_i, f_j)
int orc_scalar_gradient(orc_problem *q, int field, int anti, int mh, int f0, int f1, double *grad) {
  if (orc_field_ncomp(field) != 1) return -1;
  q->forward(field);
  const double *fld = q->f[field].data(); const int dim = q->dim;
  std::fill(grad, grad + (size_t)3 * q->nlocal, 0.0);
#pragma omp parallel for schedule(static)
  for (int ii = 0; ii < q->inum; ++ii) {
    const int i = q->ilist[ii]; double g[3] = {};
    grad_like_loop(q, ii, anti != 0, mh != 0, f0, f1, [&](int j, int k2, double gitmp, double vjtmp) { const double ijtmp = gitmp * vjtmp; g[k2] += ijtmp * (sph_op(anti != 0, fld[i], fld[j])); });
    for (int k = 0; k < dim; ++k) grad[3 * (size_t)i + k] = g[k] * 1.0;
  }
  return 0;
}

// FunctorOuterAppliedElectricPotential, functor_applied_electric_potential.h:34-96
int orc_applied_electric_potential(orc_problem *q, double *b) {
  if (!q->have_graph || q->is_filled) return -1;
  std::fill(q->val.begin(), q->val.end(), 0.0);                                                       // :47
  laplacian_matrix(q, -1.0, q->f[ORC_F_SIGMA].data(), false, false, ORC_FLUID | ORC_FILTER_MATCH, ORC_FLUID);   // :49-57
  extract_diag(q, q->diagonal);                                                                       // :58
  const double *phi = q->f[ORC_F_PHI].data();
  for (int ii = 0; ii < q->inum; ++ii) {                                                              // :64-88
    const int i = q->ilist[ii], ikind = q->kind(q->type[i]);
    b[i] = 0.0;
    if (ikind == ORC_SOLID) q->diagonal[i] = 1.0;
    else if (ikind == ORC_BUFFER_NEUMANN || ikind == ORC_BUFFER_DIRICHLET) { q->diagonal[i] = 1.0; b[i] = phi[i]; }
  }
  replace_diag(q, q->diagonal);                                                                       // :94
  return 0;
}

static void spmv_rows(const Q *q, const double *x, double *y, int nvec);

// FunctorOuterSoluteTransport, functor_solute_transport.h:47-134
int orc_solute_transport(orc_problem *q, double dt, double theta, double dcoeff, double *b) {
  if (!q->have_graph || q->is_filled) return -1;
  std::fill(q->val.begin(), q->val.end(), 0.0);                                                       // :60
  laplacian_matrix(q, dt * dcoeff, nullptr, false, false, ORC_FLUID | ORC_FILTER_MATCH, ORC_FLUID - ORC_BUFFER_NEUMANN);   // :62-70
  std::vector<double> w(q->nlocal, 0.0);
  spmv_rows(q, b, w.data(), 1);                                                                       // :88
  for (auto &v : w) v *= (1.0 - theta);                                                               // :89
  for (auto &v : q->val) v *= -theta;                                                                 // :92
  extract_diag(q, q->sld);                                                                            // :93
  for (int ii = 0; ii < q->inum; ++ii) {                                                              // :100-126
    const int i = q->ilist[ii], ikind = q->kind(q->type[i]);
    if (ikind == ORC_BUFFER_DIRICHLET || ikind == ORC_BUFFER_NEUMANN || ikind == ORC_SOLID) q->diagonal[i] = 1.0;
    else if (ikind == ORC_FLUID) { q->diagonal[i] = 1.0 + q->sld[i]; b[i] += w[i]; }
    else return -1;
  }
  replace_diag(q, q->diagonal);                                                                       // :132
  return 0;
}

// Matrix-free corrected Laplacian of a scalar field, Corrected::FunctorOuterLaplacianHelper::operator() with scalar_diff
// (functor_laplacian.h:67-277): corrected gradient of psi and of the material first, then sum a_ij (coeff (psi_i-psi_j)/r -
// e_ij . grad psi), combined as alpha (m_i lap + grad m . grad psi).  Same operation order as the reference.
static double scalar_laplacian(const Q *q, int ii, const double *fld, double alpha, const double *material, bool mh, int f0, int f1) {
  const int dim = q->dim, i = q->ilist[ii], itype = q->type[i], ikind = q->kind(itype); const double *vfrac = q->f[ORC_F_VFRAC].data();
  const double m_i = material ? material[i] : 1.0;
  if (!fyes1(f0, ikind)) return 0.0;                                                                 // :99-100
  const long long nb = q->noff[ii], ne = q->noff[ii + 1];
  std::vector<double> aijs(ne - nb, 0.0);
  double gm[3] = {}, gf[3] = {}, lap = 0.0;
  const double *G = &q->f[ORC_F_GC][(size_t)9 * i], *L = &q->f[ORC_F_LC][(size_t)6 * i];
  for (long long p = nb; p < ne; ++p) {                                                              // :111-158
    const int j = q->neigh[p] & NEIGHMASK, jtype = q->type[j], jkind = q->kind(jtype);
    const double m_j = material ? material[j] : 1.0;
    if (!fyes2(f0, f1, ikind, jkind)) continue;
    double rsq = 0.0, rij[3] = {};
    for (int k = 0; k < dim; ++k) { rij[k] = q->X(i)[k] - q->X(j)[k]; rsq += (rij[k] * rij[k]); }
    const double cutsq = q->cut2(itype, jtype);
    if (rsq < cutsq) {
      double coeff = 1.0;
      if (jkind & ORC_SOLID) coeff = q->mirror(mh, i, j, sqrt(cutsq));
      const double r = sqrt(rsq) + EPS_R, dwdr = q->kern.dval(r, q->hh(itype, jtype));
      const double vjtmp = dwdr / r * vfrac[j] * coeff;
      for (int k2 = 0; k2 < dim; ++k2) {
        double gitmp = 0.0; for (int k1 = 0; k1 < dim; ++k1) gitmp += G[k2 * dim + k1] * rij[k1];
        const double ijtmp = gitmp * vjtmp;
        if (fyes1(f0, jkind)) gm[k2] += ijtmp * (m_j - m_i);
        gf[k2] += ijtmp * (fld[j] - fld[i]);
      }
    }
  }
  for (long long p = nb; p < ne; ++p) {                                                              // :161-216 (no filter on this pass)
    const int j = q->neigh[p] & NEIGHMASK, jtype = q->type[j];
    double rsq = 0.0, rij[3] = {};
    for (int k = 0; k < dim; ++k) { rij[k] = q->X(i)[k] - q->X(j)[k]; rsq += (rij[k] * rij[k]); }
    const double cutsq = q->cut2(itype, jtype);
    if (rsq < cutsq) {
      const double r = sqrt(rsq) + EPS_R, dwdr = q->kern.dval(r, q->hh(itype, jtype));
      double eij[3]; for (int k = 0; k < dim; ++k) eij[k] = rij[k] / r;
      double aij = 0.0; const double scale_a[2] = {2.0, 1.0};
      for (int k2 = 0, op = 0; k2 < dim; ++k2) for (int k1 = 0; k1 < (k2 + 1); ++k1, ++op) aij += L[op] * eij[k1] * eij[k2] * scale_a[k1 == k2];
      aij *= 2.0 * dwdr * vfrac[j];
      double dotv = 0.0; for (int k = 0; k < dim; ++k) dotv += eij[k] * gf[k];
      const double bij = -1.0 * dotv;
      lap += aij * bij;
      aijs[p - nb] = aij;
    }
  }
  for (long long p = nb; p < ne; ++p) {                                                              // :219-263
    const int j = q->neigh[p] & NEIGHMASK, jtype = q->type[j], jkind = q->kind(jtype);
    if (!fyes2(f0, f1, ikind, jkind)) continue;
    double rsq = 0.0, rij[3] = {};
    for (int k = 0; k < dim; ++k) { rij[k] = q->X(i)[k] - q->X(j)[k]; rsq += (rij[k] * rij[k]); }
    const double cutsq = q->cut2(itype, jtype);
    if (rsq < cutsq) {
      double coeff = 1.0;
      if (jkind & ORC_SOLID) coeff = q->mirror(mh, i, j, sqrt(cutsq));
      const double r = sqrt(rsq) + EPS_R;
      const double bij = coeff * (fld[i] - fld[j]) / r;
      lap += aijs[p - nb] * bij;
    }
  }
  double dotm = 0.0; for (int k = 0; k < dim; ++k) dotm += gm[k] * gf[k];
  return alpha * (m_i * lap + dotm);                                                                 // :266-268
}

// FunctorOuterPoissonBoltzmannF, functor_poisson_boltzmann_f.h:58-88 (+ extra source, functor_poisson_boltzmann_extra_f.h:76-90)
int orc_pb_residual(orc_problem *q, int mh, int linearized, double ezcb, double psiref, double gamma, const double *extra_f, double *fout) {
  const double kappasq = 2.0 * ezcb / psiref;
  q->forward(ORC_F_PSI);                                                                              // pair_isph_corrected.cpp:446-450
  const double *psi = q->f[ORC_F_PSI].data(), *psi0 = q->f[ORC_F_PSI0].data(), *eps = q->f[ORC_F_EPS].data();
#pragma omp parallel for schedule(static)
  for (int ii = 0; ii < q->inum; ++ii) {
    const int i = q->ilist[ii], ikind = q->kind(q->type[i]);
    if (ikind == ORC_SOLID || ikind == ORC_BOUNDARY) fout[i] = (-psi[i] + psi0[i]);
    else if (ikind == ORC_BUFFER_DIRICHLET || ikind == ORC_BUFFER_NEUMANN || ikind == ORC_FLUID) {
      fout[i] = scalar_laplacian(q, ii, psi, -1.0, eps, mh != 0, ORC_FLUID, ORC_ALL);
      if (linearized) fout[i] += kappasq * (psi[i] / (1.0 + 2.0 * gamma * pow(psi[i] / 2, 2)));
      else fout[i] += kappasq * (sinh(psi[i]) / (1.0 + 2.0 * gamma * pow(sinh(psi[i] / 2.0), 2)));
    }
    if (extra_f && !(ikind & ORC_SOLID)) fout[i] += extra_f[i];
  }
  return 0;
}

// pair_isph.cpp:1017-1031: forward_comm(DeltaP) ; computeZeroMeanPressure :422-464 ; correctVelocity
// (functor_correct_velocity.h:52-78: vstar_i -= dt/rho_i grad(dp)_i, gradient filter (Fluid,Fluid), then forward_comm(Vstar)) ;
// correctPressure (functor_correct_pressure.h:29-43: p += dp or p = dp for owned AND ghost particles)
int orc_ns_correct(orc_problem *q, double dt, int anti, int incp, const double *dp_owned) {
  const int dim = q->dim; double *dp = q->f[ORC_F_DP].data(), *vstar = q->f[ORC_F_VSTAR].data(), *pr = q->f[ORC_F_PRESSURE].data();
  const double *rho = q->f[ORC_F_DENSITY].data();
  memcpy(dp, dp_owned, sizeof(double) * q->nlocal);
  q->forward(ORC_F_DP);
  if (incp) {
    int nloc = 0; double mysum = 0.0;
    for (int ii = 0; ii < q->inum; ++ii) { const int i = q->ilist[ii], ikind = q->kind(q->type[i]); if (ikind == ORC_SOLID) dp[i] = 0.0; else { mysum += dp[i]; ++nloc; } }
    const double mean_val = mysum / nloc;
    for (int i = 0; i < q->nall; ++i) dp[i] -= mean_val * (q->kind(q->type[i]) != ORC_SOLID);
  }
#pragma omp parallel for schedule(static)
  for (int ii = 0; ii < q->inum; ++ii) {
    const int i = q->ilist[ii], ikind = q->kind(q->type[i]);
    if (!fyes1(ORC_FLUID, ikind)) continue;
    double g[3] = {};
    grad_like_loop(q, ii, anti != 0, false, ORC_FLUID, ORC_FLUID, [&](int j, int k2, double gitmp, double vjtmp) { const double ijtmp = gitmp * vjtmp; g[k2] += ijtmp * (sph_op(anti != 0, dp[i], dp[j])); });
    for (int k = 0; k < dim; ++k) g[k] *= 1.0;
    for (int k = 0; k < dim; ++k) vstar[3 * (size_t)i + k] -= dt / rho[i] * g[k];
  }
  q->forward(ORC_F_VSTAR);
  for (int i = 0; i < q->nall; ++i) { if (incp) pr[i] += dp[i]; else pr[i] = dp[i]; }
  return 0;
}
int orc_set_fixed(orc_problem *q, const int *fixed_of_type) { q->fixed_of_type.assign(fixed_of_type, fixed_of_type + q->ntypes + 1); return 0; }
int orc_get_x(orc_problem *q, double *x) { memcpy(x, q->x.data(), sizeof(double) * 3 * q->nall); return 0; }

// PairISPH_Corrected::advanceTime, pair_isph_corrected.cpp:1183-1194
int orc_advance_time(orc_problem *q, double dt, int anti) {
  const int dim = q->dim; double *dp = q->f[ORC_F_DP].data(), *v = q->f[ORC_F_VELOCITY].data(), *pr = q->f[ORC_F_PRESSURE].data();
  const double *vnp1 = q->f[ORC_F_VSTAR].data();
  // FunctorOuterAdvanceTimeBegin, functor_advance_time_begin.h:52-81: dp_i = grad(p)_i . dx_i, dx = 0.5 dt (v^{n+1} + v^n); the gradient
  // operator carries FilterBinary(Fluid, Fluid) and alpha = 1 (functor_gradient.h:80-169)
#pragma omp parallel for schedule(static)
  for (int ii = 0; ii < q->inum; ++ii) {
    const int i = q->ilist[ii], ikind = q->kind(q->type[i]);
    double dx[3] = {};
    for (int k = 0; k < dim; ++k) dx[k] = 0.5 * dt * (vnp1[3 * (size_t)i + k] + v[3 * (size_t)i + k]);
    double g[3] = {};
    grad_like_loop(q, ii, anti != 0, false, ORC_FLUID, ORC_FLUID, [&](int j, int k2, double gitmp, double vjtmp) { const double ijtmp = gitmp * vjtmp; g[k2] += ijtmp * (sph_op(anti != 0, pr[i], pr[j])); });
    for (int k = 0; k < dim; ++k) g[k] *= 1.0;
    if (fyes1(ORC_FLUID, ikind)) { double s = 0.0; for (int k = 0; k < dim; ++k) s += g[k] * dx[k]; dp[i] = s; }      // util.dotVectors
    else dp[i] = 0.0;
  }
  q->forward(ORC_F_DP);                                                                    // exitFor :74-78
  // FunctorOuterAdvanceTimeEnd over owned + ghost atoms, functor_advance_time_end.h:48-66
  for (int i = 0; i < q->nall; ++i) {
    if (!q->fixed_of_type.empty() && q->fixed_of_type[q->type[i]]) { for (int k = 0; k < dim; ++k) v[3 * (size_t)i + k] = vnp1[3 * (size_t)i + k]; continue; }
    pr[i] += dp[i];
    for (int k = 0; k < dim; ++k) {
      const double delta = 0.5 * dt * (vnp1[3 * (size_t)i + k] + v[3 * (size_t)i + k]);
      q->x[3 * (size_t)i + k] += delta; v[3 * (size_t)i + k] = vnp1[3 * (size_t)i + k];
    }
  }
  return 0;
}

// Corrected::FunctorOuterBoundaryNavierSlip, functor_boundary_navier_slip.h:54-174 (iblock < 0, add_neumann_term = true)
int orc_boundary_navier_slip(orc_problem *q, double beta) {
  if (!q->have_graph) return -1;
  if (beta == 0.0) return 0;                                                               // pair_isph_corrected.cpp:922
  const int dim = q->dim; const double *vfrac = q->f[ORC_F_VFRAC].data(), *rho = q->f[ORC_F_DENSITY].data(), *nrm = q->f[ORC_F_NORMAL].data();
#pragma omp parallel for schedule(static)
  for (int ii = 0; ii < q->inum; ++ii) {
    const int i = q->ilist[ii], itype = q->type[i], ikind = q->kind(itype), row = q->row_of(i);
    if (ikind == ORC_SOLID) continue;
    if (!(ikind == ORC_FLUID || ikind == ORC_BUFFER_DIRICHLET || ikind == ORC_BUFFER_NEUMANN)) continue;   // (the reference raises an error for other kinds)
    const double *G = &q->f[ORC_F_GC][(size_t)9 * i];
    double robin_at_i = 0.0;
    for (long long p = q->noff[ii]; p < q->noff[ii + 1]; ++p) {
      const int j = q->neigh[p] & NEIGHMASK, jtype = q->type[j];
      if (q->kind(jtype) != ORC_SOLID) continue;
      double rsq = 0.0, rij[3] = {};
      for (int k = 0; k < dim; ++k) { rij[k] = q->X(i)[k] - q->X(j)[k]; rsq += (rij[k] * rij[k]); }
      if (rsq < q->cut2(itype, jtype)) {
        const double r = sqrt(rsq) + EPS_R, dwdr = q->kern.dval(r, q->hh(itype, jtype));
        double aij[3] = {};
        for (int k2 = 0; k2 < dim; ++k2) for (int k1 = 0; k1 < dim; ++k1) aij[k2] += G[k2 * dim + k1] * rij[k1];     // VIEW2(G, dim, k1, k2)
        double tmp = 0.0;
        for (int k = 0; k < dim; ++k) tmp += (nrm[3 * (size_t)i + k] + nrm[3 * (size_t)j + k]) * aij[k];
        const double robin_at_j = beta * dwdr / r * vfrac[j] / rho[i] * tmp;
        const int pos = q->find(row, q->tag[j]); if (pos >= 0) q->val[pos] += robin_at_j;                            // SumIntoGlobalValues
        robin_at_i -= robin_at_j;
      }
    }
    const int pos = q->find(row, q->tag[i]); if (pos >= 0) q->val[pos] += robin_at_i;
  }
  return 0;
}

// Corrected::FunctorOuterBoundaryDirichlet, functor_boundary_dirichlet.h:47-150
int orc_boundary_dirichlet(orc_problem *q, double *b, int lda) {
  if (!q->have_graph) return -1;
  const int dim = q->dim; const double *nrm = q->f[ORC_F_NORMAL].data();
  for (int ii = 0; ii < q->inum; ++ii) {
    const int i = q->ilist[ii], itype = q->type[i], ikind = q->kind(itype), row = q->row_of(i);
    if (ikind == ORC_SOLID) continue;
    if (ikind != ORC_FLUID) return -1;                                                     // "Particle types are not supported"
    int n_solid = 0;
    for (long long p = q->noff[ii]; p < q->noff[ii + 1]; ++p) {
      const int j = q->neigh[p] & NEIGHMASK, jtype = q->type[j];
      double rsq = 0.0; for (int k = 0; k < dim; ++k) { const double d = q->X(i)[k] - q->X(j)[k]; rsq += (d * d); }
      if (rsq < pow(q->hh(itype, jtype), 2)) n_solid += int(q->kind(jtype) == ORC_SOLID);
    }
    if (n_solid <= 0) continue;
    double xn_i = 0.0, xn_av = 0.0;
    for (int k = 0; k < dim; ++k) xn_i += q->X(i)[k] * nrm[3 * (size_t)i + k];
    std::vector<double> xn; std::vector<int> tg, idx; std::vector<double> val;
    for (long long p = q->noff[ii]; p < q->noff[ii + 1]; ++p) {
      const int j = q->neigh[p] & NEIGHMASK, jtype = q->type[j];
      double rsq = 0.0; for (int k = 0; k < dim; ++k) { const double d = q->X(i)[k] - q->X(j)[k]; rsq += (d * d); }
      if (rsq < pow(q->hh(itype, jtype), 2)) {                                              // the interface layer is h wide, not cut
        double xn_j = 0.0; for (int k = 0; k < dim; ++k) xn_j += q->X(j)[k] * nrm[3 * (size_t)i + k];
        xn_av += xn_j; tg.push_back(q->tag[j]); xn.push_back(xn_j);
      } else if (rsq < q->cut2(itype, jtype)) { val.push_back(0.0); idx.push_back(q->tag[j]); }
    }
    const int natoms_cut = (int)xn.size();
    xn_av /= natoms_cut;
    double xterm = 0.0;
    for (int j = 0; j < natoms_cut; ++j) xterm += xn[j] * (xn[j] - xn_av);
    for (int j = 0; j < natoms_cut; ++j) { val.push_back(-(xn_i - xn_av) * (xn[j] - xn_av) / xterm - 1.0 / natoms_cut); idx.push_back(tg[j]); }
    val.push_back(1.0); idx.push_back(q->tag[i]);
    for (size_t k = 0; k < idx.size(); ++k) { const int pos = q->find(row, idx[k]); if (pos >= 0) q->val[pos] = val[k]; }      // ReplaceGlobalValues, in list order
    for (int k = 0; k < dim; ++k) b[(size_t)k * lda + i] = 0;                               // VIEW2(_b, _lda, i, k)
  }
  return 0;
}

int orc_invalidate_matrix(orc_problem *q) { q->is_filled = 0; return 0; }
int orc_matrix_get(orc_problem *q, double *val) { if (!q->have_graph) return -1; memcpy(val, q->val.data(), sizeof(double) * q->val.size()); return 0; }
int orc_diag_get(orc_problem *q, double *d, double *s) {
  if (!q->have_graph) return -1;
  if (d) memcpy(d, q->diagonal.data(), sizeof(double) * q->nlocal);
  if (s) memcpy(s, q->sld.data(), sizeof(double) * q->nlocal);
  return 0;
}
int orc_spmv(orc_problem *q, const double *x, double *y, int nvec) { if (!q->have_graph) return -1; spmv_rows(q, x, y, nvec); return 0; }

}  // extern "C"
