// Operator assembly on the SELL-32 pattern: one thread per matrix row, a warp = one slice, so every load of the
// (atom, col, val) streams is a fully coalesced 128/256-byte line.  The neighbor set of a row is the set of matrix
// entries the graph kernel kept (`atom` array), i.e. exactly the pairs with rsq < cutsq of the reference loops.
//
// Replaces (paths relative to IMPLICIT-SPH/):
//   functor_volume.h:42-81, functor_gradient_correction.h:24-71, functor_laplacian_correction.h:25-153,
//   functor_normal.h:56-125, functor_laplacian_matrix.h:72-316 (+ mirror_morris_holmes.h:39-53),
//   functor_gradient_operator.h:89-170 + functor_gradient_dot_operator_matrix.h:36-79,
//   functor_divergence.h:55-124, functor_gradient.h:80-169,
//   functor_incomp_navier_stokes_poisson.h:47-181, functor_incomp_navier_stokes_helmholtz.h:48-159,
//   functor_poisson_boltzmann_jacobian.h:35-107.
// Compiled with -fmad=false: expressions keep the reference's operation order and rounding (values are compared to
// the reference at 1e-12 and near-cutoff entries cancel catastrophically in 1-s/2).  Sums over a row run in column
// order instead of neighbor-list order; that is the only reassociation.
#include "isph_internal.h"
#include <algorithm>

namespace isph {

// ---- SPH kernel functions (kernel_wendland.h:44-63, kernel_cubic.h:43-69, kernel_quintic.h:43-80) ----------------
// Integer powers rounded once, like glibc's pow(x, n) that the reference calls (kernel_wendland.h:47,58): the product
// chain is carried in double-double (error-free products through explicit FMAs) and rounded at the end.  A plain
// x*x*x is 1-2 ulp off, which the 6x6 Laplacian-correction solve would amplify by its condition number.
struct dd { double hi, lo; };
__device__ __forceinline__ dd two_prod(double a, double b) { dd r; r.hi = a * b; r.lo = __fma_rn(a, b, -r.hi); return r; }
__device__ __forceinline__ dd dd_renorm(double hi, double lo) { dd r; r.hi = hi + lo; r.lo = lo - (r.hi - hi); return r; }
__device__ __forceinline__ dd dd_mul_d(dd a, double b) { dd p = two_prod(a.hi, b); return dd_renorm(p.hi, p.lo + a.lo * b); }
__device__ __forceinline__ dd dd_mul(dd a, dd b) { dd p = two_prod(a.hi, b.hi); return dd_renorm(p.hi, p.lo + (a.hi * b.lo + a.lo * b.hi)); }
__device__ __forceinline__ double p2(double x) { return x * x; }
__device__ __forceinline__ double p3(double x) { return dd_mul_d(two_prod(x, x), x).hi; }
__device__ __forceinline__ double p4(double x) { const dd y = two_prod(x, x); return dd_mul(y, y).hi; }
__device__ __forceinline__ double p5(double x) { const dd y = two_prod(x, x); return dd_mul_d(dd_mul(y, y), x).hi; }

__device__ __forceinline__ double kern_val(const PairTab *T, int ti, int tj, double r) {
  const double s = fabs(r / T->h[ti][tj]); double v = 0.0;
  if (T->kernel == ISPH_KERNEL_WENDLAND) { v = p4(1 - 0.5 * s) * (2 * s + 1.) * (s < 2 ? 1.0 : 0.0); }
  else if (T->kernel == ISPH_KERNEL_CUBIC) { const int b = (int)floor(s); if (b == 0) v = 1.0 - 0.75 * (2 - s) * s * s; else if (b == 1) v = 0.25 * p3(2.0 - s); }
  else { const int b = (int)floor(s); if (b == 0) v += (15.0 * p5(1.0 - s)); if (b >= 0 && b <= 1) v -= (6.0 * p5(2.0 - s)); if (b >= 0 && b <= 2) v += p5(3.0 - s); }
  return v * T->kC[ti][tj];
}
__device__ __forceinline__ double kern_dval(const PairTab *T, int ti, int tj, double r) {
  const double s = fabs(r / T->h[ti][tj]); double v = 0.0;
  if (T->kernel == ISPH_KERNEL_WENDLAND) { v = -5.0 * s * p3(1 - 0.5 * s) * (s < 2 ? 1.0 : 0.0); }
  else if (T->kernel == ISPH_KERNEL_CUBIC) { const int b = (int)floor(s); if (b == 0) v = (2.25 * s - 3) * s; else if (b == 1) v = -0.75 * p2(2 - s); }
  else { const int b = (int)floor(s); if (b == 0) v -= (75.0 * p4(1 - s)); if (b >= 0 && b <= 1) v += (30.0 * p4(2 - s)); if (b >= 0 && b <= 2) v -= (5 * p4(3 - s)); }
  return v * T->kCh[ti][tj];
}

__device__ __forceinline__ double sph_op(bool anti, double fi, double fj) { return anti ? (fi + fj) : (fj - fi); }      // functor.h:9-20
// FilterBinary (filter.h:49-55) ; with ISPH_FILTER_MATCH in m0: FilterMatchBinary (filter.h:101-107), i == kind, j & mask
__device__ __forceinline__ bool fyes1(int m0, int ik) { return (m0 & ISPH_FILTER_MATCH) ? ik == (m0 & 0xff) : (ik & m0) != 0; }
__device__ __forceinline__ bool fyes2(int m0, int m1, int ik, int jk) { return fyes1(m0, ik) && (jk & m1); }

struct Dev {                         // everything a row kernel needs, passed by value
  int n, dim; const long long *slice_off; const int *row_len, *diag_k, *atom, *col; double *val;
  const int *ilist, *type, *kind, *neigh; const long long *noff; const double *x; const PairTab *T;
  const double *vfrac, *Gc, *Lc, *pnd, *normal; double morris_safe;
};

// Morris-Holmes mirror coefficient, mirror_morris_holmes.h:39-53 (r = sqrt(cutsq) at every call site)
__device__ __forceinline__ double mirror_coeff(const Dev &d, bool mh, int i, int j, int ti, int tj) {
  if (!mh) return 1.0;
  const double r = d.T->cut[ti][tj];
  const double xi_i = d.pnd[i] * d.vfrac[i], xi_j = d.pnd[j] * d.vfrac[j];
  const double d_i = 2.0 * r * (xi_i - 0.5) + ISPH_EPS_R, d_j = 2.0 * r * (xi_j - 0.5) + ISPH_EPS_R;
  return (1.0 + d_j / fmax(d_i, d.morris_safe * d.T->h[ti][tj]));
}

#define ROW_SETUP(d) \
  const int row = blockIdx.x * blockDim.x + threadIdx.x; if (row >= (d).n) return; \
  const long long base = (d).slice_off[row >> 5] + (row & 31); const int rlen = (d).row_len[row]; \
  const int i = (d).ilist[row], itype = (d).type[i], ikind = (d).kind[i]; \
  const double xi0 = (d).x[3 * (size_t)i], xi1 = (d).x[3 * (size_t)i + 1], xi2 = (d).x[3 * (size_t)i + 2]; \
  (void)ikind; (void)itype; (void)xi2;

#define PAIR_GEOM(d, DIM) \
  double rij[3] = {0.0, 0.0, 0.0}; double rsq = 0.0; \
  rij[0] = xi0 - (d).x[3 * (size_t)j]; rsq += (rij[0] * rij[0]); \
  rij[1] = xi1 - (d).x[3 * (size_t)j + 1]; rsq += (rij[1] * rij[1]); \
  if (DIM == 3) { rij[2] = xi2 - (d).x[3 * (size_t)j + 2]; rsq += (rij[2] * rij[2]); }

// The pre-computation kernels walk the ORIGINAL neighbor list in its own order with the reference's `rsq < cutsq`
// test, one thread per particle: their sums (and the 6x6 system that the Laplacian correction solves) are then
// accumulated in exactly the reference's order, so vfrac / Gc / Lc agree to the last bits instead of being limited by
// the conditioning of that system.  They need no graph (computePre precedes computeGraph).
#define LIST_SETUP(d) \
  const int row = blockIdx.x * blockDim.x + threadIdx.x; if (row >= (d).n) return; \
  const int i = (d).ilist[row], itype = (d).type[i], ikind = (d).kind[i]; \
  const double xi0 = (d).x[3 * (size_t)i], xi1 = (d).x[3 * (size_t)i + 1], xi2 = (d).x[3 * (size_t)i + 2]; \
  const long long nb = (d).noff[row], ne = (d).noff[row + 1]; (void)ikind; (void)xi2;
#define LIST_FOR(d) for (long long p = nb; p < ne; ++p)
#define LIST_PAIR(d, DIM) \
  const int j = (d).neigh[p] & ISPH_NEIGHMASK, jtype = (d).type[j]; \
  PAIR_GEOM(d, DIM) \
  if (!(rsq < (d).T->cutsq[itype][jtype])) continue;

// ---- pre-computation ---------------------------------------------------------------------------------------------
template <int DIM> __global__ void __launch_bounds__(128) k_volumes(Dev d, double *vfrac_out) {          // functor_volume.h:42-74
  LIST_SETUP(d)
  double wtmp = kern_val(d.T, itype, itype, 0.0);
  LIST_FOR(d) {
    LIST_PAIR(d, DIM)
    wtmp += kern_val(d.T, itype, jtype, sqrt(rsq));
  }
  vfrac_out[i] = 1.0 / wtmp;
}

// closed-form inverse, utils_reference.cpp:251-313 + computeDetDenseMatrix :149-172 ; column-major
template <int DIM> __device__ void invert_small(const double *A, double *B) {
#define A_(r, c) A[(c) * DIM + (r)]
#define B_(r, c) B[(c) * DIM + (r)]
  if (DIM == 2) {
    const double val = (A_(0, 0) * A_(1, 1) - A_(0, 1) * A_(1, 0));
    B_(0, 0) = A_(1, 1) / val; B_(1, 1) = A_(0, 0) / val; B_(1, 0) = -A_(1, 0) / val; B_(0, 1) = -A_(0, 1) / val;
  } else {
    const double val = (A_(0, 0) * A_(1, 1) * A_(2, 2) + A_(1, 0) * A_(2, 1) * A_(0, 2) + A_(2, 0) * A_(0, 1) * A_(1, 2)
                        - A_(2, 0) * A_(1, 1) * A_(0, 2) - A_(0, 0) * A_(2, 1) * A_(1, 2) - A_(1, 0) * A_(0, 1) * A_(2, 2));
    double v0, v1, v2;
    v0 = A_(1, 1) * A_(2, 2) - A_(2, 1) * A_(1, 2); v1 = -A_(1, 0) * A_(2, 2) + A_(2, 0) * A_(1, 2); v2 = A_(1, 0) * A_(2, 1) - A_(2, 0) * A_(1, 1);
    B_(0, 0) = v0 / val; B_(1, 0) = v1 / val; B_(2, 0) = v2 / val;
    v0 = A_(2, 1) * A_(0, 2) - A_(0, 1) * A_(2, 2); v1 = A_(0, 0) * A_(2, 2) - A_(2, 0) * A_(0, 2); v2 = -A_(0, 0) * A_(2, 1) + A_(2, 0) * A_(0, 1);
    B_(0, 1) = v0 / val; B_(1, 1) = v1 / val; B_(2, 1) = v2 / val;
    v0 = A_(0, 1) * A_(1, 2) - A_(1, 1) * A_(0, 2); v1 = -A_(0, 0) * A_(1, 2) + A_(1, 0) * A_(0, 2); v2 = A_(0, 0) * A_(1, 1) - A_(1, 0) * A_(0, 1);
    B_(0, 2) = v0 / val; B_(1, 2) = v1 / val; B_(2, 2) = v2 / val;
  }
#undef A_
#undef B_
}

template <int DIM> __global__ void __launch_bounds__(128) k_gradient_correction(Dev d, double *Gc_out) {   // functor_gradient_correction.h:24-71
  LIST_SETUP(d)
  double G[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  LIST_FOR(d) {
    LIST_PAIR(d, DIM)
    const double r = sqrt(rsq) + ISPH_EPS_R, dwdr = kern_dval(d.T, itype, jtype, r);
#pragma unroll
    for (int k2 = 0; k2 < DIM; ++k2)
#pragma unroll
      for (int k1 = 0; k1 < DIM; ++k1) G[k2 * DIM + k1] -= rij[k1] * rij[k2] * dwdr / r * d.vfrac[j];
  }
  double B[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  invert_small<DIM>(G, B);
  for (int q = 0; q < 9; ++q) Gc_out[9 * (size_t)i + q] = B[q];
}

// LU with partial pivoting (the role of DGESV, utils_reference.cpp:403), one right-hand side
template <int N> __device__ void gesv_small(double *A, double *b) {
  int piv[N];
  for (int k = 0; k < N; ++k) {
    int p = k; double mx = fabs(A[k + k * N]);
    for (int r = k + 1; r < N; ++r) if (fabs(A[r + k * N]) > mx) { mx = fabs(A[r + k * N]); p = r; }
    piv[k] = p;
    if (p != k) for (int c = 0; c < N; ++c) { const double t = A[k + c * N]; A[k + c * N] = A[p + c * N]; A[p + c * N] = t; }
    const double rinv = 1.0 / A[k + k * N];
    for (int r = k + 1; r < N; ++r) A[r + k * N] *= rinv;
    for (int c = k + 1; c < N; ++c) { const double t = A[k + c * N]; for (int r = k + 1; r < N; ++r) A[r + c * N] -= A[r + k * N] * t; }
  }
  for (int k = 0; k < N; ++k) if (piv[k] != k) { const double t = b[k]; b[k] = b[piv[k]]; b[piv[k]] = t; }
  for (int k = 0; k < N; ++k) for (int r = k + 1; r < N; ++r) b[r] -= A[r + k * N] * b[k];
  for (int k = N - 1; k >= 0; --k) { b[k] /= A[k + k * N]; for (int r = 0; r < k; ++r) b[r] -= A[r + k * N] * b[k]; }
}

// register budget: 4 resident blocks per SM (128 registers, 176 B spilled in 3-D) measured against 2 / 3 (no spills) / 5 on the 1M-row brick: 4.78 ms vs 6.55 / 4.85 / 6.13 ms
// (gpurun_out/r2_lc_occ*.json) — the spills cost less than the occupancy they buy
template <int DIM> __global__ void __launch_bounds__(128, 4) k_laplacian_correction(Dev d, double *Lc_out) {  // functor_laplacian_correction.h:25-153
  constexpr int DIMSQ = DIM * DIM, DIML = DIM * (DIM + 1) / 2;
  LIST_SETUP(d)
  double A[DIM * DIMSQ], L[DIML * DIML];
  for (int q = 0; q < DIM * DIMSQ; ++q) A[q] = 0.0;
  for (int q = 0; q < DIML * DIML; ++q) L[q] = 0.0;
  double G[DIMSQ]; for (int q = 0; q < DIMSQ; ++q) G[q] = d.Gc[9 * (size_t)i + q];
  LIST_FOR(d) {                                                                  // :44-84
    LIST_PAIR(d, DIM)
    const double r = sqrt(rsq) + ISPH_EPS_R, dwdr = kern_dval(d.T, itype, jtype, r);
    double aij[3] = {0, 0, 0};
#pragma unroll
    for (int k2 = 0; k2 < DIM; ++k2) {
#pragma unroll
      for (int k1 = 0; k1 < DIM; ++k1) aij[k2] += G[k2 * DIM + k1] * rij[k1];
      aij[k2] *= dwdr / r * d.vfrac[j];
    }
#pragma unroll
    for (int k3 = 0; k3 < DIM; ++k3)
#pragma unroll
      for (int k2 = 0; k2 < DIM; ++k2)
#pragma unroll
        for (int k1 = 0; k1 < (k2 + 1); ++k1) A[k3 * DIMSQ + k2 * DIM + k1] += aij[k3] * rij[k1] * rij[k2];
  }
  LIST_FOR(d) {                                                                  // :86-138
    LIST_PAIR(d, DIM)
    const double r = sqrt(rsq) + ISPH_EPS_R, dwdr = kern_dval(d.T, itype, jtype, r);
    double eij[3] = {0, 0, 0};
#pragma unroll
    for (int q = 0; q < DIM; ++q) eij[q] = (rij[q] / r);
    double Cm[DIMSQ]; for (int q = 0; q < DIMSQ; ++q) Cm[q] = 0.0;
#pragma unroll
    for (int k3 = 0; k3 < DIM; ++k3)
#pragma unroll
      for (int k2 = 0; k2 < DIM; ++k2)
#pragma unroll
        for (int k1 = 0; k1 < (k2 + 1); ++k1) Cm[k2 * DIM + k1] += A[k3 * DIMSQ + k2 * DIM + k1] * eij[k3];
#pragma unroll
    for (int k2 = 0; k2 < DIM; ++k2)
#pragma unroll
      for (int k1 = 0; k1 < (k2 + 1); ++k1) { Cm[k2 * DIM + k1] += rij[k1] * eij[k2]; Cm[k2 * DIM + k1] *= dwdr * d.vfrac[j]; }
    int op = 0;
#pragma unroll
    for (int k4 = 0; k4 < DIM; ++k4)
#pragma unroll
      for (int k3 = 0; k3 < (k4 + 1); ++k3, ++op) {
        int mn = 0;
#pragma unroll
        for (int k2 = 0; k2 < DIM; ++k2)
#pragma unroll
          for (int k1 = 0; k1 < (k2 + 1); ++k1, ++mn) L[op * DIML + mn] += Cm[k2 * DIM + k1] * eij[k3] * eij[k4] * (k3 == k4 ? 1.0 : 2.0);
      }
  }
  double rhs[DIML]; { int op = 0; for (int k2 = 0; k2 < DIM; ++k2) for (int k1 = 0; k1 < (k2 + 1); ++k1, ++op) rhs[op] = -double(k1 == k2); }
  gesv_small<DIML>(L, rhs);
  for (int q = 0; q < 6; ++q) Lc_out[6 * (size_t)i + q] = q < DIML ? rhs[q] : 0.0;
}

// functor_normal.h:56-125 with the filter of one of the two passes of pair_isph_corrected.cpp:404-421
template <int DIM> __global__ void __launch_bounds__(128) k_normals(Dev d, int m0, int m1, double *normal_out, double *pnd_out) {
  LIST_SETUP(d)
  if (!fyes1(m0, ikind)) return;
  double n_i[3] = {0, 0, 0}, pnd_i = 0.0;
  double G[DIM * DIM]; for (int q = 0; q < DIM * DIM; ++q) G[q] = d.Gc[9 * (size_t)i + q];
  const double orient = (ikind == ISPH_KIND_FLUID || ikind == ISPH_KIND_BUFFER_DIRICHLET || ikind == ISPH_KIND_BUFFER_NEUMANN) ? -1.0
                        : ((ikind == ISPH_KIND_SOLID || ikind == ISPH_KIND_BOUNDARY) ? 1.0 : 0.0);   // pair_isph_corrected.cpp:381-386
  LIST_FOR(d) {
    LIST_PAIR(d, DIM)
    const int jkind = d.kind[j];
    const double r = sqrt(rsq) + ISPH_EPS_R;
    if (fyes2(m0, m1, ikind, jkind)) {
      const double dwdr = kern_dval(d.T, itype, jtype, r);
#pragma unroll
      for (int k2 = 0; k2 < DIM; ++k2) {
        double gitmp = 0.0;
#pragma unroll
        for (int k1 = 0; k1 < DIM; ++k1) gitmp += G[k2 * DIM + k1] * rij[k1];
        n_i[k2] += gitmp * orient * dwdr / r * d.vfrac[j];
      }
    } else {
      pnd_i += kern_val(d.T, itype, jtype, r);
    }
  }
  pnd_i += kern_val(d.T, itype, itype, 0.0);
  double alpha = 0.0;
  for (int q = 0; q < DIM; ++q) alpha += n_i[q] * n_i[q];
  alpha = sqrt(alpha);
  if (alpha != 0.0) for (int q = 0; q < DIM; ++q) n_i[q] /= alpha;
  for (int q = 0; q < DIM; ++q) normal_out[3 * (size_t)i + q] = n_i[q];
  pnd_out[i] = pnd_i;
}

// owner -> ghost copy: col_of_atom gives the owning row of every ghost (single GPU); halo columns are filled by halo.cu
__global__ void k_forward(const int *col_of_atom, int nlocal, int nall, int nc, double *f) {
  const int a = nlocal + blockIdx.x * blockDim.x + threadIdx.x; if (a >= nall) return;
  const int o = col_of_atom[a]; if (o < 0 || o >= nlocal) return;
  for (int q = 0; q < nc; ++q) f[(size_t)a * nc + q] = f[(size_t)o * nc + q];
}

// ---- operator rows -------------------------------------------------------------------------------------------
// Corrected::FunctorOuterLaplacianMatrix<Pair,ANTI>[_MorrisHolmes]::operator(), functor_laplacian_matrix.h:72-316
// (iblock < 0, normal == NULL), fused with the PutScalar(0) that precedes it at every call site.
// FAITHFUL: the two neighbour sums every entry of the row depends on (gm = sum_j G e_ij dW V_j op(m_i,m_j), ci = sum_j a_ij e_ij,
// :153-190) are accumulated in the NEIGHBOR LIST's own order, as the reference does, in a pre-pass over the list; the per-entry
// work stays on the (column-ordered) matrix.  Without it the sums are reassociated, which is invisible relative to the row
// (1.5e-15 of the row maximum) but reaches 2e-12 RELATIVE on entries of particles that sit next to the cutoff (|a_ij| ~ 1e-8 of
// the diagonal) on random clouds — measured: tests/test_gpu_parity.py::test_assembly_parity_ragged_cloud, gpurun_out/parity_measured.json.
template <int DIM, bool ANTI, bool FAITHFUL> __global__ void __launch_bounds__(128, 8)
k_laplacian_rows(Dev d, double alpha, const double *material, bool mh, int f0, int f1) {
  ROW_SETUP(d)
  if (!fyes1(f0, ikind)) {                                                        // :88-96 (row left at zero)
    for (int k = 0; k < rlen; ++k) d.val[base + 32ll * k] = 0.0;
    return;
  }
  const double m_i = material ? material[i] : 1.0, vf_i = d.vfrac[i];
  double G[DIM * DIM], L[DIM * (DIM + 1) / 2];
  if (ANTI) { for (int q = 0; q < DIM * DIM; ++q) G[q] = (q % (DIM + 1) == 0) ? 1.0 : 0.0; int op = 0; for (int k2 = 0; k2 < DIM; ++k2) for (int k1 = 0; k1 <= k2; ++k1, ++op) L[op] = (k1 == k2); }
  else { for (int q = 0; q < DIM * DIM; ++q) G[q] = d.Gc[9 * (size_t)i + q]; for (int q = 0; q < DIM * (DIM + 1) / 2; ++q) L[q] = d.Lc[6 * (size_t)i + q]; }
  const bool self_coeff = fyes2(f0, f1, ikind, ikind);
  double gm[3] = {0, 0, 0}, ci[3] = {0, 0, 0}, diag = 0.0;
  int kself = -1;
  if (FAITHFUL) {                                                                 // gm / ci in list order (the reference's summation order)
    const long long nb = d.noff[row], ne = d.noff[row + 1];
    for (long long p = nb; p < ne; ++p) {
      const int j = d.neigh[p] & ISPH_NEIGHMASK, jtype = d.type[j];
      PAIR_GEOM(d, DIM)
      if (!(rsq < d.T->cutsq[itype][jtype])) continue;
      const int jkind = d.kind[j];
      const double m_j = material ? material[j] : 1.0;
      const double r = sqrt(rsq) + ISPH_EPS_R, dwdr = kern_dval(d.T, itype, jtype, r);
      double eij[3] = {0, 0, 0};
#pragma unroll
      for (int q = 0; q < DIM; ++q) eij[q] = rij[q] / r;
      const double vf = (ANTI ? sqrt(vf_i * d.vfrac[j]) : d.vfrac[j]), vjtmp = dwdr * vf;
#pragma unroll
      for (int k2 = 0; k2 < DIM; ++k2) {
        double gitmp = 0.0;
#pragma unroll
        for (int k1 = 0; k1 < DIM; ++k1) gitmp += G[k2 * DIM + k1] * eij[k1];
        const double ijtmp = gitmp * vjtmp;
        if (ikind & jkind) gm[k2] += ijtmp * (sph_op(ANTI, m_i, m_j));
      }
      if (!ANTI) {
        double aij = 0.0; int op = 0;
#pragma unroll
        for (int k2 = 0; k2 < DIM; ++k2)
#pragma unroll
          for (int k1 = 0; k1 < (k2 + 1); ++k1, ++op) aij += L[op] * eij[k1] * eij[k2] * (k1 == k2 ? 1.0 : 2.0);
        aij *= 2.0 * dwdr * vf;
#pragma unroll
        for (int q = 0; q < DIM; ++q) ci[q] += aij * eij[q];
      }
    }
  }
  for (int k = 0; k < rlen; ++k) {                                                // pass 1, :127-195
    const int j = d.atom[base + 32ll * k];
    if (j == i) { kself = k; continue; }
    const int jtype = d.type[j], jkind = d.kind[j];
    const double m_j = material ? material[j] : 1.0;
    PAIR_GEOM(d, DIM)
    double coeff = self_coeff;
    if (!(ikind & ISPH_KIND_SOLID) && (jkind & ISPH_KIND_SOLID)) coeff = (fyes2(f0, f1, ikind, jkind) ? mirror_coeff(d, mh, i, j, itype, jtype) : 0.0);
    const double r = sqrt(rsq) + ISPH_EPS_R, dwdr = kern_dval(d.T, itype, jtype, r);
    double eij[3] = {0, 0, 0};
#pragma unroll
    for (int q = 0; q < DIM; ++q) eij[q] = rij[q] / r;
    const double vf = (ANTI ? sqrt(vf_i * d.vfrac[j]) : d.vfrac[j]), vjtmp = dwdr * vf;
#pragma unroll
    for (int k2 = 0; k2 < DIM; ++k2) {
      double gitmp = 0.0;
#pragma unroll
      for (int k1 = 0; k1 < DIM; ++k1) gitmp += G[k2 * DIM + k1] * eij[k1];
      const double ijtmp = gitmp * vjtmp;
      if (!FAITHFUL && (ikind & jkind)) gm[k2] += ijtmp * (sph_op(ANTI, m_i, m_j));
    }
    double aij = 0.0; int op = 0;
#pragma unroll
    for (int k2 = 0; k2 < DIM; ++k2)
#pragma unroll
      for (int k1 = 0; k1 < (k2 + 1); ++k1, ++op) aij += L[op] * eij[k1] * eij[k2] * (k1 == k2 ? 1.0 : 2.0);
    aij *= 2.0 * dwdr * vf;
    if (!ANTI && !FAITHFUL) {
#pragma unroll
      for (int q = 0; q < DIM; ++q) ci[q] += aij * eij[q];
    }
    aij *= m_i * coeff / r;
    d.val[base + 32ll * k] = -aij; diag += aij;
  }
  double diag2 = 0.0;
  for (int k = 0; k < rlen; ++k) {                                                // pass 2, :211-259, then scale by alpha :267
    const int j = d.atom[base + 32ll * k];
    if (j == i) continue;
    const int jtype = d.type[j], jkind = d.kind[j];
    PAIR_GEOM(d, DIM)
    double coeff = self_coeff;
    if (!(ikind & ISPH_KIND_SOLID) && (jkind & ISPH_KIND_SOLID)) coeff = fyes2(f0, f1, ikind, jkind);
    const double r = sqrt(rsq) + ISPH_EPS_R, dwdr = kern_dval(d.T, itype, jtype, r);
    const double vf = (ANTI ? sqrt(vf_i * d.vfrac[j]) : d.vfrac[j]), vjtmp = dwdr * vf;
    double eij[3] = {0, 0, 0};
#pragma unroll
    for (int q = 0; q < DIM; ++q) eij[q] = rij[q] / r;
    double bij[3] = {0, 0, 0};
#pragma unroll
    for (int k2 = 0; k2 < DIM; ++k2)
#pragma unroll
      for (int k1 = 0; k1 < DIM; ++k1) bij[k2] += G[k2 * DIM + k1] * eij[k1];
    double dot_c = 0.0, dot_g = 0.0;
#pragma unroll
    for (int q = 0; q < DIM; ++q) dot_c += bij[q] * ci[q];
#pragma unroll
    for (int q = 0; q < DIM; ++q) dot_g += bij[q] * gm[q];
    const double tmp = coeff * (m_i * dot_c * vjtmp - dot_g * vjtmp);
    d.val[base + 32ll * k] = (d.val[base + 32ll * k] - tmp) * alpha;
    diag2 += tmp;
  }
  if (kself >= 0) d.val[base + 32ll * kself] = (diag + diag2) * alpha;
}

// Corrected::FunctorOuterGradientOperator (functor_gradient_operator.h:89-170) + FunctorOuterGradientDotOperatorMatrix
// (functor_gradient_dot_operator_matrix.h:36-79): rows alpha * sum_k vec_i^k (G_i r_ij)^k dW/dr / r V_j, self = -sum; SumInto
template <int DIM> __global__ void __launch_bounds__(128) k_gradient_dot_rows(Dev d, double alpha, const double *vec, int f0, int f1) {
  ROW_SETUP(d)
  if (!fyes1(f0, ikind)) return;
  double G[DIM * DIM]; for (int q = 0; q < DIM * DIM; ++q) G[q] = d.Gc[9 * (size_t)i + q];
  double self[3] = {0, 0, 0}, vi[3] = {0, 0, 0};
  for (int q = 0; q < DIM; ++q) vi[q] = vec[3 * (size_t)i + q];
  int kself = -1;
  for (int k = 0; k < rlen; ++k) {
    const int j = d.atom[base + 32ll * k];
    if (j == i) { kself = k; continue; }
    const int jtype = d.type[j], jkind = d.kind[j];
    if (!fyes2(f0, f1, ikind, jkind)) continue;
    PAIR_GEOM(d, DIM)
    const double r = sqrt(rsq) + ISPH_EPS_R, dwdr = kern_dval(d.T, itype, jtype, r);
    const double vjtmp = dwdr / r * d.vfrac[j] * 1.0;                           // MirrorNothing is bound (pair_isph_corrected.cpp:172,177)
    double s = 0.0;
#pragma unroll
    for (int k2 = 0; k2 < DIM; ++k2) {
      double gitmp = 0.0;
#pragma unroll
      for (int k1 = 0; k1 < DIM; ++k1) gitmp += G[k2 * DIM + k1] * rij[k1];
      const double ijtmp = gitmp * vjtmp;
      self[k2] -= ijtmp;
      s += (ijtmp * alpha) * vi[k2];
    }
    d.val[base + 32ll * k] += s;
  }
  if (kself >= 0) { double s = 0.0; for (int q = 0; q < DIM; ++q) s += (self[q] * alpha) * vi[q]; d.val[base + 32ll * kself] += s; }
}

// matrix-free gradient-like loop shared by the divergence (functor_divergence.h:55-124) and the gradient
// (functor_gradient.h:80-169): calls body(j, k2, gitmp, vjtmp)
template <int DIM, bool ANTI, class Body> __device__ __forceinline__ void
grad_like_loop(const Dev &d, int i, int itype, int ikind, double xi0, double xi1, double xi2, long long base, int rlen,
               bool mh, int f0, int f1, Body body) {
  if (!fyes1(f0, ikind)) return;
  double G[DIM * DIM];
  if (ANTI) { for (int q = 0; q < DIM * DIM; ++q) G[q] = (q % (DIM + 1) == 0) ? 1.0 : 0.0; }
  else { for (int q = 0; q < DIM * DIM; ++q) G[q] = d.Gc[9 * (size_t)i + q]; }
  const double vf_i = d.vfrac[i];
  for (int k = 0; k < rlen; ++k) {
    const int j = d.atom[base + 32ll * k]; if (j == i) continue;
    const int jtype = d.type[j], jkind = d.kind[j];
    if (!fyes2(f0, f1, ikind, jkind)) continue;
    PAIR_GEOM(d, DIM)
    double coeff = 1.0;
    if (!(ikind & ISPH_KIND_SOLID) && (jkind & ISPH_KIND_SOLID)) coeff = mirror_coeff(d, mh, i, j, itype, jtype);
    const double r = sqrt(rsq) + ISPH_EPS_R, dwdr = kern_dval(d.T, itype, jtype, r);
    const double vf = (ANTI ? sqrt(vf_i * d.vfrac[j]) : d.vfrac[j]), vjtmp = dwdr / r * vf * coeff;
#pragma unroll
    for (int k2 = 0; k2 < DIM; ++k2) {
      double gitmp = 0.0;
#pragma unroll
      for (int k1 = 0; k1 < DIM; ++k1) gitmp += G[k2 * DIM + k1] * rij[k1];
      body(j, k2, gitmp, vjtmp);
    }
  }
}

// FunctorOuterIncompNavierStokesPoisson::operator() + exitFor, functor_incomp_navier_stokes_poisson.h:111,127-181,
// and PairISPH::modifySingularMatrix pair_isph.cpp:493-520: sld = diag(A); per-row diagonal / rhs; ReplaceDiagonalValues
template <int DIM, bool ANTI> __global__ void __launch_bounds__(128)
k_poisson_rows(Dev d, bool mh, bool neumann, int singular, int first_fluid_row, const double *vstar, double *diagonal, double *sld, double *b) {
  ROW_SETUP(d)
  const int kd = d.diag_k[row];
  const double s = kd < 0 ? 0.0 : d.val[base + 32ll * kd];
  sld[row] = s;
  double diag = diagonal[i]; bool assigned = false;
  if (ikind == ISPH_KIND_SOLID) {
    if (neumann) { double nn = 0.0; for (int q = 0; q < DIM; ++q) nn += d.normal[3 * (size_t)i + q] * d.normal[3 * (size_t)i + q]; if (nn < 0.5) { diag = 1.0; assigned = true; } }
    else { diag = 1.0; assigned = true; }
    b[i] = 0.0;
  } else {
    diag = s; assigned = true;
    double div = 0.0;
    grad_like_loop<DIM, ANTI>(d, i, itype, ikind, xi0, xi1, xi2, base, rlen, mh, ISPH_KIND_FLUID, ISPH_KIND_ALL,
      [&](int j, int k2, double gitmp, double vjtmp) { div += gitmp * (sph_op(ANTI, vstar[3 * (size_t)i + k2], vstar[3 * (size_t)j + k2])) * vjtmp; });
    div *= 1.0;
    double bi = -div;
    if (row == first_fluid_row) {
      if (singular == ISPH_PINZERO) { for (int k = 0; k < rlen; ++k) d.val[base + 32ll * k] = 0.0; diag = -1.0; bi = 0.0; }
      else if (singular == ISPH_DOUBLEDIAG) diag *= 1.5;
    }
    b[i] = bi;
  }
  (void)assigned;
  diagonal[i] = diag;
  if (kd >= 0) d.val[base + 32ll * kd] = diag;
}

// FunctorOuterIncompNavierStokesHelmholtz::operator() + exitFor, functor_incomp_navier_stokes_helmholtz.h:95,108-159
template <int DIM, bool ANTI> __global__ void __launch_bounds__(128)
k_helmholtz_rows(Dev d, double dt, bool incp, double g0, double g1, double g2, const double *rho, const double *force,
                 const double *pressure, const double *w, int ld, double *diagonal, double *sld, double *b) {
  ROW_SETUP(d)
  const int kd = d.diag_k[row];
  const double s = kd < 0 ? 0.0 : d.val[base + 32ll * kd];
  sld[row] = s;
  double diag;
  if (ikind == ISPH_KIND_SOLID) { diag = 1.0; }
  else {
    diag = 1.0 + s;
    const double g[3] = {g0, g1, g2};
    double bk[3];
#pragma unroll
    for (int q = 0; q < DIM; ++q) { bk[q] = b[(size_t)q * ld + i]; bk[q] += w[(size_t)q * ld + i]; bk[q] += dt * (force[3 * (size_t)i + q] / rho[i] + g[q]); }
    if (incp) {
      double gp[3] = {0, 0, 0};
      grad_like_loop<DIM, ANTI>(d, i, itype, ikind, xi0, xi1, xi2, base, rlen, false, ISPH_KIND_FLUID, ISPH_KIND_FLUID,
        [&](int j, int k2, double gitmp, double vjtmp) { const double ijtmp = gitmp * vjtmp; gp[k2] += ijtmp * (sph_op(ANTI, pressure[i], pressure[j])); });
#pragma unroll
      for (int q = 0; q < DIM; ++q) { gp[q] *= 1.0; bk[q] += dt * (-1.0 / rho[i] * gp[q]); }
    }
#pragma unroll
    for (int q = 0; q < DIM; ++q) b[(size_t)q * ld + i] = bk[q];
  }
  diagonal[i] = diag;
  if (kd >= 0) d.val[base + 32ll * kd] = diag;
}

// FunctorOuterPoissonBoltzmannJacobian::operator() + exitFor, functor_poisson_boltzmann_jacobian.h:68-107
__global__ void k_pb_rows(Dev d, bool linearized, double kappasq, double gamma, const double *psi_, double *diagonal, const double *sld) {
  ROW_SETUP(d)
  (void)rlen;
  double diag = diagonal[i];
  if (ikind == ISPH_KIND_SOLID || ikind == ISPH_KIND_BOUNDARY) diag = -1.0;
  else if (ikind == ISPH_KIND_BUFFER_DIRICHLET || ikind == ISPH_KIND_BUFFER_NEUMANN || ikind == ISPH_KIND_FLUID) {
    diag = sld[row]; const double psi = psi_[i];
    if (linearized) {
      const double numerator = 4.0 - 2.0 * gamma * pow(psi, 2.0), denominator = pow(gamma, 2.0) * pow(psi, 4.0) + 4.0 * gamma * pow(psi, 2.0) + 4.0;
      diag += kappasq * (numerator / denominator);
    } else {
      const double numerator = 2.0 * gamma * cosh(0.5 * psi) * sinh(0.5 * psi) * sinh(psi), denominator = 2.0 * gamma * pow(sinh(0.5 * psi), 2.0) + 1.0;
      diag += kappasq * (cosh(psi) / denominator - numerator / pow(denominator, 2.0));
    }
  }
  diagonal[i] = diag;
  const int kd = d.diag_k[row]; if (kd >= 0) d.val[base + 32ll * kd] = diag;
}

// FunctorOuterAppliedElectricPotential::operator() + exitFor, functor_applied_electric_potential.h:64-96
__global__ void k_aep_rows(Dev d, const double *phi, double *diagonal, double *b) {
  ROW_SETUP(d)
  (void)rlen;
  const int kd = d.diag_k[row];
  double diag = kd < 0 ? 0.0 : d.val[base + 32ll * kd];                          // ExtractDiagonalCopy in enterFor, :58
  double bi = 0.0;
  if (ikind == ISPH_KIND_SOLID) diag = 1.0;
  else if (ikind == ISPH_KIND_BUFFER_NEUMANN || ikind == ISPH_KIND_BUFFER_DIRICHLET) { diag = 1.0; bi = phi[i]; }
  b[i] = bi; diagonal[i] = diag;
  if (kd >= 0) d.val[base + 32ll * kd] = diag;
}
// FunctorOuterSoluteTransport::operator() + exitFor, functor_solute_transport.h:100-134 (after enterFor's Scale(-theta))
__global__ void k_solute_rows(Dev d, const double *w, double *diagonal, double *sld, double *b, int *bad_kind) {
  ROW_SETUP(d)
  (void)rlen;
  const int kd = d.diag_k[row];
  const double s = kd < 0 ? 0.0 : d.val[base + 32ll * kd];
  sld[row] = s;
  double diag = diagonal[i];
  if (ikind == ISPH_KIND_BUFFER_DIRICHLET || ikind == ISPH_KIND_BUFFER_NEUMANN || ikind == ISPH_KIND_SOLID) diag = 1.0;
  else if (ikind == ISPH_KIND_FLUID) { diag = 1.0 + s; b[i] += w[i]; }
  else *bad_kind = 1;                                                             // :123-124 "Particle types are not supported"
  diagonal[i] = diag;
  if (kd >= 0) d.val[base + 32ll * kd] = diag;
}

// FunctorOuterPoissonBoltzmannF::operator() (functor_poisson_boltzmann_f.h:58-88) on top of the matrix-free corrected
// Laplacian Corrected::FunctorOuterLaplacianHelper::operator() (functor_laplacian.h:67-277; scalar field, alpha = -1,
// material = eps, filter (Fluid, All)).  The reference's Jacobian is NOT the exact derivative of this residual next to
// solids (the matrix functor weights the mirror coefficient and the material gradient differently), so the residual cannot
// be taken from the assembled matrix: the three neighbor passes are done here, in list order and with the reference's
// operation order (a_ij is recomputed in the third pass instead of being stored per neighbor).
template <int DIM> __global__ void __launch_bounds__(128)
k_pb_residual(Dev d, bool mh, bool linearized, double kappasq, double gamma, const double *psi, const double *psi0, const double *eps,
              const double *extra, double *f) {
  LIST_SETUP(d)
  const int f0 = ISPH_KIND_FLUID, f1 = ISPH_KIND_ALL;
  if (ikind == ISPH_KIND_SOLID || ikind == ISPH_KIND_BOUNDARY) f[i] = (-psi[i] + psi0[i]);
  else if (ikind == ISPH_KIND_BUFFER_DIRICHLET || ikind == ISPH_KIND_BUFFER_NEUMANN || ikind == ISPH_KIND_FLUID) {
    const double m_i = eps[i], psi_i = psi[i];
    double gm[3] = {0.0, 0.0, 0.0}, gf[3] = {0.0, 0.0, 0.0}, lap = 0.0;
    if (ikind & f0) {                                                                             // functor_laplacian.h:99-100
      const double *G = d.Gc + 9 * (size_t)i, *L = d.Lc + 6 * (size_t)i;
      LIST_FOR(d) {                                                                               // :111-158 corrected gradients of psi and eps
        const int jk = d.kind[d.neigh[p] & ISPH_NEIGHMASK];
        if (!((ikind & f0) && (jk & f1))) continue;
        LIST_PAIR(d, DIM)
        double coeff = 1.0;
        if (jk & ISPH_KIND_SOLID) coeff = mirror_coeff(d, mh, i, j, itype, jtype);
        const double r = sqrt(rsq) + ISPH_EPS_R, dwdr = kern_dval(d.T, itype, jtype, r);
        const double vjtmp = dwdr / r * d.vfrac[j] * coeff;
#pragma unroll
        for (int k2 = 0; k2 < DIM; ++k2) {
          double gitmp = 0.0;
#pragma unroll
          for (int k1 = 0; k1 < DIM; ++k1) gitmp += G[k2 * DIM + k1] * rij[k1];
          const double ijtmp = gitmp * vjtmp;
          if (jk & f0) gm[k2] += ijtmp * (eps[j] - m_i);
          gf[k2] += ijtmp * (psi[j] - psi_i);
        }
      }
      LIST_FOR(d) {                                                                               // :161-216 (no filter on this pass)
        LIST_PAIR(d, DIM)
        const double r = sqrt(rsq) + ISPH_EPS_R, dwdr = kern_dval(d.T, itype, jtype, r);
        double eij[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int k = 0; k < DIM; ++k) eij[k] = rij[k] / r;
        double aij = 0.0;
#pragma unroll
        for (int k2 = 0, op = 0; k2 < DIM; ++k2)
#pragma unroll
          for (int k1 = 0; k1 < (k2 + 1); ++k1, ++op) aij += L[op] * eij[k1] * eij[k2] * (k1 == k2 ? 1.0 : 2.0);
        aij *= 2.0 * dwdr * d.vfrac[j];
        double dotv = 0.0;
#pragma unroll
        for (int k = 0; k < DIM; ++k) dotv += eij[k] * gf[k];
        lap += aij * (-1.0 * dotv);
      }
      LIST_FOR(d) {                                                                               // :219-263
        const int jk = d.kind[d.neigh[p] & ISPH_NEIGHMASK];
        if (!((ikind & f0) && (jk & f1))) continue;
        LIST_PAIR(d, DIM)
        double coeff = 1.0;
        if (jk & ISPH_KIND_SOLID) coeff = mirror_coeff(d, mh, i, j, itype, jtype);
        const double r = sqrt(rsq) + ISPH_EPS_R, dwdr = kern_dval(d.T, itype, jtype, r);
        double eij[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int k = 0; k < DIM; ++k) eij[k] = rij[k] / r;
        double aij = 0.0;
#pragma unroll
        for (int k2 = 0, op = 0; k2 < DIM; ++k2)
#pragma unroll
          for (int k1 = 0; k1 < (k2 + 1); ++k1, ++op) aij += L[op] * eij[k1] * eij[k2] * (k1 == k2 ? 1.0 : 2.0);
        aij *= 2.0 * dwdr * d.vfrac[j];
        lap += aij * (coeff * (psi_i - psi[j]) / r);
      }
    }
    double dotm = 0.0;
#pragma unroll
    for (int k = 0; k < DIM; ++k) dotm += gm[k] * gf[k];
    double out = -1.0 * (m_i * lap + dotm);                                                       // :266-268 with alpha = -1
    if (linearized) out += kappasq * (psi_i / (1.0 + 2.0 * gamma * pow(psi_i / 2, 2.0)));
    else out += kappasq * (sinh(psi_i) / (1.0 + 2.0 * gamma * pow(sinh(psi_i / 2.0), 2.0)));
    f[i] = out;
  }
  if (extra && !(ikind & ISPH_KIND_SOLID)) f[i] += extra[i];                                      // functor_poisson_boltzmann_extra_f.h:76-90
}

// ---- the step right after the Poisson solve (pair_isph.cpp:1017-1031) -------------------------------------------
// computeZeroMeanPressure, pair_isph.cpp:422-464: solid rows are cleaned to 0, the mean over the other owned rows is removed
__global__ void __launch_bounds__(256) k_dp_sum(const int *kind, int nlocal, double *dp, double *partials) {
  double s = 0.0, cnt = 0.0;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < nlocal; i += gridDim.x * 256) { if (kind[i] == ISPH_KIND_SOLID) dp[i] = 0.0; else { s += dp[i]; cnt += 1.0; } }
  __shared__ double sm[2][8];
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
  if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = s; sm[1][threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x < 2) { double t = 0.0; for (int w = 0; w < 8; ++w) t += sm[threadIdx.x][w]; partials[2 * blockIdx.x + threadIdx.x] = t; }
}
__global__ void k_dp_sum_final(const double *partials, int nblocks, double *out) {     // fixed order: deterministic
  if (threadIdx.x < 2) { double t = 0.0; for (int b = 0; b < nblocks; ++b) t += partials[2 * b + threadIdx.x]; out[threadIdx.x] = t; }
}
__global__ void k_dp_sub_mean(const int *kind, int nall, double *dp, const double *sum_cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= nall) return;
  const double mean_val = sum_cnt[0] / sum_cnt[1];
  dp[i] -= mean_val * (kind[i] != ISPH_KIND_SOLID ? 1.0 : 0.0);
}
// FunctorOuterCorrectVelocity::operator(), functor_correct_velocity.h:52-69
template <int DIM, bool ANTI> __global__ void __launch_bounds__(128)
k_correct_velocity(Dev d, double dt, const double *rho, const double *dp, double *vstar) {
  ROW_SETUP(d)
  if (!fyes1(ISPH_KIND_FLUID, ikind)) return;
  double g[3] = {0, 0, 0};
  grad_like_loop<DIM, ANTI>(d, i, itype, ikind, xi0, xi1, xi2, base, rlen, false, ISPH_KIND_FLUID, ISPH_KIND_FLUID,
    [&](int j, int k2, double gitmp, double vjtmp) { const double ijtmp = gitmp * vjtmp; g[k2] += ijtmp * (sph_op(ANTI, dp[i], dp[j])); });
#pragma unroll
  for (int q = 0; q < DIM; ++q) { g[q] *= 1.0; vstar[3 * (size_t)i + q] -= dt / rho[i] * g[q]; }
}
// FunctorOuterCorrectPressure::operator(), functor_correct_pressure.h:29-43 (owned and ghost particles)
__global__ void k_correct_pressure(double *p, const double *dp, int nall, int incp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= nall) return;
  if (incp) p[i] += dp[i]; else p[i] = dp[i];
}


// ---- advanceTime (PairISPH_Corrected::advanceTime, pair_isph_corrected.cpp:1183-1194) -------------------------
// FunctorOuterAdvanceTimeBegin::operator(), functor_advance_time_begin.h:52-72: dp_i = grad(p)_i . 0.5 dt (v^{n+1}_i + v^n_i) on rows of
// fluid kind (0 elsewhere); the gradient functor carries FilterBinary(Fluid, Fluid) and alpha = 1
template <int DIM, bool ANTI> __global__ void __launch_bounds__(128)
k_advance_begin(Dev d, double dt, const double *v, const double *vnp1, const double *p, double *dp) {
  ROW_SETUP(d)
  double dx[3] = {0, 0, 0}, g[3] = {0, 0, 0};
#pragma unroll
  for (int q = 0; q < DIM; ++q) dx[q] = 0.5 * dt * (vnp1[3 * (size_t)i + q] + v[3 * (size_t)i + q]);
  grad_like_loop<DIM, ANTI>(d, i, itype, ikind, xi0, xi1, xi2, base, rlen, false, ISPH_KIND_FLUID, ISPH_KIND_FLUID,
    [&](int j, int k2, double gitmp, double vjtmp) { const double ijtmp = gitmp * vjtmp; g[k2] += ijtmp * (sph_op(ANTI, p[i], p[j])); });
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < DIM; ++q) { g[q] *= 1.0; s += g[q] * dx[q]; }
  dp[i] = fyes1(ISPH_KIND_FLUID, ikind) ? s : 0.0;
}
// FunctorOuterAdvanceTimeEnd::operator() over owned AND ghost atoms, functor_advance_time_end.h:48-66
struct FixedTab { int f[ISPH_MAXT]; };
__global__ void k_advance_end(int nall, int dim, double dt, const int *type, FixedTab fixed, double *v, const double *vnp1, double *p, const double *dp, double *x) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= nall) return;
  if (fixed.f[type[i]]) { for (int k = 0; k < dim; ++k) v[3 * (size_t)i + k] = vnp1[3 * (size_t)i + k]; return; }
  p[i] += dp[i];
  for (int k = 0; k < dim; ++k) {
    const double delta = 0.5 * dt * (vnp1[3 * (size_t)i + k] + v[3 * (size_t)i + k]);
    x[3 * (size_t)i + k] += delta; v[3 * (size_t)i + k] = vnp1[3 * (size_t)i + k];
  }
}

// ---- boundary-condition row modifiers of computeHelmholtz (pair_isph_corrected.cpp:918-934) --------------------
// Corrected::FunctorOuterBoundaryNavierSlip::operator(), functor_boundary_navier_slip.h:54-174 (iblock < 0, add_neumann_term): for
// rows of fluid / buffer kind, robin_j = beta dW/dr / r V_j / rho_i (n_i + n_j) . (G_i r_ij) is summed into the entries of the
// solid neighbours and -sum_j robin_j into the diagonal (SumIntoGlobalValues)
template <int DIM> __global__ void __launch_bounds__(128) k_navier_slip_rows(Dev d, double beta, const double *rho, int *bad_kind) {
  ROW_SETUP(d)
  if (ikind == ISPH_KIND_SOLID) return;
  if (!(ikind == ISPH_KIND_FLUID || ikind == ISPH_KIND_BUFFER_DIRICHLET || ikind == ISPH_KIND_BUFFER_NEUMANN)) { *bad_kind = 1; return; }
  double G[DIM * DIM]; for (int q = 0; q < DIM * DIM; ++q) G[q] = d.Gc[9 * (size_t)i + q];
  double robin_at_i = 0.0; int kself = -1;
  for (int k = 0; k < rlen; ++k) {
    const int j = d.atom[base + 32ll * k];
    if (j == i) { kself = k; continue; }
    const int jtype = d.type[j];
    if (d.kind[j] != ISPH_KIND_SOLID) continue;
    PAIR_GEOM(d, DIM)
    const double r = sqrt(rsq) + ISPH_EPS_R, dwdr = kern_dval(d.T, itype, jtype, r);
    double tmp = 0.0;
#pragma unroll
    for (int k2 = 0; k2 < DIM; ++k2) {
      double a = 0.0;
#pragma unroll
      for (int k1 = 0; k1 < DIM; ++k1) a += G[k2 * DIM + k1] * rij[k1];
      tmp += (d.normal[3 * (size_t)i + k2] + d.normal[3 * (size_t)j + k2]) * a;
    }
    const double robin_at_j = beta * dwdr / r * d.vfrac[j] / rho[i] * tmp;
    d.val[base + 32ll * k] += robin_at_j;
    robin_at_i -= robin_at_j;
  }
  if (kself >= 0) d.val[base + 32ll * kself] += robin_at_i;
}
// Corrected::FunctorOuterBoundaryDirichlet::operator(), functor_boundary_dirichlet.h:47-150: a fluid row with a solid particle within h is
// REPLACED by the least-squares extrapolation stencil along its normal over the neighbours within h (zero for the other in-cut
// neighbours, 1 on the diagonal) and its right-hand side entries are zeroed.  The two sums (mean normal coordinate, xterm) are taken
// over the neighbor list in its own order, as the reference does; the per-entry values are written on the matrix.
template <int DIM> __global__ void __launch_bounds__(128) k_dirichlet_rows(Dev d, double *b, int ldb, int *bad_kind) {
  ROW_SETUP(d)
  if (ikind == ISPH_KIND_SOLID) return;
  if (ikind != ISPH_KIND_FLUID) { *bad_kind = 1; return; }
  const long long nb = d.noff[row], ne = d.noff[row + 1];
  int n_solid = 0, natoms_cut = 0; double xn_av = 0.0;
  const double *n_i = d.normal + 3 * (size_t)i;
  for (long long p = nb; p < ne; ++p) {
    const int j = d.neigh[p] & ISPH_NEIGHMASK, jtype = d.type[j];
    PAIR_GEOM(d, DIM)
    const double hij = d.T->h[itype][jtype];
    if (rsq < hij * hij) {                                                        // pow(h, 2): exactly h * h
      n_solid += (d.kind[j] == ISPH_KIND_SOLID);
      double xn_j = 0.0;
      for (int k = 0; k < DIM; ++k) xn_j += d.x[3 * (size_t)j + k] * n_i[k];
      xn_av += xn_j; ++natoms_cut;
    }
  }
  if (n_solid <= 0) return;
  double xn_i = 0.0;
  xn_i += xi0 * n_i[0]; xn_i += xi1 * n_i[1]; if (DIM == 3) xn_i += xi2 * n_i[2];
  xn_av /= natoms_cut;
  double xterm = 0.0;
  for (long long p = nb; p < ne; ++p) {
    const int j = d.neigh[p] & ISPH_NEIGHMASK, jtype = d.type[j];
    PAIR_GEOM(d, DIM)
    const double hij = d.T->h[itype][jtype];
    if (rsq < hij * hij) { double xn_j = 0.0; for (int k = 0; k < DIM; ++k) xn_j += d.x[3 * (size_t)j + k] * n_i[k]; xterm += xn_j * (xn_j - xn_av); }
  }
  for (int k = 0; k < rlen; ++k) {
    const int j = d.atom[base + 32ll * k];
    if (j == i) { d.val[base + 32ll * k] = 1.0; continue; }
    const int jtype = d.type[j];
    PAIR_GEOM(d, DIM)
    const double hij = d.T->h[itype][jtype];
    double v = 0.0;
    if (rsq < hij * hij) { double xn_j = 0.0; for (int kk = 0; kk < DIM; ++kk) xn_j += d.x[3 * (size_t)j + kk] * n_i[kk]; v = -(xn_i - xn_av) * (xn_j - xn_av) / xterm - 1.0 / natoms_cut; }
    d.val[base + 32ll * k] = v;
  }
  for (int k = 0; k < DIM; ++k) b[(size_t)k * ldb + i] = 0.0;
}

__global__ void k_recip(const double *a, double *o, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) o[i] = 1.0 / a[i]; }
__global__ void k_mul(const double *a, const double *b, double *o, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) o[i] = a[i] * b[i]; }
__global__ void k_scale_vec(double *a, double s, int n, int ld, int nvec) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) for (int q = 0; q < nvec; ++q) a[(size_t)q * ld + i] *= s; }

// ---- host drivers ------------------------------------------------------------------------------------------------
static Dev make_dev(Ctx *c, bool need_graph = true) {
  Matrix &A = c->A;
  ISPH_REQUIRE(c->have_pair && c->have_atoms && c->have_neigh, "pair_coeff, atoms and neighbors must be set first");
  if (need_graph && (!A.built || A.external)) { ISPH_REQUIRE(!A.external, "this operation needs a matrix built from a neighbor list"); graph_build(c); }
  Dev d; d.n = need_graph ? A.n : c->inum; d.neigh = c->neigh.p; d.noff = c->noff.p; d.dim = c->tab.dim; d.slice_off = A.slice_off.p; d.row_len = A.row_len.p; d.diag_k = A.diag_k.p; d.atom = A.atom.p; d.col = A.col.p; d.val = A.val.p;
  d.ilist = c->ilist.p; d.type = c->type.p; d.kind = c->kind.p; d.x = c->x.p; d.T = c->d_tab.p;
  d.vfrac = c->field[ISPH_F_VFRAC].p; d.Gc = c->field[ISPH_F_GC].p; d.Lc = c->field[ISPH_F_LC].p; d.pnd = c->field[ISPH_F_PND].p; d.normal = c->field[ISPH_F_NORMAL].p;
  d.morris_safe = c->tab.morris_safe;
  return d;
}
#define GRID(c) ceil_div((c)->A.n, 128), 128, 0, (c)->stream
#define LGRID(c) ceil_div((c)->inum, 128), 128, 0, (c)->stream

void forward_comm(Ctx *c, int field) {
  static const int nc[ISPH_F_COUNT] = {1, 9, 6, 3, 1, 1, 1, 1, 3, 3, 3, 1, 1, 1, 1, 1, 1};
  if (c->nghost == 0) return;
  if (c->nranks > 1) halo_forward_field(c, field, nc[field]);      // ghosts owned by other ranks
  k_forward<<<ceil_div(c->nghost, 256), 256, 0, c->stream>>>(c->col_of_atom.p, c->nlocal, c->nall, nc[field], c->field[field].p); ++c->launches;
}

void compute_volumes(Ctx *c) {
  Dev d = make_dev(c, false); c->tic("computeVolumes");
  if (d.dim == 2) k_volumes<2><<<LGRID(c)>>>(d, c->field[ISPH_F_VFRAC].p); else k_volumes<3><<<LGRID(c)>>>(d, c->field[ISPH_F_VFRAC].p);
  ++c->launches;
  forward_comm(c, ISPH_F_VFRAC);                                                   // functor_volume.h:76-81
  c->toc("computeVolumes");
}
void compute_gradient_correction(Ctx *c) {
  Dev d = make_dev(c, false); c->tic("computeGradientCorrection");
  if (d.dim == 2) k_gradient_correction<2><<<LGRID(c)>>>(d, c->field[ISPH_F_GC].p); else k_gradient_correction<3><<<LGRID(c)>>>(d, c->field[ISPH_F_GC].p);
  ++c->launches; c->toc("computeGradientCorrection");
}
void compute_laplacian_correction(Ctx *c) {
  Dev d = make_dev(c, false); c->tic("computeLaplacianCorrection");
  if (d.dim == 2) k_laplacian_correction<2><<<LGRID(c)>>>(d, c->field[ISPH_F_LC].p); else k_laplacian_correction<3><<<LGRID(c)>>>(d, c->field[ISPH_F_LC].p);
  ++c->launches; c->toc("computeLaplacianCorrection");
}
void compute_normals(Ctx *c) {
  Dev d = make_dev(c, false); c->tic("computeNormals");
  const int masks[2][2] = {{ISPH_KIND_FLUID, ISPH_KIND_SOLID}, {ISPH_KIND_SOLID, ISPH_KIND_FLUID}};
  for (int p = 0; p < 2; ++p) {
    if (d.dim == 2) k_normals<2><<<LGRID(c)>>>(d, masks[p][0], masks[p][1], c->field[ISPH_F_NORMAL].p, c->field[ISPH_F_PND].p);
    else k_normals<3><<<LGRID(c)>>>(d, masks[p][0], masks[p][1], c->field[ISPH_F_NORMAL].p, c->field[ISPH_F_PND].p);
    ++c->launches;
  }
  forward_comm(c, ISPH_F_NORMAL); forward_comm(c, ISPH_F_PND);                      // pair_isph_corrected.cpp:425-427
  c->toc("computeNormals");
}

void assemble_laplacian(Ctx *c, double alpha, const double *mat, bool anti, bool mh, int f0, int f1) {
  Dev d = make_dev(c);
  // list-order neighbour sums (bit-faithful; default) or everything on the column-ordered matrix (ISPH_ASM_REASSOC=1: ~1/3 less work, entries next to the cutoff then differ by up to ~2e-12 relative)
  static const bool reassoc = getenv("ISPH_ASM_REASSOC") != nullptr;
#define LAP(DIM_, ANTI_) do { if (reassoc) k_laplacian_rows<DIM_, ANTI_, false><<<GRID(c)>>>(d, alpha, mat, mh, f0, f1); else k_laplacian_rows<DIM_, ANTI_, true><<<GRID(c)>>>(d, alpha, mat, mh, f0, f1); } while (0)
  if (d.dim == 2) { if (anti) LAP(2, true); else LAP(2, false); }
  else { if (anti) LAP(3, true); else LAP(3, false); }
#undef LAP
  ++c->launches;
  matrix_merge_duplicates(c);
  c->A.is_filled = 1;                                                              // exitFor, functor_laplacian_matrix.h:322-326
}
void assemble_gradient_dot(Ctx *c, double alpha, const double *vec, int f0, int f1) {
  Dev d = make_dev(c);
  if (d.dim == 2) k_gradient_dot_rows<2><<<GRID(c)>>>(d, alpha, vec, f0, f1); else k_gradient_dot_rows<3><<<GRID(c)>>>(d, alpha, vec, f0, f1);
  ++c->launches;
  matrix_merge_duplicates(c);
}

void ns_poisson(Ctx *c, double dt, bool anti, int singular, bool mh) {
  Dev d = make_dev(c);
  ISPH_REQUIRE(c->A.is_filled == 0, "FunctorIncompNavierStokesPoisson:: A is already filled");      // functor_incomp_navier_stokes_poisson.h:58-59
  ISPH_REQUIRE(c->b_nvec >= 1 && c->bs.p, "isph_ns_poisson: create the load multivector first");
  c->tic("computePoisson");
  int f1; bool neumann;
  if (singular == ISPH_NOT_SINGULAR) { f1 = ISPH_KIND_ALL; neumann = false; } else { f1 = ISPH_KIND_FLUID; neumann = true; }   // :72-87
  c->wk.ensure(c->nall);
  k_recip<<<ceil_div(c->nall, 256), 256, 0, c->stream>>>(c->field[ISPH_F_DENSITY].p, c->wk.p, c->nall); ++c->launches;            // :89-92
  assemble_laplacian(c, -dt, c->wk.p, anti, false, ISPH_KIND_FLUID, f1);                                                          // :94-97
  if (neumann) assemble_gradient_dot(c, -dt, c->field[ISPH_F_NORMAL].p, ISPH_KIND_SOLID, ISPH_KIND_ALL);                          // :100-109
  double *b = c->bs.p, *diag = c->A.diagonal.p, *sld = c->A.sld.p; const double *vstar = c->field[ISPH_F_VSTAR].p;
#define PR(D, AN) k_poisson_rows<D, AN><<<GRID(c)>>>(d, mh, neumann, singular, c->rank == 0 ? c->first_fluid_row : -1, vstar, diag, sld, b)
  if (d.dim == 2) { if (anti) PR(2, true); else PR(2, false); } else { if (anti) PR(3, true); else PR(3, false); }
#undef PR
  ++c->launches;
  c->toc("computePoisson");
}

void ns_helmholtz(Ctx *c, double dt, double theta, bool anti, bool mh, bool incp, const double *g) {
  Dev d = make_dev(c);
  ISPH_REQUIRE(c->A.is_filled == 0, "FunctorIncompNavierStokesHelmholtz:: A is already filled");
  ISPH_REQUIRE(c->b_nvec == d.dim && c->bs.p, "isph_ns_helmholtz: the load multivector must have dim columns holding v^n");
  c->tic("computeHelmholtz");
  c->wk.ensure((size_t)c->nall + (size_t)c->ld * 3);
  double *mu = c->wk.p, *w = c->wk.p + c->nall;
  k_mul<<<ceil_div(c->nall, 256), 256, 0, c->stream>>>(c->field[ISPH_F_VISCOSITY].p, c->field[ISPH_F_DENSITY].p, mu, c->nall); ++c->launches;   // :69-72
  assemble_laplacian(c, dt, mu, anti, mh, ISPH_KIND_FLUID, ISPH_KIND_ALL);                       // :74-77
  matrix_left_scale_dev(c, c->field[ISPH_F_DENSITY].p, true);                                    // :80-83  LeftScale(1/rho)
  spmv(c, c->bs.p, w, d.dim, c->ld, c->ld);                                                      // :90     w = A b
  k_scale_vec<<<ceil_div(c->nlocal, 256), 256, 0, c->stream>>>(w, 1.0 - theta, c->nlocal, c->ld, d.dim); ++c->launches;   // :91
  matrix_scale(c, -theta);                                                                       // :94
  const double gg[3] = {g ? g[0] : 0.0, g ? g[1] : 0.0, g ? g[2] : 0.0};
#define HR(D, AN) k_helmholtz_rows<D, AN><<<GRID(c)>>>(d, dt, incp, gg[0], gg[1], gg[2], c->field[ISPH_F_DENSITY].p, c->field[ISPH_F_FORCE].p, \
                                                      c->field[ISPH_F_PRESSURE].p, w, c->ld, c->A.diagonal.p, c->A.sld.p, c->bs.p)
  if (d.dim == 2) { if (anti) HR(2, true); else HR(2, false); } else { if (anti) HR(3, true); else HR(3, false); }
#undef HR
  ++c->launches;
  c->toc("computeHelmholtz");
}

void ns_correct(Ctx *c, double dt, bool anti, bool incp, const double *dp_host) {
  Dev d = make_dev(c);
  double *dp = c->field[ISPH_F_DP].p;
  c->tic("correctVelocityPressure");
  if (dp_host) CUDA_CHECK(cudaMemcpyAsync(dp, dp_host, sizeof(double) * c->nlocal, cudaMemcpyHostToDevice, c->stream));
  else { ISPH_REQUIRE(c->x_nvec >= 1 && c->xs.p, "isph_ns_correct: no solution vector"); CUDA_CHECK(cudaMemcpyAsync(dp, c->xs.p, sizeof(double) * c->nlocal, cudaMemcpyDeviceToDevice, c->stream)); }
  forward_comm(c, ISPH_F_DP);                                                                     // pair_isph.cpp:1017-1019
  if (incp) {                                                                                     // :1022-1023
    c->red.ensure(4096); c->hbuf.ensure(8192);
    const int nb = std::min(592, ceil_div(c->nlocal, 256)); double *sc = c->hbuf.p + 4100;
    k_dp_sum<<<nb, 256, 0, c->stream>>>(c->kind.p, c->nlocal, dp, c->red.p);
    k_dp_sum_final<<<1, 32, 0, c->stream>>>(c->red.p, nb, sc); c->launches += 2;
    if (c->nranks > 1) halo_allreduce(c, sc, 2);
    k_dp_sub_mean<<<ceil_div(c->nall, 256), 256, 0, c->stream>>>(c->kind.p, c->nall, dp, sc); ++c->launches;
  }
  const double *rho = c->field[ISPH_F_DENSITY].p; double *vstar = c->field[ISPH_F_VSTAR].p;
#define CV(D, AN) k_correct_velocity<D, AN><<<GRID(c)>>>(d, dt, rho, dp, vstar)
  if (d.dim == 2) { if (anti) CV(2, true); else CV(2, false); } else { if (anti) CV(3, true); else CV(3, false); }
#undef CV
  ++c->launches;
  forward_comm(c, ISPH_F_VSTAR);                                                                  // functor_correct_velocity.h:71-78
  k_correct_pressure<<<ceil_div(c->nall, 256), 256, 0, c->stream>>>(c->field[ISPH_F_PRESSURE].p, dp, c->nall, incp ? 1 : 0); ++c->launches;
  c->toc("correctVelocityPressure");
}

void pb_jacobian(Ctx *c, bool mh, bool linearized, double ezcb, double psiref, double gamma) {
  Dev d = make_dev(c);
  c->tic("computeJacobianPoissonBoltzmann");
  const double kappasq = 2.0 * ezcb / psiref;                                                    // functor_poisson_boltzmann_jacobian.h:44
  if (c->A.is_filled == 0) {                                                                      // :50-65
    assemble_laplacian(c, -1.0, c->field[ISPH_F_EPS].p, false, mh, ISPH_KIND_FLUID, ISPH_KIND_ALL);
    matrix_extract_diag_dev(c, c->A.sld.p);
    c->A.is_filled = 1;
  }
  k_pb_rows<<<GRID(c)>>>(d, linearized, kappasq, gamma, c->field[ISPH_F_PSI].p, c->A.diagonal.p, c->A.sld.p); ++c->launches;
  c->toc("computeJacobianPoissonBoltzmann");
}

void applied_electric_potential(Ctx *c) {
  Dev d = make_dev(c);
  ISPH_REQUIRE(c->A.is_filled == 0, "FunctorAppliedElectricPotential:: A is already filled");     // functor_applied_electric_potential.h:43-44
  ISPH_REQUIRE(c->b_nvec == 1 && c->bs.p, "isph_applied_electric_potential: create the load vector first");
  c->tic("computeAppliedElectricPotential");
  assemble_laplacian(c, -1.0, c->field[ISPH_F_SIGMA].p, false, false, ISPH_KIND_FLUID | ISPH_FILTER_MATCH, ISPH_KIND_FLUID);   // :49-57
  k_aep_rows<<<GRID(c)>>>(d, c->field[ISPH_F_PHI].p, c->A.diagonal.p, c->bs.p); ++c->launches;
  c->toc("computeAppliedElectricPotential");
}

void solute_transport(Ctx *c, double dt, double theta, double dcoeff) {
  Dev d = make_dev(c);
  ISPH_REQUIRE(c->A.is_filled == 0, "FunctorSoluteTransport:: A is already filled");              // functor_solute_transport.h:56-57
  ISPH_REQUIRE(c->b_nvec == 1 && c->bs.p, "isph_solute_transport: the load vector must hold the concentration");
  c->tic("computeSoluteTransportSpecies");
  c->wk.ensure((size_t)c->nall + (size_t)c->ld * 3); c->flag.ensure(16);
  double *w = c->wk.p + c->nall;
  assemble_laplacian(c, dt * dcoeff, nullptr, false, false, ISPH_KIND_FLUID | ISPH_FILTER_MATCH, ISPH_KIND_FLUID - ISPH_KIND_BUFFER_NEUMANN);   // :62-70
  spmv(c, c->bs.p, w, 1, c->ld, c->ld);                                                          // :88     w = A c^n
  k_scale_vec<<<ceil_div(c->nlocal, 256), 256, 0, c->stream>>>(w, 1.0 - theta, c->nlocal, c->ld, 1); ++c->launches;              // :89
  matrix_scale(c, -theta);                                                                       // :92
  CUDA_CHECK(cudaMemsetAsync(c->flag.p + 7, 0, sizeof(int), c->stream));
  k_solute_rows<<<GRID(c)>>>(d, w, c->A.diagonal.p, c->A.sld.p, c->bs.p, c->flag.p + 7); ++c->launches;
  int bad = 0; CUDA_CHECK(cudaMemcpyAsync(&bad, c->flag.p + 7, sizeof(int), cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream));
  c->toc("computeSoluteTransportSpecies");
  ISPH_REQUIRE(!bad, "FunctorSoluteTransport:: Particle types are not supported");
}

// PairISPH_Corrected::computeF, pair_isph_corrected.cpp:438-485: psi is communicated to the ghosts, then the functor runs
void pb_residual(Ctx *c, bool mh, bool linearized, double ezcb, double psiref, double gamma, const double *d_extra, double *d_f) {
  Dev d = make_dev(c, false);
  c->tic("computeFPoissonBoltzmann");
  forward_comm(c, ISPH_F_PSI);
  const double kappasq = 2.0 * ezcb / psiref;
  const double *psi = c->field[ISPH_F_PSI].p, *psi0 = c->field[ISPH_F_PSI0].p, *eps = c->field[ISPH_F_EPS].p;
  if (d.dim == 2) k_pb_residual<2><<<LGRID(c)>>>(d, mh, linearized, kappasq, gamma, psi, psi0, eps, d_extra, d_f);
  else k_pb_residual<3><<<LGRID(c)>>>(d, mh, linearized, kappasq, gamma, psi, psi0, eps, d_extra, d_f);
  ++c->launches;
  c->toc("computeFPoissonBoltzmann");
}

void advance_time(Ctx *c, double dt, bool anti) {
  Dev d = make_dev(c);
  c->tic("advanceTime");
  double *v = c->field[ISPH_F_VELOCITY].p, *p = c->field[ISPH_F_PRESSURE].p, *dp = c->field[ISPH_F_DP].p; const double *vnp1 = c->field[ISPH_F_VSTAR].p;
#define AB(D, AN) k_advance_begin<D, AN><<<GRID(c)>>>(d, dt, v, vnp1, p, dp)
  if (d.dim == 2) { if (anti) AB(2, true); else AB(2, false); } else { if (anti) AB(3, true); else AB(3, false); }
#undef AB
  ++c->launches;
  forward_comm(c, ISPH_F_DP);                                                                     // functor_advance_time_begin.h:74-78
  FixedTab ft; for (int t = 0; t < ISPH_MAXT; ++t) ft.f[t] = c->fixed_of_type[t];
  k_advance_end<<<ceil_div(c->nall, 256), 256, 0, c->stream>>>(c->nall, d.dim, dt, c->type.p, ft, v, vnp1, p, dp, c->x.p); ++c->launches;
  c->A.built = false; c->A.is_filled = 0;                                                         // the particles moved: graph and matrix belong to the old positions
  c->toc("advanceTime");
}

static void check_kind_flag(Ctx *c, const char *what) {
  int bad = 0; CUDA_CHECK(cudaMemcpyAsync(&bad, c->flag.p + 5, sizeof(int), cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream));
  ISPH_REQUIRE(bad == 0, std::string(what) + ":: Particle types are not supported");
}
void boundary_navier_slip(Ctx *c, double beta) {
  Dev d = make_dev(c);
  if (beta == 0.0) return;                                                                        // pair_isph_corrected.cpp:922
  c->flag.ensure(16); CUDA_CHECK(cudaMemsetAsync(c->flag.p + 5, 0, sizeof(int), c->stream));
  if (d.dim == 2) k_navier_slip_rows<2><<<GRID(c)>>>(d, beta, c->field[ISPH_F_DENSITY].p, c->flag.p + 5);
  else k_navier_slip_rows<3><<<GRID(c)>>>(d, beta, c->field[ISPH_F_DENSITY].p, c->flag.p + 5);
  ++c->launches;
  check_kind_flag(c, "FunctorBoundaryNavierSlip");
}
void boundary_dirichlet(Ctx *c) {
  Dev d = make_dev(c);
  ISPH_REQUIRE(c->b_nvec == d.dim && c->bs.p, "isph_boundary_dirichlet: the load multivector must have dim columns");
  c->flag.ensure(16); CUDA_CHECK(cudaMemsetAsync(c->flag.p + 5, 0, sizeof(int), c->stream));
  if (d.dim == 2) k_dirichlet_rows<2><<<GRID(c)>>>(d, c->bs.p, c->ld, c->flag.p + 5); else k_dirichlet_rows<3><<<GRID(c)>>>(d, c->bs.p, c->ld, c->flag.p + 5);
  ++c->launches;
  check_kind_flag(c, "FunctorBoundaryDirichlet");
}

}  // namespace isph
