// Multilevel preconditioner standing in for PrecondWrapper_ML (precond_ml.h:17-172; `Precond Package` = "ML" is the reference's
// default, pair_isph.cpp:325-329,359-361): create() builds the hierarchy from the assembled matrix, the Krylov solver calls the V-cycle
// as ApplyInverse, free() drops it — every solve, as the reference does (solver_lin_belos.h:153,190).
//
// ML itself is third-party code that is neither vendored nor pinned, so nothing here can be pinned against it ("parity unpinned" by
// construction; the CPU restatement is oracle/amg_oracle.h).  What is built is the data-parallel member of ML's own option space:
//   aggregation: MIS (distance-2 maximal independent set on the strength graph, priorities = hash of the particle tag), aggregates never
//     cross a rank ("Uncoupled"); strength = ML's criterion a_ij^2 > eps^2 |a_ii a_jj| ("aggregation: threshold");
//   prolongator: piecewise constants ("aggregation: damping factor" 0 = non-smoothed aggregation), Galerkin operators P^T A P;
//   smoothers: Chebyshev (Ifpack/ML recurrence on D^-1 A, lambda_max from a power method) or damped Jacobi on every level, the coarsest
//     level included (what PrecondWrapper_ML::setNullVector selects for singular problems, precond_ml.h:118-120);
//   cycle: V, with a scaled coarse-grid correction.
// The reference's default smoother (symmetric Gauss-Seidel, precond_ml.h:53) is sequential within a rank and is rejected by name.
//
// B200 mapping.  The finest level is the SELL-32 matrix itself (SpMV at ~97 % of the HBM peak, halo over NVLink); with the default
// threshold the 3-D lattice coarsens ~38x per level, so every coarser level is < 1 % of the finest and is kept REPLICATED on every GPU:
// the setup all-gathers the level-1 rows once, a V-cycle costs one all-reduce of the level-1 residual, and the coarse levels need no
// halo plan of their own.  Setup passes over the fine matrix: strength graph (1), row compression to aggregate columns (1), power
// method ("eigen-analysis: iterations").  Every choice is deterministic and independent of the local numbering (priorities and
// tie-breaks hash global ids; sums run in fixed orders), so all ranks build the same replicated levels and the CPU oracle reproduces
// the same aggregates.
#include "isph_internal.h"
#include <cub/cub.cuh>

namespace isph {

void halo_allgather_bytes(Ctx *c, const void *send, void *recv, size_t bytes_per_rank);   // halo.cu

namespace {

const int VB = 256;
inline int vgrid(long long n) { long long g = (n + VB - 1) / VB; return (int)(g < 1 ? 1 : (g > 1184 ? 1184 : g)); }
inline int tgrid(long long n, int b = 256) { return (int)((n + b - 1) / b > 0 ? (n + b - 1) / b : 1); }

struct SellView { const long long *so; const int *col; const double *val; const int *rl; };
struct CsrView { const int *rp; const int *col; const double *val; };
__device__ __forceinline__ void row_span(const SellView &M, int r, long long &base, int &len, int &stride) { base = M.so[r >> 5] + (r & 31); len = M.rl[r]; stride = 32; }
__device__ __forceinline__ void row_span(const CsrView &M, int r, long long &base, int &len, int &stride) { base = M.rp[r]; len = M.rp[r + 1] - M.rp[r]; stride = 1; }

__device__ __forceinline__ unsigned long long mix64(unsigned long long t, int salt) {
  unsigned long long z = t + 0x9E3779B97F4A7C15ULL * (unsigned long long)(salt + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; z = z ^ (z >> 31);
  return z;
}

// ---- strength graph: column-major ELL (entry k of row r at k*n + r) of the strong neighbours and their |a_ij| as float ----------
template <class M> __global__ void __launch_bounds__(128) k_amg_diag(M A, int n, double *d) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x; if (r >= n) return;
  long long base; int len, st; row_span(A, r, base, len, st); double s = 0.0;
  for (int e = 0; e < len; ++e) if (A.col[base + (long long)e * st] == r) s += A.val[base + (long long)e * st];
  d[r] = s;
}
template <class M> __global__ void __launch_bounds__(128) k_amg_strength(M A, int n, int nown, const double *d, const int *blk, double theta2, int *sg, float *sw, int *cnt) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x; if (r >= n) return;
  long long base; int len, st; row_span(A, r, base, len, st);
  const double dr = d[r]; const int br = blk ? blk[r] : 0; int k = 0;
  auto visit = [&](int c, double a) {
    if (c == r || c >= nown) return;                                 // halo columns belong to another rank: never strong
    if (blk && blk[c] != br) return;
    if (a * a > theta2 * fabs(dr * d[c])) { sg[(size_t)k * n + r] = c; sw[(size_t)k * n + r] = (float)fabs(a); ++k; }
  };
  int e = 0;
  for (; e + 4 <= len; e += 4) {                                     // four independent (column, value) loads in flight, visited in stored order
    const long long q = base + (long long)e * st;
    const int c0 = A.col[q], c1 = A.col[q + st], c2 = A.col[q + 2ll * st], c3 = A.col[q + 3ll * st];
    const double a0 = A.val[q], a1 = A.val[q + st], a2 = A.val[q + 2ll * st], a3 = A.val[q + 3ll * st];
    visit(c0, a0); visit(c1, a1); visit(c2, a2); visit(c3, a3);
  }
  for (; e < len; ++e) visit(A.col[base + (long long)e * st], A.val[base + (long long)e * st]);
  cnt[r] = k;
}

// ---- distance-2 maximal independent set, synchronous rounds.  state: 0 undecided, 1 root, -1 out, -2 no strong connection ----------
// tok = what a row shows its neighbours: all ones once it is a root, its priority while undecided, 0 when out / without strong connection
// (one 8-byte gather per neighbour per sweep instead of state + priority)
__global__ void __launch_bounds__(VB) k_amg_mis_init(int n, const int *gid, const int *cnt, unsigned long long *key, int *state, unsigned long long *tok) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x; if (r >= n) return;
  const unsigned g = (unsigned)gid[r];
  const unsigned long long k = ((mix64((unsigned long long)g, 3) >> 32) << 32) | g;
  key[r] = k; state[r] = cnt[r] == 0 ? -2 : 0; tok[r] = cnt[r] == 0 ? 0ULL : k;
}
__global__ void __launch_bounds__(VB) k_amg_mis_m1(int n, const int *sg, const int *cnt, const unsigned long long *tok, unsigned long long *m1) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x; if (r >= n) return;
  unsigned long long m = tok[r]; const int nk = cnt[r]; int k = 0;
  for (; k + 4 <= nk; k += 4) {                                      // four neighbour tokens in flight
    const int j0 = sg[(size_t)k * n + r], j1 = sg[(size_t)(k + 1) * n + r], j2 = sg[(size_t)(k + 2) * n + r], j3 = sg[(size_t)(k + 3) * n + r];
    const unsigned long long t0 = tok[j0], t1 = tok[j1], t2 = tok[j2], t3 = tok[j3];
    const unsigned long long a = t0 > t1 ? t0 : t1, b = t2 > t3 ? t2 : t3, c2 = a > b ? a : b; m = c2 > m ? c2 : m;
  }
  for (; k < nk; ++k) { const unsigned long long t = tok[sg[(size_t)k * n + r]]; m = t > m ? t : m; }
  m1[r] = m;
}
__global__ void __launch_bounds__(VB) k_amg_mis_m2(int n, const int *sg, const int *cnt, const unsigned long long *key, int *state, unsigned long long *tok, const unsigned long long *m1, int *undecided) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x; if (r >= n) return;
  if (state[r] != 0) return;                                        // only a row's own thread writes its state / token; neighbours are read through m1 (snapshot)
  unsigned long long m = m1[r]; const int nk = cnt[r]; int k = 0;
  for (; k + 4 <= nk; k += 4) {
    const int j0 = sg[(size_t)k * n + r], j1 = sg[(size_t)(k + 1) * n + r], j2 = sg[(size_t)(k + 2) * n + r], j3 = sg[(size_t)(k + 3) * n + r];
    const unsigned long long t0 = m1[j0], t1 = m1[j1], t2 = m1[j2], t3 = m1[j3];
    const unsigned long long a = t0 > t1 ? t0 : t1, b = t2 > t3 ? t2 : t3, c2 = a > b ? a : b; m = c2 > m ? c2 : m;
  }
  for (; k < nk; ++k) { const unsigned long long t = m1[sg[(size_t)k * n + r]]; m = t > m ? t : m; }
  if (m == ~0ULL) { state[r] = -1; tok[r] = 0ULL; } else if (m == key[r]) { state[r] = 1; tok[r] = ~0ULL; } else atomicAdd(undecided, 1);
}
__global__ void __launch_bounds__(VB) k_amg_flag(int n, const int *state, const int *agg, int what, int *flag) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x; if (r > n) return;
  flag[r] = r < n && (what == 0 ? state[r] == 1 : (agg[r] < 0 && state[r] != -2)) ? 1 : 0;
}
__global__ void __launch_bounds__(VB) k_amg_number(int n, const int *flag, const int *scan, int offset, int *agg, int *root_of, int init) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x; if (r >= n) return;
  if (flag[r]) { const int a = offset + scan[r]; agg[r] = a; root_of[a] = r; } else if (init) agg[r] = -1;
}
// join the aggregate of the strongest aggregated neighbour (snapshot agg_in -> agg_out); ties: the neighbour with the larger priority
__global__ void __launch_bounds__(VB) k_amg_join(int n, const int *sg, const float *sw, const int *cnt, const unsigned long long *key, const int *state, const int *agg_in, int *agg_out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x; if (r >= n) return;
  int a = agg_in[r];
  if (a < 0 && state[r] != -2) {
    int best = -1; float bw = -1.0f; unsigned long long bk = 0;
    for (int k = 0; k < cnt[r]; ++k) {
      const int j = sg[(size_t)k * n + r]; const int aj = agg_in[j]; if (aj < 0) continue;
      const float w = sw[(size_t)k * n + r]; const unsigned long long kj = key[j];
      if (best < 0 || w > bw || (w == bw && kj > bk)) { best = aj; bw = w; bk = kj; }
    }
    if (best >= 0) a = best;
  }
  agg_out[r] = a;
}
__global__ void __launch_bounds__(VB) k_amg_gather_int(int n, const int *idx, const int *src, int *dst) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) dst[i] = src[idx[i]]; }
__global__ void __launch_bounds__(VB) k_amg_sortkeys(int n, const int *agg, int nc, int *key, int *val, int *count) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x; if (r >= n) return;
  const int a = agg[r]; key[r] = a < 0 ? nc : a; val[r] = r; if (a >= 0) atomicAdd(count + a, 1);
}
// aggregate id of every column as the coarse (global) index: owned rows offset by this rank's first aggregate, halo columns from their owners
__global__ void __launch_bounds__(VB) k_amg_agg_to_double(int n, const int *agg, int offset, double *out) { const int r = blockIdx.x * blockDim.x + threadIdx.x; if (r < n) out[r] = agg[r] < 0 ? -1.0 : (double)(agg[r] + offset); }
__global__ void __launch_bounds__(VB) k_amg_double_to_agg(int n, const double *in, int *out) { const int r = blockIdx.x * blockDim.x + threadIdx.x; if (r < n) out[r] = (int)in[r]; }

// ---- Galerkin operator of plain aggregation, two deterministic steps -------------------------------------------------------------------
// (1) every fine row compressed to its aggregate columns: Q = A P, sums in stored entry order, list in order of first appearance
template <class M, int KQ> __global__ void __launch_bounds__(128) k_amg_compress(M A, int n, const int *aggc, const int *aggr, int *qcnt, int *qj, double *qv, int *overflow) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x; if (r >= n) return;
  if (aggr[r] < 0) { qcnt[r] = 0; return; }
  long long base; int len, st; row_span(A, r, base, len, st);
  // (summing runs of consecutive entries with the same aggregate in registers before touching the list was measured: 14.4 ms against 12.3 ms
  // for this loop on 8M rows — the extra divergence costs more than the saved local-memory traffic; not kept)
  int J[KQ]; double S[KQ]; int m = 0;
  auto visit = [&](int jc, double a) {
    if (jc < 0) return;
    int k = 0; for (; k < m; ++k) if (J[k] == jc) break;
    if (k < m) S[k] += a; else if (m < KQ) { J[m] = jc; S[m] = a; ++m; } else *overflow = 1;
  };
  int e = 0;
  for (; e + 4 <= len; e += 4) {                                     // four (column -> aggregate) gathers in flight; the list is updated in stored order
    const long long q = base + (long long)e * st;
    const int c0 = A.col[q], c1 = A.col[q + st], c2 = A.col[q + 2ll * st], c3 = A.col[q + 3ll * st];
    const double a0 = A.val[q], a1 = A.val[q + st], a2 = A.val[q + 2ll * st], a3 = A.val[q + 3ll * st];
    const int j0 = aggc[c0], j1 = aggc[c1], j2 = aggc[c2], j3 = aggc[c3];
    visit(j0, a0); visit(j1, a1); visit(j2, a2); visit(j3, a3);
  }
  for (; e < len; ++e) visit(aggc[A.col[base + (long long)e * st]], A.val[base + (long long)e * st]);
  qcnt[r] = m;
  for (int k = 0; k < m; ++k) { qj[(size_t)k * n + r] = J[k]; qv[(size_t)k * n + r] = S[k]; }
}
// (2) the rows of an aggregate merged, one warp per coarse row: members in ascending row order, their Q entries in list order.  The set
// of columns goes through a shared-memory hash, is sorted, and every sum is accumulated by the lane that owns its slot, in stream order.
template <int HS, int WPB, bool FILL> __global__ void __launch_bounds__(32 * WPB) k_amg_merge(int nc, int n, const int *moff, const int *mem, const int *qcnt, const int *qj, const double *qv,
                                                                                 int *ccnt, const int *crp, int *cci, double *cva, int *overflow) {
  __shared__ int hk[WPB][HS]; __shared__ int lst[WPB][HS]; __shared__ double acc[FILL ? WPB : 1][FILL ? HS : 1]; __shared__ int cnts[WPB];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, I = blockIdx.x * WPB + w;
  if (I >= nc) return;
  int *H = hk[w], *Ls = lst[w];
  for (int i = lane; i < HS; i += 32) H[i] = -1;
  if (lane == 0) cnts[w] = 0;
  __syncwarp();
  const int m0 = moff[I], m1 = moff[I + 1];
  // the column set: a lane per member (32 members' rows in flight at once; the order of insertion is irrelevant for a set)
  for (int b = m0; b < m1; b += 32) {
    const int my_row = b + lane < m1 ? mem[b + lane] : 0, my_qc = b + lane < m1 ? qcnt[my_row] : 0;
    int qmax = my_qc; for (int o = 16; o > 0; o >>= 1) qmax = max(qmax, __shfl_xor_sync(0xffffffffu, qmax, o));
    for (int k = 0; k < qmax; ++k) {
      if (k < my_qc) {
        const int J = qj[(size_t)k * n + my_row]; unsigned h = ((unsigned)J * 2654435761u) & (HS - 1); bool placed = false;
        for (int probe = 0; probe < HS && !placed; ++probe) {
          const int old = atomicCAS(&H[h], -1, J);
          if (old == -1) { atomicAdd(&cnts[w], 1); placed = true; }
          else if (old == J) placed = true;
          else h = (h + 1) & (HS - 1);
        }
        if (!placed) atomicExch(&cnts[w], HS);                      // table full: reported as overflow below (never a silently dropped column)
      }
      __syncwarp();
      if (cnts[w] > HS * 3 / 4) { if (lane == 0) { *overflow = 1; if (!FILL) ccnt[I] = 0; } return; }
    }
  }
  const int nd = cnts[w];
  if (!FILL) { if (lane == 0) ccnt[I] = nd; return; }
  int P = 32; while (P < nd) P <<= 1;
  { int pos = 0;                                                     // compact the table (order irrelevant: sorted next)
    for (int i0 = 0; i0 < HS; i0 += 32) { const int kv = H[i0 + lane]; const unsigned b = __ballot_sync(0xffffffffu, kv >= 0);
      if (kv >= 0) Ls[pos + __popc(b & ((1u << lane) - 1))] = kv; pos += __popc(b); }
    for (int i = nd + lane; i < P; i += 32) Ls[i] = 0x7fffffff; }
  __syncwarp();
  for (int k = 2; k <= P; k <<= 1) for (int j = k >> 1; j > 0; j >>= 1) {
    for (int i = lane; i < P; i += 32) { const int x = i ^ j; if (x > i) { const int a = Ls[i], b = Ls[x]; if ((a > b) == ((i & k) == 0)) { Ls[i] = b; Ls[x] = a; } } }
    __syncwarp();
  }
  double *Ac = acc[w];
  for (int i = lane; i < nd; i += 32) Ac[i] = 0.0;
  __syncwarp();
  // the sums, members in ascending row order, entries of a member in list order: (row, length) of 32 members are fetched at once and the
  // first 32 entries of the next member are loaded while the current one is accumulated (one memory latency per member instead of three)
  for (int b = m0; b < m1; b += 32) {
    const int nb = m1 - b < 32 ? m1 - b : 32;
    const int my_row = lane < nb ? mem[b + lane] : 0, my_qc = lane < nb ? qcnt[my_row] : 0;
    int row = __shfl_sync(0xffffffffu, my_row, 0), qc = __shfl_sync(0xffffffffu, my_qc, 0);
    int J = 0; double v = 0.0; if (lane < qc) { J = qj[(size_t)lane * n + row]; v = qv[(size_t)lane * n + row]; }
    for (int t = 0; t < nb; ++t) {
      int rown = 0, qcn = 0, Jn = 0; double vn = 0.0;
      if (t + 1 < nb) { rown = __shfl_sync(0xffffffffu, my_row, t + 1); qcn = __shfl_sync(0xffffffffu, my_qc, t + 1); if (lane < qcn) { Jn = qj[(size_t)lane * n + rown]; vn = qv[(size_t)lane * n + rown]; } }
      for (int k0 = 0; k0 < qc; k0 += 32) {
        const int k = k0 + lane; int slot = -1;
        if (k0 > 0 && k < qc) { J = qj[(size_t)k * n + row]; v = qv[(size_t)k * n + row]; }
        if (k < qc) { int lo = 0, hi = nd - 1; while (lo < hi) { const int mid = (lo + hi) >> 1; if (Ls[mid] < J) lo = mid + 1; else hi = mid; } slot = lo; }
        const int na = qc - k0 < 32 ? qc - k0 : 32;
        for (int u = 0; u < na; ++u) { const int sl = __shfl_sync(0xffffffffu, slot, u); const double vt = __shfl_sync(0xffffffffu, v, u); if ((sl & 31) == lane) Ac[sl] += vt; }
      }
      row = rown; qc = qcn; J = Jn; v = vn;
    }
  }
  __syncwarp();
  const int o = crp[I];
  for (int i = lane; i < nd; i += 32) { cci[o + i] = Ls[i]; cva[o + i] = Ac[i]; }
}

// ---- vector / smoother kernels ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VB) k_amg_invdiag(const double *d, double *inv, int n) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) inv[i] = d[i] != 0.0 ? 1.0 / d[i] : 0.0; }
__global__ void __launch_bounds__(VB) k_amg_hashvec(double *y, const int *gid, int n) {
  for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) y[i] = 2.0 * ((double)(mix64((unsigned long long)(unsigned)gid[i], 7) >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
}
__global__ void __launch_bounds__(VB) k_amg_scale(double *x, const double *y, const double *d, double s, int n) { for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) x[i] = (d ? d[i] : 1.0) * y[i] * s; }
__global__ void __launch_bounds__(VB) k_amg_dot3(const double *a, const double *b, const double *d, int n, double *partials) {      // a.(d b), a.a, (d b).(d b)
  double s0 = 0, s1 = 0, s2 = 0;
  for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) { const double y = (d ? d[i] : 1.0) * b[i]; s0 += a[i] * y; s1 += a[i] * a[i]; s2 += y * y; }
  __shared__ double sm[3][VB / 32];
  for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = s0; sm[1][threadIdx.x >> 5] = s1; sm[2][threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x < 3) { double s = 0; for (int w = 0; w < VB / 32; ++w) s += sm[threadIdx.x][w]; partials[blockIdx.x * 3 + threadIdx.x] = s; }
}
// smoother updates given V = A x (finest level: the product comes from the SELL SpMV).  MODE 0: zero start  w = D^-1 r c1 ; x = w
// 1: restart from x  w = D^-1 (r - V) c1 ; x += w     2: Chebyshev step  w = c1 w + c2 D^-1 (r - V) ; x += w     3: Jacobi  x += c1 D^-1 (r - V)
template <int MODE> __global__ void __launch_bounds__(VB) k_amg_update(const double *r, const double *V, const double *invdiag, double c1, double c2, double *W, double *x, int n) {
  for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) {
    if (MODE == 0) { const double w = invdiag[i] * r[i] * c1; W[i] = w; x[i] = w; }
    if (MODE == 1) { const double w = invdiag[i] * (r[i] - V[i]) * c1; W[i] = w; x[i] += w; }
    if (MODE == 2) { double w = W[i] * c1; w += c2 * invdiag[i] * (r[i] - V[i]); W[i] = w; x[i] += w; }
    if (MODE == 3) x[i] += c1 * invdiag[i] * (r[i] - V[i]);
  }
}
// coarse levels (CSR, replicated): product and update in one kernel, 8 lanes per row, x ping-pongs (xin is gathered by other rows).
// MODE as above, plus 4: xout = r - A xin (residual for the restriction), 5: xout = D^-1 A xin (power method)
template <int MODE> __global__ void __launch_bounds__(VB) k_amg_csr(CsrView A, int n, const double *xin, const double *r, const double *invdiag, double c1, double c2, double *W, double *xout) {
  const int t = blockIdx.x * VB + threadIdx.x, row = t >> 3, l = t & 7;
  double s = 0.0;
  if (row < n) for (int q = A.rp[row] + l; q < A.rp[row + 1]; q += 8) s += A.val[q] * xin[A.col[q]];
  s += __shfl_xor_sync(0xffffffffu, s, 4); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 1);
  if (row >= n || l != 0) return;
  if (MODE == 1) { const double w = invdiag[row] * (r[row] - s) * c1; W[row] = w; xout[row] = xin[row] + w; }
  if (MODE == 2) { double w = W[row] * c1; w += c2 * invdiag[row] * (r[row] - s); W[row] = w; xout[row] = xin[row] + w; }
  if (MODE == 3) xout[row] = xin[row] + c1 * invdiag[row] * (r[row] - s);
  if (MODE == 4) xout[row] = r[row] - s;
  if (MODE == 5) xout[row] = invdiag[row] * s;
}
// restriction: rc[I] = sum over the members (ascending) of res; entries of other ranks' aggregates are written as zero (they arrive
// through the all-reduce).  V == nullptr: the pre-smoother did nothing, res = r
__global__ void __launch_bounds__(VB) k_amg_restrict(int nc_total, int first, int nc_mine, const int *moff, const int *mem, const double *r, const double *V, double *rc) {
  const int I = blockIdx.x * blockDim.x + threadIdx.x; if (I >= nc_total) return;
  const int a = I - first; double s = 0.0;
  if (a >= 0 && a < nc_mine) for (int q = moff[a]; q < moff[a + 1]; ++q) { const int m = mem[q]; s += V ? r[m] - V[m] : r[m]; }
  rc[I] = s;
}
__global__ void __launch_bounds__(VB) k_amg_prolong(int n, const int *agg, int offset, const double *ec, double scale, double *x) {
  for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) { const int a = agg[i]; if (a >= 0) x[i] += scale * ec[a + offset]; }
}
// coarsest level, "coarse: type" = Amesos-KLU: x = Ainv b with the explicit inverse (n <= 1024), one warp per row, fixed shuffle tree
__global__ void __launch_bounds__(VB) k_amg_dense_apply(int n, const double *Ainv, const double *b, double *x) {
  const int row = (blockIdx.x * VB + threadIdx.x) >> 5, lane = threadIdx.x & 31; if (row >= n) return;
  double s = 0.0; for (int k = lane; k < n; k += 32) s += Ainv[(size_t)row * n + k] * b[k];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) x[row] = s;
}
__global__ void __launch_bounds__(VB) k_amg_rowlen(int n, const int *rp, int *len) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) len[i] = rp[i + 1] - rp[i]; }

}  // namespace

struct AmgLevel {
  int n = 0, nc = 0, nc_mine = 0, first = 0; long long nnz = 0; double lmax = 0.0;         // nc = coarse size (all ranks), first/nc_mine = this rank's aggregates (level 0)
  DevBuf<int> rp, ci, gid, agg, root_of, moff, mem; DevBuf<double> va, invdiag, xa, xb, b, w;
  ~AmgLevel() { rp.release(); ci.release(); gid.release(); agg.release(); root_of.release(); moff.release(); mem.release(); va.release(); invdiag.release(); xa.release(); xb.release(); b.release(); w.release(); }
};
struct AmgData {
  std::vector<AmgLevel *> L; int nlev = 0; bool ready = false, direct = false; DevBuf<double> coarse_inv;
  // setup scratch (grow-only, shared by all levels)
  DevBuf<int> sg, cnt, state, flag, scan, agg2, skey, skey2, sval, ccnt, qcnt, qj, aggc, rlen, gbuf_i; DevBuf<float> sw; DevBuf<unsigned long long> key, m1, tok; DevBuf<double> d, qv, aggd, gbuf_d, red;
  DevBuf<char> tmp; DevBuf<int> ctr; DevBuf<long long> cnt2;
  DevBuf<double> t0, w0;                           // finest-level work vectors (length ld: the product needs the halo tail)
  std::map<std::string, double> setup_ms;
  ~AmgData() {
    for (auto *l : L) delete l;
    sg.release(); cnt.release(); state.release(); flag.release(); scan.release(); agg2.release(); skey.release(); skey2.release(); sval.release(); ccnt.release(); qcnt.release(); qj.release();
    aggc.release(); rlen.release(); gbuf_i.release(); sw.release(); key.release(); m1.release(); tok.release(); d.release(); qv.release(); aggd.release(); gbuf_d.release(); red.release();
    tmp.release(); ctr.release(); cnt2.release(); t0.release(); w0.release(); coarse_inv.release();
  }
};

namespace {

#define LAUNCH(c) (++(c)->launches)

template <class T> T d2h(Ctx *c, const T *p) { T v; CUDA_CHECK(cudaMemcpyAsync(&v, p, sizeof(T), cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream)); return v; }

void exclusive_scan(Ctx *c, AmgData *D, const int *in, int *out, int n) {
  size_t tb = 0; cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, n, c->stream); D->tmp.ensure(tb);
  tb = D->tmp.cap; cub::DeviceScan::ExclusiveSum(D->tmp.p, tb, in, out, n, c->stream); LAUNCH(c);
}
int reduce_max(Ctx *c, AmgData *D, const int *in, int n) {
  D->ctr.ensure(8); size_t tb = 0; cub::DeviceReduce::Max(nullptr, tb, in, D->ctr.p + 4, n, c->stream); D->tmp.ensure(tb);
  tb = D->tmp.cap; cub::DeviceReduce::Max(D->tmp.p, tb, in, D->ctr.p + 4, n, c->stream); LAUNCH(c);
  return d2h(c, D->ctr.p + 4);
}

// local (this GPU) or global (all ranks, finest level) a.(d b), a.a, (d b).(d b)
void dot3(Ctx *c, AmgData *D, const double *a, const double *b, const double *d, int n, bool global, double out[3]) {
  const int g = vgrid(n) < 592 ? vgrid(n) : 592; D->red.ensure((size_t)592 * 3 + 8);
  k_amg_dot3<<<g, VB, 0, c->stream>>>(a, b, d, n, D->red.p); LAUNCH(c);
  std::vector<double> h((size_t)g * 3);
  CUDA_CHECK(cudaMemcpyAsync(h.data(), D->red.p, sizeof(double) * g * 3, cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream));
  out[0] = out[1] = out[2] = 0.0;
  for (int q = 0; q < g; ++q) for (int k = 0; k < 3; ++k) out[k] += h[(size_t)q * 3 + k];
  if (global && c->nranks > 1) {
    double *dv = D->red.p + 592 * 3; CUDA_CHECK(cudaMemcpyAsync(dv, out, 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    halo_allreduce(c, dv, 3); CUDA_CHECK(cudaMemcpyAsync(out, dv, 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream));
  }
}

CsrView view(AmgLevel *L) { return CsrView{L->rp.p, L->ci.p, L->va.p}; }
SellView view(Matrix &A) { return SellView{A.slice_off.p, A.col.p, A.val.p, A.row_len.p}; }

// lambda_max of D^-1 A by the power method (ML "eigen-analysis", Ifpack_Chebyshev::PowerMethod), start vector = hash of the global ids
double power_method(Ctx *c, AmgData *D, AmgLevel *L, int lev, int iters) {
  const int n = L->n; if (n == 0) return 1.0;
  double d[3], lmax = 1.0;
  if (lev == 0) {
    double *x = D->w0.p, *y = D->t0.p; const int ld = c->ld;
    k_amg_hashvec<<<vgrid(n), VB, 0, c->stream>>>(x, L->gid.p, n); LAUNCH(c);
    dot3(c, D, x, x, nullptr, n, true, d);
    k_amg_scale<<<vgrid(n), VB, 0, c->stream>>>(x, x, nullptr, 1.0 / sqrt(d[1]), n); LAUNCH(c);
    for (int it = 0; it < iters; ++it) {
      spmv(c, x, y, 1, ld, ld);
      dot3(c, D, x, y, L->invdiag.p, n, true, d); lmax = d[0] / d[1];
      k_amg_scale<<<vgrid(n), VB, 0, c->stream>>>(x, y, L->invdiag.p, 1.0 / sqrt(d[2]), n); LAUNCH(c);
    }
  } else {
    double *x = L->xa.p, *y = L->xb.p;
    k_amg_hashvec<<<vgrid(n), VB, 0, c->stream>>>(x, L->gid.p, n); LAUNCH(c);
    dot3(c, D, x, x, nullptr, n, false, d);
    k_amg_scale<<<vgrid(n), VB, 0, c->stream>>>(x, x, nullptr, 1.0 / sqrt(d[1]), n); LAUNCH(c);
    for (int it = 0; it < iters; ++it) {
      k_amg_csr<5><<<tgrid((long long)n * 8), VB, 0, c->stream>>>(view(L), n, x, nullptr, L->invdiag.p, 0.0, 0.0, nullptr, y); LAUNCH(c);
      dot3(c, D, x, y, nullptr, n, false, d); lmax = d[0] / d[1];
      k_amg_scale<<<vgrid(n), VB, 0, c->stream>>>(x, y, nullptr, 1.0 / sqrt(d[2]), n); LAUNCH(c);
    }
  }
  return lmax;
}

struct PhaseTimer {                       // host wall-clock of a setup phase (the phases synchronise anyway: counts come back to the host)
  Ctx *c; AmgData *D; const char *name; cudaEvent_t a, b;
  PhaseTimer(Ctx *c_, AmgData *D_, const char *n) : c(c_), D(D_), name(n) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, c->stream); }
  ~PhaseTimer() { cudaEventRecord(b, c->stream); cudaEventSynchronize(b); float ms = 0; cudaEventElapsedTime(&ms, a, b); D->setup_ms[name] += ms; cudaEventDestroy(a); cudaEventDestroy(b); }
};

// aggregates of level `lev` (L->agg, L->root_of, L->nc_mine); M = its matrix view.  nown = rows/columns owned here.
template <class M> void aggregate(Ctx *c, AmgData *D, AmgLevel *L, int lev, M A, int maxrow, const double *diag, const int *blk) {
  const int n = L->n; const double th = c->pp.ml_threshold;
  D->sg.ensure((size_t)n * maxrow); D->sw.ensure((size_t)n * maxrow); D->cnt.ensure(n); D->state.ensure(n); D->key.ensure(n); D->m1.ensure(n); D->tok.ensure(n);
  D->flag.ensure(n + 1); D->scan.ensure(n + 1); D->agg2.ensure(n); D->ctr.ensure(8); L->agg.ensure(n); L->root_of.ensure(n);
  k_amg_strength<<<tgrid(n, 128), 128, 0, c->stream>>>(A, n, n, diag, blk, th * th, D->sg.p, D->sw.p, D->cnt.p); LAUNCH(c);
  k_amg_mis_init<<<tgrid(n), VB, 0, c->stream>>>(n, L->gid.p, D->cnt.p, D->key.p, D->state.p, D->tok.p); LAUNCH(c);
  for (int round = 0; round < 64; ++round) {
    CUDA_CHECK(cudaMemsetAsync(D->ctr.p, 0, sizeof(int), c->stream));
    k_amg_mis_m1<<<tgrid(n), VB, 0, c->stream>>>(n, D->sg.p, D->cnt.p, D->tok.p, D->m1.p); LAUNCH(c);
    k_amg_mis_m2<<<tgrid(n), VB, 0, c->stream>>>(n, D->sg.p, D->cnt.p, D->key.p, D->state.p, D->tok.p, D->m1.p, D->ctr.p); LAUNCH(c);
    if (d2h(c, D->ctr.p) == 0) break;
    ISPH_REQUIRE(round < 63, "ML stand-in: the independent-set rounds did not terminate");
  }
  k_amg_flag<<<tgrid(n + 1), VB, 0, c->stream>>>(n, D->state.p, nullptr, 0, D->flag.p); LAUNCH(c);
  exclusive_scan(c, D, D->flag.p, D->scan.p, n + 1);
  const int nroot = d2h(c, D->scan.p + n);
  k_amg_number<<<tgrid(n), VB, 0, c->stream>>>(n, D->flag.p, D->scan.p, 0, L->agg.p, L->root_of.p, 1); LAUNCH(c);
  int *a = L->agg.p, *b = D->agg2.p;
  for (int pass = 0; pass < 3; ++pass) { k_amg_join<<<tgrid(n), VB, 0, c->stream>>>(n, D->sg.p, D->sw.p, D->cnt.p, D->key.p, D->state.p, a, b); LAUNCH(c); std::swap(a, b); }
  if (a != L->agg.p) CUDA_CHECK(cudaMemcpyAsync(L->agg.p, a, sizeof(int) * n, cudaMemcpyDeviceToDevice, c->stream));
  k_amg_flag<<<tgrid(n + 1), VB, 0, c->stream>>>(n, D->state.p, L->agg.p, 1, D->flag.p); LAUNCH(c);           // leftovers become singletons, numbered behind the roots
  exclusive_scan(c, D, D->flag.p, D->scan.p, n + 1);
  const int nleft = d2h(c, D->scan.p + n);
  if (nleft) { k_amg_number<<<tgrid(n), VB, 0, c->stream>>>(n, D->flag.p, D->scan.p, nroot, L->agg.p, L->root_of.p, 0); LAUNCH(c); }
  L->nc_mine = nroot + nleft;
  (void)lev;
}

// member lists of the aggregates: rows sorted by aggregate (stable radix sort: ascending row order inside an aggregate)
void member_lists(Ctx *c, AmgData *D, AmgLevel *L) {
  const int n = L->n, nc = L->nc_mine;
  D->skey.ensure(n); D->skey2.ensure(n); D->sval.ensure(n); L->mem.ensure(n); L->moff.ensure(nc + 2); D->ccnt.ensure(nc + 2);
  CUDA_CHECK(cudaMemsetAsync(D->ccnt.p, 0, sizeof(int) * (nc + 2), c->stream));
  k_amg_sortkeys<<<tgrid(n), VB, 0, c->stream>>>(n, L->agg.p, nc, D->skey.p, D->sval.p, D->ccnt.p); LAUNCH(c);
  int bits = 1; while ((1ll << bits) <= nc) ++bits;
  size_t tb = 0; cub::DeviceRadixSort::SortPairs(nullptr, tb, D->skey.p, D->skey2.p, D->sval.p, L->mem.p, n, 0, bits, c->stream); D->tmp.ensure(tb);
  tb = D->tmp.cap; cub::DeviceRadixSort::SortPairs(D->tmp.p, tb, D->skey.p, D->skey2.p, D->sval.p, L->mem.p, n, 0, bits, c->stream); LAUNCH(c);
  exclusive_scan(c, D, D->ccnt.p, L->moff.p, nc + 1);
}

// Galerkin operator of level `lev` -> local rows (nc_mine) with global coarse column ids, in D->rlen / Cn->ci / Cn->va (CSR by Cn->rp)
template <class M> bool galerkin(Ctx *c, AmgData *D, AmgLevel *L, AmgLevel *Cn, M A, const int *aggc) {
  const int n = L->n, nc = L->nc_mine; const int KQ = 48, KQ2 = 160, KQ3 = 512;
  D->qcnt.ensure(n); D->qj.ensure((size_t)n * KQ); D->qv.ensure((size_t)n * KQ); D->ctr.ensure(8);
  CUDA_CHECK(cudaMemsetAsync(D->ctr.p + 1, 0, sizeof(int), c->stream));
  k_amg_compress<M, KQ><<<tgrid(n, 128), 128, 0, c->stream>>>(A, n, aggc, L->agg.p, D->qcnt.p, D->qj.p, D->qv.p, D->ctr.p + 1); LAUNCH(c);
  if (d2h(c, D->ctr.p + 1)) {                                             // a row reaches more than 48 aggregates (small aggregates under a wide stencil): the wide variants
    D->qj.ensure((size_t)n * KQ2); D->qv.ensure((size_t)n * KQ2);
    CUDA_CHECK(cudaMemsetAsync(D->ctr.p + 1, 0, sizeof(int), c->stream));
    k_amg_compress<M, KQ2><<<tgrid(n, 128), 128, 0, c->stream>>>(A, n, aggc, L->agg.p, D->qcnt.p, D->qj.p, D->qv.p, D->ctr.p + 1); LAUNCH(c);
    if (d2h(c, D->ctr.p + 1)) {
      if ((size_t)n * KQ3 * 12 > ((size_t)16 << 30)) return false;        // more than 160 aggregates per row on a level too large to give every row 512 slots
      D->qj.ensure((size_t)n * KQ3); D->qv.ensure((size_t)n * KQ3);
      CUDA_CHECK(cudaMemsetAsync(D->ctr.p + 1, 0, sizeof(int), c->stream));
      k_amg_compress<M, KQ3><<<tgrid(n, 128), 128, 0, c->stream>>>(A, n, aggc, L->agg.p, D->qcnt.p, D->qj.p, D->qv.p, D->ctr.p + 1); LAUNCH(c);
      if (d2h(c, D->ctr.p + 1)) return false;
    }
  }
  D->rlen.ensure(nc + 2); Cn->rp.ensure(nc + 2);
  // coarse rows of up to 384 columns: 4 warps per block, 512-slot tables; wider rows (dense small levels): one warp per block, 2048 slots
  bool wide = false;
  for (int attempt = 0; attempt < 2; ++attempt) {
    CUDA_CHECK(cudaMemsetAsync(D->rlen.p, 0, sizeof(int) * (nc + 2), c->stream)); CUDA_CHECK(cudaMemsetAsync(D->ctr.p + 1, 0, sizeof(int), c->stream));
    if (!wide) { k_amg_merge<512, 4, false><<<tgrid(nc, 4), 128, 0, c->stream>>>(nc, n, L->moff.p, L->mem.p, D->qcnt.p, D->qj.p, D->qv.p, D->rlen.p, nullptr, nullptr, nullptr, D->ctr.p + 1); LAUNCH(c); }
    else { k_amg_merge<2048, 1, false><<<tgrid(nc, 1), 32, 0, c->stream>>>(nc, n, L->moff.p, L->mem.p, D->qcnt.p, D->qj.p, D->qv.p, D->rlen.p, nullptr, nullptr, nullptr, D->ctr.p + 1); LAUNCH(c); }
    if (!d2h(c, D->ctr.p + 1)) break;
    if (wide) return false;                                                // a coarse row with more than 1536 columns: stop coarsening here
    wide = true;
  }
  exclusive_scan(c, D, D->rlen.p, Cn->rp.p, nc + 1);
  const long long nnz = d2h(c, Cn->rp.p + nc);
  Cn->ci.ensure(nnz + 1); Cn->va.ensure(nnz + 1);
  if (!wide) { k_amg_merge<512, 4, true><<<tgrid(nc, 4), 128, 0, c->stream>>>(nc, n, L->moff.p, L->mem.p, D->qcnt.p, D->qj.p, D->qv.p, nullptr, Cn->rp.p, Cn->ci.p, Cn->va.p, D->ctr.p + 1); LAUNCH(c); }
  else { k_amg_merge<2048, 1, true><<<tgrid(nc, 1), 32, 0, c->stream>>>(nc, n, L->moff.p, L->mem.p, D->qcnt.p, D->qj.p, D->qv.p, nullptr, Cn->rp.p, Cn->ci.p, Cn->va.p, D->ctr.p + 1); LAUNCH(c); }
  Cn->nnz = nnz;
  return true;
}

}  // namespace

// dense inverse by Gauss-Jordan elimination with partial pivoting — the same routine as oracle/amg_oracle.h::dense_inverse (row-major n x n)
static bool dense_inverse(int n, std::vector<double> &a, std::vector<double> &inv) {
  inv.assign((size_t)n * n, 0.0); for (int i = 0; i < n; ++i) inv[(size_t)i * n + i] = 1.0;
  for (int col = 0; col < n; ++col) {
    int piv = col; double best = std::fabs(a[(size_t)col * n + col]);
    for (int r = col + 1; r < n; ++r) { const double v = std::fabs(a[(size_t)r * n + col]); if (v > best) { best = v; piv = r; } }
    if (best == 0.0) return false;
    if (piv != col) for (int k = 0; k < n; ++k) { std::swap(a[(size_t)col * n + k], a[(size_t)piv * n + k]); std::swap(inv[(size_t)col * n + k], inv[(size_t)piv * n + k]); }
    const double d = 1.0 / a[(size_t)col * n + col];
    for (int k = 0; k < n; ++k) { a[(size_t)col * n + k] *= d; inv[(size_t)col * n + k] *= d; }
    for (int r = 0; r < n; ++r) if (r != col) { const double f = a[(size_t)r * n + col]; if (f == 0.0) continue;
      for (int k = 0; k < n; ++k) { a[(size_t)r * n + k] -= f * a[(size_t)col * n + k]; inv[(size_t)r * n + k] -= f * inv[(size_t)col * n + k]; } }
  }
  return true;
}

void amg_free(Ctx *c) { if (c->amg) c->amg->ready = false; }
void amg_destroy(Ctx *c) { if (c->amg) { cudaSetDevice(c->device); delete c->amg; } c->amg = nullptr; }

void amg_create(Ctx *c) {
  const PrecondParams &pp = c->pp;
  ISPH_REQUIRE(pp.ml_smoother == "Chebyshev" || pp.ml_smoother == "Jacobi",
               "ML stand-in: smoother: type must be Chebyshev or Jacobi (" + pp.ml_smoother + " is sequential within a rank and is not provided)");
  // "coarse: type": Amesos-* = a direct solve of the coarsest operator (explicit inverse, the reference's default Amesos-KLU, precond_ml.h:55);
  // for a singular problem the reference itself replaces it by the smoother (PrecondWrapper_ML::setNullVector, precond_ml.h:118-120, called
  // from solveProblem, solver_lin_belos.h:150-151) — mirrored here.  Otherwise the coarse solver is the smoother.
  const bool coarse_direct = pp.ml_coarse.rfind("Amesos", 0) == 0;
  ISPH_REQUIRE(coarse_direct || pp.ml_coarse == pp.ml_smoother, "ML stand-in: coarse: type must be Amesos-KLU (direct) or equal smoother: type (" + pp.ml_coarse + " is not provided)");
  ISPH_REQUIRE(pp.ml_agg_damping == 0.0, "ML stand-in: aggregation: damping factor must be 0 (non-smoothed aggregation)");
  ISPH_REQUIRE(pp.ml_max_levels >= 1 && pp.ml_max_levels <= 16, "ML stand-in: max levels must be in 1..16");
  if (!c->amg) c->amg = new AmgData();
  AmgData *D = c->amg; D->ready = false; D->setup_ms.clear();
  Matrix &A = c->A; const int n = A.n, ld = c->ld;
  while ((int)D->L.size() < pp.ml_max_levels) D->L.push_back(new AmgLevel());
  D->t0.ensure(ld); D->w0.ensure(ld);
  AmgLevel *F = D->L[0]; F->n = n; F->nnz = A.nnz; F->first = 0; F->nc = F->nc_mine = 0;
  F->gid.ensure(n); F->invdiag.ensure(n); D->d.ensure(ld);
  if (A.external || !c->have_atoms) { ISPH_REQUIRE(c->nranks == 1, "ML stand-in: an external matrix has no global row ids across ranks");
    std::vector<int> g(n); for (int i = 0; i < n; ++i) g[i] = i + 1; CUDA_CHECK(cudaMemcpyAsync(F->gid.p, g.data(), sizeof(int) * n, cudaMemcpyHostToDevice, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream)); }
  else CUDA_CHECK(cudaMemcpyAsync(F->gid.p, c->tag.p, sizeof(int) * n, cudaMemcpyDeviceToDevice, c->stream));
  int nlev = 1;
  for (int lev = 0;; ++lev) {
    AmgLevel *L = D->L[lev]; const int nl = L->n;
    // smoother data of this level
    { PhaseTimer pt(c, D, "eigen");
      if (lev == 0) matrix_extract_diag_dev(c, D->d.p);
      else { D->d.ensure(nl); k_amg_diag<<<tgrid(nl, 128), 128, 0, c->stream>>>(view(L), nl, D->d.p); LAUNCH(c); L->invdiag.ensure(nl); L->xa.ensure(nl); L->xb.ensure(nl); L->b.ensure(nl); L->w.ensure(nl); }
      k_amg_invdiag<<<tgrid(nl), VB, 0, c->stream>>>(D->d.p, L->invdiag.p, nl); LAUNCH(c);
      L->lmax = pp.ml_smoother == "Chebyshev" || pp.ml_coarse == "Chebyshev" ? power_method(c, D, L, lev, pp.ml_eig_iters) : 1.0; }
    L->nc = L->nc_mine = 0;
    long long nl_all = nl;
    if (lev == 0 && c->nranks > 1) { double v = (double)nl; D->red.ensure(592 * 3 + 8); double *dv = D->red.p + 592 * 3 + 4; CUDA_CHECK(cudaMemcpyAsync(dv, &v, sizeof(double), cudaMemcpyHostToDevice, c->stream));
      halo_allreduce(c, dv, 1); nl_all = (long long)d2h(c, dv); }
    if (lev + 1 >= pp.ml_max_levels || nl_all <= pp.ml_max_coarse) break;
    // aggregates
    int maxrow;
    { PhaseTimer pt(c, D, "aggregate");
      if (lev == 0) { maxrow = reduce_max(c, D, A.row_len.p, nl); aggregate(c, D, L, lev, view(A), maxrow, D->d.p, c->have_blocks ? c->block_of_row.p : nullptr); }
      else { D->rlen.ensure(nl + 2); k_amg_rowlen<<<tgrid(nl), VB, 0, c->stream>>>(nl, L->rp.p, D->rlen.p); LAUNCH(c); maxrow = reduce_max(c, D, D->rlen.p, nl); aggregate(c, D, L, lev, view(L), maxrow, D->d.p, nullptr); } }
    // global numbering of the aggregates (finest level across ranks: rank-major, as the rows are)
    std::vector<long long> counts(c->nranks, 0);
    if (lev == 0 && c->nranks > 1) {
      D->cnt2.ensure((size_t)c->nranks * 2 + 2); long long mine = L->nc_mine;
      CUDA_CHECK(cudaMemcpyAsync(D->cnt2.p + c->nranks, &mine, sizeof(long long), cudaMemcpyHostToDevice, c->stream));
      halo_allgather_bytes(c, D->cnt2.p + c->nranks, D->cnt2.p, sizeof(long long));
      CUDA_CHECK(cudaMemcpyAsync(counts.data(), D->cnt2.p, sizeof(long long) * c->nranks, cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream));
    } else counts[0] = L->nc_mine;
    long long nc_all = 0, first = 0; for (int r = 0; r < c->nranks; ++r) { if (r < c->rank) first += counts[r]; nc_all += counts[r]; }
    if (lev > 0) { first = 0; nc_all = L->nc_mine; }
    if (nc_all == 0 || nc_all >= nl_all * 9 / 10) break;                   // no coarsening left: this level is the coarsest
    L->first = (int)first; L->nc = (int)nc_all;
    AmgLevel *Cn = D->L[lev + 1];
    bool ok;
    { PhaseTimer pt(c, D, "galerkin");
      member_lists(c, D, L);
      // aggregate (global coarse index) of every column
      const int ncols = lev == 0 ? A.ncols : nl; D->aggc.ensure(ncols > nl ? ncols : nl);
      if (lev == 0 && c->nranks > 1) {
        D->aggd.ensure(ld);
        k_amg_agg_to_double<<<tgrid(nl), VB, 0, c->stream>>>(nl, L->agg.p, L->first, D->aggd.p); LAUNCH(c);
        halo_exchange(c, D->aggd.p, 1, ld);
        k_amg_double_to_agg<<<tgrid(ncols), VB, 0, c->stream>>>(ncols, D->aggd.p, D->aggc.p); LAUNCH(c);
      } else CUDA_CHECK(cudaMemcpyAsync(D->aggc.p, L->agg.p, sizeof(int) * nl, cudaMemcpyDeviceToDevice, c->stream));
      ok = lev == 0 ? galerkin(c, D, L, Cn, view(A), D->aggc.p) : galerkin(c, D, L, Cn, view(L), D->aggc.p);
      if (lev == 0 && c->nranks > 1) {                                      // the decision to stop coarsening must be the same on every rank (collectives follow)
        double bad = ok ? 0.0 : 1.0; double *dv = D->red.p + 592 * 3 + 4; CUDA_CHECK(cudaMemcpyAsync(dv, &bad, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        halo_allreduce(c, dv, 1); ok = d2h(c, dv) == 0.0;
      }
      if (!ok && c->rank == 0) fprintf(stderr, ">> isph_b200 ML stand-in: level %d is not coarsened further (a row reaches more than 512 aggregates or a coarse row would hold more than %d columns); it becomes the coarsest level\n", lev, 1536);
      if (ok) {
        Cn->gid.ensure(nc_all + 1);
        if (lev == 0 && c->nranks > 1) {
          // replicate level 1: every rank's rows (lengths, ids of the roots, columns, values) all-gathered in padded chunks, then packed rank-major
          const int R = c->nranks; long long mine[2] = {L->nc_mine, Cn->nnz};
          D->cnt2.ensure((size_t)R * 4 + 4);
          CUDA_CHECK(cudaMemcpyAsync(D->cnt2.p + 2 * R, mine, 2 * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
          halo_allgather_bytes(c, D->cnt2.p + 2 * R, D->cnt2.p, 2 * sizeof(long long));
          std::vector<long long> all((size_t)R * 2);
          CUDA_CHECK(cudaMemcpyAsync(all.data(), D->cnt2.p, sizeof(long long) * R * 2, cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream));
          long long maxr = 0, maxz = 0, totz = 0; for (int r = 0; r < R; ++r) { maxr = std::max(maxr, all[2 * r]); maxz = std::max(maxz, all[2 * r + 1]); totz += all[2 * r + 1]; }
          // own chunk: [rlen | gid] ints, [ci] ints, [va] doubles
          D->skey.ensure(2 * maxr + 2); D->gbuf_i.ensure((size_t)R * (2 * maxr + maxz) + 4); D->gbuf_d.ensure((size_t)R * maxz + 4);
          k_amg_gather_int<<<tgrid(L->nc_mine), VB, 0, c->stream>>>(L->nc_mine, L->root_of.p, L->gid.p, D->skey.p + maxr); LAUNCH(c);
          CUDA_CHECK(cudaMemcpyAsync(D->skey.p, D->rlen.p, sizeof(int) * L->nc_mine, cudaMemcpyDeviceToDevice, c->stream));
          int *g_rg = D->gbuf_i.p, *g_ci = D->gbuf_i.p + (size_t)R * 2 * maxr;
          halo_allgather_bytes(c, D->skey.p, g_rg, sizeof(int) * 2 * maxr);
          Cn->ci.ensure(maxz + 1, true); Cn->va.ensure(maxz + 1, true);
          halo_allgather_bytes(c, Cn->ci.p, g_ci, sizeof(int) * maxz);
          halo_allgather_bytes(c, Cn->va.p, D->gbuf_d.p, sizeof(double) * maxz);
          Cn->ci.ensure(totz + 1); Cn->va.ensure(totz + 1); Cn->rp.ensure(nc_all + 2); D->rlen.ensure(nc_all + 2);
          long long ro = 0, zo = 0;
          for (int r = 0; r < R; ++r) {
            const long long nr = all[2 * r], nz = all[2 * r + 1];
            if (nr) { CUDA_CHECK(cudaMemcpyAsync(D->rlen.p + ro, g_rg + (size_t)r * 2 * maxr, sizeof(int) * nr, cudaMemcpyDeviceToDevice, c->stream));
                      CUDA_CHECK(cudaMemcpyAsync(Cn->gid.p + ro, g_rg + (size_t)r * 2 * maxr + maxr, sizeof(int) * nr, cudaMemcpyDeviceToDevice, c->stream)); }
            if (nz) { CUDA_CHECK(cudaMemcpyAsync(Cn->ci.p + zo, g_ci + (size_t)r * maxz, sizeof(int) * nz, cudaMemcpyDeviceToDevice, c->stream));
                      CUDA_CHECK(cudaMemcpyAsync(Cn->va.p + zo, D->gbuf_d.p + (size_t)r * maxz, sizeof(double) * nz, cudaMemcpyDeviceToDevice, c->stream)); }
            ro += nr; zo += nz;
          }
          CUDA_CHECK(cudaMemsetAsync(D->rlen.p + nc_all, 0, sizeof(int), c->stream));
          exclusive_scan(c, D, D->rlen.p, Cn->rp.p, (int)nc_all + 1);
          Cn->nnz = totz;
        } else { k_amg_gather_int<<<tgrid(L->nc_mine), VB, 0, c->stream>>>(L->nc_mine, L->root_of.p, L->gid.p, Cn->gid.p); LAUNCH(c); }
        Cn->n = (int)nc_all;
      }
    }
    if (!ok) { L->nc = L->nc_mine = 0; break; }
    ++nlev;
  }
  D->nlev = nlev; D->direct = false;
  if (coarse_direct && !c->is_singular && nlev > 1 && D->L[nlev - 1]->n > 1024 && c->rank == 0)
    fprintf(stderr, ">> isph_b200 ML stand-in: the coarsest level has %d rows — too many for the direct coarse solve (coarse: type = Amesos-KLU); the smoother is used on it\n", D->L[nlev - 1]->n);
  if (coarse_direct && !c->is_singular && nlev > 1 && D->L[nlev - 1]->n <= 1024) {
    AmgLevel *L = D->L[nlev - 1]; const int nc = L->n;
    std::vector<int> rp(nc + 1), ci(L->nnz); std::vector<double> va(L->nnz), a((size_t)nc * nc, 0.0), inv;
    CUDA_CHECK(cudaMemcpyAsync(rp.data(), L->rp.p, sizeof(int) * (nc + 1), cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaMemcpyAsync(ci.data(), L->ci.p, sizeof(int) * L->nnz, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(va.data(), L->va.p, sizeof(double) * L->nnz, cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < nc; ++i) for (int q = rp[i]; q < rp[i + 1]; ++q) a[(size_t)i * nc + ci[q]] += va[q];
    ISPH_REQUIRE(dense_inverse(nc, a, inv), "ML stand-in: the coarsest operator is singular — a direct coarse solve (coarse: type = Amesos-KLU) needs a non-singular problem");
    D->coarse_inv.ensure((size_t)nc * nc); CUDA_CHECK(cudaMemcpyAsync(D->coarse_inv.p, inv.data(), sizeof(double) * nc * nc, cudaMemcpyHostToDevice, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream));
    D->direct = true;
  }
  D->ready = true;
}

namespace {

struct Cheb { double c1first, s1, delta; };
Cheb cheb_consts(double lmax, double ratio) { const double alpha = lmax / ratio, beta = 1.1 * lmax, delta = 2.0 / (beta - alpha), theta = 0.5 * (beta + alpha); return Cheb{1.0 / theta, theta * delta, delta}; }

// `degree` smoothing steps on a coarse level: b = right-hand side, x = current iterate (ping-pong with `other`), zero = start from x = 0.
// Returns the buffer that holds the result.
double *smooth_csr(Ctx *c, AmgLevel *L, const double *b, double *x, double *other, bool zero, int degree, double ratio, bool jacobi, double damping) {
  const int n = L->n; const int g8 = tgrid((long long)n * 8); CsrView A = view(L);
  if (degree <= 0) { if (zero) CUDA_CHECK(cudaMemsetAsync(x, 0, sizeof(double) * n, c->stream)); return x; }
  if (jacobi) {
    for (int s = 0; s < degree; ++s) {
      if (s == 0 && zero) { k_amg_update<0><<<vgrid(n), VB, 0, c->stream>>>(b, nullptr, L->invdiag.p, damping, 0.0, L->w.p, x, n); LAUNCH(c); continue; }
      k_amg_csr<3><<<g8, VB, 0, c->stream>>>(A, n, x, b, L->invdiag.p, damping, 0.0, nullptr, other); LAUNCH(c); std::swap(x, other);
    }
    return x;
  }
  const Cheb k = cheb_consts(L->lmax, ratio);
  if (zero) { k_amg_update<0><<<vgrid(n), VB, 0, c->stream>>>(b, nullptr, L->invdiag.p, k.c1first, 0.0, L->w.p, x, n); LAUNCH(c); }
  else { k_amg_csr<1><<<g8, VB, 0, c->stream>>>(A, n, x, b, L->invdiag.p, k.c1first, 0.0, L->w.p, other); LAUNCH(c); std::swap(x, other); }
  double rhok = 1.0 / k.s1;
  for (int deg = 0; deg < degree - 1; ++deg) {
    const double rhokp1 = 1.0 / (2.0 * k.s1 - rhok), d1 = rhokp1 * rhok, d2 = 2.0 * rhokp1 * k.delta; rhok = rhokp1;
    k_amg_csr<2><<<g8, VB, 0, c->stream>>>(A, n, x, b, L->invdiag.p, d1, d2, L->w.p, other); LAUNCH(c); std::swap(x, other);
  }
  return x;
}

// V-cycle on the replicated coarse levels: right-hand side in L->b, returns the buffer with the correction
double *vcycle_csr(Ctx *c, AmgData *D, int lev) {
  const PrecondParams &pp = c->pp; AmgLevel *L = D->L[lev]; const int n = L->n; const bool jac = pp.ml_smoother == "Jacobi";
  if (lev + 1 == D->nlev) {
    if (D->direct) { k_amg_dense_apply<<<tgrid((long long)n * 32), VB, 0, c->stream>>>(n, D->coarse_inv.p, L->b.p, L->xa.p); LAUNCH(c); return L->xa.p; }
    return smooth_csr(c, L, L->b.p, L->xa.p, L->xb.p, true, pp.ml_coarse_sweeps, pp.ml_coarse_alpha, jac, pp.ml_damping);
  }
  AmgLevel *Cn = D->L[lev + 1];
  double *x = smooth_csr(c, L, L->b.p, L->xa.p, L->xb.p, true, pp.ml_level_sweeps, pp.ml_level_alpha, jac, pp.ml_damping);
  double *other = x == L->xa.p ? L->xb.p : L->xa.p;
  const double *res = nullptr;
  if (pp.ml_level_sweeps > 0) { k_amg_csr<4><<<tgrid((long long)n * 8), VB, 0, c->stream>>>(view(L), n, x, L->b.p, nullptr, 0.0, 0.0, nullptr, other); LAUNCH(c); res = other; }
  // restriction of res (already r - A x): V = nullptr form with r := res
  k_amg_restrict<<<tgrid(L->nc), VB, 0, c->stream>>>(L->nc, 0, L->nc, L->moff.p, L->mem.p, res ? res : L->b.p, nullptr, Cn->b.p); LAUNCH(c);
  const double *ec = vcycle_csr(c, D, lev + 1);
  k_amg_prolong<<<vgrid(n), VB, 0, c->stream>>>(n, L->agg.p, 0, ec, pp.ml_level_scale, x); LAUNCH(c);
  return smooth_csr(c, L, L->b.p, x, other, false, pp.ml_level_sweeps, pp.ml_level_alpha, jac, pp.ml_damping);
}

}  // namespace

// z = M^-1 r : one V-cycle (ML_Epetra::MultiLevelPreconditioner::ApplyInverse as wrapped by Belos::EpetraPrecOp, solver_lin_belos.h:155)
void amg_apply(Ctx *c, const double *r, double *z) {
  AmgData *D = c->amg; ISPH_REQUIRE(D && D->ready, "ML stand-in: preconditioner not created");
  const PrecondParams &pp = c->pp; AmgLevel *L = D->L[0]; const int n = L->n, ld = c->ld, g = vgrid(n); const bool jac = pp.ml_smoother == "Jacobi";
  double *t = D->t0.p, *W = D->w0.p; const double *inv = L->invdiag.p;
  const Cheb k = cheb_consts(L->lmax, D->nlev == 1 ? pp.ml_coarse_alpha : pp.ml_alpha);      // a hierarchy of one level: the finest level is the coarsest ("coarse: *" parameters)
  auto smooth = [&](bool zero, int degree) {
    if (degree <= 0) { if (zero) CUDA_CHECK(cudaMemsetAsync(z, 0, sizeof(double) * n, c->stream)); return; }
    if (jac) {
      for (int s = 0; s < degree; ++s) {
        if (s == 0 && zero) { k_amg_update<0><<<g, VB, 0, c->stream>>>(r, nullptr, inv, pp.ml_damping, 0.0, W, z, n); LAUNCH(c); continue; }
        spmv(c, z, t, 1, ld, ld); k_amg_update<3><<<g, VB, 0, c->stream>>>(r, t, inv, pp.ml_damping, 0.0, W, z, n); LAUNCH(c);
      }
      return;
    }
    if (zero) { k_amg_update<0><<<g, VB, 0, c->stream>>>(r, nullptr, inv, k.c1first, 0.0, W, z, n); LAUNCH(c); }
    else { spmv(c, z, t, 1, ld, ld); k_amg_update<1><<<g, VB, 0, c->stream>>>(r, t, inv, k.c1first, 0.0, W, z, n); LAUNCH(c); }
    double rhok = 1.0 / k.s1;
    for (int deg = 0; deg < degree - 1; ++deg) {
      spmv(c, z, t, 1, ld, ld);
      const double rhokp1 = 1.0 / (2.0 * k.s1 - rhok), d1 = rhokp1 * rhok, d2 = 2.0 * rhokp1 * k.delta; rhok = rhokp1;
      k_amg_update<2><<<g, VB, 0, c->stream>>>(r, t, inv, d1, d2, W, z, n); LAUNCH(c);
    }
  };
  if (D->nlev == 1) { smooth(true, pp.ml_coarse_sweeps); return; }
  AmgLevel *Cn = D->L[1];
  smooth(true, pp.ml_pre);
  if (pp.ml_pre > 0) spmv(c, z, t, 1, ld, ld);
  k_amg_restrict<<<tgrid(L->nc), VB, 0, c->stream>>>(L->nc, L->first, L->nc_mine, L->moff.p, L->mem.p, r, pp.ml_pre > 0 ? t : nullptr, Cn->b.p); LAUNCH(c);
  if (c->nranks > 1) halo_allreduce(c, Cn->b.p, L->nc);                        // every rank's slice of the level-1 residual: the one exchange of the cycle
  const double *ec = vcycle_csr(c, D, 1);
  k_amg_prolong<<<g, VB, 0, c->stream>>>(n, L->agg.p, L->first, ec, pp.ml_scale, z); LAUNCH(c);
  smooth(false, pp.ml_post);
}

// sizes of the hierarchy of the last create (tests, bench): returns the number of levels
int amg_info(Ctx *c, int *rows, long long *nnz, double *lmax, int cap) {
  AmgData *D = c->amg; if (!D) return 0;
  for (int l = 0; l < D->nlev && l < cap; ++l) { rows[l] = D->L[l]->n; nnz[l] = D->L[l]->nnz; lmax[l] = D->L[l]->lmax; }
  return D->nlev;
}
void amg_aggregates(Ctx *c, int *agg_host) {
  AmgData *D = c->amg; ISPH_REQUIRE(D && D->nlev >= 1, "ML stand-in: no hierarchy");
  AmgLevel *L = D->L[0];
  if (D->nlev == 1) { for (int i = 0; i < L->n; ++i) agg_host[i] = -1; return; }
  CUDA_CHECK(cudaMemcpyAsync(agg_host, L->agg.p, sizeof(int) * L->n, cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < L->n; ++i) if (agg_host[i] >= 0) agg_host[i] += L->first;
}
double amg_setup_ms(Ctx *c, const char *phase) { AmgData *D = c->amg; if (!D) return 0.0; auto it = D->setup_ms.find(phase); return it == D->setup_ms.end() ? 0.0 : it->second; }

}  // namespace isph
