// Block-Jacobi ILU(k): Ifpack::Create("ILU", A, overlap 0) with "fact: level-of-fill" k (precond_ifpack.h:50-75 is the
// call site; Ifpack_ILU itself is third-party and restated in oracle/krylov_oracle.cpp).  k = 0 (BASELINE's ILU(0)) is
// the all-device path; for k > 0 (the reference's own default is 1, precond_ifpack.h:38) the level-of-fill pattern comes
// from a host pass (isph_iluk_symbolic_host, Ifpack_IlukGraph's level rule) and the same numeric kernels run on it.
//   * block = the set of rows one Ifpack/MPI rank would own ("block_of_row", default: all local rows); columns outside
//     the row's block are dropped (Ifpack_LocalFilter), so blocks are independent;
//   * factors in Ifpack's form: L unit-lower (l_ij = a_ij / d_j), D stored inverted, U unit-upper scaled by 1/d_i;
//     row-wise IKJ elimination in ascending column order, multiplier taken before the 1/d_j scaling (Ifpack_ILU::Compute);
//   * apply = L solve, D^-1 scale, U solve (Ifpack_ILU::ApplyInverse).
// GPU execution is level-scheduled: row i of the factorisation / forward solve can run once every row j < i of its L
// pattern is done; rows are bucketed by dependency level and one persistent cooperative kernel walks the levels with a
// grid-wide barrier in between (one warp per row, lanes over the row's entries).  The critical path is the number of
// levels (12 n_block - 11 for an open 3-D lattice block in lexicographic order, = the number of rows for a block that
// is periodic in itself), so this kernel is latency-bound by design — DESIGN.md reports level counts.
#include "isph_internal.h"
#include <cooperative_groups.h>
#include <cub/cub.cuh>
#include <algorithm>
#include <queue>
#include <thread>

namespace cg = cooperative_groups;
#define ILU_TB 1024

namespace isph {

struct IluData {
  int n = 0; long long nnz = 0; int nlev_l = 0, nlev_u = 0, maxw_l = 0, maxw_u = 0;
  DevBuf<int> rp, ci, dpos, order_l, order_u, lptr_l, lptr_u, cnt, lev, hist, rp0, ci0; DevBuf<double> fv, fv0, dinv, y; DevBuf<char> tmp;
  int grid_f = 1, grid_s = 1, maxlen = 0, tb_f = ILU_TB; bool sync_free = true; size_t smem_f = 0; DevBuf<int> fault;
  // level-ordered split copy of the factors for the apply: position idx of the forward (backward) sweep owns the contiguous
  // entries Lrp[idx]..Lrp[idx+1] (Urp..) of row order_l[idx] (order_u[idx]) — no row-pointer / diagonal-position indirection
  DevBuf<int> Lrp, Lci, Urp, Uci, plen; DevBuf<double> Lfv, Ufv, Udinv; bool lv = false; int grid_lv = 1;
  // "Overlap Level" 1 across ranks: the factored problem has n = n_own + nhalo rows (owned rows, then the imported rows of the halo
  // particles in halo-slot order); r / z of an apply are extended with the imported / exported halo part
  bool ext = false; int n_own = 0; OverlapRows ov; DevBuf<double> rext, zext;
  DevBuf<int> slot_ref;   // halo slot referenced by an owned row?  Ifpack_OverlappingRowMatrix extends a rank by the off-rank columns of ITS rows (the column map of A); LAMMPS ghosts that no owned
                          // row reaches (corners of the ghost shell) have a halo slot here but are not part of the overlap: their rows are replaced by identity rows that nothing couples to
};

// ---- block-restricted row-major copy of A --------------------------------------------------------------------------
__global__ void k_ilu_count(const long long *slice_off, const int *row_len, const int *col, const int *blk, int n, int climit, int *cnt) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x; if (row >= n) return;
  const long long base = slice_off[row >> 5] + (row & 31); const int rlen = row_len[row]; const int mb = blk ? blk[row] : 0;
  int c = 0, prev = -1;
  for (int k = 0; k < rlen; ++k) { const int cc = col[base + 32ll * k]; if (cc < climit && cc != prev && (!blk || blk[cc] == mb)) ++c; prev = cc; }
  cnt[row] = c;
}
// rows of the halo particles (overlap 1): sorted (column, value) segments from halo_import_rows; duplicate columns are summed
__global__ void k_ilu_mark_ref(const long long *slice_off, const int *row_len, const int *col, int n_own, int *slot_ref) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x; if (row >= n_own) return;
  const long long base = slice_off[row >> 5] + (row & 31); const int rlen = row_len[row];
  for (int k = 0; k < rlen; ++k) { const int cc = col[base + 32ll * k]; if (cc >= n_own) slot_ref[cc - n_own] = 1; }
}
__global__ void k_ilu_zero_unref(double *rext_halo, const int *slot_ref, int nh) { const int s = blockIdx.x * blockDim.x + threadIdx.x; if (s < nh && !slot_ref[s]) rext_halo[s] = 0.0; }
__global__ void k_ilu_ext_count(const long long *off, const int *cols, const int *slot_ref, int nh, int *cnt) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x; if (s >= nh) return;
  if (!slot_ref[s]) { cnt[s] = 1; return; }
  int c = 0, prev = -1;
  for (long long j = off[s]; j < off[s + 1]; ++j) { const int cc = cols[j]; if (cc == 0x7fffffff) break; if (cc != prev) ++c; prev = cc; }
  cnt[s] = c;
}
__global__ void k_ilu_ext_fill(const long long *off, const int *cols, const double *vals, const int *slot_ref, int nh, int n_own, const int *rp, int *ci, double *fv, int *dpos) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x; if (s >= nh) return;
  const int row = n_own + s; int o = rp[row], prev = -1, dp = -1;
  if (!slot_ref[s]) { ci[o] = row; fv[o] = 1.0; dpos[row] = o; return; }
  for (long long j = off[s]; j < off[s + 1]; ++j) {
    const int cc = cols[j]; if (cc == 0x7fffffff) break;
    if (cc != prev) { ci[o] = cc; fv[o] = vals[j]; if (cc == row) dp = o; ++o; } else fv[o - 1] += vals[j];
    prev = cc;
  }
  dpos[row] = dp;
}
__global__ void k_ilu_fill(const long long *slice_off, const int *row_len, const int *col, const double *val, const int *blk, int n,
                           int climit, const int *rp, int *ci, double *fv, int *dpos) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x; if (row >= n) return;
  const long long base = slice_off[row >> 5] + (row & 31); const int rlen = row_len[row]; const int mb = blk ? blk[row] : 0;
  int o = rp[row], prev = -1, dp = -1;
  for (int k = 0; k < rlen; ++k) {
    const int cc = col[base + 32ll * k];
    if (cc < climit && cc != prev && (!blk || blk[cc] == mb)) { ci[o] = cc; fv[o] = val[base + 32ll * k]; if (cc == row) dp = o; ++o; }
    prev = cc;
  }
  dpos[row] = dp;
}

// values of the block-restricted matrix scattered into the (larger) level-of-fill pattern; fill entries start at zero
__global__ void k_ilu_expand(const int *rp0, const int *ci0, const double *fv0, const int *rp1, const int *ci1, int n, double *fv1, int *dpos1) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x; if (row >= n) return;
  int a = rp0[row], dp = -1; const int ae = rp0[row + 1];
  for (int q = rp1[row]; q < rp1[row + 1]; ++q) {
    const int cc = ci1[q];
    while (a < ae && ci0[a] < cc) ++a;
    fv1[q] = (a < ae && ci0[a] == cc) ? fv0[a] : 0.0;
    if (cc == row) dp = q;
  }
  dpos1[row] = dp;
}

// ---- numeric factorisation, one warp per row, levels separated by grid barriers ------------------------------------
__device__ __forceinline__ int find_col(const int *ci, int lo, int hi, int key) {   // position of key in ci[lo,hi) or -1
  while (lo < hi) { const int mid = (lo + hi) >> 1; const int v = ci[mid]; if (v < key) lo = mid + 1; else hi = mid; }
  return lo;
}
__global__ void __launch_bounds__(ILU_TB) k_ilu_factor(const int *rp, const int *ci, const int *dpos, double *fv, double *dinv,
                                                    const int *order, const int *lptr, int nlev) {
  cg::grid_group grid = cg::this_grid();
  const int lane = threadIdx.x & 31, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int l = 0; l < nlev; ++l) {
    for (int idx = lptr[l] + gw; idx < lptr[l + 1]; idx += nw) {
      const int i = order[idx], b = rp[i], e = rp[i + 1], dp = dpos[i];
      for (int q = b; q < dp; ++q) {                       // strictly-lower entries, ascending column
        const int j = ci[q];
        const double multiplier = fv[q];
        __syncwarp();
        if (lane == 0) fv[q] = multiplier * __ldcg(dinv + j);
        const int ub = dpos[j] + 1, ue = rp[j + 1];
        for (int u = ub + lane; u < ue; u += 32) {
          const int k = ci[u]; const int pos = find_col(ci, q + 1, e, k);
          if (pos < e && ci[pos] == k) fv[pos] -= multiplier * __ldcg(fv + u);      // row j was finished in an earlier level (other SM): read through L2
        }
        __syncwarp();
      }
      const double d = 1.0 / fv[dp];
      __syncwarp();
      if (lane == 0) dinv[i] = d;
      for (int u = dp + 1 + lane; u < e; u += 32) fv[u] *= d;
    }
    __threadfence();
    grid.sync();
  }
}

// ---- apply: L solve (forward levels), D^-1, U solve (backward levels) ----------------------------------------------
__global__ void __launch_bounds__(ILU_TB) k_ilu_solve(const int *rp, const int *ci, const int *dpos, const double *fv, const double *dinv,
                                                   const int *order_l, const int *lptr_l, int nlev_l, const int *order_u, const int *lptr_u, int nlev_u,
                                                   const double *r, double *y, double *z) {
  cg::grid_group grid = cg::this_grid();
  const int lane = threadIdx.x & 31, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int l = 0; l < nlev_l; ++l) {
    for (int idx = lptr_l[l] + gw; idx < lptr_l[l + 1]; idx += nw) {
      const int i = order_l[idx], b = rp[i], dp = dpos[i];
      double s = 0.0;
      for (int q = b + lane; q < dp; q += 32) s += fv[q] * __ldcg(y + ci[q]);
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) y[i] = r[i] - s;
    }
    __threadfence();
    grid.sync();
  }
  for (int l = 0; l < nlev_u; ++l) {
    for (int idx = lptr_u[l] + gw; idx < lptr_u[l + 1]; idx += nw) {
      const int i = order_u[idx], e = rp[i + 1], dp = dpos[i];
      double s = 0.0;
      for (int q = dp + 1 + lane; q < e; q += 32) s += fv[q] * __ldcg(z + ci[q]);
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) z[i] = y[i] * dinv[i] - s;
    }
    __threadfence();
    grid.sync();
  }
}

// ---- synchronisation-free variants (default) -----------------------------------------------------------------------
// No grid barriers: the data is its own "ready" flag.  dinv / y / z are preset to NaN (memset 0xff); a consumer polls the
// value it depends on until it is a number.  Persistent, co-resident warps (cooperative launch) take rows in level order,
// so the lowest unfinished row never waits on an unfinished one: no deadlock.  The critical path becomes
// (#levels x one L2 round trip) instead of (#levels x one grid barrier).  Waits are bounded; a timeout or a computed NaN
// is published as 0 with the fault word raised, so that nobody downstream spins on it.
__device__ int g_ilu_backoff = 0;      // ns to sleep after a failed poll (0: spin); rows far ahead of the wavefront then stop hammering L2
// (ld/st.relaxed.gpu instead of the volatile accesses — LDG/STG .STRONG.GPU instead of .STRONG.SYS in SASS — was measured on the 1M-row / 8-brick
// case: 4.58 ms against 4.47 ms per apply with both behind a run-time switch, i.e. no gain from the narrower scope; the switch itself cost 1.3 ms
// per apply (3.19 ms without it).  ncu source view of this loop: ~21 polls per row on the first dependency of the forward sweep.  Not kept.)
__device__ __forceinline__ double wait_value(const double *p, int *fault) {
  const volatile double *vp = p; double v = *vp; int spins = 0;
  while (v != v) {
    if (g_ilu_backoff) __nanosleep(g_ilu_backoff);
    if ((++spins & 1023) == 0) {                               // once a fault is raised anywhere, every wait drains immediately
      if (*reinterpret_cast<volatile int *>(fault)) return 0.0;
      if (spins > (1 << 22)) { *fault = 1; return 0.0; }
    }
    v = *vp;
  }
  return v;
}
__device__ __forceinline__ double publishable(double v, int *fault) { if (v != v) { *fault = 2; return 0.0; } return v; }

// shared memory per warp: the row's columns and values (maxlen each); row j's U part is read through L2
__global__ void __launch_bounds__(ILU_TB) k_ilu_factor_sf(const int *rp, const int *ci, const int *dpos, double *fv, double *dinv,
                                                          const int *order, int n, int maxlen, int *fault) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  double *s_val = reinterpret_cast<double *>(smem_raw) + (size_t)wib * maxlen;
  int *s_col = reinterpret_cast<int *>(smem_raw + (size_t)(blockDim.x >> 5) * maxlen * sizeof(double)) + (size_t)wib * maxlen;
  for (int idx = gw; idx < n; idx += nw) {
    const int i = order[idx], b = rp[i], len = rp[i + 1] - b, dp = dpos[i] - b;
    for (int t = lane; t < len; t += 32) { s_col[t] = ci[b + t]; s_val[t] = fv[b + t]; }
    __syncwarp();
    for (int q = 0; q < dp; ++q) {                          // strictly-lower entries, ascending column
      const int j = s_col[q];
      double dj = 0.0;
      if (lane == 0) dj = wait_value(dinv + j, fault);      // row j finished (its U part was fenced before dinv[j] was published)
      dj = __shfl_sync(0xffffffffu, dj, 0);
      __threadfence();
      const double multiplier = s_val[q];
      const int ub = dpos[j] + 1, ue = rp[j + 1];
      for (int u = ub + lane; u < ue; u += 32) {
        const int k = __ldcg(ci + u);
        int lo = q + 1, hi = len;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_col[mid] < k) lo = mid + 1; else hi = mid; }
        if (lo < len && s_col[lo] == k) s_val[lo] -= multiplier * __ldcg(fv + u);
      }
      __syncwarp();
      if (lane == 0) s_val[q] = multiplier * dj;
      __syncwarp();
    }
    const double d = 1.0 / s_val[dp];
    __syncwarp();
    for (int t = lane; t < len; t += 32) fv[b + t] = t > dp ? s_val[t] * d : s_val[t];
    __threadfence();
    __syncwarp();
    if (lane == 0) *reinterpret_cast<volatile double *>(dinv + i) = publishable(d, fault);
  }
}


// (Round 2 also measured a factorisation kernel with the pivot metadata (dpos[j], rp[j+1], a first look at dinv[j]) fetched for all
// pivots of a row at once and the next pivot row requested while the current one is applied: 171 ms against 171 ms on the 8M-row /
// 64-block problem, 54 against 51 ms on 1M rows / 8 blocks — the row-to-row dependency wait, not the per-pivot round trips, is what
// the factorisation spends its time in.  Not kept.)

__global__ void __launch_bounds__(ILU_TB) k_ilu_solve_sf(const int *rp, const int *ci, const int *dpos, const double *fv, const double *dinv,
                                                         const int *order_l, const int *order_u, int n, const double *r, double *y, double *z, int *fault) {
  const int lane = threadIdx.x & 31, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int idx = gw; idx < n; idx += nw) {                   // L solve
    const int i = order_l[idx], b = rp[i], dp = dpos[i];
    double s = 0.0;
    for (int q = b + lane; q < dp; q += 32) s += fv[q] * wait_value(y + ci[q], fault);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) *reinterpret_cast<volatile double *>(y + i) = publishable(r[i] - s, fault);
  }
  for (int idx = gw; idx < n; idx += nw) {                   // D^-1 and U solve
    const int i = order_u[idx], e = rp[i + 1], dp = dpos[i];
    double s = 0.0;
    for (int q = dp + 1 + lane; q < e; q += 32) s += fv[q] * wait_value(z + ci[q], fault);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) { const double yi = wait_value(y + i, fault); *reinterpret_cast<volatile double *>(z + i) = publishable(yi * dinv[i] - s, fault); }
  }
}

// ---- level-ordered apply (default) ---------------------------------------------------------------------------------
// The sync-free solve above spends ~4 dependent global loads per row (order -> rp/dpos -> ci/fv -> y[ci]) with one row in
// flight per warp: ~6 us per row and warp, 15 % of the HBM roofline on the 8M-row problem (BENCH r2: ilu_roofline).  Here the
// factors are copied once per factorisation into two arrays in SWEEP order (L parts in forward-level order, U parts + D^-1 in
// backward-level order), so position idx of a sweep finds its entries without the row-pointer / diagonal indirection, and the
// sweep is software-pipelined: while a warp waits for the dependencies of row idx it already holds the entries of row idx + nw
// and the extents of row idx + 2 nw in registers.  Same data-is-the-flag protocol (NaN = not solved yet), same arithmetic order
// within a row (lanes over entries, shuffle tree), so the result is bit-identical to k_ilu_solve_sf.
__global__ void k_ilu_perm_len(const int *rp, const int *dpos, const int *order, int n, int upper, int *len) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x; if (idx >= n) return;
  const int i = order[idx]; len[idx] = upper ? rp[i + 1] - dpos[i] - 1 : dpos[i] - rp[i];
  if (idx == 0) len[n] = 0;
}
__global__ void __launch_bounds__(256) k_ilu_perm_fill(const int *rp, const int *ci, const int *dpos, const double *fv, const double *dinv, const int *order, int n, int upper,
                                                       const int *prp, int *pci, double *pfv, double *pdinv) {
  const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31; if (idx >= n) return;
  const int i = order[idx], src = upper ? dpos[i] + 1 : rp[i], dst = prp[idx], len = prp[idx + 1] - dst;
  for (int t = lane; t < len; t += 32) { pci[dst + t] = ci[src + t]; pfv[dst + t] = fv[src + t]; }
  if (upper && lane == 0) pdinv[idx] = dinv[i];
}
// one sweep: out[row] = (UPPER ? in[row] * dinv : in[row]) - sum_k pfv[k] * out[pci[k]] ; rows in level order
template <bool UPPER> __device__ __forceinline__ void ilu_sweep_lv(const int *__restrict__ prp, const int *__restrict__ pci, const double *__restrict__ pfv, const int *__restrict__ order,
                                                                  const double *__restrict__ pdinv, int n, const double *in, double *out, int *fault, int gw, int nw, int lane) {
  int idx = gw;
  // pipeline registers: row A = current (entries loaded), row B = next (extents loaded)
  int iA = 0, bA = 0, eA = 0, iB = 0, bB = 0, eB = 0; int cA0 = 0, cA1 = 0; double fA0 = 0.0, fA1 = 0.0, rhsA = 0.0, dA = 1.0;
  if (idx < n) { iA = __ldg(order + idx); bA = __ldg(prp + idx); eA = __ldg(prp + idx + 1); }
  if (idx + nw < n) { iB = __ldg(order + idx + nw); bB = __ldg(prp + idx + nw); eB = __ldg(prp + idx + nw + 1); }
  if (idx < n) {
    if (bA + lane < eA) { cA0 = __ldcs(pci + bA + lane); fA0 = __ldcs(pfv + bA + lane); }
    if (bA + lane + 32 < eA) { cA1 = __ldcs(pci + bA + lane + 32); fA1 = __ldcs(pfv + bA + lane + 32); }
    if (lane == 0) { rhsA = UPPER ? wait_value(in + iA, fault) : in[iA]; if (UPPER) dA = __ldg(pdinv + idx); }
  }
  while (idx < n) {
    // (1) extents of row idx + 2 nw, (2) entries of row idx + nw: both in flight while row idx waits for its dependencies
    int iC = 0, bC = 0, eC = 0; const int idxC = idx + 2 * nw;
    if (idxC < n) { iC = __ldg(order + idxC); bC = __ldg(prp + idxC); eC = __ldg(prp + idxC + 1); }
    int cB0 = 0, cB1 = 0; double fB0 = 0.0, fB1 = 0.0, rhsB = 0.0, dB = 1.0;
    if (idx + nw < n) {
      if (bB + lane < eB) { cB0 = __ldcs(pci + bB + lane); fB0 = __ldcs(pfv + bB + lane); }
      if (bB + lane + 32 < eB) { cB1 = __ldcs(pci + bB + lane + 32); fB1 = __ldcs(pfv + bB + lane + 32); }
      if (lane == 0 && !UPPER) rhsB = in[iB];
      if (lane == 0 && UPPER) dB = __ldg(pdinv + idx + nw);
    }
    // row idx: same per-lane accumulation order as k_ilu_solve_sf (q = b + lane, b + lane + 32, ...)
    double s = 0.0;
    if (bA + lane < eA) s += fA0 * wait_value(out + cA0, fault);
    if (bA + lane + 32 < eA) s += fA1 * wait_value(out + cA1, fault);
    for (int q = bA + lane + 64; q < eA; q += 32) s += pfv[q] * wait_value(out + pci[q], fault);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) *reinterpret_cast<volatile double *>(out + iA) = publishable(UPPER ? rhsA * dA - s : rhsA - s, fault);
    // rotate
    if (UPPER && lane == 0 && idx + nw < n) rhsB = wait_value(in + iB, fault);           // y of the next row (forward sweep of this launch: published by some warp)
    iA = iB; bA = bB; eA = eB; cA0 = cB0; cA1 = cB1; fA0 = fB0; fA1 = fB1; rhsA = rhsB; dA = dB;
    iB = iC; bB = bC; eB = eC;
    idx += nw;
  }
}
__global__ void __launch_bounds__(ILU_TB) k_ilu_solve_lv(const int *Lrp, const int *Lci, const double *Lfv, const int *order_l, const int *Urp, const int *Uci, const double *Ufv,
                                                         const int *order_u, const double *Udinv, int n, const double *r, double *y, double *z, int *fault) {
  const int lane = threadIdx.x & 31, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  ilu_sweep_lv<false>(Lrp, Lci, Lfv, order_l, nullptr, n, r, y, fault, gw, nw, lane);
  ilu_sweep_lv<true>(Urp, Uci, Ufv, order_u, Udinv, n, y, z, fault, gw, nw, lane);
}

// (Sub-warp rows — 16 or 8 lanes per row, 2 or 4 rows per warp, to hold a whole wide level in flight — were measured on the 8M-row /
// 64-block problem: 13.9 ms and 21.1 ms per apply against 12.1 ms for the kernel above and 10.3 ms for k_ilu_solve_sf; the spinning
// groups of a warp serialise each other.  Not kept.  Which kernel runs is decided by the width of the levels, see ilu_create.)

// ---- dependency levels on the device -----------------------------------------------------------------------------
// level[i] = 1 + max(level[j] : j in L(i)) (0 without dependencies).  One warp per row in dependency order (ascending
// rows for L, descending for U): lanes read the levels of the row's dependencies and spin until they are published.
// CTAs are dispatched in index order, so a waiting warp only ever waits for rows of CTAs that were dispatched before its
// own (resident or finished) — the scheme of the "synchronisation-free" triangular solvers.  The spin is bounded; on a
// timeout the host-side pass below is used instead.
__global__ void __launch_bounds__(256) k_ilu_levels(const int *rp, const int *ci, const int *dpos, int n, int lower, int *level, int *maxlev, int *fault) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31; if (w >= n) return;
  const int i = lower ? w : n - 1 - w;
  if (dpos[i] < 0) { if (lane == 0) { *fault = 2; level[i] = 0; } return; }          // a row without a diagonal entry
  const int b = lower ? rp[i] : dpos[i] + 1, e = lower ? dpos[i] : rp[i + 1];
  int m = -1;
  for (int q = b + lane; q < e; q += 32) {
    const volatile int *p = level + ci[q]; int l; long long spins = 0;
    while ((l = *p) < 0) { if (++spins > (1ll << 26)) { *fault = 1; l = 0; break; } }
    m = max(m, l);
  }
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) { level[i] = m + 1; __threadfence(); atomicMax(maxlev, m + 1); }
}
__global__ void k_ilu_hist(const int *level, int n, int *hist) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) atomicAdd(hist + level[i] + 1, 1); }
__global__ void k_ilu_scatter(const int *level, int n, int *cursor, int *order) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) order[atomicAdd(cursor + level[i], 1)] = i; }

static void level_sets(int n, const std::vector<int> &rp, const std::vector<int> &ci, const std::vector<int> &dpos, bool lower,
                       std::vector<int> &order, std::vector<int> &lptr, int &maxw) {
  std::vector<int> lev(n, 0); int nlev = 0;
  if (lower) { for (int i = 0; i < n; ++i) { int m = -1; for (int q = rp[i]; q < dpos[i]; ++q) m = std::max(m, lev[ci[q]]); lev[i] = m + 1; nlev = std::max(nlev, lev[i] + 1); } }
  else { for (int i = n - 1; i >= 0; --i) { int m = -1; for (int q = dpos[i] + 1; q < rp[i + 1]; ++q) m = std::max(m, lev[ci[q]]); lev[i] = m + 1; nlev = std::max(nlev, lev[i] + 1); } }
  lptr.assign(nlev + 1, 0);
  for (int i = 0; i < n; ++i) ++lptr[lev[i] + 1];
  maxw = 0; for (int l = 0; l < nlev; ++l) { maxw = std::max(maxw, lptr[l + 1]); lptr[l + 1] += lptr[l]; }
  order.resize(n); std::vector<int> pos(lptr.begin(), lptr.end() - 1);
  for (int i = 0; i < n; ++i) order[pos[lev[i]]++] = i;
}

// Level-of-fill pattern, Ifpack_IlukGraph's rule: entries of the matrix have level 0; eliminating row i with pivot row k
// creates (i,j) for every strictly-upper (k,j) with level(i,k) + level(k,j) + 1, kept when that is <= fill.  Rows ascending,
// pivots of a row in ascending column order (a heap holds the pending lower columns), levels of finished rows kept per entry.
// Level 1 (Ifpack's default) needs no levels of earlier rows: (i,j) is a level-1 entry iff some level-0 pivot k < min(i,j) has
// (i,k) and (k,j) in the matrix, so every row's pattern is the union of its own row and the upper parts of its lower
// neighbours' ORIGINAL rows — rows are independent and are formed by host threads.
static void ilu1_symbolic_parallel(int n, const int *rp, const int *ci, std::vector<int> &frp, std::vector<int> &fci) {
  unsigned T = std::thread::hardware_concurrency(); T = std::max(1u, std::min(T, 32u)); if (n < 4096) T = 1;
  std::vector<std::vector<int>> cols(T), lens(T);
  auto work = [&](unsigned t) {
    const int r0 = (int)((long long)n * t / T), r1 = (int)((long long)n * (t + 1) / T);
    std::vector<int> mark(n, -1), touched;
    for (int i = r0; i < r1; ++i) {
      touched.clear();
      for (int q = rp[i]; q < rp[i + 1]; ++q) { const int cc = ci[q]; if (mark[cc] != i) { mark[cc] = i; touched.push_back(cc); } }
      for (int q = rp[i]; q < rp[i + 1]; ++q) {
        const int k = ci[q]; if (k >= i) continue;
        for (int u = rp[k]; u < rp[k + 1]; ++u) { const int j = ci[u]; if (j > k && mark[j] != i) { mark[j] = i; touched.push_back(j); } }
      }
      std::sort(touched.begin(), touched.end());
      cols[t].insert(cols[t].end(), touched.begin(), touched.end()); lens[t].push_back((int)touched.size());
    }
  };
  std::vector<std::thread> th; for (unsigned t = 1; t < T; ++t) th.emplace_back(work, t);
  work(0); for (auto &x : th) x.join();
  frp.assign(1, 0); fci.clear();
  for (unsigned t = 0; t < T; ++t) { for (int l : lens[t]) frp.push_back(frp.back() + l); fci.insert(fci.end(), cols[t].begin(), cols[t].end()); }
}

static void iluk_symbolic(int n, const int *rp, const int *ci, int fill, std::vector<int> &frp, std::vector<int> &fci) {
  if (fill == 1) { ilu1_symbolic_parallel(n, rp, ci, frp, fci); return; }
  std::vector<unsigned char> flev; std::vector<int> ubeg(n, 0), lev(n, -1), touched;
  std::priority_queue<int, std::vector<int>, std::greater<int>> pend;
  frp.assign(1, 0); fci.clear(); fci.reserve((size_t)rp[n] * (fill + 1));
  for (int i = 0; i < n; ++i) {
    touched.clear();
    for (int q = rp[i]; q < rp[i + 1]; ++q) { const int cc = ci[q]; if (lev[cc] < 0) { lev[cc] = 0; touched.push_back(cc); if (cc < i) pend.push(cc); } }
    while (!pend.empty()) {
      const int k = pend.top(); pend.pop(); const int lk = lev[k];
      for (int u = ubeg[k]; u < frp[k + 1]; ++u) {
        const int nl = lk + flev[u] + 1; if (nl > fill) continue;
        const int j = fci[u];
        if (lev[j] < 0) { lev[j] = nl; touched.push_back(j); if (j < i) pend.push(j); } else if (nl < lev[j]) lev[j] = nl;
      }
    }
    std::sort(touched.begin(), touched.end());
    ubeg[i] = (int)fci.size();
    for (int cc : touched) { if (cc <= i) ++ubeg[i]; fci.push_back(cc); flev.push_back((unsigned char)lev[cc]); lev[cc] = -1; }
    frp.push_back((int)fci.size());
  }
}

static int coop_grid(Ctx *c, const void *fn, int max_width) {
  int per_sm = 0, sms = 0;
  CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, ILU_TB, 0));
  CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
  // at most ONE fat CTA per SM: the cost of a grid-wide barrier grows with the number of CTAs (850 CTAs: ~10 us per
  // level, measured), and a level of an 8-brick 1M-row problem holds only ~1700 rows anyway
  const int cap = std::max(1, std::min(per_sm, 1) * sms), want = std::max(1, (max_width + ILU_TB / 32 - 1) / (ILU_TB / 32));
  return std::min(cap, want);
}

void ilu_create(Ctx *c) {
  Matrix &A = c->A; const int n_own = A.n;
  matrix_merge_duplicates(c);
  if (!c->ilu) c->ilu = new IluData();
  IluData &I = *c->ilu;
  // "Overlap Level" 1 across ranks (validated in precond_create: ILU, no sub-blocks): the rows of the halo particles join the local problem
  const bool ext = c->nranks > 1 && c->pp.overlap >= 1;
  if (ext) {
    c->tic("iluImportRows");
    const int nh = halo_ncols(c) - n_own; I.slot_ref.ensure(nh + 1); CUDA_CHECK(cudaMemsetAsync(I.slot_ref.p, 0, sizeof(int) * (nh + 1), c->stream));
    k_ilu_mark_ref<<<ceil_div(n_own, 128), 128, 0, c->stream>>>(A.slice_off.p, A.row_len.p, A.col.p, n_own, I.slot_ref.p); ++c->launches;
    halo_import_rows(c, &I.ov, I.slot_ref.p);
    c->toc("iluImportRows");
  }
  const int n = n_own + (ext ? I.ov.nhalo : 0), climit = ext ? n : n_own;
  I.n = n; I.n_own = n_own; I.ext = ext;
  const int *blk = (c->have_blocks && !ext) ? c->block_of_row.p : nullptr;
  I.cnt.ensure(n + 1); I.rp.ensure(n + 1); I.dpos.ensure(n); I.dinv.ensure(n); I.y.ensure(std::max(c->ld, n));
  c->tic("iluPattern");
  k_ilu_count<<<ceil_div(n_own, 128), 128, 0, c->stream>>>(A.slice_off.p, A.row_len.p, A.col.p, blk, n_own, climit, I.cnt.p); ++c->launches;
  if (ext && I.ov.nhalo) { k_ilu_ext_count<<<ceil_div(I.ov.nhalo, 128), 128, 0, c->stream>>>(I.ov.off_r.p, I.ov.col_r2.p, I.slot_ref.p, I.ov.nhalo, I.cnt.p + n_own); ++c->launches; }
  CUDA_CHECK(cudaMemsetAsync(I.cnt.p + n, 0, sizeof(int), c->stream));
  size_t tb = 0; cub::DeviceScan::ExclusiveSum(nullptr, tb, I.cnt.p, I.rp.p, n + 1, c->stream);
  I.tmp.ensure(tb);
  cub::DeviceScan::ExclusiveSum(I.tmp.p, tb, I.cnt.p, I.rp.p, n + 1, c->stream); ++c->launches;
  // level 0 (the all-device path): only two integers come back — the number of stored entries and the longest row
  std::vector<int> rp;
  { int *d_max = c->flag.p + 6; size_t tbm = 0; cub::DeviceReduce::Max(nullptr, tbm, I.cnt.p, d_max, n, c->stream); I.tmp.ensure(tbm);
    tbm = I.tmp.cap; cub::DeviceReduce::Max(I.tmp.p, tbm, I.cnt.p, d_max, n, c->stream); ++c->launches;
    int h2[2] = {0, 0};
    CUDA_CHECK(cudaMemcpyAsync(&h2[0], I.rp.p + n, sizeof(int), cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaMemcpyAsync(&h2[1], d_max, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    I.nnz = h2[0]; I.maxlen = std::max(1, h2[1]); }
  if (c->pp.fill > 0) { rp.resize(n + 1); CUDA_CHECK(cudaMemcpyAsync(rp.data(), I.rp.p, sizeof(int) * (n + 1), cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream)); }
  I.ci.ensure(I.nnz); I.fv.ensure(I.nnz);
  I.sync_free = !getenv("ISPH_ILU_BARRIER"); I.fault.ensure(4); CUDA_CHECK(cudaMemsetAsync(I.fault.p, 0, 4 * sizeof(int), c->stream));
  k_ilu_fill<<<ceil_div(n_own, 128), 128, 0, c->stream>>>(A.slice_off.p, A.row_len.p, A.col.p, A.val.p, blk, n_own, climit, I.rp.p, I.ci.p, I.fv.p, I.dpos.p); ++c->launches;
  if (ext && I.ov.nhalo) { k_ilu_ext_fill<<<ceil_div(I.ov.nhalo, 128), 128, 0, c->stream>>>(I.ov.off_r.p, I.ov.col_r2.p, I.ov.val_r2.p, I.slot_ref.p, I.ov.nhalo, n_own, I.rp.p, I.ci.p, I.fv.p, I.dpos.p); ++c->launches; }
  if (c->pp.fill > 0) {                                          // level-of-fill pattern (host), values expanded on the device
    std::vector<int> ci0(I.nnz), frp, fci;
    CUDA_CHECK(cudaMemcpyAsync(ci0.data(), I.ci.p, sizeof(int) * I.nnz, cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream));
    iluk_symbolic(n, rp.data(), ci0.data(), c->pp.fill, frp, fci);
    std::swap(I.rp0, I.rp); std::swap(I.ci0, I.ci); std::swap(I.fv0, I.fv);               // level-0 arrays become the source of the expansion
    I.nnz = frp[n]; I.rp.ensure(n + 1); I.ci.ensure(I.nnz); I.fv.ensure(I.nnz);
    CUDA_CHECK(cudaMemcpyAsync(I.rp.p, frp.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(I.ci.p, fci.data(), sizeof(int) * I.nnz, cudaMemcpyHostToDevice, c->stream));
    k_ilu_expand<<<ceil_div(n, 128), 128, 0, c->stream>>>(I.rp0.p, I.ci0.p, I.fv0.p, I.rp.p, I.ci.p, n, I.fv.p, I.dpos.p); ++c->launches;
    CUDA_CHECK(cudaStreamSynchronize(c->stream));               // frp / fci are host temporaries
    rp = frp;
    I.maxlen = 1; for (int i = 0; i < n; ++i) I.maxlen = std::max(I.maxlen, rp[i + 1] - rp[i]);
  }
  c->toc("iluPattern"); c->tic("iluLevels");
  // dependency levels + level sets, on the device (no download of the pattern)
  I.order_l.ensure(n); I.order_u.ensure(n); I.lev.ensure(n); I.hist.ensure(n + 2); I.lptr_l.ensure(n + 2); I.lptr_u.ensure(n + 2);
  bool device_ok = !getenv("ISPH_ILU_HOST_LEVELS");
  for (int pass = 0; pass < 2 && device_ok; ++pass) {
    const int lower = pass == 0; int *order = lower ? I.order_l.p : I.order_u.p, *lptr = lower ? I.lptr_l.p : I.lptr_u.p;
    CUDA_CHECK(cudaMemsetAsync(I.lev.p, 0xff, sizeof(int) * n, c->stream));
    CUDA_CHECK(cudaMemsetAsync(c->flag.p, 0, 4 * sizeof(int), c->stream));                 // [0] max level, [1] fault
    k_ilu_levels<<<ceil_div((long long)n * 32, 256), 256, 0, c->stream>>>(I.rp.p, I.ci.p, I.dpos.p, n, lower, I.lev.p, c->flag.p, c->flag.p + 1); ++c->launches;
    int h[2]; CUDA_CHECK(cudaMemcpyAsync(h, c->flag.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream));
    ISPH_REQUIRE(h[1] != 2, "ILU: a row has no diagonal entry");
    if (h[1]) { device_ok = false; break; }
    const int nlev = h[0] + 1;
    CUDA_CHECK(cudaMemsetAsync(I.hist.p, 0, sizeof(int) * (nlev + 1), c->stream));
    k_ilu_hist<<<ceil_div(n, 256), 256, 0, c->stream>>>(I.lev.p, n, I.hist.p); ++c->launches;
    size_t tb2 = 0; cub::DeviceScan::InclusiveSum(nullptr, tb2, I.hist.p, lptr, nlev + 1, c->stream); I.tmp.ensure(tb2);
    cub::DeviceScan::InclusiveSum(I.tmp.p, tb2, I.hist.p, lptr, nlev + 1, c->stream); ++c->launches;     // lptr[l] = first position of level l
    std::vector<int> hh(nlev + 1);
    CUDA_CHECK(cudaMemcpyAsync(hh.data(), I.hist.p, sizeof(int) * (nlev + 1), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(I.hist.p, lptr, sizeof(int) * (nlev + 1), cudaMemcpyDeviceToDevice, c->stream));   // cursors
    k_ilu_scatter<<<ceil_div(n, 256), 256, 0, c->stream>>>(I.lev.p, n, I.hist.p, order); ++c->launches;
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    int mw = 0; for (int l = 1; l <= nlev; ++l) mw = std::max(mw, hh[l]);
    if (lower) { I.nlev_l = nlev; I.maxw_l = mw; } else { I.nlev_u = nlev; I.maxw_u = mw; }
  }
  if (!device_ok) {   // host pass over the pattern (fallback)
    std::vector<int> ci(I.nnz), dpos(n);
    if (rp.empty()) { rp.resize(n + 1); CUDA_CHECK(cudaMemcpyAsync(rp.data(), I.rp.p, sizeof(int) * (n + 1), cudaMemcpyDeviceToHost, c->stream)); }
    CUDA_CHECK(cudaMemcpyAsync(ci.data(), I.ci.p, sizeof(int) * I.nnz, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(dpos.data(), I.dpos.p, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n; ++i) ISPH_REQUIRE(dpos[i] >= 0, "ILU: a row has no diagonal entry");
    std::vector<int> ol, pl, ou, pu;
    level_sets(n, rp, ci, dpos, true, ol, pl, I.maxw_l); level_sets(n, rp, ci, dpos, false, ou, pu, I.maxw_u);
    I.nlev_l = (int)pl.size() - 1; I.nlev_u = (int)pu.size() - 1;
    CUDA_CHECK(cudaMemcpyAsync(I.order_l.p, ol.data(), sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(I.order_u.p, ou.data(), sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(I.lptr_l.p, pl.data(), sizeof(int) * pl.size(), cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(I.lptr_u.p, pu.data(), sizeof(int) * pu.size(), cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
  }
  c->toc("iluLevels");
  if (I.sync_free) {
    int sms = 0, per_sm = 0; CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    { const char *e = getenv("ISPH_ILU_BACKOFF"); const int bo = (e && *e) ? atoi(e) : 0; CUDA_CHECK(cudaMemcpyToSymbolAsync(g_ilu_backoff, &bo, sizeof(int), 0, cudaMemcpyHostToDevice, c->stream)); }
    // factorisation: per-warp staging of the row in shared memory (12 B per entry); shrink the CTA until it fits
    const size_t per_entry = 12; const void *fk = (const void *)k_ilu_factor_sf;
    I.tb_f = ILU_TB; while (I.tb_f > 32 && (size_t)(I.tb_f / 32) * I.maxlen * per_entry > 200 * 1024) I.tb_f >>= 1;
    I.smem_f = (size_t)(I.tb_f / 32) * I.maxlen * per_entry; ISPH_REQUIRE(I.smem_f <= 200 * 1024, "ILU: a row is too long for the factorisation kernel");
    CUDA_CHECK(cudaFuncSetAttribute(fk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)I.smem_f));
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fk, I.tb_f, I.smem_f)); ISPH_REQUIRE(per_sm >= 1, "ILU: factorisation kernel does not fit an SM");
    I.grid_f = std::max(1, std::min(per_sm * sms, ceil_div((long long)n * 32, I.tb_f)));
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ilu_solve_sf, ILU_TB, 0)); ISPH_REQUIRE(per_sm >= 1, "ILU: solve kernel does not fit an SM");
    I.grid_s = std::max(1, std::min(per_sm * sms, ceil_div((long long)n * 32, ILU_TB)));
    CUDA_CHECK(cudaMemsetAsync(I.dinv.p, 0xff, sizeof(double) * n, c->stream));           // NaN = "row not factored yet"
    const int *rpp = I.rp.p, *cip = I.ci.p, *dpp = I.dpos.p, *ord = I.order_l.p; double *fv = I.fv.p, *dinv = I.dinv.p; int nn = n, ml = I.maxlen; int *flt = I.fault.p;
    void *args[] = {&rpp, &cip, &dpp, &fv, &dinv, &ord, &nn, &ml, &flt};
    c->tic("iluFactor");
    CUDA_CHECK(cudaLaunchCooperativeKernel(fk, dim3(I.grid_f), dim3(I.tb_f), args, I.smem_f, c->stream)); ++c->launches;
    c->toc("iluFactor"); c->tic("iluPermute");
    // level-ordered split copy for the apply (stream order: after the factorisation)
    // narrow levels (fewer rows per level than resident warps: the 1M-row / 8-block case, 1600 rows per level) are latency-bound per
    // row and gain from the level-ordered pipelined kernel (3.2 ms against 5.2 ms per apply); wide levels (8M rows / 64 blocks, 12500
    // rows per level) are bound by the level-to-level hand-over and run faster on the plain kernel (10.3 against 12.1 ms).
    // ISPH_ILU_APPLY = sf | lv forces one of them.
    { const double rows_per_level = (double)n / std::max(1, std::max(I.nlev_l, I.nlev_u)); const int resident_warps = 2 * sms * (ILU_TB / 32);
      I.lv = rows_per_level < 0.5 * resident_warps;
      const char *e = getenv("ISPH_ILU_APPLY"); if (e && !strcmp(e, "sf")) I.lv = false; else if (e && !strcmp(e, "lv")) I.lv = true; }
    if (I.lv) {
      I.plen.ensure(n + 1); I.Lrp.ensure(n + 1); I.Urp.ensure(n + 1); I.Lci.ensure(I.nnz); I.Uci.ensure(I.nnz); I.Lfv.ensure(I.nnz); I.Ufv.ensure(I.nnz); I.Udinv.ensure(n);
      for (int upper = 0; upper < 2; ++upper) {
        const int *ord = upper ? I.order_u.p : I.order_l.p; int *prp = upper ? I.Urp.p : I.Lrp.p;
        k_ilu_perm_len<<<ceil_div(n, 256), 256, 0, c->stream>>>(I.rp.p, I.dpos.p, ord, n, upper, I.plen.p);
        size_t tb3 = 0; cub::DeviceScan::ExclusiveSum(nullptr, tb3, I.plen.p, prp, n + 1, c->stream); I.tmp.ensure(tb3);
        tb3 = I.tmp.cap; cub::DeviceScan::ExclusiveSum(I.tmp.p, tb3, I.plen.p, prp, n + 1, c->stream);
        k_ilu_perm_fill<<<ceil_div((long long)n * 32, 256), 256, 0, c->stream>>>(I.rp.p, I.ci.p, I.dpos.p, I.fv.p, I.dinv.p, ord, n, upper, prp, upper ? I.Uci.p : I.Lci.p,
                                                                                upper ? I.Ufv.p : I.Lfv.p, I.Udinv.p);
        c->launches += 3;
      }
      CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ilu_solve_lv, ILU_TB, 0)); ISPH_REQUIRE(per_sm >= 1, "ILU: level-ordered solve kernel does not fit an SM");
      I.grid_lv = std::max(1, std::min(per_sm * sms, ceil_div((long long)n * 32, ILU_TB)));
    }
    c->toc("iluPermute");
    return;
  }
  I.grid_f = coop_grid(c, (const void *)k_ilu_factor, I.maxw_l);
  I.grid_s = coop_grid(c, (const void *)k_ilu_solve, std::max(I.maxw_l, I.maxw_u));
  const int *rpp = I.rp.p, *cip = I.ci.p, *dpp = I.dpos.p, *ord = I.order_l.p, *lp = I.lptr_l.p; double *fv = I.fv.p, *dinv = I.dinv.p; int nlev = I.nlev_l;
  void *args[] = {&rpp, &cip, &dpp, &fv, &dinv, &ord, &lp, &nlev};
  CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)k_ilu_factor, dim3(I.grid_f), dim3(ILU_TB), args, 0, c->stream)); ++c->launches;
}

void ilu_free(Ctx *c) { (void)c; /* buffers are grow-only and reused by the next create() (rebuilt every solve, solver_lin_belos.h:153,190) */ }

bool ilu_fault(Ctx *c) {
  if (!c->ilu || !c->ilu->fault.p) return false;
  int f = 0; cudaMemcpy(&f, c->ilu->fault.p, sizeof(int), cudaMemcpyDeviceToHost); return f != 0;
}

void ilu_info(Ctx *c, long long *nnz, int *nlev_l, int *nlev_u, int *maxlen) {
  *nnz = 0; *nlev_l = *nlev_u = *maxlen = 0; if (!c->ilu) return;
  *nnz = c->ilu->nnz; *nlev_l = c->ilu->nlev_l; *nlev_u = c->ilu->nlev_u; *maxlen = c->ilu->maxlen;
}

static void ilu_apply_launch(Ctx *c, const double *r, double *z);
static void ilu_apply_core(Ctx *c, const double *r, double *z);
void ilu_apply(Ctx *c, const double *r, double *z) {
  IluData &I = *c->ilu;
  if (!I.ext) { ilu_apply_core(c, r, z); return; }
  // additive Schwarz with one level of overlap, combine mode Add (Ifpack_AdditiveSchwarz::ApplyInverse): import r at the halo rows,
  // solve the extended local problem, keep the owned part and add the halo part of the solution into the owners' rows
  I.rext.ensure(I.n + 32); I.zext.ensure(I.n + 32);
  CUDA_CHECK(cudaMemcpyAsync(I.rext.p, r, sizeof(double) * I.n_own, cudaMemcpyDeviceToDevice, c->stream));
  halo_exchange(c, I.rext.p, 1, I.n);
  if (I.ov.nhalo) { k_ilu_zero_unref<<<ceil_div(I.ov.nhalo, 256), 256, 0, c->stream>>>(I.rext.p + I.n_own, I.slot_ref.p, I.ov.nhalo); ++c->launches; }   // not part of the overlap: solves to 0, exports 0
  ilu_apply_core(c, I.rext.p, I.zext.p);
  CUDA_CHECK(cudaMemcpyAsync(z, I.zext.p, sizeof(double) * I.n_own, cudaMemcpyDeviceToDevice, c->stream));
  halo_export_add(c, I.zext.p + I.n_own, z);
}
static void ilu_apply_core(Ctx *c, const double *r, double *z) {
  if (!c->prof_spmv) { ilu_apply_launch(c, r, z); return; }
  // per-launch device timing on the launching stream (bench.py: achieved GB/s of the triangular solves)
  if (c->pprof_used + 2 > c->pprof_ev.size()) { const size_t o = c->pprof_ev.size(); c->pprof_ev.resize(o + 512, nullptr); for (size_t q = o; q < c->pprof_ev.size(); ++q) CUDA_CHECK(cudaEventCreate(&c->pprof_ev[q])); }
  cudaEvent_t e0 = c->pprof_ev[c->pprof_used++], e1 = c->pprof_ev[c->pprof_used++];
  CUDA_CHECK(cudaEventRecord(e0, c->stream)); ilu_apply_launch(c, r, z); CUDA_CHECK(cudaEventRecord(e1, c->stream));
}
static void ilu_apply_launch(Ctx *c, const double *r, double *z) {
  IluData &I = *c->ilu;
  if (I.sync_free && I.lv) {
    CUDA_CHECK(cudaMemsetAsync(I.y.p, 0xff, sizeof(double) * I.n, c->stream)); CUDA_CHECK(cudaMemsetAsync(z, 0xff, sizeof(double) * I.n, c->stream));   // NaN = "not solved yet"
    const int *a0 = I.Lrp.p, *a1 = I.Lci.p, *a3 = I.order_l.p, *a4 = I.Urp.p, *a5 = I.Uci.p, *a7 = I.order_u.p; const double *a2 = I.Lfv.p, *a6 = I.Ufv.p, *a8 = I.Udinv.p; double *y = I.y.p; int nn = I.n; int *flt = I.fault.p;
    void *args[] = {&a0, &a1, &a2, &a3, &a4, &a5, &a6, &a7, &a8, &nn, &r, &y, &z, &flt};
    CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)k_ilu_solve_lv, dim3(I.grid_lv), dim3(ILU_TB), args, 0, c->stream)); ++c->launches;
    return;
  }
  if (I.sync_free) {
    CUDA_CHECK(cudaMemsetAsync(I.y.p, 0xff, sizeof(double) * I.n, c->stream)); CUDA_CHECK(cudaMemsetAsync(z, 0xff, sizeof(double) * I.n, c->stream));   // NaN = "not solved yet"
    const int *rpp = I.rp.p, *cip = I.ci.p, *dpp = I.dpos.p, *ol = I.order_l.p, *ou = I.order_u.p; const double *fv = I.fv.p, *dinv = I.dinv.p; double *y = I.y.p; int nn = I.n; int *flt = I.fault.p;
    void *args[] = {&rpp, &cip, &dpp, &fv, &dinv, &ol, &ou, &nn, &r, &y, &z, &flt};
    CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)k_ilu_solve_sf, dim3(I.grid_s), dim3(ILU_TB), args, 0, c->stream)); ++c->launches;
    return;
  }
  const int *rpp = I.rp.p, *cip = I.ci.p, *dpp = I.dpos.p, *ol = I.order_l.p, *pl = I.lptr_l.p, *ou = I.order_u.p, *pu = I.lptr_u.p;
  const double *fv = I.fv.p, *dinv = I.dinv.p; double *y = I.y.p; int nl = I.nlev_l, nu = I.nlev_u;
  void *args[] = {&rpp, &cip, &dpp, &fv, &dinv, &ol, &pl, &nl, &ou, &pu, &nu, &r, &y, &z};
  CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)k_ilu_solve, dim3(I.grid_s), dim3(ILU_TB), args, 0, c->stream)); ++c->launches;
}

void ilu_destroy(Ctx *c) {
  if (!c->ilu) return; IluData &I = *c->ilu;
  I.rp.release(); I.ci.release(); I.dpos.release(); I.order_l.release(); I.order_u.release(); I.lptr_l.release(); I.lptr_u.release(); I.cnt.release();
  I.ov.release(); I.rext.release(); I.zext.release(); I.slot_ref.release();
  I.Lrp.release(); I.Lci.release(); I.Urp.release(); I.Uci.release(); I.plen.release(); I.Lfv.release(); I.Ufv.release(); I.Udinv.release();
  I.fv.release(); I.dinv.release(); I.y.release(); I.tmp.release(); I.rp0.release(); I.ci0.release(); I.fv0.release(); I.lev.release(); I.hist.release(); I.fault.release(); delete c->ilu; c->ilu = nullptr;
}

}  // namespace isph

extern "C" {
// Pure host code (no CUDA): the level-of-fill pattern of ILU(k) for a CSR pattern with ascending columns (CPU-side test
// hook).  Returns ISPH_FAILURE when `cap` is too small; *nnz_out is set either way.
int isph_iluk_symbolic_host(int n, const int *rowptr, const int *col, int fill, int *rowptr_out, int *col_out, long long cap, long long *nnz_out) {
  if (n < 0 || !rowptr || !col || fill < 0 || fill > 255 || !nnz_out) return ISPH_FAILURE;
  std::vector<int> frp, fci; isph::iluk_symbolic(n, rowptr, col, fill, frp, fci);
  *nnz_out = (long long)fci.size();
  if (!rowptr_out || !col_out || cap < (long long)fci.size()) return ISPH_FAILURE;
  memcpy(rowptr_out, frp.data(), sizeof(int) * (n + 1)); memcpy(col_out, fci.data(), sizeof(int) * fci.size());
  return ISPH_SUCCESS;
}
}
