// Krylov solvers behind SolverLin_Belos::solveProblem (solver_lin_belos.h:130-222) with the parameter list of
// setParameters (:224-264): right-preconditioned (flexible) GMRES(m) with DGKS orthogonalisation and PCG, the
// PoissonProjection operator (solver_lin.h:130-140) for singular problems, null-vector handling of
// solver_lin.cpp:59-77.  Belos itself is third-party code that is not vendored with the reference; its semantics are
// restated in oracle/krylov_oracle.cpp (header there) and this file implements the same algorithm on the device:
//   * every vector operation is a fused, grid-stride kernel with warp-shuffle reductions; per-block partial sums are
//     combined in block order by the last block to finish (deterministic, no atomics on doubles);
//   * all Krylov scalars (Hessenberg column, Givens rotations, DGKS decision, alpha/beta of CG) stay on the device;
//     the Hessenberg least-squares update runs in a single warp; the host only reads the implicit residual, one
//     iteration late, from pinned memory, so the GPU never waits for the host inside a restart cycle;
//   * with several ranks the partial results are summed with one ncclAllReduce per reduction (halo.cu).
// No tensor cores: nothing here is a dense contraction (largest dense object: the 51x50 Hessenberg).
#include "isph_internal.h"

namespace isph {

static const int VB = 256;              // threads per block of the vector kernels
static const double DEP_TOL = 0.70710678118654752440;   // DGKSOrthoManager dep_tol = 1/sqrt(2)

// layout of the small device scalar block `hbuf`
enum { S_H = 0, S_H2 = 64, S_G = 128, S_CS = 192, S_SN = 256, S_Y = 320, S_OLD = 384, S_NEW1 = 385, S_NEW2 = 386, S_PROJ = 387,
       S_INV = 388, S_RES = 389, S_ALPHA = 390, S_BETA = 391, S_RZ = 392, S_PAP = 393, S_TMP = 394, S_HM = 448 /* H: 64 x 64 */, S_TOTAL = 448 + 64 * 64 };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-reduce NACC per-thread accumulators, store the block's partials, and let the last block to finish add the
// partials of all blocks (fixed assignment of blocks to lanes + fixed shuffle tree => deterministic) into out[0..nacc)
template <int NACC> __device__ void reduce_finish(double (&acc)[NACC], int nacc, double *partials, unsigned *counter, double *out, const P2PRed &pr) {
  __shared__ double sm[NACC][VB / 32];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NACC; ++k) { const double v = warp_sum(acc[k]); if (lane == 0) sm[k][warp] = v; }
  __syncthreads();
  if (threadIdx.x < nacc) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < VB / 32; ++w) s += sm[threadIdx.x][w];
    partials[(size_t)blockIdx.x * NACC + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last) {
    __threadfence();
    for (int k = warp; k < nacc; k += VB / 32) {
      double s = 0.0;
      for (unsigned b = lane; b < gridDim.x; b += 32) s += __ldcg(partials + (size_t)b * NACC + k);
      s = warp_sum(s);
      if (lane == 0) out[k] = s;
    }
    if (threadIdx.x == 0) *counter = 0u;
    if (pr.nranks > 1) p2p_allreduce_block(pr, out, nacc);      // sum over ranks through the NVLink mailboxes
  }
}

// the pre-projection squared norm of pass 0 is stored right behind the nv coefficients (one allreduce covers both)
__device__ __forceinline__ bool dgks_second(const double *S, int nv) { return S[S_NEW1] < DEP_TOL * S[S_H + nv]; }

// out[0] = sum a_i b_i (b == nullptr: a_i a_i)
__global__ void __launch_bounds__(VB) k_dot(const double *a, const double *b, int n, double *partials, unsigned *counter, double *out, P2PRed pr) {
  double acc[1] = {0.0};
  for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) acc[0] += a[i] * (b ? b[i] : a[i]);
  reduce_finish<1>(acc, 1, partials, counter, out, pr);
}

// Classical Gram-Schmidt coefficients.  grid = (row chunks, vector groups): block (bx, g) owns a CONTIGUOUS chunk of
// rows and the G basis vectors [gG, gG+G): every thread streams one w value pair and G basis value pairs per step
// (128-bit loads), keeps G+1 accumulators in registers, and a block touches only G+1 pages — the first version, where
// every thread walked all <= 51 vectors, ran at ~1.4 TB/s (profiles/r01_launches_c2_first.txt).
//   pass 0: h[k] = V_k . w' (k < nv), h[nv] = w'.w'  with  w' = w - proj * nvec  (PoissonProjection tail on the fly)
//   pass 1: skipped unless the DGKS test asks for a second pass; h2[k] = V_k . w
template <int G> __global__ void __launch_bounds__(VB)
k_multidot(const double *__restrict__ V, int ld, int nv, const double *__restrict__ w, const double *__restrict__ nvec, int n,
           double *S, int pass, double *partials, unsigned *counters, P2PRed pr) {
  if (pass == 1 && !dgks_second(S, nv)) return;
  const int g = blockIdx.y, k0 = g * G, cnt = min(G, nv - k0);
  const double proj = (pass == 0 && nvec) ? S[S_PROJ] : 0.0;
  int chunk = (n + gridDim.x - 1) / gridDim.x; chunk = (chunk + 1) & ~1;
  const int r0 = blockIdx.x * chunk, r1 = min(n, r0 + chunk);
  const double *Vg = V + (size_t)k0 * ld;
  double acc[G + 1];
#pragma unroll
  for (int k = 0; k <= G; ++k) acc[k] = 0.0;
  for (int i = r0 + 2 * threadIdx.x; i < r1; i += 2 * VB) {
    double2 wi;
    if (i + 1 < r1) { wi = *reinterpret_cast<const double2 *>(w + i); if (pass == 0 && nvec) { const double2 nn = *reinterpret_cast<const double2 *>(nvec + i); wi.x -= proj * nn.x; wi.y -= proj * nn.y; } }
    else { wi.x = w[i]; if (pass == 0 && nvec) wi.x -= proj * nvec[i]; wi.y = 0.0; }
    // all G loads are issued unconditionally and up front (vectors past the group's end alias its last one: L1 hits,
    // results discarded) so that 16 independent 128-bit loads are in flight per thread; a predicated load/use chain
    // here made the kernel latency-bound (~3 TB/s)
    double2 v[G];
    if (i + 1 < r1) {
#pragma unroll
      for (int k = 0; k < G; ++k) v[k] = *reinterpret_cast<const double2 *>(Vg + (size_t)min(k, cnt - 1) * ld + i);
    } else {
#pragma unroll
      for (int k = 0; k < G; ++k) { v[k].x = Vg[(size_t)min(k, cnt - 1) * ld + i]; v[k].y = 0.0; }
    }
#pragma unroll
    for (int k = 0; k < G; ++k) { acc[k] += v[k].x * wi.x; acc[k] += v[k].y * wi.y; }
    if (g == 0) { acc[G] += wi.x * wi.x; acc[G] += wi.y * wi.y; }
  }
  __shared__ double sm[G + 1][VB / 32];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k <= G; ++k) { const double v = warp_sum(acc[k]); if (lane == 0) sm[k][warp] = v; }
  __syncthreads();
  double *mine = partials + ((size_t)g * gridDim.x + blockIdx.x) * (G + 1);
  if (threadIdx.x <= G) { double s = 0.0; for (int q = 0; q < VB / 32; ++q) s += sm[threadIdx.x][q]; mine[threadIdx.x] = s; }
  __threadfence(); __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(counters + g, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last) {
    __threadfence();
    const double *grp = partials + (size_t)g * gridDim.x * (G + 1);
    for (int k = warp; k <= G; k += VB / 32) {
      if (!(k < cnt || (k == G && g == 0 && pass == 0))) continue;
      double s = 0.0;
      for (unsigned b = lane; b < gridDim.x; b += 32) s += __ldcg(grp + (size_t)b * (G + 1) + k);
      s = warp_sum(s);
      if (lane == 0) { if (k == G) S[S_H + nv] = s; else S[(pass == 0 ? S_H : S_H2) + k0 + k] = s; }
    }
    if (threadIdx.x == 0) counters[g] = 0u;
    if (pr.nranks > 1) {                                   // the last group to finish exchanges h[0..nv) (+ the norm) with the peers
      __shared__ bool all_done;
      __threadfence(); __syncthreads();
      if (threadIdx.x == 0) { all_done = (atomicAdd(counters + 6, 1u) == gridDim.y - 1); if (all_done) counters[6] = 0u; }   // flag word 15
      __syncthreads();
      if (all_done) { __threadfence(); p2p_allreduce_block(pr, S + (pass == 0 ? S_H : S_H2), pass == 0 ? nv + 1 : nv); }
    }
  }
}

// w <- w' - sum_k h_k V_k ; new = ||w||^2.  One contiguous row chunk per block, 128-bit accesses.
__global__ void __launch_bounds__(VB)
k_cgs_update(const double *__restrict__ V, int ld, int nv, double *__restrict__ w, const double *__restrict__ nvec, int n,
             double *S, int pass, double *partials, unsigned *counter, int rev, P2PRed pr) {
  if (pass == 1 && !dgks_second(S, nv)) return;
  __shared__ double sh[64];
  if (threadIdx.x < nv) sh[threadIdx.x] = S[(pass == 0 ? S_H : S_H2) + threadIdx.x];
  __syncthreads();
  const double proj = (pass == 0 && nvec) ? S[S_PROJ] : 0.0;
  int chunk = (n + gridDim.x - 1) / gridDim.x; chunk = (chunk + 1) & ~1;
  (void)rev;   // reversed sweeps (to catch the tail of the previous sweep in L2) were measured: no gain on B200, and the
               // indexed loop they need costs ~40 % on this kernel
  const int r0 = blockIdx.x * chunk, r1 = min(n, r0 + chunk);
  double acc[1] = {0.0};
  for (int i = r0 + 2 * threadIdx.x; i < r1; i += 2 * VB) {
    if (i + 1 < r1) {
      double2 wi = *reinterpret_cast<const double2 *>(w + i);
      if (pass == 0 && nvec) { const double2 nn = *reinterpret_cast<const double2 *>(nvec + i); wi.x -= proj * nn.x; wi.y -= proj * nn.y; }
#pragma unroll 8
      for (int k = 0; k < nv; ++k) { const double2 v = *reinterpret_cast<const double2 *>(V + (size_t)k * ld + i); wi.x -= sh[k] * v.x; wi.y -= sh[k] * v.y; }
      *reinterpret_cast<double2 *>(w + i) = wi; acc[0] += wi.x * wi.x; acc[0] += wi.y * wi.y;
    } else {
      double wi = w[i]; if (pass == 0 && nvec) wi -= proj * nvec[i];
      for (int k = 0; k < nv; ++k) wi -= sh[k] * V[(size_t)k * ld + i];
      w[i] = wi; acc[0] += wi * wi;
    }
  }
  reduce_finish<1>(acc, 1, partials, counter, S + (pass == 0 ? S_NEW1 : S_NEW2), pr);
}

// Fused first Gram-Schmidt update + second-pass coefficients: ONE sweep over the basis instead of two.
//   w1 = w' - sum_k h_k V_k   (pass-0 update, ||w1||^2 -> S_NEW1)   and   h2[k] = V_k . w1   (pass-1 dots, speculative:
//   whether the DGKS test uses them is decided afterwards by k_cgs_update(pass 1) / k_givens from S_NEW1).
// Each block owns a contiguous row chunk and walks it in tiles of UT rows: the tile of all nv basis vectors is staged in
// shared memory (<= 52 x 64 x 8 B = 26 KB), w1 of the tile is formed from it, and the same staged values feed the dots.
static const int UT = 64;
__global__ void __launch_bounds__(VB)
k_update_dot(const double *__restrict__ V, int ld, int nv, double *__restrict__ w, const double *__restrict__ nvec, int n,
             double *S, double *partials, unsigned *counter) {
  extern __shared__ __align__(16) double smem[];
  double *tile = smem;                     // [nv][UT]
  double *w1 = smem + (size_t)nv * UT;     // [UT]
  double *psum = w1 + UT;                  // [VB/UT][UT]
  __shared__ double sh[64], red[VB / 32];
  __shared__ bool last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < nv) sh[tid] = S[S_H + tid];
  const double proj = nvec ? S[S_PROJ] : 0.0;
  int chunk = (n + gridDim.x - 1) / gridDim.x; chunk = (chunk + UT - 1) / UT * UT;
  const int r0 = blockIdx.x * chunk, r1 = min(n, r0 + chunk);
  double acc[7] = {0, 0, 0, 0, 0, 0, 0}, nd = 0.0;
  __syncthreads();
  for (int t0 = r0; t0 < r1; t0 += UT) {
    const int tl = min(UT, r1 - t0);
    for (int e = tid; e < nv * (UT / 2); e += VB) {            // stage the tile, 128-bit loads (t0 and ld are even)
      const int k = e / (UT / 2), r = 2 * (e % (UT / 2));
      double2 v = make_double2(0.0, 0.0);
      if (r + 1 < tl) v = *reinterpret_cast<const double2 *>(V + (size_t)k * ld + t0 + r); else if (r < tl) v.x = V[(size_t)k * ld + t0 + r];
      *reinterpret_cast<double2 *>(tile + (size_t)k * UT + r) = v;
    }
    __syncthreads();
    { const int r = tid & (UT - 1), gq = tid / UT;               // 4 thread groups share the k loop of one row
      double s = 0.0;
      for (int k = gq; k < nv; k += VB / UT) s += sh[k] * tile[(size_t)k * UT + r];
      psum[gq * UT + r] = s; }
    __syncthreads();
    if (tid < UT) {
      double wi = 0.0;
      if (tid < tl) { wi = w[t0 + tid]; if (nvec) wi -= proj * nvec[t0 + tid]; double ps = 0.0; for (int q = 0; q < VB / UT; ++q) ps += psum[q * UT + tid]; wi -= ps; w[t0 + tid] = wi; nd += wi * wi; }
      w1[tid] = wi;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 7; ++q) { const int k = warp + 8 * q; if (k < nv) { for (int r = lane; r < UT; r += 32) acc[q] += tile[(size_t)k * UT + r] * w1[r]; } }
    __syncthreads();
  }
  double *mine = partials + (size_t)blockIdx.x * 64;
#pragma unroll
  for (int q = 0; q < 7; ++q) { const int k = warp + 8 * q; const double v = warp_sum(acc[q]); if (lane == 0 && k < nv) mine[k] = v; }
  nd = warp_sum(nd); if (lane == 0) red[warp] = nd;
  __syncthreads();
  if (tid == 0) { double t = 0.0; for (int q = 0; q < (UT + 31) / 32; ++q) t += red[q]; mine[63] = t; }
  __threadfence(); __syncthreads();
  if (tid == 0) last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last) {
    __threadfence();
    for (int k = warp; k < 64; k += VB / 32) {
      if (!(k < nv || k == 63)) continue;
      double s = 0.0;
      for (unsigned b = lane; b < gridDim.x; b += 32) s += __ldcg(partials + (size_t)b * 64 + k);
      s = warp_sum(s);
      if (lane == 0) { if (k == 63) { S[S_NEW1] = s; S[S_H2 + nv] = s; } else S[S_H2 + k] = s; }
    }
    if (tid == 0) *counter = 0u;
  }
}

// Hessenberg column j: DGKS bookkeeping, Givens rotations, implicit residual (BlockGmresIter::updateLSQR); one warp:
// lanes stage the column and the rotations in shared memory, lane 0 runs the (inherently sequential) recurrence there
__global__ void k_givens(double *S, int j, int m, double *host_res, int slot) {
  __shared__ double h[64], cs[64], sn[64];
  const int lane = threadIdx.x;
  const bool second = dgks_second(S, j + 1);
  for (int k = lane; k <= j; k += 32) { h[k] = S[S_H + k] + (second ? S[S_H2 + k] : 0.0); cs[k] = S[S_CS + k]; sn[k] = S[S_SN + k]; }
  __syncwarp();
  if (lane == 0) {
    const double newDot = second ? S[S_NEW2] : S[S_NEW1];
    const double hn = sqrt(newDot);
    S[S_INV] = hn > 0.0 ? 1.0 / hn : 0.0;
    h[j + 1] = hn;
    for (int k = 0; k < j; ++k) {                       // previous rotations
      const double a = h[k], b = h[k + 1];
      h[k] = cs[k] * a + sn[k] * b; h[k + 1] = -sn[k] * a + cs[k] * b;
    }
    const double a = h[j], b = h[j + 1], rr = hypot(a, b);
    const double c_ = rr == 0.0 ? 1.0 : a / rr, s_ = rr == 0.0 ? 0.0 : b / rr;
    h[j] = rr; h[j + 1] = 0.0;
    const double gj = S[S_G + j];
    S[S_CS + j] = c_; S[S_SN + j] = s_; S[S_G + j + 1] = -s_ * gj; S[S_G + j] = c_ * gj;
    const double res = fabs(s_ * gj);
    S[S_RES] = res;
    host_res[slot] = res;
    __threadfence_system();
  }
  __syncwarp();
  for (int k = lane; k <= j; k += 32) S[S_HM + k * 64 + j] = h[k];
}

// y = H^-1 g for the first ncol columns
__global__ void k_backsolve(double *S, int ncol) {
  if (threadIdx.x != 0) return;
  double *g = S + S_G, *y = S + S_Y, *H = S + S_HM;
  for (int k = ncol - 1; k >= 0; --k) { double s = g[k]; for (int l = k + 1; l < ncol; ++l) s -= H[k * 64 + l] * y[l]; y[k] = s / H[k * 64 + k]; }
}

// x += sum_k y_k Z_k
__global__ void __launch_bounds__(VB) k_update_x(double *x, const double *Z, int ld, int ncol, const double *S, int n) {
  __shared__ double sy[64];
  if (threadIdx.x < ncol) sy[threadIdx.x] = S[S_Y + threadIdx.x];
  __syncthreads();
  for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) {
    double xi = x[i];
#pragma unroll 8
    for (int k = 0; k < ncol; ++k) xi += sy[k] * Z[(size_t)k * ld + i];
    x[i] = xi;
  }
}
// t = sum_k y_k V_k
__global__ void __launch_bounds__(VB) k_combine(double *t, const double *V, int ld, int ncol, const double *S, int n) {
  __shared__ double sy[64];
  if (threadIdx.x < ncol) sy[threadIdx.x] = S[S_Y + threadIdx.x];
  __syncthreads();
  for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) {
    double s = 0.0;
    for (int k = 0; k < ncol; ++k) s += sy[k] * V[(size_t)k * ld + i];
    t[i] = s;
  }
}

// v <- w * S[S_INV] (in place) and, for Jacobi, z <- damping * invdiag * v in the same pass
__global__ void __launch_bounds__(VB) k_normalize_prec(double *w, const double *S, const double *invdiag, double damping, double *z, int n) {
  const double s = S[S_INV];
  for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) { const double v = w[i] * s; w[i] = v; if (z) z[i] = invdiag ? damping * invdiag[i] * v : v; }
}
// r = b - t (t may be null: r = b) ; out = ||r||^2
__global__ void __launch_bounds__(VB) k_residual(const double *b, const double *t, double *r, int n, double *partials, unsigned *counter, double *out, P2PRed pr) {
  double acc[1] = {0.0};
  for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) { const double v = b[i] - (t ? t[i] : 0.0); r[i] = v; acc[0] += v * v; }
  reduce_finish<1>(acc, 1, partials, counter, out, pr);
}
// v0 = r / beta ; g = (beta, 0, ...)
__global__ void __launch_bounds__(VB) k_start_cycle(const double *r, double *v0, double *S, double beta, int n) {
  if (blockIdx.x == 0 && threadIdx.x < 64) S[S_G + threadIdx.x] = threadIdx.x == 0 ? beta : 0.0;
  for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) v0[i] = r[i] / beta;
}
// y <- y - (*coef) * nvec  (projection tail / x,b clean-up) ; sign -1 uses +coef
__global__ void __launch_bounds__(VB) k_axpy_dev(double *y, const double *x, const double *coef, double sign, int n) {
  const double a = sign * (*coef);
  for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) y[i] += a * x[i];
}
__global__ void __launch_bounds__(VB) k_fill(double *y, double v, int n) { for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) y[i] = v; }
__global__ void __launch_bounds__(VB) k_mask_to_vec(const int *mask, double *nv, int n) { for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) nv[i] = mask ? (double)mask[i] : 1.0; }
__global__ void __launch_bounds__(VB) k_scale_by(double *y, double s, int n) { for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) y[i] *= s; }
__global__ void __launch_bounds__(VB) k_random(double *y, const int *tag, int n, int salt) {     // Epetra Random() stand-in: per-tag hash in (-1,1)
  for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) {
    unsigned long long z = (unsigned long long)(tag ? tag[i] : i + 1) + 0x9E3779B97F4A7C15ULL * (unsigned long long)(salt + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; z = z ^ (z >> 31);
    y[i] = 2.0 * ((double)(z >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
  }
}

// ---- PCG kernels -------------------------------------------------------------------------------------------------
// pAp = p.Ap ; alpha = rz / pAp   (alpha computed by whoever consumes it, so that an allreduce can sit in between)
__global__ void __launch_bounds__(VB) k_cg_update(double *x, double *r, const double *p, const double *Ap, double *S, int n, double *partials, unsigned *counter, P2PRed pr) {
  const double alpha = S[S_RZ] / S[S_PAP];
  double acc[1] = {0.0};
  for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) { x[i] += alpha * p[i]; const double ri = r[i] - alpha * Ap[i]; r[i] = ri; acc[0] += ri * ri; }
  reduce_finish<1>(acc, 1, partials, counter, S + S_TMP, pr);
}
__global__ void k_cg_publish(double *S, double *host_res, int slot) { if (threadIdx.x == 0) { const double res = sqrt(S[S_TMP]); S[S_RES] = res; host_res[slot] = res; __threadfence_system(); } }
// z = damping * invdiag * r (Jacobi) fused with rz_new = r.z ; for other preconditioners z is given and only the dot is taken
__global__ void __launch_bounds__(VB) k_cg_precdot(const double *r, double *z, const double *invdiag, double damping, int n, double *partials, unsigned *counter, double *out, P2PRed pr) {
  double acc[1] = {0.0};
  for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) { double zi; if (invdiag) { zi = damping * invdiag[i] * r[i]; z[i] = zi; } else zi = z[i]; acc[0] += r[i] * zi; }
  reduce_finish<1>(acc, 1, partials, counter, out, pr);
}
// p = z + beta p with beta = rz_new / rz ; then rz <- rz_new (done by block 0 after everyone has read it: separate tiny kernel)
__global__ void __launch_bounds__(VB) k_cg_direction(double *p, const double *z, const double *S, int n, int first) {
  const double beta = first ? 0.0 : S[S_BETA] / S[S_RZ];
  for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) p[i] = first ? z[i] : z[i] + beta * p[i];
}
__global__ void k_cg_shift(double *S) { if (threadIdx.x == 0) S[S_RZ] = S[S_BETA]; }

// ---------------------------------------------------------------------------------------------------------------
static int vgrid(Ctx *c, int n) { (void)c; int g = ceil_div(n, VB); return g < 592 ? (g < 1 ? 1 : g) : 592; }   // 148 SMs x 4 CTAs

void solver_prepare_vectors(Ctx *c) {
  const int need = c->A.ncols > c->A.n ? c->A.ncols : c->A.n;
  c->ld = (need + 31) / 32 * 32;
}


static void dot_dev(Ctx *c, const double *a, const double *b, int n, double *out) {
  P2PRed pr = halo_p2p_ticket(c);
  k_dot<<<vgrid(c, n), VB, 0, c->stream>>>(a, b, n, c->red.p, (unsigned *)c->flag.p + 8, out, pr); ++c->launches;
  if (c->nranks > 1 && pr.nranks <= 1) halo_allreduce(c, out, 1);
}
static double read_scalar(Ctx *c, const double *d) {
  double v; CUDA_CHECK(cudaMemcpyAsync(c->h_scal.p, d, sizeof(double), cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream));
  v = c->h_scal.p[0]; return v;
}

// operator apply: y = A x, or PoissonProjection::Apply  y = A x ; y -= (y.n) n  (solver_lin.h:130-140).
// With `defer` the projection coefficient is left in S[S_PROJ] for the orthogonalisation kernels to apply on the fly.
static void op_apply(Ctx *c, const double *x, double *y, bool defer) {
  double *S = c->hbuf.p;
  spmv(c, x, y, 1, c->ld, c->ld, c->is_singular ? c->nullvec.p : nullptr, S + S_PROJ);     // (y.n) reduced in the SpMV epilogue
  if (c->is_singular) {
    if (!defer) { k_axpy_dev<<<vgrid(c, c->A.n), VB, 0, c->stream>>>(y, c->nullvec.p, S + S_PROJ, -1.0, c->A.n); ++c->launches; }
  }
}

static void apply_prec(Ctx *c, bool use_prec, const double *r, double *z) {
  if (use_prec) precond_apply(c, r, z);
  else CUDA_CHECK(cudaMemcpyAsync(z, r, sizeof(double) * c->A.n, cudaMemcpyDeviceToDevice, c->stream));
}

static void launch_multidot(Ctx *c, const double *V, int nv, const double *w, int pass) {
  const int n = c->A.n; double *S = c->hbuf.p; const double *nv_ = c->is_singular ? c->nullvec.p : nullptr; unsigned *cnt = (unsigned *)c->flag.p + 9;
  const int G = nv <= 4 ? 4 : (nv <= 8 ? 8 : 16);            // short bases: do not pay for 16 (aliased) loads per thread
  const int groups = (nv + G - 1) / G;
  int gx = 592 / groups; if (gx < 148) gx = 148; { const int mx = ceil_div(n, 2 * VB); if (gx > mx) gx = mx < 1 ? 1 : mx; }
  P2PRed pr = halo_p2p_ticket(c);
  if (G == 4) k_multidot<4><<<dim3(gx, groups), VB, 0, c->stream>>>(V, c->ld, nv, w, nv_, n, S, pass, c->red.p, cnt, pr);
  else if (G == 8) k_multidot<8><<<dim3(gx, groups), VB, 0, c->stream>>>(V, c->ld, nv, w, nv_, n, S, pass, c->red.p, cnt, pr);
  else k_multidot<16><<<dim3(gx, groups), VB, 0, c->stream>>>(V, c->ld, nv, w, nv_, n, S, pass, c->red.p, cnt, pr);
  ++c->launches;
  if (c->nranks > 1 && pr.nranks <= 1) halo_allreduce(c, S + (pass == 0 ? S_H : S_H2), pass == 0 ? nv + 1 : nv);
}

static void dbg(Ctx *c, const char *what) {      // ISPH_DEBUG_SYNC=1: synchronise after every phase and name the one that faulted
  static const bool on = getenv("ISPH_DEBUG_SYNC") != nullptr; if (!on) return;
  cudaError_t e = cudaStreamSynchronize(c->stream); if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("device fault after ") + what + ": " + cudaGetErrorString(e));
}

static int gmres_solve(Ctx *c, bool use_prec, double *x, double *b, int *iters_out, double *relres_out) {
  const int n = c->A.n, ld = c->ld, m = c->sp.num_blocks; const bool flex = c->sp.flexible;
  ISPH_REQUIRE(m >= 1 && m <= 51, "Num Blocks must be in 1..51");
  double *S = c->hbuf.p, *V = c->V.p, *Z = c->Z.p, *r = c->wk.p; unsigned *cnt = (unsigned *)c->flag.p + 8;
  const double *nvp = c->is_singular ? c->nullvec.p : nullptr;
  const bool jacobi_fused = use_prec && c->prec_kind == 1 && c->pp.sweeps == 1;
  // fused update+second-pass dots (k_update_dot): measured 109-141 us vs 48+45 us for the two streaming kernels on B200
  // (the staged tile is not pipelined), so it stays opt-in until it carries a cp.async double buffer
  static const int opt_fuse = getenv("ISPH_FUSE") ? 1 : 0;
  const int opt_rev = 0;   // reversed sweeps were measured: no L2 reuse gain on B200 (both dies stream concurrently), kept off
  std::vector<cudaEvent_t> ev(m);
  for (auto &e : ev) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  int iters = 0, restarts = 0; bool converged = false, first = true; double scale = 0.0, res = 0.0;
  const int g = vgrid(c, n);
  while (true) {
    // r = b - Op x ; beta = ||r||
    if (!(first && c->init_type == ISPH_INIT_ZERO)) op_apply(c, x, V + (size_t)ld, false);
    { P2PRed pr = halo_p2p_ticket(c);
      k_residual<<<g, VB, 0, c->stream>>>(b, (first && c->init_type == ISPH_INIT_ZERO) ? nullptr : V + (size_t)ld, r, n, c->red.p, cnt, S + S_TMP, pr); ++c->launches;
      if (c->nranks > 1 && pr.nranks <= 1) halo_allreduce(c, S + S_TMP, 1); }
    const double beta = sqrt(read_scalar(c, S + S_TMP));
    if (first) { scale = beta; first = false; }
    res = beta;
    if (scale == 0.0 || res / scale <= c->sp.tol) { converged = true; break; }
    k_start_cycle<<<g, VB, 0, c->stream>>>(r, V, S, beta, n); ++c->launches;
    // z_0 = M^-1 v_0
    apply_prec(c, use_prec, V, Z);
    int ncol = 0; bool stop = false;
    int j = 0;
    for (; j < m; ++j) {
      double *zj = flex ? Z + (size_t)j * ld : Z, *vn = V + (size_t)(j + 1) * ld;
      dbg(c, "prologue");
      { ProfScope ps(c, "op_apply"); op_apply(c, zj, vn, true); } dbg(c, "op_apply");         // w = A z_j (projection coefficient deferred)
      { ProfScope ps(c, "multidot0"); launch_multidot(c, V, j + 1, vn, 0); } dbg(c, "multidot0");
      if (opt_fuse) {
        ProfScope ps(c, "update0+dot1");
        const int nvj = j + 1; const size_t sm = ((size_t)nvj * UT + (1 + VB / UT) * UT) * sizeof(double);
        k_update_dot<<<g, VB, sm, c->stream>>>(V, ld, nvj, vn, nvp, n, S, c->red.p, cnt); ++c->launches;
        if (c->nranks > 1) { halo_allreduce(c, S + S_H2, nvj + 1); CUDA_CHECK(cudaMemcpyAsync(S + S_NEW1, S + S_H2 + nvj, sizeof(double), cudaMemcpyDeviceToDevice, c->stream)); }
      } else {
        { ProfScope ps(c, "update0"); P2PRed pr = halo_p2p_ticket(c); k_cgs_update<<<g, VB, 0, c->stream>>>(V, ld, j + 1, vn, nvp, n, S, 0, c->red.p, cnt, opt_rev, pr); ++c->launches; if (c->nranks > 1 && pr.nranks <= 1) halo_allreduce(c, S + S_NEW1, 1); }
        { ProfScope ps(c, "multidot1"); launch_multidot(c, V, j + 1, vn, 1); }
      }
      dbg(c, "update0/dot1");
      { ProfScope ps(c, "update1"); P2PRed pr = halo_p2p_ticket(c); k_cgs_update<<<g, VB, 0, c->stream>>>(V, ld, j + 1, vn, nullptr, n, S, 1, c->red.p, cnt, opt_rev, pr); ++c->launches; if (c->nranks > 1 && pr.nranks <= 1) halo_allreduce(c, S + S_NEW2, 1); }
      dbg(c, "update1");
      { ProfScope ps(c, "givens"); k_givens<<<1, 32, 0, c->stream>>>(S, j, m, c->h_scal.p + 8, iters + 1); ++c->launches; }
      dbg(c, "givens");
      CUDA_CHECK(cudaEventRecord(ev[j], c->stream));
      ++iters;
      if (j + 1 < m) {                                           // prepare the next Arnoldi step before looking at the residual
        ProfScope ps(c, "normalize_prec");
        double *zn = flex ? Z + (size_t)(j + 1) * ld : Z;
        if (jacobi_fused) { k_normalize_prec<<<g, VB, 0, c->stream>>>(vn, S, c->invdiag.p, c->pp.damping, zn, n); ++c->launches; }
        else { k_normalize_prec<<<g, VB, 0, c->stream>>>(vn, S, nullptr, 1.0, nullptr, n); ++c->launches; apply_prec(c, use_prec, vn, zn); }
      }
      // look at the residual of the PREVIOUS step (already finished on the device): no pipeline bubble
      if (j >= 1) {
        CUDA_CHECK(cudaEventSynchronize(ev[j - 1]));
        res = c->h_scal.p[8 + iters - 1];
        if (res / scale <= c->sp.tol) { converged = true; stop = true; ncol = j; --iters; break; }
        if (iters - 1 >= c->sp.max_iters) { stop = true; ncol = j; --iters; break; }
      }
    }
    if (!stop) {                                                  // last step of the cycle (or m == 1)
      CUDA_CHECK(cudaEventSynchronize(ev[m - 1]));
      res = c->h_scal.p[8 + iters]; ncol = m;
      if (res / scale <= c->sp.tol) { converged = true; stop = true; }
      else if (iters >= c->sp.max_iters) stop = true;
    }
    // x += Z y (flexible) or x += M^-1 (V y)
    k_backsolve<<<1, 32, 0, c->stream>>>(S, ncol); ++c->launches;
    if (flex) { k_update_x<<<g, VB, 0, c->stream>>>(x, Z, ld, ncol, S, n); ++c->launches; }
    else { k_combine<<<g, VB, 0, c->stream>>>(r, V, ld, ncol, S, n); ++c->launches; apply_prec(c, use_prec, r, Z);
           k_fill<<<1, 32, 0, c->stream>>>(S + S_TMP, 1.0, 1); ++c->launches;
           k_axpy_dev<<<g, VB, 0, c->stream>>>(x, Z, S + S_TMP, 1.0, n); ++c->launches; }
    if (stop) break;
    if (restarts >= c->sp.max_restarts) break;
    ++restarts;
  }
  for (auto &e : ev) cudaEventDestroy(e);
  *iters_out = iters; *relres_out = scale > 0.0 ? res / scale : 0.0;
  return converged ? 1 : 0;
}

static int cg_solve(Ctx *c, bool use_prec, double *x, double *b, int *iters_out, double *relres_out) {
  const int n = c->A.n, ld = c->ld, g = vgrid(c, n);
  double *S = c->hbuf.p, *r = c->wk.p, *z = c->V.p, *p = c->V.p + (size_t)ld, *Ap = c->V.p + (size_t)2 * ld; unsigned *cnt = (unsigned *)c->flag.p + 8;
  const bool jacobi_fused = use_prec && c->prec_kind == 1 && c->pp.sweeps == 1;
  cudaEvent_t ev[2]; for (auto &e : ev) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  // R = b - A x ; Z = M^-1 R ; P = Z ; rz = R.Z
  if (c->init_type != ISPH_INIT_ZERO) op_apply(c, x, Ap, false);
  { P2PRed pr = halo_p2p_ticket(c);
    k_residual<<<g, VB, 0, c->stream>>>(b, c->init_type == ISPH_INIT_ZERO ? nullptr : Ap, r, n, c->red.p, cnt, S + S_TMP, pr); ++c->launches;
    if (c->nranks > 1 && pr.nranks <= 1) halo_allreduce(c, S + S_TMP, 1); }
  const double scale = sqrt(read_scalar(c, S + S_TMP)); double res = scale; int iters = 0; bool converged = false;
  if (scale == 0.0 || res / scale <= c->sp.tol) converged = true;
  else {
    if (!jacobi_fused) apply_prec(c, use_prec, r, z);
    { P2PRed pr = halo_p2p_ticket(c);
      k_cg_precdot<<<g, VB, 0, c->stream>>>(r, z, jacobi_fused ? c->invdiag.p : nullptr, jacobi_fused ? c->pp.damping : 1.0, n, c->red.p, cnt, S + S_RZ, pr); ++c->launches;
      if (c->nranks > 1 && pr.nranks <= 1) halo_allreduce(c, S + S_RZ, 1); }
    k_cg_direction<<<g, VB, 0, c->stream>>>(p, z, S, n, 1); ++c->launches;
    while (true) {
      ++iters;
      op_apply(c, p, Ap, false);
      dot_dev(c, p, Ap, n, S + S_PAP);
      { P2PRed pr = halo_p2p_ticket(c); k_cg_update<<<g, VB, 0, c->stream>>>(x, r, p, Ap, S, n, c->red.p, cnt, pr); ++c->launches; if (c->nranks > 1 && pr.nranks <= 1) halo_allreduce(c, S + S_TMP, 1); }
      k_cg_publish<<<1, 32, 0, c->stream>>>(S, c->h_scal.p + 8, iters); ++c->launches;
      CUDA_CHECK(cudaEventRecord(ev[iters & 1], c->stream));
      // next direction, enqueued before the residual of this step is inspected
      if (!jacobi_fused) apply_prec(c, use_prec, r, z);
      { P2PRed pr = halo_p2p_ticket(c);
        k_cg_precdot<<<g, VB, 0, c->stream>>>(r, z, jacobi_fused ? c->invdiag.p : nullptr, jacobi_fused ? c->pp.damping : 1.0, n, c->red.p, cnt, S + S_BETA, pr); ++c->launches;
        if (c->nranks > 1 && pr.nranks <= 1) halo_allreduce(c, S + S_BETA, 1); }
      k_cg_direction<<<g, VB, 0, c->stream>>>(p, z, S, n, 0); ++c->launches;
      k_cg_shift<<<1, 32, 0, c->stream>>>(S); ++c->launches;
      CUDA_CHECK(cudaEventSynchronize(ev[iters & 1]));
      res = c->h_scal.p[8 + iters];
      if (res / scale <= c->sp.tol) { converged = true; break; }
      if (iters >= c->sp.max_iters) break;
    }
  }
  for (auto &e : ev) cudaEventDestroy(e);
  *iters_out = iters; *relres_out = scale > 0.0 ? res / scale : 0.0;
  return converged ? 1 : 0;
}

// SolverLin_Belos::solveProblem, solver_lin_belos.h:130-222
void solver_solve(Ctx *c, bool use_prec, const char *label) {
  Matrix &A = c->A; ISPH_REQUIRE(A.built, "solveProblem: no matrix (setMatrix)");
  ISPH_REQUIRE(c->x_nvec >= 1 && c->b_nvec == c->x_nvec && c->xs.p && c->bs.p, "solveProblem: create the solution and load multivectors first");
  const int n = A.n, ld = c->ld, m = c->sp.num_blocks, g = vgrid(c, n);
  const bool is_cg = c->sp.solver_type == "Block CG";
  ISPH_REQUIRE(is_cg || c->sp.solver_type == "Block GMRES", "Solver Type must be \"Block GMRES\" or \"Block CG\" (Recycling GMRES is not implemented)");
  ISPH_REQUIRE(c->sp.block_size == 1, "Block Size must be 1");
  c->prof_phases = getenv("ISPH_PROFILE") != nullptr;
  std::string tname = std::string("solve") + (label ? label : "");
  c->tic(tname.c_str());
  c->hbuf.ensure(S_TOTAL); c->flag.ensure(16); c->red.ensure((size_t)4 * 592 * 17 + 1024 + (size_t)A.nslices / 8 + 64);
  c->wk.ensure((size_t)ld + c->nall + 3 * (size_t)ld);
  c->h_scal.ensure(16 + c->sp.max_iters + m + 8);
  c->V.ensure((size_t)(is_cg ? 3 : m + 1) * ld);
  c->Z.ensure((size_t)(is_cg ? 1 : (c->sp.flexible ? m : 1)) * ld);
  CUDA_CHECK(cudaMemsetAsync(c->flag.p + 8, 0, 8 * sizeof(int), c->stream));
  // initial solution (setInitialSolution, solver_lin.cpp:141-147): applied here, on the device
  const size_t xl = (size_t)ld * c->x_nvec;
  if (c->init_type == ISPH_INIT_ZERO) CUDA_CHECK(cudaMemsetAsync(c->xs.p, 0, sizeof(double) * xl, c->stream));
  else if (c->init_type == ISPH_INIT_VALUE) { k_fill<<<vgrid(c, (int)xl), VB, 0, c->stream>>>(c->xs.p, c->init_val, (int)xl); ++c->launches; }
  else if (c->init_type == ISPH_INIT_RANDOM) { for (int q = 0; q < c->x_nvec; ++q) { k_random<<<g, VB, 0, c->stream>>>(c->xs.p + (size_t)q * ld, A.external ? nullptr : c->tag.p, n, 11 + q); ++c->launches; } }
  else if (c->x_host) {     // caller's x is the initial guess (Helmholtz: x = v, pair_isph.cpp:932-941)
    for (int q = 0; q < c->x_nvec; ++q) CUDA_CHECK(cudaMemcpyAsync(c->xs.p + (size_t)q * ld, c->x_host + (size_t)q * c->x_lda, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
  }
  if (c->b_host && !c->b_owned) for (int q = 0; q < c->b_nvec; ++q) CUDA_CHECK(cudaMemcpyAsync(c->bs.p + (size_t)q * ld, c->b_host + (size_t)q * c->b_lda, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
  double *S = c->hbuf.p;
  if (c->is_singular) {     // createNullVector (solver_lin.cpp:59-77) ; b -= (b.n) n (solver_lin_belos.h:138-144)
    c->nullvec.ensure(ld);
    k_mask_to_vec<<<g, VB, 0, c->stream>>>(c->have_mask ? c->mask.p : nullptr, c->nullvec.p, n); ++c->launches;
    dot_dev(c, c->nullvec.p, nullptr, n, S + S_TMP);
    const double nrm = sqrt(read_scalar(c, S + S_TMP));
    k_scale_by<<<g, VB, 0, c->stream>>>(c->nullvec.p, 1.0 / nrm, n); ++c->launches;
    for (int q = 0; q < c->b_nvec; ++q) { dot_dev(c, c->bs.p + (size_t)q * ld, c->nullvec.p, n, S + S_TMP);
      k_axpy_dev<<<g, VB, 0, c->stream>>>(c->bs.p + (size_t)q * ld, c->nullvec.p, S + S_TMP, -1.0, n); ++c->launches; }
  }
  if (use_prec) precond_create(c);                               // prec->create(), solver_lin_belos.h:153
  int iters_tot = 0, conv_all = 1; double relres = 0.0;
  for (int q = 0; q < c->x_nvec; ++q) {                          // block size 1: right-hand sides are solved one after another
    int it = 0; double rr = 0.0;
    const int ok = is_cg ? cg_solve(c, use_prec, c->xs.p + (size_t)q * ld, c->bs.p + (size_t)q * ld, &it, &rr)
                         : gmres_solve(c, use_prec, c->xs.p + (size_t)q * ld, c->bs.p + (size_t)q * ld, &it, &rr);
    iters_tot += it; conv_all &= ok; relres = rr > relres ? rr : relres;
  }
  if (use_prec) precond_free(c);                                 // prec->free(), :186-191
  if (!conv_all) {                                               // :197-213 : not an error, report ||b - A x|| / ||b|| (collective: every rank takes part)
    double rn = 0.0, bn = 0.0;
    spmv(c, c->xs.p, c->wk.p, 1, ld, ld);
    { P2PRed pr = halo_p2p_ticket(c); k_residual<<<g, VB, 0, c->stream>>>(c->bs.p, c->wk.p, c->wk.p, n, c->red.p, (unsigned *)c->flag.p + 8, S + S_TMP, pr); ++c->launches;
      if (c->nranks > 1 && pr.nranks <= 1) halo_allreduce(c, S + S_TMP, 1); }
    rn = sqrt(read_scalar(c, S + S_TMP));
    dot_dev(c, c->bs.p, nullptr, n, S + S_TMP); bn = sqrt(read_scalar(c, S + S_TMP));
    if (c->rank == 0) fprintf(stderr, ">> isph_b200::Status - Failed to converge! %s  ||r|| / ||b|| = %6.4e\n", label ? label : " ", bn > 0 ? rn / bn : rn);
  }
  if (c->is_singular) {                                          // x -= (x.n) n, :215-219
    for (int q = 0; q < c->x_nvec; ++q) { dot_dev(c, c->xs.p + (size_t)q * ld, c->nullvec.p, n, S + S_TMP);
      k_axpy_dev<<<g, VB, 0, c->stream>>>(c->xs.p + (size_t)q * ld, c->nullvec.p, S + S_TMP, -1.0, n); ++c->launches; }
  }
  if (c->x_host) for (int q = 0; q < c->x_nvec; ++q)             // x is a View of caller memory (solver_lin.cpp:52-58)
    CUDA_CHECK(cudaMemcpyAsync(c->x_host + (size_t)q * c->x_lda, c->xs.p + (size_t)q * ld, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  c->toc(tname.c_str());
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  if (c->prof_phases && c->rank == 0) {
    fprintf(stderr, "[isph profile] %s: %d iterations\n", tname.c_str(), iters_tot);
    for (auto &kv : c->phase_ev) { size_t u = c->phase_used[kv.first]; double tot = 0.0; for (size_t q = 0; q + 1 < u; q += 2) { float ms = 0.f; cudaEventElapsedTime(&ms, kv.second[q], kv.second[q + 1]); tot += ms; }
      fprintf(stderr, "[isph profile]   %-16s %6zu x  avg %8.2f us  total %8.3f ms\n", kv.first.c_str(), u / 2, u ? 1e3 * tot / (u / 2) : 0.0, tot); c->phase_used[kv.first] = 0; }
  }
  if (c->nranks > 1) ISPH_REQUIRE(!halo_fault(c), "peer exchange timed out: a rank stopped responding");
  c->last_iters = iters_tot; c->last_converged = conv_all; c->last_relres = relres;
  c->init_type = -1;
}

}  // namespace isph
