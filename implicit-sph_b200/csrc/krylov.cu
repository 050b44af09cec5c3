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
#include <cuda.h>                                                // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

namespace isph {

static const int VB = 256;              // threads per block of the vector kernels

// Programmatic dependent launch: the kernel may be scheduled while the previous kernel of the stream is still draining (its CTAs
// become resident as SMs free up and wait at `griddepcontrol.wait`, which every kernel launched this way executes before it reads
// anything the previous kernel wrote).  Hides the launch latency and the ramp of the four kernels of an Arnoldi step behind the
// tails (last-block reduction, peer all-reduce) of their predecessors.  ISPH_NO_PDL=1: plain launches.
template <class... KA, class... A> static void launch_pdl(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A... args) {
  static const bool pdl = getenv("ISPH_NO_PDL") == nullptr;
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg)); cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, static_cast<KA>(args)...));
}
#define PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#define PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
static const double DEP_TOL = 0.70710678118654752440;   // DGKSOrthoManager dep_tol = 1/sqrt(2)

// layout of the small device scalar block `hbuf`
enum { S_H = 0, S_H2 = 64, S_G = 128, S_CS = 192, S_SN = 256, S_Y = 320, S_OLD = 384, S_NEW1 = 385, S_NEW2 = 386, S_PROJ = 387,
       S_INV = 388, S_RES = 389, S_ALPHA = 390, S_BETA = 391, S_RZ = 392, S_PAP = 393, S_TMP = 394, S_NSEC = 395, S_HM = 448 /* H: 64 x 64 */, S_TOTAL = 448 + 64 * 64 };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-reduce NACC per-thread accumulators, store the block's partials, and let the last block to finish add the
// partials of all blocks (fixed assignment of blocks to lanes + fixed shuffle tree => deterministic) into out[0..nacc)
template <int NACC> __device__ bool reduce_finish(double (&acc)[NACC], int nacc, double *partials, unsigned *counter, double *out, const P2PRed &pr) {
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
  return last;                                                   // true in every thread of the one block that holds the final sums
}

// diagonal preconditioners: Ifpack_PointRelaxation (Jacobi) forms (damping * invdiag) * v, Ifpack_Chebyshev of degree 1 forms
// (invdiag * v) / theta — the association is kept so that both stay bit-identical to the oracle
__device__ __forceinline__ double diag_prec(double invd, double v, double s, int post) { return post ? invd * v * s : s * invd * v; }

// ---- DGKS bookkeeping without extra reductions ----------------------------------------------------------------------
// Belos' DGKS manager takes four to five global reductions per Arnoldi step (oldDot, Q^T w, newDot, [second pass], norm).
// Here an Arnoldi step takes TWO: the pass-0 message carries h = V^T y together with y.y and (singular problems) n.y, the
// pass-1 message carries h2 = V^T w1 together with w1.w1.  Everything else follows from Pythagoras, which is exact up to
// rounding because the basis is orthonormal:
//   PoissonProjection tail (solver_lin.h:135-137): w' = y - (n.y) n is never formed on its own; n is treated as one more
//     (unit, orthogonal to V up to rounding) direction of the same classical Gram-Schmidt sweep;
//   oldDot = ||w'||^2 = y.y - (n.y)^2 ;  newDot = ||w1||^2 ~ oldDot - sum h_k^2  — used ONLY for the DGKS decision
//     newDot < dep_tol * oldDot (a threshold test; when cancellation is severe the estimate is tiny or negative and the
//     second pass is taken, as it should be);
//   the norm that enters the Hessenberg matrix: second pass taken -> ||w2||^2 = w1.w1 - sum h2_k^2 (h2 is O(eps) relative
//     to w1: no cancellation); not taken -> oldDot - sum h_k^2 with newDot >= 0.707 oldDot (relative error <= 2 eps).
// Message layout: S_H[0..nv) = h, S_H[nv] = y.y, S_H[nv+1] = n.y ; S_H2[0..nv) = h2, S_H2[nv] = w1.w1.
struct DgksNorm { bool second; double hn2; };
__device__ __forceinline__ DgksNorm dgks_norm(const double *S, int nv, bool singular) {      // serial, fixed order: identical wherever it is evaluated
  double old = S[S_H + nv]; if (singular) { const double p = S[S_H + nv + 1]; old -= p * p; }
  double s = 0.0; for (int k = 0; k < nv; ++k) { const double h = S[S_H + k]; s += h * h; }
  const double new1 = old - s;
  DgksNorm d; d.second = new1 < DEP_TOL * old; d.hn2 = new1;
  if (d.second) { double s2 = 0.0; for (int k = 0; k < nv; ++k) { const double h = S[S_H2 + k]; s2 += h * h; } d.hn2 = S[S_H2 + nv] - s2; }
  if (!(d.hn2 > 0.0)) d.hn2 = 0.0;
  return d;
}
__device__ __forceinline__ bool dgks_second(const double *S, int nv, bool singular) {
  double old = S[S_H + nv]; if (singular) { const double p = S[S_H + nv + 1]; old -= p * p; }
  double s = 0.0; for (int k = 0; k < nv; ++k) { const double h = S[S_H + k]; s += h * h; }
  return old - s < DEP_TOL * old;
}

// out[0] = sum a_i b_i (b == nullptr: a_i a_i)
__global__ void __launch_bounds__(VB) k_dot(const double *a, const double *b, int n, double *partials, unsigned *counter, double *out, P2PRed pr) {
  double acc[1] = {0.0};
  for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) acc[0] += a[i] * (b ? b[i] : a[i]);
  reduce_finish<1>(acc, 1, partials, counter, out, pr);
}

// Classical Gram-Schmidt coefficients.  grid = (row chunks, vector groups): block (bx, g) owns a CONTIGUOUS chunk of
// rows and the G basis vectors [gG, gG+G): every thread streams one w value pair and G basis value pairs per step
// (128-bit loads), keeps G+2 accumulators in registers, and a block touches only G+2 pages — the first version, where
// every thread walked all <= 51 vectors, ran at ~1.4 TB/s (profiles/r01_launches_c2_first.txt).
//   pass 0: h[k] = V_k . y (k < nv), h[nv] = y.y, h[nv+1] = n.y (group 0; n = null vector of a singular problem)
//   pass 1: skipped unless the DGKS test asks for a second pass; h2[k] = V_k . w1, h2[nv] = w1.w1
template <int G> __global__ void __launch_bounds__(VB, G == 16 ? 2 : (G == 8 ? 3 : 5))
k_multidot(const double *__restrict__ V, int ld, int nv, const double *__restrict__ w, const double *__restrict__ nvec, int n,
           double *S, int pass, int rev, double *partials, unsigned *counters, P2PRed pr) {
  __shared__ bool go;
  PDL_WAIT(); PDL_TRIGGER();                                 // launched dependent on the SpMV (pass 0); the sweep that follows may start its prologue
  if (pass == 1) { if (threadIdx.x == 0) go = dgks_second(S, nv, nvec != nullptr); __syncthreads(); if (!go) return; }
  const int g = blockIdx.y, k0 = g * G, cnt = min(G, nv - k0);
  const bool with_n = (pass == 0 && nvec != nullptr && g == 0);
  int chunk = (n + gridDim.x - 1) / gridDim.x; chunk = (chunk + 1) & ~1;
  const int r0 = (rev ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x) * chunk, r1 = min(n, r0 + chunk);
  const double *Vg = V + (size_t)k0 * ld;
  double acc[G + 2];
#pragma unroll
  for (int k = 0; k < G + 2; ++k) acc[k] = 0.0;
  // rev: the chunks and the rows inside a chunk are walked backwards (the sweeps of an Arnoldi step alternate direction so that
  // each starts on what the previous one left in L2, see k_update_dot_tma)
  const int i_lo = r0 + 2 * (int)threadIdx.x, nit = i_lo < r1 ? (r1 - i_lo + 2 * VB - 1) / (2 * VB) : 0;
  for (int t = 0; t < nit; ++t) {
    const int i = i_lo + (rev ? nit - 1 - t : t) * 2 * VB;
    double2 wi, nn = make_double2(0.0, 0.0);
    if (i + 1 < r1) { wi = *reinterpret_cast<const double2 *>(w + i); if (with_n) nn = *reinterpret_cast<const double2 *>(nvec + i); }
    else { wi.x = w[i]; wi.y = 0.0; if (with_n) nn.x = nvec[i]; }
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
    if (g == 0) { acc[G] += wi.x * wi.x; acc[G] += wi.y * wi.y; acc[G + 1] += nn.x * wi.x; acc[G + 1] += nn.y * wi.y; }
  }
  __shared__ double sm[G + 2][VB / 32];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < G + 2; ++k) { const double v = warp_sum(acc[k]); if (lane == 0) sm[k][warp] = v; }
  __syncthreads();
  double *mine = partials + ((size_t)g * gridDim.x + blockIdx.x) * (G + 2);
  if (threadIdx.x < G + 2) { double s = 0.0; for (int q = 0; q < VB / 32; ++q) s += sm[threadIdx.x][q]; mine[threadIdx.x] = s; }
  __threadfence(); __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(counters + g, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last) {
    __threadfence();
    const double *grp = partials + (size_t)g * gridDim.x * (G + 2);
    for (int k = warp; k < G + 2; k += VB / 32) {
      if (!(k < cnt || (k == G && g == 0) || (k == G + 1 && with_n))) continue;
      double s = 0.0;
      for (unsigned b = lane; b < gridDim.x; b += 32) s += __ldcg(grp + (size_t)b * (G + 2) + k);
      s = warp_sum(s);
      if (lane == 0) S[(pass == 0 ? S_H : S_H2) + (k < G ? k0 + k : nv + (k - G))] = s;
    }
    if (threadIdx.x == 0) counters[g] = 0u;
    if (pr.nranks > 1) {                                   // the last group to finish exchanges the whole message with the peers
      __shared__ bool all_done;
      __threadfence(); __syncthreads();
      if (threadIdx.x == 0) { all_done = (atomicAdd(counters + 6, 1u) == gridDim.y - 1); if (all_done) counters[6] = 0u; }   // flag word 15
      __syncthreads();
      if (all_done) { __threadfence(); p2p_allreduce_block(pr, S + (pass == 0 ? S_H : S_H2), pass == 0 ? nv + (nvec ? 2 : 1) : nv + 1); }
    }
  }
}

// first Gram-Schmidt update: w1 = y - (n.y) n - sum_k h_k V_k.  One contiguous row chunk per block, 128-bit accesses; no
// reduction (||w1||^2 travels with the pass-1 message, see above).
__global__ void __launch_bounds__(VB)
k_cgs_update(const double *__restrict__ V, int ld, int nv, double *__restrict__ w, const double *__restrict__ nvec, int n, const double *S, int rev) {
  __shared__ double sh[64];
  PDL_WAIT(); PDL_TRIGGER();
  if (threadIdx.x < nv) sh[threadIdx.x] = S[S_H + threadIdx.x];
  __syncthreads();
  const double proj = nvec ? S[S_H + nv + 1] : 0.0;
  int chunk = (n + gridDim.x - 1) / gridDim.x; chunk = (chunk + 1) & ~1;
  const int r0 = (rev ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x) * chunk, r1 = min(n, r0 + chunk);
  const int i_lo = r0 + 2 * (int)threadIdx.x, nit = i_lo < r1 ? (r1 - i_lo + 2 * VB - 1) / (2 * VB) : 0;
  for (int t = 0; t < nit; ++t) {
    const int i = i_lo + (rev ? nit - 1 - t : t) * 2 * VB;
    if (i + 1 < r1) {
      double2 wi = *reinterpret_cast<const double2 *>(w + i);
      if (nvec) { const double2 nn = *reinterpret_cast<const double2 *>(nvec + i); wi.x -= proj * nn.x; wi.y -= proj * nn.y; }
#pragma unroll 8
      for (int k = 0; k < nv; ++k) { const double2 v = *reinterpret_cast<const double2 *>(V + (size_t)k * ld + i); wi.x -= sh[k] * v.x; wi.y -= sh[k] * v.y; }
      *reinterpret_cast<double2 *>(w + i) = wi;
    } else {
      double wi = w[i]; if (nvec) wi -= proj * nvec[i];
      for (int k = 0; k < nv; ++k) wi -= sh[k] * V[(size_t)k * ld + i];
      w[i] = wi;
    }
  }
}

// First Gram-Schmidt update AND second-pass coefficients in ONE sweep over the basis (the two kernels above read V twice):
//   w1 = y - (n.y) n - sum_k h_k V_k ,  h2[k] = V_k . w1 ,  h2[nv] = w1.w1
// The block walks its row chunk in tiles of UT rows.  Warp q owns the basis vectors k = q, q+8, ... (<= KPW of them): it
// loads their tile values ONCE into registers, contributes its partial sum of h_k V_k to shared memory, the first UT threads
// finish w1 for the tile (written back, kept in shared memory), and every warp then forms its vectors' dot products with w1
// from the SAME registers.  Traffic 8 n (nv + 3) instead of 8 n (2 nv + 4).  The dots are only formed when the DGKS test
// (known from the pass-0 message) asks for a second pass.  Used for nv > 8; shorter bases keep the two streaming kernels.
static const int KPW = 7;                                       // 8 warps x 7 vectors >= 51 basis vectors
template <int HV> __global__ void __launch_bounds__(VB, HV == 1 ? 3 : 2)     // HV = 64-row halves per tile
k_update_dot(const double *__restrict__ V, int ld, int nv, double *__restrict__ w, const double *__restrict__ nvec, int n, int rev,
             double *S, double *partials, unsigned *counter, P2PRed pr) {
  constexpr int UT = 64 * HV;
  __shared__ __align__(16) double s_part[VB / 32][UT];
  __shared__ __align__(16) double s_w1[UT];
  __shared__ double s_h[64], s_red[UT / 32];
  __shared__ bool s_go, s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  PDL_WAIT(); PDL_TRIGGER();
  if (tid == 0) s_go = dgks_second(S, nv, nvec != nullptr);
  if (tid < nv) s_h[tid] = S[S_H + tid];
  __syncthreads();
  const bool go = s_go;
  const double proj = nvec ? S[S_H + nv + 1] : 0.0;
  int chunk = (n + gridDim.x - 1) / gridDim.x; chunk = (chunk + UT - 1) / UT * UT;
  const int r0 = min(n, (rev ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x) * chunk), r1 = min(n, r0 + chunk), ntl = (r1 - r0 + UT - 1) / UT;
  double acc[KPW], nrm = 0.0;
#pragma unroll
  for (int q = 0; q < KPW; ++q) acc[q] = 0.0;
  for (int tl = 0; tl < ntl; ++tl) {
    const int t0 = r0 + (rev ? ntl - 1 - tl : tl) * UT;
    // the tile's w (and n) values are requested together with the basis values: one memory latency per tile, not two
    double wv = 0.0, nn = 0.0;
    if (tid < UT && t0 + tid < r1) { wv = w[t0 + tid]; if (nvec) nn = nvec[t0 + tid]; }
    double2 v[KPW][HV], p[HV];
#pragma unroll
    for (int hf = 0; hf < HV; ++hf) p[hf] = make_double2(0.0, 0.0);
#pragma unroll
    for (int q = 0; q < KPW; ++q) {
      const int k = warp + (VB / 32) * q;
#pragma unroll
      for (int hf = 0; hf < HV; ++hf) {
        const int i = t0 + 64 * hf + 2 * lane;                   // this lane's rows: two in each half of the tile
        v[q][hf] = make_double2(0.0, 0.0);
        if (k < nv) {
          const double *vk = V + (size_t)k * ld + i;
          if (i + 1 < r1) v[q][hf] = *reinterpret_cast<const double2 *>(vk); else if (i < r1) v[q][hf].x = vk[0];
        }
      }
    }
#pragma unroll
    for (int q = 0; q < KPW; ++q) {
      const int k = warp + (VB / 32) * q;
      if (k < nv) { const double h = s_h[k];
#pragma unroll
        for (int hf = 0; hf < HV; ++hf) { p[hf].x += h * v[q][hf].x; p[hf].y += h * v[q][hf].y; } }
    }
#pragma unroll
    for (int hf = 0; hf < HV; ++hf) *reinterpret_cast<double2 *>(&s_part[warp][64 * hf + 2 * lane]) = p[hf];
    __syncthreads();
    if (tid < UT) {
      double w1 = 0.0;
      if (t0 + tid < r1) {
        w1 = wv - proj * nn;
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < VB / 32; ++q) s += s_part[q][tid];
        w1 -= s; w[t0 + tid] = w1; nrm += w1 * w1;
      }
      s_w1[tid] = w1;
    }
    __syncthreads();
    if (go) {
#pragma unroll
      for (int hf = 0; hf < HV; ++hf) {
        const double2 ww = *reinterpret_cast<const double2 *>(&s_w1[64 * hf + 2 * lane]);
#pragma unroll
        for (int q = 0; q < KPW; ++q) { acc[q] += v[q][hf].x * ww.x; acc[q] += v[q][hf].y * ww.y; }
      }
    }
  }
  if (!go) return;                                               // block-uniform: the DGKS decision is the same everywhere (and on every rank)
  double *mine = partials + (size_t)blockIdx.x * 64;
#pragma unroll
  for (int q = 0; q < KPW; ++q) { const double s = warp_sum(acc[q]); const int k = warp + (VB / 32) * q; if (lane == 0 && k < nv) mine[k] = s; }
  if (tid < UT) { const double s = warp_sum(nrm); if (lane == 0) s_red[warp] = s; }
  __syncthreads();
  if (tid == 0) { double t = 0.0; for (int q = 0; q < UT / 32; ++q) t += s_red[q]; mine[63] = t; }
  __threadfence(); __syncthreads();
  if (tid == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence();
    for (int k = warp; k < 64; k += VB / 32) {
      if (!(k < nv || k == 63)) continue;
      double s = 0.0;
      for (unsigned b = lane; b < gridDim.x; b += 32) s += __ldcg(partials + (size_t)b * 64 + k);
      s = warp_sum(s);
      if (lane == 0) S[S_H2 + (k == 63 ? nv : k)] = s;
    }
    if (tid == 0) *counter = 0u;
    if (pr.nranks > 1) { __threadfence(); p2p_allreduce_block(pr, S + S_H2, nv + 1); }
  }
}

// ---- the same fused sweep as a TMA pipeline (default for nv > 8) -----------------------------------------------------------
// w1 = y - (n.y) n - sum_k h_k V_k ,  h2[k] = V_k . w1 ,  h2[nv] = w1.w1        (one read of the basis, as above)
// k_update_dot stages every tile through registers and meets two block barriers per tile with nothing in flight behind them
// (ncu: 24 % warp occupancy, barrier + long-scoreboard stalls: profiles/r01_prof_update_dot_c2_ncu.txt).  Here the Krylov basis is
// described to the TMA unit as a 2-D tensor [basis vector][row] (cuTensorMapEncodeTiled, one map per basis size) and ONE
// cp.async.bulk.tensor instruction brings a tile — TT rows of the nv basis vectors AND of y, which is row nv of the same array —
// into shared memory, completion counted by an mbarrier per stage, NS stages deep.  Warp 0 is the producer (waits for a free stage,
// arms the barrier, issues the copy); warps 1-4 are the consumers: thread-per-row update out of shared memory, a named barrier
// among the 128 consumer threads, warp-per-vector dot products out of the SAME tile, then one arrive per warp frees the stage.
// While a tile is being worked on, the copies of the next NS-1 tiles are in flight, so the barriers no longer drain the memory
// pipeline.  (A first version issued nv+2 one-dimensional cp.async.bulk copies per tile from a compute thread: the ~50 serial
// UBLKCP issues per tile made it 1.8x SLOWER than the register-tile kernel, gpurun_out/r2_ud_*_c2.err.)
// Sweep direction: `rev` walks the row chunks and the tiles inside a chunk backwards.  The orthogonalisation kernels of a step
// alternate direction, so each starts on the rows the previous one touched last — the part of the basis still in the 126 MB L2.
static const int TT = 128;                                       // rows per tile = consumer threads per CTA
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_tile_2d(void *dst, const CUtensorMap *tm, int c0, int c1, unsigned long long *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__global__ void __launch_bounds__(TT + 32)
k_update_dot_tma(const __grid_constant__ CUtensorMap tm, int nv, double *__restrict__ w, const double *__restrict__ nvec, int n, int ns, int rev,
                 double *S, double *partials, unsigned *counter, P2PRed pr) {
  extern __shared__ __align__(128) unsigned char dsm[];
  __shared__ __align__(8) unsigned long long full[4], empty[4];
  __shared__ __align__(16) double s_w1[2][TT];
  __shared__ double s_h[64], s_red[TT / 32];
  __shared__ bool s_go, s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t stage_doubles = (size_t)(nv + 1) * TT + (nvec ? TT : 0);           // [nv basis segments][y][n]
  double *stage0 = reinterpret_cast<double *>(dsm);
  int chunk = (n + gridDim.x - 1) / gridDim.x; chunk = (chunk + TT - 1) / TT * TT;
  const int cb = rev ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
  const int r0 = min(n, cb * chunk), r1 = min(n, r0 + chunk), ntiles = (r1 - r0 + TT - 1) / TT;
  if (tid == 0) { for (int q = 0; q < ns; ++q) { mbar_init(&full[q], 1); mbar_init(&empty[q], TT / 32); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  PDL_WAIT(); PDL_TRIGGER();                                     // everything below reads what the pass-0 kernel produced (coefficients, y)
  if (tid == 0) s_go = dgks_second(S, nv, nvec != nullptr);
  if (tid < nv) s_h[tid] = S[S_H + tid];
  __syncthreads();
  const bool go = s_go;
  constexpr int KW = 13;                                         // 4 consumer warps x 13 vectors >= 51 basis vectors
  double acc[KW], nrm = 0.0;
#pragma unroll
  for (int q = 0; q < KW; ++q) acc[q] = 0.0;
  if (warp == 0) {                                               // ---- producer
    if (lane == 0) {
      const unsigned bytes = (unsigned)(((size_t)(nv + 1) * TT + (nvec ? TT : 0)) * sizeof(double));
      for (int it = 0; it < ntiles; ++it) {
        const int q = it % ns, tile = rev ? ntiles - 1 - it : it, t0 = r0 + tile * TT;
        if (it >= ns) mbar_wait(&empty[q], (unsigned)(((it / ns) - 1) & 1));
        double *dst = stage0 + (size_t)q * stage_doubles;
        mbar_expect_tx(&full[q], bytes);
        tma_tile_2d(dst, &tm, t0, 0, &full[q]);
        if (nvec) bulk_g2s(dst + (size_t)(nv + 1) * TT, nvec + t0, TT * sizeof(double), &full[q]);     // the null vector is allocated with ld >= t0 + TT entries
      }
    }
  } else {                                                       // ---- consumers: threads 32..159
    const int ct = tid - 32, cw = warp - 1;
    const double proj = nvec ? S[S_H + nv + 1] : 0.0;
    for (int it = 0; it < ntiles; ++it) {
      const int q = it % ns, tile = rev ? ntiles - 1 - it : it, t0 = r0 + tile * TT, rows = min(TT, r1 - t0);
      const double *sv = stage0 + (size_t)q * stage_doubles;
      mbar_wait(&full[q], (unsigned)((it / ns) & 1));
      double w1 = 0.0;
      if (ct < rows) {                                           // thread-per-row update out of shared memory (conflict-free: consecutive rows)
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0; int k = 0;
        for (; k + 4 <= nv; k += 4) { s0 += s_h[k] * sv[(size_t)k * TT + ct]; s1 += s_h[k + 1] * sv[(size_t)(k + 1) * TT + ct];
                                      s2 += s_h[k + 2] * sv[(size_t)(k + 2) * TT + ct]; s3 += s_h[k + 3] * sv[(size_t)(k + 3) * TT + ct]; }
        for (; k < nv; ++k) s0 += s_h[k] * sv[(size_t)k * TT + ct];
        w1 = sv[(size_t)nv * TT + ct]; if (nvec) w1 -= proj * sv[(size_t)(nv + 1) * TT + ct];
        w1 -= ((s0 + s1) + (s2 + s3)); w[t0 + ct] = w1; nrm += w1 * w1;
      }
      if (go) {
        double *sw = s_w1[it & 1];
        sw[ct] = w1;
        asm volatile("bar.sync 1, %0;" ::"n"(TT) : "memory");    // the 128 consumer threads: w1 of the tile is complete
        double ww[TT / 32];
#pragma unroll
        for (int i = 0; i < TT / 32; ++i) ww[i] = sw[lane + 32 * i];       // rows past the tile's end carry w1 = 0 ...
#pragma unroll
        for (int j = 0; j < KW; ++j) {
          const int k = cw + (TT / 32) * j;
          if (k < nv) {
            const double *vk = sv + (size_t)k * TT;
#pragma unroll
            for (int i = 0; i < TT / 32; ++i) { const int r = lane + 32 * i; if (r < rows) acc[j] += vk[r] * ww[i]; }   // ... and their basis cells are never read
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[q]);                     // this warp is done with stage q
    }
  }
  if (!go) return;                                               // block-uniform: the DGKS decision is the same everywhere (and on every rank)
  double *mine = partials + (size_t)blockIdx.x * 64;
  if (warp > 0) {
    const int cw = warp - 1;
#pragma unroll
    for (int j = 0; j < KW; ++j) { const double s = warp_sum(acc[j]); const int k = cw + (TT / 32) * j; if (lane == 0 && k < nv) mine[k] = s; }
    const double s = warp_sum(nrm); if (lane == 0) s_red[cw] = s;
  }
  __syncthreads();
  if (tid == 0) { double t = 0.0; for (int q = 0; q < TT / 32; ++q) t += s_red[q]; mine[63] = t; }
  __threadfence(); __syncthreads();
  if (tid == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence();
    for (int k = warp; k < 64; k += (TT + 32) / 32) {
      if (!(k < nv || k == 63)) continue;
      double s = 0.0;
      for (unsigned b = lane; b < gridDim.x; b += 32) s += __ldcg(partials + (size_t)b * 64 + k);
      s = warp_sum(s);
      if (lane == 0) S[S_H2 + (k == 63 ? nv : k)] = s;
    }
    if (tid == 0) *counter = 0u;
    if (pr.nranks > 1) { __threadfence(); p2p_allreduce_block(pr, S + S_H2, nv + 1); }
  }
}

// Hessenberg column j: DGKS bookkeeping, Givens rotations, implicit residual (BlockGmresIter::updateLSQR); one warp:
// lanes stage the column and the rotations in shared memory, lane 0 runs the (inherently sequential) recurrence there
__device__ void givens_step(double *S, int j, bool singular, double *host_res, int slot) {
  __shared__ double h[64], cs[64], sn[64];
  const int lane = threadIdx.x & 31;
  const DgksNorm d = dgks_norm(S, j + 1, singular);
  for (int k = lane; k <= j; k += 32) { h[k] = S[S_H + k] + (d.second ? S[S_H2 + k] : 0.0); cs[k] = S[S_CS + k]; sn[k] = S[S_SN + k]; }
  __syncwarp();
  if (lane == 0) {
    const double hn = sqrt(d.hn2);
    h[j + 1] = hn;
    if (d.second) S[S_NSEC] += 1.0;                     // statistics: Arnoldi steps that took the DGKS second pass
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
__global__ void k_givens(double *S, int j, int singular, double *host_res, int slot) { givens_step(S, j, singular != 0, host_res, slot); }

// End of an Arnoldi step in ONE sweep: second Gram-Schmidt update (when the DGKS test asked for it), normalisation
// v_{j+1} = w2 / ||w2|| (norm from dgks_norm, no reduction) and — Jacobi — the next preconditioned vector
// z_{j+1} = damping * D^-1 v_{j+1}.  Warp 0 of block 0 also does the Hessenberg/Givens step of this column.
// store z_i into the staging buffers of the peers that have row i on their halo list
__device__ __forceinline__ void prepush_row(const PrePush &pp, int hslot, int b, int e, double v) {
  v = p2p_payload(v);
  for (; b < e; ++b) { const int code = __ldg(pp.sd + b); *reinterpret_cast<volatile double *>(pp.plan->peer[code >> 28] + (size_t)hslot * 3 * pp.plan->cap + (code & 0x0fffffff)) = v; }
}
__global__ void __launch_bounds__(VB)
k_finish(const double *__restrict__ V, int ld, int nv, double *__restrict__ w, int n, double *S, int singular,
         const double *__restrict__ invdiag, double damping, int post, double *__restrict__ z, double *host_res, int slot, int rev, PrePush pp) {
  PDL_WAIT(); PDL_TRIGGER();
  if (blockIdx.x == 0) {      // block 0 is dedicated to the (sequential, ~10 us) Hessenberg/Givens step: hidden behind the sweep
    if (threadIdx.x < 32) givens_step(S, nv - 1, singular != 0, host_res, slot);
    return;
  }
  __shared__ double sh[64]; __shared__ double s_inv; __shared__ int s_second;
  if (threadIdx.x == 0) { const DgksNorm d = dgks_norm(S, nv, singular != 0); const double hn = sqrt(d.hn2); s_inv = hn > 0.0 ? 1.0 / hn : 0.0; s_second = d.second ? 1 : 0; }
  if (threadIdx.x < nv) sh[threadIdx.x] = S[S_H2 + threadIdx.x];
  __syncthreads();
  const double inv = s_inv; const int nk = s_second ? nv : 0;
  const int nb = gridDim.x - 1, bx = rev ? nb - (int)blockIdx.x : (int)blockIdx.x - 1, hslot = (int)(pp.seq % MB_SLOTS);
  int chunk = (n + nb - 1) / nb; chunk = (chunk + 1) & ~1;
  const int r0 = bx * chunk, r1 = min(n, r0 + chunk);
  const int i_lo = r0 + 2 * (int)threadIdx.x, nit = i_lo < r1 ? (r1 - i_lo + 2 * VB - 1) / (2 * VB) : 0;
  for (int t = 0; t < nit; ++t) {
    const int i = i_lo + (rev ? nit - 1 - t : t) * 2 * VB;
    if (i + 1 < r1) {
      double2 wi = *reinterpret_cast<const double2 *>(w + i);
#pragma unroll 8
      for (int k = 0; k < nk; ++k) { const double2 v = *reinterpret_cast<const double2 *>(V + (size_t)k * ld + i); wi.x -= sh[k] * v.x; wi.y -= sh[k] * v.y; }
      wi.x *= inv; wi.y *= inv;
      *reinterpret_cast<double2 *>(w + i) = wi;
      if (z) { double2 zi = wi; if (invdiag) { const double2 dd = *reinterpret_cast<const double2 *>(invdiag + i); zi.x = diag_prec(dd.x, wi.x, damping, post); zi.y = diag_prec(dd.y, wi.y, damping, post); } *reinterpret_cast<double2 *>(z + i) = zi;
        if (pp.plan) { const int s0 = __ldg(pp.sp + i), s1 = __ldg(pp.sp + i + 1), s2 = __ldg(pp.sp + i + 2); if (s2 > s0) { prepush_row(pp, hslot, s0, s1, zi.x); prepush_row(pp, hslot, s1, s2, zi.y); } } }
    } else {
      double wi = w[i];
      for (int k = 0; k < nk; ++k) wi -= sh[k] * V[(size_t)k * ld + i];
      wi *= inv; w[i] = wi;
      if (z) { const double zi = invdiag ? diag_prec(invdiag[i], wi, damping, post) : wi; z[i] = zi; if (pp.plan) prepush_row(pp, hslot, __ldg(pp.sp + i), __ldg(pp.sp + i + 1), zi); }
    }
  }
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
// One PCG iteration = 4 launches, 3 reductions: [SpMV with p.Ap in its epilogue] -> k_cg_update (x, r, ||r||^2, publishes the
// residual) -> k_cg_precdot (diagonal preconditioner fused, r.z) -> k_cg_direction (p = z + beta p; pushes the halo rows of p).
// r.z lives in two slots used alternately (rz_cur / rz_new), so no kernel has to shift it.
__global__ void __launch_bounds__(VB) k_cg_update(double *x, double *r, const double *p, const double *Ap, double *S, int rz_cur, int n, double *partials, unsigned *counter, P2PRed pr,
                                                  double *host_res, int slot) {
  const double alpha = S[rz_cur] / S[S_PAP];
  double acc[1] = {0.0};
  for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) { x[i] += alpha * p[i]; const double ri = r[i] - alpha * Ap[i]; r[i] = ri; acc[0] += ri * ri; }
  const bool last = reduce_finish<1>(acc, 1, partials, counter, S + S_TMP, pr);
  if (last && host_res && threadIdx.x == 0) { const double res = sqrt(S[S_TMP]); S[S_RES] = res; host_res[slot] = res; __threadfence_system(); }
}
__global__ void k_cg_publish(double *S, double *host_res, int slot) { if (threadIdx.x == 0) { const double res = sqrt(S[S_TMP]); S[S_RES] = res; host_res[slot] = res; __threadfence_system(); } }
// z = diagonal preconditioner applied to r, fused with out = r.z ; for other preconditioners z is given and only the dot is taken
__global__ void __launch_bounds__(VB) k_cg_precdot(const double *r, double *z, const double *invdiag, double damping, int post, int n, double *partials, unsigned *counter, double *out, P2PRed pr) {
  double acc[1] = {0.0};
  for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) { double zi; if (invdiag) { zi = diag_prec(invdiag[i], r[i], damping, post); z[i] = zi; } else zi = z[i]; acc[0] += r[i] * zi; }
  reduce_finish<1>(acc, 1, partials, counter, out, pr);
}
// p = z + beta p with beta = rz_new / rz_cur ; p is the next SpMV input: its halo rows leave from here
__global__ void __launch_bounds__(VB) k_cg_direction(double *p, const double *z, const double *S, int rz_cur, int rz_new, int n, int first, PrePush pp) {
  const double beta = first ? 0.0 : S[rz_new] / S[rz_cur];
  const int hslot = (int)(pp.seq % MB_SLOTS);
  for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) {
    const double pi = first ? z[i] : z[i] + beta * p[i]; p[i] = pi;
    if (pp.plan) { const int s0 = __ldg(pp.sp + i), s1 = __ldg(pp.sp + i + 1); if (s1 > s0) prepush_row(pp, hslot, s0, s1, pi); }
  }
}

// ---------------------------------------------------------------------------------------------------------------
static int vgrid(Ctx *c, int n) {      // default 148 SMs x 16 chunks (measured best for the streaming updates on 1M rows: 6.2 TB/s vs 4.9 at 592); ISPH_VGRID overrides
  (void)c; static const int cap = getenv("ISPH_VGRID") ? atoi(getenv("ISPH_VGRID")) : 2368;
  int g = ceil_div(n, 2 * VB); return g < cap ? (g < 1 ? 1 : g) : cap;
}

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

// preconditioners that are one diagonal scaling and can ride on another sweep: Jacobi with one sweep, Chebyshev of degree 1
struct DiagPrec { bool on; const double *invdiag; double scale; int post; };
static DiagPrec diag_prec_of(Ctx *c, bool use_prec) {
  DiagPrec d{false, nullptr, 1.0, 0};
  if (!use_prec) return d;
  if (c->prec_kind == 1 && c->pp.sweeps == 1) { d.on = true; d.invdiag = c->invdiag.p; d.scale = c->pp.damping; d.post = 0; }
  else if (c->prec_kind == 2 && c->pp.cheb_degree == 1) {        // Ifpack_Chebyshev, degree 1, zero start: z = invDiag * r / theta
    const double lmax = c->last_lmax, alpha = lmax / c->pp.cheb_ratio, beta = 1.1 * lmax, theta = 0.5 * (beta + alpha);
    d.on = true; d.invdiag = c->invdiag.p; d.scale = 1.0 / theta; d.post = 1;
  }
  return d;
}

static void launch_multidot(Ctx *c, const double *V, int nv, const double *w, int pass, int rev) {
  const int n = c->A.n; double *S = c->hbuf.p; const double *nv_ = c->is_singular ? c->nullvec.p : nullptr; unsigned *cnt = (unsigned *)c->flag.p + 9;
  const int G = nv <= 4 ? 4 : (nv <= 8 ? 8 : 16);            // short bases: do not pay for 16 (aliased) loads per thread
  const int groups = (nv + G - 1) / G;
  // one wave of resident CTAs: 2 / 3 / 5 CTAs per SM for the 128 / 72 / 48-register variants (a second, partial wave cost 5-10 %)
  static const int mdenv = getenv("ISPH_MDGRID") ? atoi(getenv("ISPH_MDGRID")) : 0;
  const int mdcap = mdenv ? mdenv : (G == 16 ? 296 : (G == 8 ? 444 : 740));
  int gx = mdcap / groups; if (gx < 74) gx = 74; { const int mx = ceil_div(n, 2 * VB); if (gx > mx) gx = mx < 1 ? 1 : mx; }
  P2PRed pr = halo_p2p_ticket(c);
  if (G == 4) launch_pdl(k_multidot<4>, dim3(gx, groups), dim3(VB), 0, c->stream, V, c->ld, nv, w, nv_, n, S, pass, rev, c->red.p, cnt, pr);
  else if (G == 8) launch_pdl(k_multidot<8>, dim3(gx, groups), dim3(VB), 0, c->stream, V, c->ld, nv, w, nv_, n, S, pass, rev, c->red.p, cnt, pr);
  else launch_pdl(k_multidot<16>, dim3(gx, groups), dim3(VB), 0, c->stream, V, c->ld, nv, w, nv_, n, S, pass, rev, c->red.p, cnt, pr);
  ++c->launches;
  // NCCL fallback (no peer access).  Pass 1 is conditional on the device: a rank-independent decision (it is taken from the
  // already reduced pass-0 message), so every rank either contributes fresh sums or the same stale, unused ones.
  if (c->nranks > 1 && pr.nranks <= 1) halo_allreduce(c, S + (pass == 0 ? S_H : S_H2), pass == 0 ? nv + (nv_ ? 2 : 1) : nv + 1);
}

// fused update + second-pass sweep, TMA pipeline; launched with programmatic stream serialization so that its launch and prologue
// overlap the tail (last-block reduction, peer all-reduce) of the pass-0 kernel before it.  Tensor maps: the basis array
// V[(m+1)][ld] as a rank-2 tensor {ld (contiguous), m + 1}, box {TT rows, nv + 1 vectors}; one map per basis size, re-encoded
// only when the basis buffer or its leading dimension changes.
struct TmaMaps { const double *V = nullptr; int ld = 0, rows = 0; CUtensorMap map[64]; bool ok[64] = {}; };
static bool tma_map_for(Ctx *c, TmaMaps &T, const double *V, int ld, int rows_total, int nv, CUtensorMap *out) {
  typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                               CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn enc = nullptr; static bool tried = false;
  if (!tried) { tried = true; void *fn = nullptr; cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess) enc = (EncodeFn)fn; else cudaGetLastError(); }
  if (!enc) return false;
  if (T.V != V || T.ld != ld || T.rows != rows_total) { T.V = V; T.ld = ld; T.rows = rows_total; memset(T.ok, 0, sizeof(T.ok)); }
  if (!T.ok[nv]) {
    const cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)rows_total}, gstr[1] = {(cuuint64_t)ld * sizeof(double)};
    const cuuint32_t box[2] = {(cuuint32_t)TT, (cuuint32_t)(nv + 1)}, estr[2] = {1, 1};
    if (enc(&T.map[nv], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(V), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
    T.ok[nv] = true;
  }
  (void)c; *out = T.map[nv]; return true;
}
static bool launch_update_dot_tma(Ctx *c, const double *V, int ld, int rows_total, int nv, double *w, const double *nvp, int n, int rev, double *S, unsigned *counter, P2PRed pr) {
  static const int grid_env = getenv("ISPH_UDGRID") ? atoi(getenv("ISPH_UDGRID")) : 0, ns_env = getenv("ISPH_UD_STAGES") ? atoi(getenv("ISPH_UD_STAGES")) : 0;
  static const bool pdl = getenv("ISPH_NO_PDL") == nullptr;
  static int sms = 0; static bool attr_set = false; static TmaMaps maps;
  if (w != V + (size_t)nv * ld) return false;                   // y must be row nv of the basis array (it is: v_{j+1} is formed in place)
  CUtensorMap tm; if (!tma_map_for(c, maps, V, ld, rows_total, nv, &tm)) return false;
  if (!attr_set) { CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    CUDA_CHECK(cudaFuncSetAttribute(k_update_dot_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024)); attr_set = true; }
  const size_t stage = ((size_t)(nv + 1) * TT + (nvp ? TT : 0)) * sizeof(double);
  int ns = ns_env ? ns_env : (int)((size_t)(106 * 1024) / stage); ns = std::max(2, std::min(4, ns));      // two CTAs of <= 106 KB per SM
  const int grid = std::max(1, std::min(grid_env ? grid_env : 2 * sms, ceil_div(n, TT)));
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(TT + 32); cfg.dynamicSmemBytes = stage * ns; cfg.stream = c->stream;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_update_dot_tma, tm, nv, w, nvp, n, ns, rev, S, c->red.p, counter, pr));
  return true;
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
  const DiagPrec dp = diag_prec_of(c, use_prec); const bool jacobi_fused = dp.on;
  const int sing = c->is_singular ? 1 : 0;
  static const bool fuse_ud = getenv("ISPH_NO_FUSE_UD") == nullptr;
  // TMA-pipelined sweep vs the register-tile kernel: measured (gpurun_out/r2_v2_*: 8M rows 365 vs 381 us per sweep, 1M rows 65 vs 62 us,
  // where the fixed costs of a launch dominate both) => TMA from 3M rows per GPU up; ISPH_UD_TMA=0 / 1 forces one of them
  static const int ud_tma_env = getenv("ISPH_UD_TMA") ? atoi(getenv("ISPH_UD_TMA")) : -1;
  const bool ud_tma = ud_tma_env >= 0 ? ud_tma_env != 0 : n >= 3000000;
  // alternating sweep direction (each sweep starts on the rows the previous one touched last): measured neutral on B200 — the L2 does not
  // keep the tail of a 200+ MB stream in a usable way (gpurun_out/r2_v2_oldpp_* vs r2_v2_old_*) — so it is off unless ISPH_PINGPONG=1
  static const bool pingpong = getenv("ISPH_PINGPONG") != nullptr;
  int sweep_dir = 0; auto sweep = [&]() { if (!pingpong) return 0; sweep_dir ^= 1; return sweep_dir; };
  static const int ud_hv = getenv("ISPH_UD_HV") ? atoi(getenv("ISPH_UD_HV")) : 2;            // 64-row halves per tile of the fused sweep
  static const int udcap = getenv("ISPH_UDGRID") ? atoi(getenv("ISPH_UDGRID")) : (ud_hv == 1 ? 444 : 296);       // 148 SMs x resident CTAs
  const int gud = std::max(1, std::min(udcap, ceil_div(n, 64 * ud_hv)));
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
      { ProfScope ps(c, "op_apply"); spmv(c, zj, vn, 1, ld, ld); } dbg(c, "op_apply");          // y = A z_j ; the PoissonProjection tail rides on the Gram-Schmidt sweep
      // ISPH_PINGPONG=1: the sweeps over the basis alternate direction
      { ProfScope ps(c, "multidot0"); launch_multidot(c, V, j + 1, vn, 0, sweep()); } dbg(c, "multidot0");
      if (fuse_ud && j + 1 > 8) {                                // one sweep: first update + second-pass coefficients
        ProfScope ps(c, "update0+dot1"); P2PRed pr = halo_p2p_ticket(c); const int rv = sweep();
        if (!(ud_tma && launch_update_dot_tma(c, V, ld, m + 1, j + 1, vn, nvp, n, rv, S, cnt + 6, pr))) {                     // flag word 14
          if (ud_hv == 1) launch_pdl(k_update_dot<1>, dim3(gud), dim3(VB), 0, c->stream, V, ld, j + 1, vn, nvp, n, rv, S, c->red.p, cnt + 6, pr);
          else launch_pdl(k_update_dot<2>, dim3(gud), dim3(VB), 0, c->stream, V, ld, j + 1, vn, nvp, n, rv, S, c->red.p, cnt + 6, pr);
        }
        ++c->launches;
        if (c->nranks > 1 && pr.nranks <= 1) halo_allreduce(c, S + S_H2, j + 2);
      } else {
        { ProfScope ps(c, "update0"); launch_pdl(k_cgs_update, dim3(g), dim3(VB), 0, c->stream, V, ld, j + 1, vn, nvp, n, S, sweep()); ++c->launches; }
        { ProfScope ps(c, "multidot1"); launch_multidot(c, V, j + 1, vn, 1, sweep()); }
      }
      dbg(c, "update0/dot1");
      ++iters;
      if (j + 1 < m) {                                           // second update + normalisation + next preconditioned vector + Givens, one sweep
        ProfScope ps(c, "finish");
        double *zn = flex ? Z + (size_t)(j + 1) * ld : Z;
        PrePush pp; pp.plan = nullptr; pp.sp = pp.sd = nullptr; pp.seq = 0;
        if (jacobi_fused) { halo_prepush_begin(c, zn, &pp);       // z_{j+1} is the next SpMV input: its halo rows leave from this kernel
          launch_pdl(k_finish, dim3(g + 1), dim3(VB), 0, c->stream, V, ld, j + 1, vn, n, S, sing, dp.invdiag, dp.scale, dp.post, zn, c->h_scal.p + 8, iters, sweep(), pp); ++c->launches; }
        else { launch_pdl(k_finish, dim3(g + 1), dim3(VB), 0, c->stream, V, ld, j + 1, vn, n, S, sing, (const double *)nullptr, 1.0, 0, use_prec ? (double *)nullptr : zn, c->h_scal.p + 8, iters, sweep(), pp); ++c->launches; }
      } else { k_givens<<<1, 32, 0, c->stream>>>(S, j, sing, c->h_scal.p + 8, iters); ++c->launches; }   // last column of the cycle: v_{m} is never used
      dbg(c, "finish");
      CUDA_CHECK(cudaEventRecord(ev[j], c->stream));
      if (j + 1 < m && !jacobi_fused && use_prec) { ProfScope ps(c, "precond"); apply_prec(c, true, vn, flex ? Z + (size_t)(j + 1) * ld : Z); }
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
  // a push whose SpMV was never issued (converged one step late): take it anyway so that the staging cells are re-armed
  if (c->prepush_x) { halo_wait_unstage(c, const_cast<double *>(c->prepush_x), c->prepush_seq); c->prepush_x = nullptr; }
  *iters_out = iters; *relres_out = scale > 0.0 ? res / scale : 0.0;
  return converged ? 1 : 0;
}

static int cg_solve(Ctx *c, bool use_prec, double *x, double *b, int *iters_out, double *relres_out) {
  const int n = c->A.n, ld = c->ld, g = vgrid(c, n);
  double *S = c->hbuf.p, *r = c->wk.p, *z = c->V.p, *p = c->V.p + (size_t)ld, *Ap = c->V.p + (size_t)2 * ld; unsigned *cnt = (unsigned *)c->flag.p + 8;
  const DiagPrec dp = diag_prec_of(c, use_prec);
  const int RZ[2] = {S_RZ, S_BETA};                              // r.z of the current / next iteration, used alternately
  cudaEvent_t ev[2]; for (auto &e : ev) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  auto precdot = [&](int slot) {
    if (!dp.on) apply_prec(c, use_prec, r, z);
    P2PRed pr = halo_p2p_ticket(c);
    k_cg_precdot<<<g, VB, 0, c->stream>>>(r, z, dp.on ? dp.invdiag : nullptr, dp.scale, dp.post, n, c->red.p, cnt, S + slot, pr); ++c->launches;
    if (c->nranks > 1 && pr.nranks <= 1) halo_allreduce(c, S + slot, 1);
  };
  auto direction = [&](int cur, int nxt, int first) {
    PrePush pp; halo_prepush_begin(c, p, &pp);                   // p is the next SpMV input: its halo rows leave from this kernel
    k_cg_direction<<<g, VB, 0, c->stream>>>(p, z, S, cur, nxt, n, first, pp); ++c->launches;
  };
  // R = b - A x ; Z = M^-1 R ; P = Z ; rz = R.Z
  if (c->init_type != ISPH_INIT_ZERO) op_apply(c, x, Ap, false);
  { P2PRed pr = halo_p2p_ticket(c);
    k_residual<<<g, VB, 0, c->stream>>>(b, c->init_type == ISPH_INIT_ZERO ? nullptr : Ap, r, n, c->red.p, cnt, S + S_TMP, pr); ++c->launches;
    if (c->nranks > 1 && pr.nranks <= 1) halo_allreduce(c, S + S_TMP, 1); }
  const double scale = sqrt(read_scalar(c, S + S_TMP)); double res = scale; int iters = 0; bool converged = false;
  if (scale == 0.0 || res / scale <= c->sp.tol) converged = true;
  else {
    precdot(RZ[0]);
    direction(RZ[0], RZ[0], 1);
    while (true) {
      ++iters;
      const int cur = RZ[(iters - 1) & 1], nxt = RZ[iters & 1];
      if (c->is_singular) { op_apply(c, p, Ap, false); dot_dev(c, p, Ap, n, S + S_PAP); }
      else spmv(c, p, Ap, 1, ld, ld, p, S + S_PAP);               // Ap = A p with p.Ap reduced in the SpMV epilogue
      { P2PRed pr = halo_p2p_ticket(c); const bool fold = c->nranks == 1 || pr.nranks > 1;       // residual published by the kernel that reduces it
        k_cg_update<<<g, VB, 0, c->stream>>>(x, r, p, Ap, S, cur, n, c->red.p, cnt, pr, fold ? c->h_scal.p + 8 : nullptr, iters); ++c->launches;
        if (!fold) { halo_allreduce(c, S + S_TMP, 1); k_cg_publish<<<1, 32, 0, c->stream>>>(S, c->h_scal.p + 8, iters); ++c->launches; } }
      CUDA_CHECK(cudaEventRecord(ev[iters & 1], c->stream));
      // next direction, enqueued before the residual of this step is inspected
      precdot(nxt);
      direction(cur, nxt, 0);
      CUDA_CHECK(cudaEventSynchronize(ev[iters & 1]));
      res = c->h_scal.p[8 + iters];
      if (res / scale <= c->sp.tol) { converged = true; break; }
      if (iters >= c->sp.max_iters) break;
    }
  }
  for (auto &e : ev) cudaEventDestroy(e);
  // a push whose SpMV was never issued: take it anyway so that the staging cells are re-armed
  if (c->prepush_x) { halo_wait_unstage(c, const_cast<double *>(c->prepush_x), c->prepush_seq); c->prepush_x = nullptr; }
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
  ISPH_REQUIRE(!(is_cg && use_prec && !c->prec_parent && c->pp.type == "ML" && c->pp.ml_pre != c->pp.ml_post),
               "Block CG needs a symmetric preconditioner: give the ML stand-in the same number of pre- and post-smoothing sweeps (smoother: sweeps)");
  c->prof_phases = getenv("ISPH_PROFILE") != nullptr;
  halo_recover(c);                                               // a peer wait that timed out in an earlier solve: re-arm the slots on all ranks
  std::string tname = std::string("solve") + (label ? label : "");
  c->tic(tname.c_str());
  c->hbuf.ensure(S_TOTAL); c->flag.ensure(16); c->red.ensure((size_t)8 * 592 * 18 + 1024 + (size_t)A.nslices / 8 + 64);
  c->wk.ensure((size_t)ld + c->nall + 3 * (size_t)ld);
  c->h_scal.ensure(16 + c->sp.max_iters + m + 8);
  c->V.ensure((size_t)(is_cg ? 3 : m + 1) * ld);
  c->Z.ensure((size_t)(is_cg ? 1 : (c->sp.flexible ? m : 1)) * ld);
  CUDA_CHECK(cudaMemsetAsync(c->flag.p + 8, 0, 8 * sizeof(int), c->stream));
  CUDA_CHECK(cudaMemsetAsync(c->hbuf.p + S_NSEC, 0, sizeof(double), c->stream));
  // initial solution (setInitialSolution, solver_lin.cpp:141-147): applied here, on the device
  const size_t xl = (size_t)ld * c->x_nvec;
  if (c->init_type == ISPH_INIT_ZERO) CUDA_CHECK(cudaMemsetAsync(c->xs.p, 0, sizeof(double) * xl, c->stream));
  else if (c->init_type == ISPH_INIT_VALUE) { k_fill<<<vgrid(c, (int)xl), VB, 0, c->stream>>>(c->xs.p, c->init_val, (int)xl); ++c->launches; }
  else if (c->init_type == ISPH_INIT_RANDOM) { for (int q = 0; q < c->x_nvec; ++q) { k_random<<<g, VB, 0, c->stream>>>(c->xs.p + (size_t)q * ld, A.external ? nullptr : c->tag.p, n, 11 + q); ++c->launches; } }
  else if (c->x_host) {     // caller's x is the initial guess (Helmholtz: x = v, pair_isph.cpp:932-941)
    for (int q = 0; q < c->x_nvec; ++q) CUDA_CHECK(cudaMemcpyAsync(c->xs.p + (size_t)q * ld, c->x_host + (size_t)q * c->x_lda, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
  }
  load_from_host(c); c->b_dev_fresh = false;                     // borrowed b: the host View is uploaded unless a device functor wrote the load vector since the last solve
  double *S = c->hbuf.p;
  if (c->is_singular) {     // createNullVector (solver_lin.cpp:59-77) ; b -= (b.n) n (solver_lin_belos.h:138-144)
    c->nullvec.ensure((size_t)ld + 256);                        // + one tile: the TMA sweep copies whole tiles of it
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
  CUDA_CHECK(cudaMemcpyAsync(c->h_scal.p + 1, c->hbuf.p + S_NSEC, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  c->toc(tname.c_str());
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  c->last_second_passes = (long long)c->h_scal.p[1];
  if (c->prof_phases && c->rank == 0) {
    fprintf(stderr, "[isph profile] %s: %d iterations\n", tname.c_str(), iters_tot);
    for (auto &kv : c->phase_ev) { size_t u = c->phase_used[kv.first]; double tot = 0.0; for (size_t q = 0; q + 1 < u; q += 2) { float ms = 0.f; cudaEventElapsedTime(&ms, kv.second[q], kv.second[q + 1]); tot += ms; }
      fprintf(stderr, "[isph profile]   %-16s %6zu x  avg %8.2f us  total %8.3f ms\n", kv.first.c_str(), u / 2, u ? 1e3 * tot / (u / 2) : 0.0, tot); c->phase_used[kv.first] = 0; }
  }
  if (c->nranks > 1) ISPH_REQUIRE(!halo_fault(c), "peer exchange timed out: a rank stopped responding");
  if (use_prec && c->prec_kind == 3) ISPH_REQUIRE(!ilu_fault(c), "ILU(0): a dependency wait timed out or a factor entry is not a number (zero pivot?)");
  c->last_iters = iters_tot; c->last_converged = conv_all; c->last_relres = relres;
  c->init_type = -1;
}

// ---- Newton's method for the Poisson-Boltzmann problem ------------------------------------------------------------------
// What NOX does for PairISPH::computePoissonBoltzmann (pair_isph.cpp:572-600) with the reference's default lists
// (solver_nox_impl.h:76-160): line search "Full Step", direction Newton, stop when NormF <= tol_f AND NormUpdate <= tol_update
// (NOX::StatusTest::NormF / NormUpdate, 2-norms scaled by sqrt(N)), at most max_newton iterations.  NOX itself is third-party
// and out of scope; what is on the path is what each iteration calls back into: computeF (pb_residual), computeJacobian
// (pb_jacobian) and the Jacobian solve, which runs through solver_solve with the context's Krylov / preconditioner lists.
// psi (field ISPH_F_PSI) is the initial guess and the result; F stays in the load vector.
__global__ void __launch_bounds__(VB) k_add_rows(double *field, const double *dx, int n) { for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) field[i] += dx[i]; }

void pb_newton(Ctx *c, bool mh, bool linearized, double ezcb, double psiref, double gamma, const double *d_extra, int max_newton, double tol_f, double tol_update,
               bool use_prec, int *newton_iters, int *linear_iters, double *normf_out, int *converged_out) {
  ISPH_REQUIRE(c->A.built && !c->A.external, "isph_pb_newton: build the graph first");
  ISPH_REQUIRE(c->x_nvec == 1 && c->b_nvec == 1 && c->xs.p && c->bs.p, "isph_pb_newton: create one solution and one load vector first");
  const int n = c->A.n, g = vgrid(c, n);
  c->hbuf.ensure(S_TOTAL); c->flag.ensure(16); c->red.ensure((size_t)8 * 592 * 18 + 1024 + (size_t)c->A.nslices / 8 + 64); c->h_scal.ensure(16 + c->sp.max_iters + c->sp.num_blocks + 8);
  CUDA_CHECK(cudaMemsetAsync(c->flag.p + 8, 0, 8 * sizeof(int), c->stream));
  double *S = c->hbuf.p;
  double nglobal = (double)n;
  if (c->nranks > 1) { CUDA_CHECK(cudaMemcpyAsync(S + S_TMP, &nglobal, sizeof(double), cudaMemcpyHostToDevice, c->stream)); halo_allreduce(c, S + S_TMP, 1); nglobal = read_scalar(c, S + S_TMP); }
  int k = 0, lin_total = 0; bool converged = false; double nf = 0.0, nup = 0.0;
  for (;; ++k) {
    pb_residual(c, mh, linearized, ezcb, psiref, gamma, d_extra, c->bs.p);                       // F(psi_k)
    dot_dev(c, c->bs.p, nullptr, n, S + S_TMP); nf = sqrt(read_scalar(c, S + S_TMP) / nglobal);
    if (!(nf == nf)) break;                                                                       // NOX::StatusTest::FiniteValue
    if (k > 0 && nf <= tol_f && nup <= tol_update) { converged = true; break; }
    if (k >= max_newton) break;
    pb_jacobian(c, mh, linearized, ezcb, psiref, gamma);                                          // J(psi_k): diagonal refreshed, Laplacian block kept
    k_scale_by<<<g, VB, 0, c->stream>>>(c->bs.p, -1.0, n); ++c->launches;                         // J dpsi = -F
    c->init_type = ISPH_INIT_ZERO; c->b_dev_fresh = true;         // -F was formed on the device
    solver_solve(c, use_prec, "PoissonBoltzmannJacobian"); lin_total += c->last_iters;
    k_add_rows<<<g, VB, 0, c->stream>>>(c->field[ISPH_F_PSI].p, c->xs.p, n); ++c->launches;       // full step
    dot_dev(c, c->xs.p, nullptr, n, S + S_TMP); nup = sqrt(read_scalar(c, S + S_TMP) / nglobal);
  }
  forward_comm(c, ISPH_F_PSI);                                                                    // pair_isph.cpp:595-598
  *newton_iters = k; *linear_iters = lin_total; *normf_out = nf; *converged_out = converged ? 1 : 0;
}

}  // namespace isph
