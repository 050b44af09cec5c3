// SpMV on the SELL-32 layout: y = A x (Epetra_CrsMatrix::Multiply/Apply as called by Belos and by
// functor_incomp_navier_stokes_helmholtz.h:90; solver_lin.h:133).
//
// One thread per row, one warp per slice: the warp's loads of `col` (128 B) and `val` (256 B) for entry k of its 32
// rows are single fully-used cache lines, streamed with evict-first hints so that they do not displace x from L2/L1;
// x is gathered through the read-only path.  Because rows follow the particle order, the k-th neighbours of 32
// consecutive rows are (nearly) consecutive columns, so the gather touches 2-3 lines per warp instead of 32 —
// this is what keeps the kernel HBM-bound rather than L1-wavefront-bound ("SpMV").
// Algorithmic traffic: 12 B per stored entry + 8 B x + 8 B y (+ slice metadata) per row.
#include "isph_internal.h"

namespace isph {

__device__ __forceinline__ void spmv_dot_finish(double v, double *partials, unsigned *counter, double *out, const P2PRed &pr) {
  __shared__ double sm[8]; __shared__ bool last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0.0; for (int w = 0; w < 8; ++w) t += sm[w]; partials[blockIdx.x] = t; __threadfence(); last = (atomicAdd(counter, 1u) == gridDim.x - 1); }
  __syncthreads();
  if (last) {
    __threadfence();
    double t = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += 256) t += __ldcg(partials + b);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    __syncthreads();
    if (lane == 0) sm[warp] = t;
    __syncthreads();
    if (threadIdx.x == 0) { double r = 0.0; for (int w = 0; w < 8; ++w) r += sm[w]; *out = r; *counter = 0u; }
    if (pr.nranks > 1) p2p_allreduce_block(pr, out, 1);      // sum over ranks through the NVLink mailboxes
  }
}

// DOT: additionally out[0] = sum_i y_i d_i (the (y.n) of PoissonProjection::Apply, solver_lin.h:133-137) reduced in the
// epilogue: block partials in block order, combined by the last block to finish — saves a pass over y per iteration.
template <int NV, bool DOT> __global__ void __launch_bounds__(256, 8)     // 32 registers: 8 CTAs = 2048 threads per SM
k_spmv_sell(const long long *__restrict__ slice_off, const int *__restrict__ slice_len, const int *__restrict__ col,
            const double *__restrict__ val, int n, int nslices, const double *__restrict__ x, int ldx, double *__restrict__ y, int ldy,
            const double *__restrict__ dvec, double *partials, unsigned *counter, double *out, P2PRed pr) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x, s = row >> 5;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // the reduction kernel that follows may be scheduled into the tail of this grid (it waits for it)
  if (!DOT && s >= nslices) return;
  // DOT: warps past the last slice do no row work (slen = 0) but fall through to the ONE spmv_dot_finish call site below, so
  // every thread of the block meets the same barrier instructions
  const bool active = s < nslices;
  const long long base = active ? slice_off[s] + (row & 31) : 0;
  const int slen = active ? slice_len[s] : 0;
  const int *cp = col + base; const double *vp = val + base;
  double acc[NV];
#pragma unroll
  for (int q = 0; q < NV; ++q) acc[q] = 0.0;
  int k = 0;
  for (; k + 4 <= slen; k += 4) {
    const int c0 = __ldcs(cp + 32 * (k + 0)), c1 = __ldcs(cp + 32 * (k + 1)), c2 = __ldcs(cp + 32 * (k + 2)), c3 = __ldcs(cp + 32 * (k + 3));
    const double v0 = __ldcs(vp + 32 * (k + 0)), v1 = __ldcs(vp + 32 * (k + 1)), v2 = __ldcs(vp + 32 * (k + 2)), v3 = __ldcs(vp + 32 * (k + 3));
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      const double *xq = x + (size_t)q * ldx;
      acc[q] += v0 * __ldg(xq + c0); acc[q] += v1 * __ldg(xq + c1); acc[q] += v2 * __ldg(xq + c2); acc[q] += v3 * __ldg(xq + c3);
    }
  }
  for (; k < slen; ++k) {
    const int c0 = __ldcs(cp + 32 * k); const double v0 = __ldcs(vp + 32 * k);
#pragma unroll
    for (int q = 0; q < NV; ++q) acc[q] += v0 * __ldg(x + (size_t)q * ldx + c0);
  }
  if (row < n) {
#pragma unroll
    for (int q = 0; q < NV; ++q) y[(size_t)q * ldy + row] = acc[q];
  }
  if (DOT) spmv_dot_finish(row < n ? acc[0] * dvec[row] : 0.0, partials, counter, out, pr);
}


// (Two variants were built and measured in round 2 and are not kept.  (1) The column stream as 16-bit deltas along a row with
// 32-bit escapes — 10 B per stored entry instead of 12: SLOWER, 250 us against 226 us per launch on the 1M-row brick and 1.77 ms
// against 1.71 ms on 8M rows (gpurun_out/r2_c16_*.json vs r2_c32_*.json): the kernel issues the same number of memory requests
// for fewer bytes and gains a serial add chain per row, i.e. at 95 % of the HBM peak it is also at the limit of its request rate.
// (2) A persistent variant in which every warp draws its next slice from a global ticket counter, measured on the 1M-row brick:
// 233 us per launch against 226 us for this kernel — the hardware CTA scheduler already balances at the granularity that matters,
// and consecutive slices on one SM share the x gather in L1.  gpurun_out/r2_dyn*_c2.json; not kept.)

// With several ranks the off-rank x entries (Epetra_Import) land behind the owned rows of x before the multiply starts —
// pushed by the kernel that produced x (krylov.cu: k_finish) or by halo_exchange — so the multiply itself is the same
// branch-free kernel on one GPU and on eight (reading the staging buffer from inside the SpMV cost 12 % of its bandwidth,
// and splitting it into a halo-free and a halo phase inside one kernel was slower still: profiles/r01_halo_variants.md).
void spmv(Ctx *c, const double *x, double *y, int nvec, int ldx, int ldy, const double *dot_vec, double *dot_out) {
  Matrix &A = c->A; ISPH_REQUIRE(A.built, "spmv: no matrix");
  const int grid = ceil_div((long long)A.nslices * 32, 256);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (c->prof_spmv) {      // per-launch device timing on the launching stream (bench.py roofline); with several ranks it includes the import
    if (c->prof_used + 2 > c->prof_ev.size()) { c->prof_ev.resize(c->prof_ev.size() + 512, nullptr); for (size_t q = c->prof_used; q < c->prof_ev.size(); ++q) if (!c->prof_ev[q]) CUDA_CHECK(cudaEventCreate(&c->prof_ev[q])); }
    e0 = c->prof_ev[c->prof_used++]; e1 = c->prof_ev[c->prof_used++]; CUDA_CHECK(cudaEventRecord(e0, c->stream));
  }
  if (c->nranks > 1 && nvec == 1 && c->prepush_x == x) { halo_wait_unstage(c, const_cast<double *>(x), c->prepush_seq); c->prepush_x = nullptr; }   // pushed by the producer of x
  else if (c->nranks > 1) halo_exchange(c, const_cast<double *>(x), nvec, ldx);
  P2PRed none; none.tab = nullptr; none.seq = 0; none.nranks = 1;
#define SPMV_ARGS(xx, yy, dv, pr) A.slice_off.p, A.slice_len.p, A.col.p, A.val.p, A.n, A.nslices, xx, ldx, yy, ldy, dv, c->red.p, (unsigned *)c->flag.p + 13, dot_out, pr
  int done = 0;
  while (done < nvec) {
    const int nv = nvec - done >= 3 ? 3 : (nvec - done >= 2 ? 2 : 1);
    const double *xx = x + (size_t)done * ldx; double *yy = y + (size_t)done * ldy;
    if (nv == 3) k_spmv_sell<3, false><<<grid, 256, 0, c->stream>>>(SPMV_ARGS(xx, yy, nullptr, none));
    else if (nv == 2) k_spmv_sell<2, false><<<grid, 256, 0, c->stream>>>(SPMV_ARGS(xx, yy, nullptr, none));
    else if (dot_vec && nvec == 1) {
      ISPH_REQUIRE(c->red.cap >= (size_t)grid, "spmv: reduction workspace too small");
      P2PRed pr = halo_p2p_ticket(c);
      k_spmv_sell<1, true><<<grid, 256, 0, c->stream>>>(SPMV_ARGS(xx, yy, dot_vec, pr));
      if (c->nranks > 1 && pr.nranks <= 1) halo_allreduce(c, dot_out, 1);
    }
    else k_spmv_sell<1, false><<<grid, 256, 0, c->stream>>>(SPMV_ARGS(xx, yy, nullptr, none));
    ++c->launches; done += nv;
  }
#undef SPMV_ARGS
  if (e1) CUDA_CHECK(cudaEventRecord(e1, c->stream));
}

}  // namespace isph
