// Device-side pieces of the NVLink peer-memory exchange (see halo.cu for the set-up and the protocol).
//
// Every rank owns a mailbox in its own HBM that all peers can address through cudaIpc mappings.  A reduction / halo
// "call" carries a sequence number that is identical on all ranks (every rank executes the same sequence of calls on
// its stream); a slot ring of MB_SLOTS entries is indexed by seq % MB_SLOTS.  Producer: NVLink stores of the payload into
// every peer's slot, __threadfence_system(), then a monotonic flag (= seq + 1).  Consumer: polls the flags in its OWN
// mailbox (local memory, written by the peers), then reads the payloads in rank order, so every rank forms the
// bit-identical sum.  A rank can never overwrite a slot a peer still reads: to get MB_SLOTS calls ahead it would need
// the peer's contributions to the calls in between, which the peer only issues after it has consumed the older call.
#pragma once
#include <cuda_runtime.h>

#define MB_STRIDE 72          /* 64 payload doubles + 1 flag word (+ padding), in doubles */
#define MB_SLOTS 4
#define ISPH_MAX_RANKS 8

namespace isph {

struct P2PTab {               // lives in DEVICE memory (one per context): indexed dynamically, so it must not be a kernel
  double *box[ISPH_MAX_RANKS];   // parameter (a by-value struct that is indexed at run time is copied to local memory by
  double *mine;                  // every thread of the kernel - measured: -8 % on the SpMV)
  int nranks, rank;
  int *fault;
};
struct P2PRed {               // passed by value to kernels; tab == nullptr means "no exchange"
  const P2PTab *tab;
  unsigned long long seq;
  int nranks;                 // copy of tab->nranks (1 when disabled) for cheap host/device tests
};

// All threads of ONE block (>= 64 threads) call this; vals[0..count) (global memory, count <= 64) is replaced by the
// sum over ranks.  Used in the "last block" epilogue of the reduction kernels, so no other block of the grid is waiting.
static __device__ __noinline__ void p2p_allreduce_block(const P2PRed &r, double *vals, int count) {
  const P2PTab *T = r.tab; const int nr = T->nranks, me = T->rank;
  const int slot = (int)(r.seq % MB_SLOTS), t = threadIdx.x;
  __syncthreads();
  if (t < count) {
    const double v = __ldcg(vals + t);
    for (int p = 0; p < nr; ++p) T->box[p][(size_t)(slot * nr + me) * MB_STRIDE + t] = v;
    __threadfence_system();
  }
  __syncthreads();
  if (t < nr) {
    volatile unsigned long long *f = reinterpret_cast<volatile unsigned long long *>(T->box[t] + (size_t)(slot * nr + me) * MB_STRIDE + 64);
    *f = r.seq + 1;                                              // my contribution to rank t is complete
    volatile unsigned long long *w = reinterpret_cast<volatile unsigned long long *>(T->mine + (size_t)(slot * nr + t) * MB_STRIDE + 64);
    long long spins = 0;
    while (*w < r.seq + 1) { if (++spins > (1ll << 31)) { *T->fault = 1; break; } }   // bounded: a dead peer must not hang the GPU
    __threadfence_system();
  }
  __syncthreads();
  if (t < count) {
    double s = 0.0;
    for (int p = 0; p < nr; ++p) s += __ldcv(T->mine + (size_t)(slot * nr + p) * MB_STRIDE + t);
    vals[t] = s;
  }
  __syncthreads();
}

}  // namespace isph
