// Device-side pieces of the NVLink peer-memory exchange (see halo.cu for the set-up).
//
// Every rank owns a mailbox (small all-reduces) and a halo staging buffer in its own HBM that all peers can address
// through cudaIpc mappings.  A reduction / halo "call" carries a sequence number that is identical on all ranks (every
// rank executes the same sequence of calls on its stream); a ring of MB_SLOTS slots is indexed by seq % MB_SLOTS.
//
// Protocol: THE DATA IS ITS OWN FLAG.  Every 8-byte cell of a slot rests at a sentinel bit pattern (all ones: a NaN that
// arithmetic never produces; a payload that happens to carry it is canonicalised first).  Producer: one plain NVLink
// store per value — no fence, no flag, no counter.  Consumer: polls the cells in its OWN buffer (local memory) until they
// differ from the sentinel, uses them, and writes the sentinel back for the slot's next use.  One one-way NVLink store
// latency per exchange instead of store -> system fence -> flag -> poll.  All-reduce payloads are summed in rank order,
// so every rank forms the bit-identical sum.  A rank can never overwrite a cell a peer has not consumed yet: to get
// MB_SLOTS calls ahead it would need the peer's contributions to the calls in between, which the peer only issues after
// it has consumed (and re-armed) the older call.  For the halo staging slots this argument needs every receiver of a peer to
// also send to that peer (halo_setup checks the symmetry of the exchange matrix on all ranks and falls back to NCCL send/recv
// otherwise).  Waits are bounded in wall-clock time: a dead peer raises the fault word, after which every wait on this GPU
// drains immediately.
#pragma once
#include <cuda_runtime.h>

#define MB_STRIDE 64          /* payload doubles per (slot, rank) cell block */
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

#define ISPH_SENTINEL 0xFFFFFFFFFFFFFFFFull
__device__ __forceinline__ double p2p_payload(double v) {      // a value that may be stored into a peer cell
  return (unsigned long long)__double_as_longlong(v) == ISPH_SENTINEL ? __longlong_as_double(0x7FF8000000000000ll) : v;
}
// poll a cell of this rank's own buffer until a peer's value has landed, then re-arm it.  The wait is bounded by WALL-CLOCK
// time (%globaltimer; fault[1] = limit in milliseconds, ISPH_P2P_TIMEOUT_MS, default 20 s), not by a poll count: a peer that
// is merely late to enqueue (host I/O, a first-touch cudaMalloc) is waited for; a dead one raises the fault word fault[0], after
// which every wait on this GPU drains immediately, the running solve reports the failure (solver_solve) and the next solve
// re-arms all slots collectively (halo_recover).
__device__ __forceinline__ unsigned long long p2p_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ double p2p_take(double *cell, int *fault) {
  volatile unsigned long long *q = reinterpret_cast<volatile unsigned long long *>(cell);
  unsigned long long bits = *q, t0 = 0; int spins = 0;
  while (bits == ISPH_SENTINEL) {
    if ((++spins & 1023) == 0) {
      volatile int *f = reinterpret_cast<volatile int *>(fault);
      if (f[0]) return 0.0;
      const unsigned long long now = p2p_now();
      if (t0 == 0) t0 = now; else if (now - t0 > (unsigned long long)f[1] * 1000000ull) { f[0] = 1; return 0.0; }
    }
    bits = *q;
  }
  *q = ISPH_SENTINEL;
  return __longlong_as_double((long long)bits);
}

// All threads of ONE block (>= 64 threads) call this; vals[0..count) (global memory, count <= 64) is replaced by the
// sum over ranks.  Used in the "last block" epilogue of the reduction kernels, so no other block of the grid is waiting.
static __device__ __noinline__ void p2p_allreduce_block(const P2PRed &r, double *vals, int count) {
  const P2PTab *T = r.tab; const int nr = T->nranks, me = T->rank;
  const int slot = (int)(r.seq % MB_SLOTS), t = threadIdx.x;
  __syncthreads();
  if (t < count) {
    const double v = p2p_payload(__ldcg(vals + t));
    for (int p = 0; p < nr; ++p) *reinterpret_cast<volatile double *>(T->box[p] + (size_t)(slot * nr + me) * MB_STRIDE + t) = v;
    double s = 0.0;
    for (int p = 0; p < nr; ++p) s += p2p_take(T->mine + (size_t)(slot * nr + p) * MB_STRIDE + t, T->fault);
    vals[t] = s;
  }
  __syncthreads();
}

}  // namespace isph

