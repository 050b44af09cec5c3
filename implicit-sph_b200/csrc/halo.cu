// Multi-GPU plumbing: one process per GPU, rows partitioned along LAMMPS subdomains (Epetra_Map of owned tags,
// pair_isph.cpp:1258-1259).  Replaces what Epetra does through MPI underneath the reference's calls:
//   * column map / importer  -> halo plan: every ghost atom whose tag is owned by another rank gets a halo column
//     (distinct remote tags, grouped by owner so that each peer's data lands contiguously behind the owned rows);
//   * Epetra_Import in CrsMatrix::Multiply -> the requested x entries are stored straight into the peers' halo staging
//     buffers over NVLink (peer memory mapped with cudaIpc; protocol in p2p_device.cuh: the data is its own flag), either
//     by the kernel that PRODUCES x (k_finish in krylov.cu pushes z_{j+1} while it writes it, so the import overlaps the
//     end of the Arnoldi step) or by a small gather kernel; a second small kernel takes the arrived values behind the
//     owned rows of x and re-arms the slot; the SpMV itself stays branch-free.  Fallback (no peer access): pack kernel
//     + grouped ncclSend/ncclRecv into the halo tail of x;
//   * Epetra_MpiComm::SumAll -> the last block of every reduction kernel exchanges its partial sums through peer
//     mailboxes (p2p_device.cuh); fallback ncclAllReduce;
//   * comm->forward_comm_pair (owner -> ghost field copy) -> the same plan with ncomp doubles per particle (NCCL).
// NCCL is loaded with dlopen so that the library also loads on hosts without it (single-GPU use, CPU-side ABI tests);
// it remains the bootstrap (unique id from the host program, all-gathers of plan data and of the IPC handles).
// The plan itself (isph_halo_plan_host) is pure host code without any CUDA call: tests/test_multirank_cpu.py drives it
// with a gloo transport at world_size 2.
#include "isph_internal.h"
#include "p2p_device.cuh"
#include <nccl.h>
#include <dlfcn.h>
#include <algorithm>
#include <tuple>
#include <cub/cub.cuh>

namespace isph {

struct NcclApi {
  void *dl = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool load() {
    if (dl) return true;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) { dl = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (dl) break; }
    if (!dl) return false;
#define L(sym) *(void **)(&sym) = dlsym(dl, "nccl" #sym); if (!sym) return false;
    L(GetUniqueId) L(CommInitRank) L(CommDestroy) L(AllReduce) L(AllGather) L(Send) L(Recv) L(GroupStart) L(GroupEnd) L(GetErrorString)
#undef L
    return true;
  }
};
static NcclApi g_nccl;
#define NCCL_CHECK(expr) do { ncclResult_t _r = (expr); if (_r != ncclSuccess) throw std::runtime_error(std::string(#expr) + ": " + g_nccl.GetErrorString(_r)); } while (0)

// stand-alone small all-reduce (reductions whose kernel does not carry the exchange in its epilogue)
__global__ void k_allreduce_p2p(P2PRed r, double *buf, int count) { p2p_allreduce_block(r, buf, count); }

// gather the x entries the peers asked for and store them into THEIR staging buffers over NVLink (fire and forget)
__global__ void __launch_bounds__(256) k_halo_push(const HaloDev *hp, unsigned long long seq, const double *x, int ldx, int nvec, const int *send_idx, int nsend) {
  const int slot = (int)(seq % MB_SLOTS);
  const int k = blockIdx.x * blockDim.x + threadIdx.x; if (k >= nsend) return;
  int p = 0; while (k >= hp->send_off[p + 1]) ++p;
  const int s = send_idx[k];
  double *dst = hp->peer[p] + (size_t)slot * 3 * hp->cap + hp->dst_off[p] + (k - hp->send_off[p]);
  for (int q = 0; q < nvec; ++q) *reinterpret_cast<volatile double *>(dst + (size_t)q * hp->cap) = p2p_payload(x[(size_t)q * ldx + s]);
}
// take the arrived halo values out of this rank's staging slot (poll until each has landed, re-arm the cell) and put them
// behind the owned rows of x
__global__ void __launch_bounds__(256) k_halo_unstage(const HaloDev *hp, unsigned long long seq, double *x, int ldx, int nlocal, int nhalo, int nvec) {
  const int slot = (int)(seq % MB_SLOTS);
  const int k = blockIdx.x * blockDim.x + threadIdx.x; if (k >= nhalo) return;
  double *cell = hp->mine + (size_t)slot * 3 * hp->cap + k;
  for (int q = 0; q < nvec; ++q) x[(size_t)q * ldx + nlocal + k] = p2p_take(cell + (size_t)q * hp->cap, hp->fault);
}

// ---- device-side plan (the role of Epetra's column map construction): distinct remote (owner, tag) pairs numbered in
// (owner, tag) order.  Same result as isph_halo_plan_host (the pure host restatement kept for the CPU tests).
__global__ void k_plan_keys(const int *tag, const int *owner, int nlocal, int nghost, int rank, int nranks, unsigned long long *key, int *val, int *err) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x; if (g >= nghost) return;
  const int o = owner[g]; unsigned long long k = ~0ull;
  if (o < 0 || o >= nranks) atomicAdd(err, 1);                   // a ghost whose tag no rank owns
  else if (o != rank) k = ((unsigned long long)(unsigned)o << 32) | (unsigned)tag[nlocal + g];
  key[g] = k; val[g] = g;
}
__global__ void k_plan_heads(const unsigned long long *key, int n, int *head) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  head[i] = (key[i] != ~0ull && (i == 0 || key[i] != key[i - 1])) ? 1 : 0;
}
__global__ void k_plan_assign(const unsigned long long *key, const int *val, const int *head, const int *inc, const int *owner, const int *owner_idx, int n, int nlocal,
                              int rank, int *ghost_col, int *request, int *recv_count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  const int g = val[i];
  if (key[i] == ~0ull) { if (owner[g] == rank) ghost_col[g] = owner_idx[g]; return; }
  const int slot = inc[i] - 1;
  ghost_col[g] = nlocal + slot;
  if (head[i]) { request[slot] = owner_idx[g]; atomicAdd(recv_count + (int)(key[i] >> 32), 1); }
}
// ghost_tag/ghost_owner/ghost_owner_idx: device arrays of nghost entries (ghost_tag = tag + nlocal).  Results on the device:
// ghost_col[nghost], request[<= nghost], counts[0..R) = values received per peer, counts[R] = number of unowned ghosts.
struct PlanWork { DevBuf<unsigned long long> gkey, gkey2; DevBuf<int> gval, gval2, head, inc; DevBuf<char> tmp;
  void release() { gkey.release(); gkey2.release(); gval.release(); gval2.release(); head.release(); inc.release(); tmp.release(); } };
static void plan_device(Ctx *c, PlanWork *h, const int *d_tag_all, const int *d_owner, const int *d_idx, int nl, int ng, int *d_ghost_col, int *d_request, int *d_counts) {
  const int R = c->nranks, B = 256;
  CUDA_CHECK(cudaMemsetAsync(d_counts, 0, sizeof(int) * (R + 1), c->stream));
  if (ng == 0) return;
  h->gkey.ensure(ng); h->gkey2.ensure(ng); h->gval.ensure(ng); h->gval2.ensure(ng); h->head.ensure(ng); h->inc.ensure(ng);
  k_plan_keys<<<ceil_div(ng, B), B, 0, c->stream>>>(d_tag_all, d_owner, nl, ng, c->rank, R, h->gkey.p, h->gval.p, d_counts + R);
  size_t t1 = 0, t2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, t1, h->gkey.p, h->gkey2.p, h->gval.p, h->gval2.p, ng, 0, 64, c->stream);
  cub::DeviceScan::InclusiveSum(nullptr, t2, h->head.p, h->inc.p, ng, c->stream);
  h->tmp.ensure(std::max(t1, t2));
  t1 = h->tmp.cap; cub::DeviceRadixSort::SortPairs(h->tmp.p, t1, h->gkey.p, h->gkey2.p, h->gval.p, h->gval2.p, ng, 0, 64, c->stream);
  k_plan_heads<<<ceil_div(ng, B), B, 0, c->stream>>>(h->gkey2.p, ng, h->head.p);
  t2 = h->tmp.cap; cub::DeviceScan::InclusiveSum(h->tmp.p, t2, h->head.p, h->inc.p, ng, c->stream);
  k_plan_assign<<<ceil_div(ng, B), B, 0, c->stream>>>(h->gkey2.p, h->gval2.p, h->head.p, h->inc.p, d_owner, d_idx, ng, nl, c->rank, d_ghost_col, d_request, d_counts);
  c->launches += 5;
}

struct Halo {
  ncclComm_t comm = nullptr; bool inited = false;
  // peer-memory exchange
  bool p2p = false; double *mbox = nullptr; double *mpeer[ISPH_MAX_RANKS] = {}; void *mopened[ISPH_MAX_RANKS] = {};
  double *hbox = nullptr; double *hpeer[ISPH_MAX_RANKS] = {}; void *hopened[ISPH_MAX_RANKS] = {}; long long hcap = 0;
  unsigned long long seq = 0, hseq = 0; int *fault = nullptr; std::vector<int> dst_off; P2PTab *d_tab = nullptr;
  // plan
  int nhalo = 0, nsend = 0;
  std::vector<int> recv_count, recv_off, send_count, send_off;
  DevBuf<int> send_idx, itmp, itmp2; DevBuf<long long> owner_tab; DevBuf<double> sendbuf, fieldbuf;
  // set-up workspace: grow-only like every other buffer of the path (isph_internal.h DevBuf) — nothing is allocated or freed per step
  DevBuf<int> mytags, alltags, gown, gidx, req, gcol_saved; PlanWork pw;
  PinBuf<int> h_small;                                           // pinned landing zone of the few integers the host needs
  // plan cache: the plan depends only on (nlocal, nghost, tags); when no rank's tag set changed since the last set-up it is reused
  bool plan_valid = false; unsigned long long plan_hash = 0; int plan_nl = -1, plan_ng = -1; long long setups = 0, reuses = 0;
  // device-resident exchange plan (halo kernels, push from the producer)
  DevBuf<HaloDev> d_plan; HaloDev hd_host; bool plan_ok = false;
  DevBuf<int> row_sp, row_sd, row_cur; DevBuf<char> cubtmp; bool rows_ok = false;      // per-row send list (push from the producer)
};

// per-row send list: count, scan, fill (order within a row is irrelevant)
__global__ void k_rowsend_count(const int *send_idx, int nsend, int *cnt) { const int k = blockIdx.x * blockDim.x + threadIdx.x; if (k < nsend) atomicAdd(cnt + send_idx[k], 1); }
__global__ void k_rowsend_fill(const HaloDev *hp, const int *send_idx, int nsend, int *cursor, int *sd) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x; if (k >= nsend) return;
  int p = 0; while (k >= hp->send_off[p + 1]) ++p;
  sd[atomicAdd(cursor + send_idx[k], 1)] = (p << 28) | (hp->dst_off[p] + (k - hp->send_off[p]));
}
__global__ void k_owner_tab(const int *all_tags, const int *nloc_all, int maxn, int nranks, int max_tag, long long *tab) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; if (t >= (long long)nranks * maxn) return;
  const int r = (int)(t / maxn), i = (int)(t % maxn); if (i >= nloc_all[r]) return;
  const int tag = all_tags[t]; if (tag <= max_tag) tab[tag] = ((long long)r << 32) | (unsigned)i;
}
__global__ void k_ghost_owner(const int *tag, const int *col_of_atom, const long long *tab, int nlocal, int nghost, int rank, int *owner, int *idx) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x; if (g >= nghost) return;
  const int a = nlocal + g, c = col_of_atom[a];
  if (c >= 0) { owner[g] = rank; idx[g] = c; return; }
  const long long e = tab[tag[a]];
  if (e < 0) { owner[g] = -1; idx[g] = -1; } else { owner[g] = (int)(e >> 32); idx[g] = (int)(e & 0xffffffffll); }
}
__global__ void k_pack(const double *x, const int *send_idx, int nsend, int nc, double *buf) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x; if (k >= nsend) return;
  const int s = send_idx[k];
  for (int q = 0; q < nc; ++q) buf[(size_t)k * nc + q] = x[(size_t)s * nc + q];
}
__global__ void k_ghost_from_halo(const int *col_of_atom, int nlocal, int nall, int nc, const double *halobuf, double *f) {
  const int a = nlocal + blockIdx.x * blockDim.x + threadIdx.x; if (a >= nall) return;
  const int c = col_of_atom[a]; if (c < nlocal) return;
  for (int q = 0; q < nc; ++q) f[(size_t)a * nc + q] = halobuf[(size_t)(c - nlocal) * nc + q];
}

// map `mine` (a cudaMalloc'ed buffer of this rank) into every peer: all-gather of the IPC handles over NCCL.
// Collective; returns false (on every rank alike) when any rank could not open a peer mapping.
static bool ipc_share(Ctx *c, Halo *h, double *mine, double **peer, void **opened) {
  const int R = c->nranks;
  cudaIpcMemHandle_t mh; bool ok = cudaIpcGetMemHandle(&mh, mine) == cudaSuccess; if (!ok) { cudaGetLastError(); memset(&mh, 0, sizeof(mh)); }
  char *dsend = nullptr, *dall = nullptr; const size_t hs = sizeof(mh);
  CUDA_CHECK(cudaMalloc(&dsend, hs)); CUDA_CHECK(cudaMalloc(&dall, hs * R));
  CUDA_CHECK(cudaMemcpy(dsend, &mh, hs, cudaMemcpyHostToDevice));
  NCCL_CHECK(g_nccl.AllGather(dsend, dall, hs, ncclChar, h->comm, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream));
  std::vector<cudaIpcMemHandle_t> all(R); CUDA_CHECK(cudaMemcpy(all.data(), dall, hs * R, cudaMemcpyDeviceToHost));
  cudaFree(dsend); cudaFree(dall);
  for (int p = 0; p < R && ok; ++p) {
    if (p == c->rank) { peer[p] = mine; continue; }
    void *ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, all[p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
    opened[p] = ptr; peer[p] = (double *)ptr;
  }
  double *flag = nullptr; CUDA_CHECK(cudaMalloc(&flag, sizeof(double))); const double bad = ok ? 0.0 : 1.0;   // every rank must take the same path
  CUDA_CHECK(cudaMemcpy(flag, &bad, sizeof(double), cudaMemcpyHostToDevice));
  NCCL_CHECK(g_nccl.AllReduce(flag, flag, 1, ncclDouble, ncclSum, h->comm, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream));
  double tot = 1.0; CUDA_CHECK(cudaMemcpy(&tot, flag, sizeof(double), cudaMemcpyDeviceToHost)); cudaFree(flag);
  return tot == 0.0;
}
static void ipc_close(void **opened) { for (int p = 0; p < ISPH_MAX_RANKS; ++p) if (opened[p]) { cudaIpcCloseMemHandle(opened[p]); opened[p] = nullptr; } }

static Halo *get(Ctx *c) {
  if (!c->halo) c->halo = new Halo();
  Halo *h = c->halo;
  if (!h->inited) {
    ISPH_REQUIRE(g_nccl.load(), "nranks > 1: libnccl.so.2 could not be loaded");
    ISPH_REQUIRE(c->have_nccl_id, "nranks > 1: no NCCL unique id");
    ncclUniqueId id; memcpy(&id, c->nccl_id, sizeof(id) < 128 ? sizeof(id) : 128);
    NCCL_CHECK(g_nccl.CommInitRank(&h->comm, c->nranks, id, c->rank));
    h->inited = true;
    if (!getenv("ISPH_NO_P2P") && c->nranks <= ISPH_MAX_RANKS) {   // peer mailboxes for the small all-reduces
      const size_t bytes = sizeof(double) * MB_SLOTS * c->nranks * MB_STRIDE;
      CUDA_CHECK(cudaMalloc(&h->mbox, bytes)); CUDA_CHECK(cudaMemset(h->mbox, 0xff, bytes));     // every cell armed (sentinel)
      { const int lim[4] = {0, getenv("ISPH_P2P_TIMEOUT_MS") ? std::max(1, atoi(getenv("ISPH_P2P_TIMEOUT_MS"))) : 20000, 0, 0};   // fault[0] = raised, fault[1] = wait limit in ms
        CUDA_CHECK(cudaMalloc(&h->fault, sizeof(lim))); CUDA_CHECK(cudaMemcpy(h->fault, lim, sizeof(lim), cudaMemcpyHostToDevice)); }
      h->p2p = ipc_share(c, h, h->mbox, h->mpeer, h->mopened);
      if (h->p2p) { P2PTab t; memset(&t, 0, sizeof(t)); for (int p = 0; p < c->nranks; ++p) t.box[p] = h->mpeer[p]; t.mine = h->mbox; t.nranks = c->nranks; t.rank = c->rank; t.fault = h->fault;
        CUDA_CHECK(cudaMalloc(&h->d_tab, sizeof(t))); CUDA_CHECK(cudaMemcpy(h->d_tab, &t, sizeof(t), cudaMemcpyHostToDevice)); }
    }
  }
  return h;
}

P2PRed halo_p2p_ticket(Ctx *c) {
  P2PRed r; r.tab = nullptr; r.seq = 0; r.nranks = 1;
  if (c->nranks <= 1) return r;
  Halo *h = get(c); if (!h->p2p) return r;
  r.tab = h->d_tab; r.seq = h->seq++; r.nranks = c->nranks;
  return r;
}

static void exchange(Ctx *c, Halo *h, const double *sendbuf, double *recv_base, int nc) {
  NCCL_CHECK(g_nccl.GroupStart());
  for (int p = 0; p < c->nranks; ++p) {
    if (h->send_count[p]) NCCL_CHECK(g_nccl.Send(sendbuf + (size_t)h->send_off[p] * nc, (size_t)h->send_count[p] * nc, ncclDouble, p, h->comm, c->stream));
    if (h->recv_count[p]) NCCL_CHECK(g_nccl.Recv(recv_base + (size_t)h->recv_off[p] * nc, (size_t)h->recv_count[p] * nc, ncclDouble, p, h->comm, c->stream));
  }
  NCCL_CHECK(g_nccl.GroupEnd());
}

// After a peer wait timed out (fault word raised) the slot rings of the ranks are out of step.  Collective, called at the start
// of every solve: if ANY rank saw a fault, all ranks drain, re-arm every cell, reset the sequence numbers and clear the fault.
void halo_recover(Ctx *c) {
  if (c->nranks <= 1) return; Halo *h = get(c); if (!h->p2p) return;
  h->itmp.ensure((size_t)c->nranks + 8); h->h_small.ensure(64);
  CUDA_CHECK(cudaMemcpyAsync(h->itmp.p, h->fault, sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
  NCCL_CHECK(g_nccl.AllReduce(h->itmp.p, h->itmp.p, 1, ncclInt, ncclMax, h->comm, c->stream));
  CUDA_CHECK(cudaMemcpyAsync(h->h_small.p, h->itmp.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream));
  if (h->h_small.p[0] == 0) return;
  CUDA_CHECK(cudaMemsetAsync(h->mbox, 0xff, sizeof(double) * MB_SLOTS * c->nranks * MB_STRIDE, c->stream));
  if (h->hbox) CUDA_CHECK(cudaMemsetAsync(h->hbox, 0xff, sizeof(double) * ((size_t)MB_SLOTS * 3 * h->hcap), c->stream));
  CUDA_CHECK(cudaMemsetAsync(h->fault, 0, sizeof(int), c->stream));
  h->seq = 0; h->hseq = 0; c->prepush_x = nullptr;
  NCCL_CHECK(g_nccl.AllReduce(h->itmp.p, h->itmp.p, 1, ncclInt, ncclMax, h->comm, c->stream));      // nobody pushes before everybody has re-armed
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

void halo_setup(Ctx *c) {
  Halo *h = get(c); const int R = c->nranks, nl = c->nlocal, ng = c->nghost;
  c->tic("haloSetup");
  static const bool no_cache = getenv("ISPH_NO_PLAN_CACHE") != nullptr;
  h->itmp.ensure((size_t)R + 8); h->itmp2.ensure((size_t)R * (R + 1) + 8); h->h_small.ensure((size_t)R * (R + 1) + 64);
  // ---- plan cache: a collective decision (one rank's new tag set changes the other ranks' send lists)
  { const int changed = (no_cache || !h->plan_valid || h->plan_hash != c->tag_hash || h->plan_nl != nl || h->plan_ng != ng) ? 1 : 0;
    h->h_small.p[0] = changed;
    CUDA_CHECK(cudaMemcpyAsync(h->itmp2.p, h->h_small.p, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    NCCL_CHECK(g_nccl.AllReduce(h->itmp2.p, h->itmp2.p, 1, ncclInt, ncclMax, h->comm, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(h->h_small.p + 1, h->itmp2.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (h->h_small.p[1] == 0) {                                   // same tags everywhere: only the ghost columns have to be put back
      if (ng) CUDA_CHECK(cudaMemcpyAsync(c->col_of_atom.p + nl, h->gcol_saved.p, sizeof(int) * ng, cudaMemcpyDeviceToDevice, c->stream));
      ++h->reuses; c->toc("haloSetup"); return;
    } }
  h->plan_valid = false; ++h->setups;
  // owned-tag directory: allgather (padded to the largest rank) -> tag -> (owner rank, owner-local index)
  h->h_small.p[0] = nl;
  CUDA_CHECK(cudaMemcpyAsync(h->itmp2.p, h->h_small.p, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  NCCL_CHECK(g_nccl.AllGather(h->itmp2.p, h->itmp.p, 1, ncclInt, h->comm, c->stream));
  CUDA_CHECK(cudaMemcpyAsync(h->h_small.p, h->itmp.p, sizeof(int) * R, cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream));
  const int maxn = *std::max_element(h->h_small.p, h->h_small.p + R);
  h->mytags.ensure(maxn); h->alltags.ensure((size_t)R * maxn);
  CUDA_CHECK(cudaMemsetAsync(h->mytags.p, 0, sizeof(int) * maxn, c->stream));
  CUDA_CHECK(cudaMemcpyAsync(h->mytags.p, c->tag.p, sizeof(int) * nl, cudaMemcpyDeviceToDevice, c->stream));
  NCCL_CHECK(g_nccl.AllGather(h->mytags.p, h->alltags.p, maxn, ncclInt, h->comm, c->stream));
  h->owner_tab.ensure((size_t)c->max_tag + 1);
  CUDA_CHECK(cudaMemsetAsync(h->owner_tab.p, 0xff, sizeof(long long) * ((size_t)c->max_tag + 1), c->stream));
  k_owner_tab<<<ceil_div((long long)R * maxn, 256), 256, 0, c->stream>>>(h->alltags.p, h->itmp.p, maxn, R, c->max_tag, h->owner_tab.p); ++c->launches;
  h->gown.ensure(ng + 1); h->gidx.ensure(ng + 1); h->req.ensure(ng + 1); h->gcol_saved.ensure(ng + 1);
  k_ghost_owner<<<ceil_div(ng, 256), 256, 0, c->stream>>>(c->tag.p, c->col_of_atom.p, h->owner_tab.p, nl, ng, c->rank, h->gown.p, h->gidx.p); ++c->launches;
  // the plan itself: sorted and numbered on the device; only R + 1 integers per rank come back to the host
  int *d_counts = h->itmp.p;                                      // [R] received per peer, [R] = unowned ghosts
  plan_device(c, &h->pw, c->tag.p, h->gown.p, h->gidx.p, nl, ng, c->col_of_atom.p + nl, h->req.p, d_counts);
  if (ng) CUDA_CHECK(cudaMemcpyAsync(h->gcol_saved.p, c->col_of_atom.p + nl, sizeof(int) * ng, cudaMemcpyDeviceToDevice, c->stream));
  // who needs what from me: allgather the (request count, error) rows -> M[r][p] = how many values rank r receives from rank p
  NCCL_CHECK(g_nccl.AllGather(d_counts, h->itmp2.p, R + 1, ncclInt, h->comm, c->stream));
  int *M = h->h_small.p;
  CUDA_CHECK(cudaMemcpyAsync(M, h->itmp2.p, sizeof(int) * R * (R + 1), cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream));
  for (int r = 0; r < R; ++r) ISPH_REQUIRE(M[(size_t)r * (R + 1) + R] == 0, "halo plan: a ghost atom's tag is owned by no rank");
  h->recv_count.assign(R, 0); h->send_count.assign(R, 0); h->send_off.assign(R + 1, 0); h->recv_off.assign(R + 1, 0); h->dst_off.assign(R, 0);
  long long max_halo = 0; bool symmetric = true;
  for (int p = 0; p < R; ++p) {
    h->recv_count[p] = M[(size_t)c->rank * (R + 1) + p];
    h->send_count[p] = M[(size_t)p * (R + 1) + c->rank]; h->send_off[p + 1] = h->send_off[p] + h->send_count[p]; h->recv_off[p + 1] = h->recv_off[p] + h->recv_count[p];
    long long tot = 0; int before_me = 0;                  // where my block starts in rank p's staging buffer
    for (int q = 0; q < R; ++q) { if (q < c->rank) before_me += M[(size_t)p * (R + 1) + q]; tot += M[(size_t)p * (R + 1) + q];
      if ((M[(size_t)p * (R + 1) + q] > 0) != (M[(size_t)q * (R + 1) + p] > 0)) symmetric = false; }
    h->dst_off[p] = before_me; max_halo = std::max(max_halo, tot);
  }
  h->nhalo = h->recv_off[R]; h->nsend = h->send_off[R];
  h->send_idx.ensure(h->nsend + 1);
  NCCL_CHECK(g_nccl.GroupStart());
  for (int p = 0; p < R; ++p) {
    if (h->recv_count[p]) NCCL_CHECK(g_nccl.Send(h->req.p + h->recv_off[p], h->recv_count[p], ncclInt, p, h->comm, c->stream));
    if (h->send_count[p]) NCCL_CHECK(g_nccl.Recv(h->send_idx.p + h->send_off[p], h->send_count[p], ncclInt, p, h->comm, c->stream));
  }
  NCCL_CHECK(g_nccl.GroupEnd());
  h->sendbuf.ensure((size_t)h->nsend * 9 + 8); h->fieldbuf.ensure((size_t)h->nhalo * 9 + 8);
  // halo staging buffers in peer-addressable memory; every rank sees the same max_halo, so growth is collective
  if (h->p2p && max_halo > h->hcap) {
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    ipc_close(h->hopened); if (h->hbox) cudaFree(h->hbox);
    h->hcap = max_halo + max_halo / 4 + 1024;
    const size_t bytes = sizeof(double) * ((size_t)MB_SLOTS * 3 * h->hcap);
    CUDA_CHECK(cudaMalloc(&h->hbox, bytes)); CUDA_CHECK(cudaMemset(h->hbox, 0xff, bytes));     // every cell armed (sentinel)
    h->hseq = 0;
    if (!ipc_share(c, h, h->hbox, h->hpeer, h->hopened)) { cudaFree(h->hbox); h->hbox = nullptr; h->hcap = 0; }
  }
  h->plan_ok = false; h->rows_ok = false;
  // the slot-ring safety of the staging buffers needs every receiver of a peer to also send to it (p2p_device.cuh); the exchange
  // matrix is known to all ranks, so all take the same decision.  R > ISPH_MAX_RANKS never has peer buffers (get()).
  if (h->p2p && h->hbox && symmetric && R <= ISPH_MAX_RANKS) {   // device copy of the plan for the halo kernels
    HaloDev &hd = h->hd_host; memset(&hd, 0, sizeof(hd));
    for (int p = 0; p < R; ++p) { hd.peer[p] = h->hpeer[p]; hd.send_off[p] = h->send_off[p]; hd.dst_off[p] = h->dst_off[p]; hd.recv_cnt[p] = h->recv_count[p]; }
    for (int p = R; p <= ISPH_MAX_RANKS; ++p) hd.send_off[p] = h->send_off[R];
    hd.mine = h->hbox; hd.nranks = R; hd.rank = c->rank; hd.cap = h->hcap; hd.fault = h->fault;
    h->d_plan.ensure(1);
    CUDA_CHECK(cudaMemcpyAsync(h->d_plan.p, &hd, sizeof(hd), cudaMemcpyHostToDevice, c->stream));   // hd_host is only rewritten after the next set-up's first synchronisation
    h->plan_ok = true;
    if (!getenv("ISPH_NO_PREPUSH") && h->nsend > 0 && h->hcap < (1ll << 28)) {
      h->row_sp.ensure(nl + 2); h->row_cur.ensure(nl + 2); h->row_sd.ensure(h->nsend + 1);
      CUDA_CHECK(cudaMemsetAsync(h->row_cur.p, 0, sizeof(int) * (nl + 1), c->stream));
      k_rowsend_count<<<ceil_div(h->nsend, 256), 256, 0, c->stream>>>(h->send_idx.p, h->nsend, h->row_cur.p); ++c->launches;
      size_t tb = 0; cub::DeviceScan::ExclusiveSum(nullptr, tb, h->row_cur.p, h->row_sp.p, nl + 1, c->stream); h->cubtmp.ensure(tb);
      tb = h->cubtmp.cap; cub::DeviceScan::ExclusiveSum(h->cubtmp.p, tb, h->row_cur.p, h->row_sp.p, nl + 1, c->stream); ++c->launches;
      CUDA_CHECK(cudaMemcpyAsync(h->row_cur.p, h->row_sp.p, sizeof(int) * (nl + 1), cudaMemcpyDeviceToDevice, c->stream));
      k_rowsend_fill<<<ceil_div(h->nsend, 256), 256, 0, c->stream>>>(h->d_plan.p, h->send_idx.p, h->nsend, h->row_cur.p, h->row_sd.p); ++c->launches;
      h->rows_ok = true;
    }
  }
  h->plan_valid = true; h->plan_hash = c->tag_hash; h->plan_nl = nl; h->plan_ng = ng;
  c->toc("haloSetup");
}

bool halo_prepush_begin(Ctx *c, const double *x_next, PrePush *pp) {
  pp->plan = nullptr; pp->sp = pp->sd = nullptr; pp->seq = 0;
  if (c->nranks <= 1) return false;
  Halo *h = get(c);
  if (!h->plan_ok || !h->rows_ok) return false;
  pp->plan = h->d_plan.p; pp->sp = h->row_sp.p; pp->sd = h->row_sd.p; pp->seq = h->hseq++;
  c->prepush_x = x_next; c->prepush_seq = pp->seq;
  return true;
}
void halo_wait_unstage(Ctx *c, double *x, unsigned long long seq) {
  Halo *h = get(c);
  if (h->nhalo) { k_halo_unstage<<<ceil_div(h->nhalo, 256), 256, 0, c->stream>>>(h->d_plan.p, seq, x, 0, c->nlocal, h->nhalo, 1); ++c->launches; }
}

int halo_ncols(Ctx *c) { return c->nlocal + (c->halo ? c->halo->nhalo : 0); }
void halo_counts(Ctx *c, int *nhalo, int *nsend, int *npeers) {
  *nhalo = *nsend = *npeers = 0; if (!c->halo) return;
  *nhalo = c->halo->nhalo; *nsend = c->halo->nsend;
  for (size_t p = 0; p < c->halo->send_count.size(); ++p) if (c->halo->send_count[p] > 0) ++*npeers;
}

// Halo import for an SpMV (Epetra_Import): the off-rank entries of x land behind its owned rows.
void halo_exchange(Ctx *c, double *x, int nvec, int ldx) {
  Halo *h = get(c);
  if (h->nhalo == 0 && h->nsend == 0) return;
  if (h->plan_ok && nvec <= 3) {
    const unsigned long long seq = h->hseq++;
    if (h->nsend) { k_halo_push<<<ceil_div(h->nsend, 256), 256, 0, c->stream>>>(h->d_plan.p, seq, x, ldx, nvec, h->send_idx.p, h->nsend); ++c->launches; }
    if (h->nhalo) { k_halo_unstage<<<ceil_div(h->nhalo, 256), 256, 0, c->stream>>>(h->d_plan.p, seq, x, ldx, c->nlocal, h->nhalo, nvec); ++c->launches; }
    return;
  }
  for (int q = 0; q < nvec; ++q) {
    double *xq = x + (size_t)q * ldx;
    if (h->nsend) { k_pack<<<ceil_div(h->nsend, 256), 256, 0, c->stream>>>(xq, h->send_idx.p, h->nsend, 1, h->sendbuf.p + (size_t)q * h->nsend); ++c->launches; }
    exchange(c, h, h->sendbuf.p + (size_t)q * h->nsend, xq + c->nlocal, 1);
  }
}

void halo_forward_field(Ctx *c, int field, int nc) {
  Halo *h = get(c); if (h->nhalo == 0 && h->nsend == 0) return;
  double *f = c->field[field].p;
  if (h->nsend) { k_pack<<<ceil_div(h->nsend, 256), 256, 0, c->stream>>>(f, h->send_idx.p, h->nsend, nc, h->sendbuf.p); ++c->launches; }
  exchange(c, h, h->sendbuf.p, h->fieldbuf.p, nc);
  k_ghost_from_halo<<<ceil_div(c->nghost, 256), 256, 0, c->stream>>>(c->col_of_atom.p, c->nlocal, c->nall, nc, h->fieldbuf.p, f); ++c->launches;
}

void halo_allreduce(Ctx *c, double *buf, int count) {
  Halo *h = get(c);
  if (h->p2p && count <= 64) {
    P2PRed r = halo_p2p_ticket(c);
    k_allreduce_p2p<<<1, 64, 0, c->stream>>>(r, buf, count); ++c->launches;
    return;
  }
  NCCL_CHECK(g_nccl.AllReduce(buf, buf, count, ncclDouble, ncclSum, h->comm, c->stream));
}

// equal-size all-gather of raw bytes (setup of the replicated coarse levels, amg.cu); one rank: a copy
void halo_allgather_bytes(Ctx *c, const void *send, void *recv, size_t bytes) {
  if (c->nranks <= 1) { CUDA_CHECK(cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, c->stream)); return; }
  Halo *h = get(c);
  NCCL_CHECK(g_nccl.AllGather(send, recv, bytes, ncclChar, h->comm, c->stream));
}

// ---- "Overlap Level" 1 (Ifpack_OverlappingRowMatrix / Ifpack_AdditiveSchwarz, precond_ifpack.h:35-43) ------------------------------
// The extended local problem of a rank = its owned rows + the rows of its halo columns (one level of overlap), restricted to that set.
// Owned rows need nothing from outside (every column of an owned row is an owned or a halo column).  The rows of the halo particles live
// on their owners: each owner packs, for every row on its send list, (global tag, value) of the stored entries; lengths, tags and values
// travel with grouped ncclSend/ncclRecv along the halo plan; the receiver maps tags to its own column ids (entries outside the extended
// set are dropped: Ifpack_LocalFilter) and sorts every row by column (cub segmented sort).  Once per preconditioner creation.
__global__ void k_coltag(const int *tag, const int *col_of_atom, int nall, int *coltag, int *tag2col) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x; if (a >= nall) return;
  const int c = col_of_atom[a]; if (c < 0) return;
  coltag[c] = tag[a]; tag2col[tag[a]] = c;                        // all copies of a particle carry the same tag and column
}
__global__ void k_rows_len(const int *row_len, const int *send_idx, int nsend, int *len) { const int k = blockIdx.x * blockDim.x + threadIdx.x; if (k < nsend) len[k] = row_len[send_idx[k]]; }
__global__ void k_rows_pack(const long long *slice_off, const int *col, const double *val, const int *send_idx, const long long *off, const int *len, int nsend, const int *coltag, int *tags, double *vals) {
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31; if (k >= nsend) return;
  const int r = send_idx[k]; const long long base = slice_off[r >> 5] + (r & 31), o = off[k];
  for (int e = lane; e < len[k]; e += 32) { tags[o + e] = coltag[col[base + 32ll * e]]; vals[o + e] = val[base + 32ll * e]; }
}
__global__ void k_rows_map(const int *tags, long long ntot, const int *tag2col, int max_tag, int n_own, const int *slot_ref, int *cols) {
  const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; if (j >= ntot) return;
  const int t = tags[j]; int c = (t >= 0 && t <= max_tag) ? tag2col[t] : -1;
  if (c >= n_own && slot_ref && !slot_ref[c - n_own]) c = -1;      // a halo slot no owned row reaches is outside the extended set
  cols[j] = c < 0 ? 0x7fffffff : c;
}
__global__ void k_add_rows_from(double *z, const int *send_idx, const double *buf, int k0, int k1) { const int k = k0 + blockIdx.x * blockDim.x + threadIdx.x; if (k < k1) z[send_idx[k]] += buf[k]; }

void halo_import_rows(Ctx *c, OverlapRows *out, const int *slot_ref) {
  Halo *h = get(c); Matrix &A = c->A; const int R = c->nranks, nl = c->nlocal, nh = h->nhalo, ns = h->nsend;
  out->nhalo = nh;
  out->coltag.ensure(A.ncols + 1); out->tag2col.ensure((size_t)c->max_tag + 2);
  CUDA_CHECK(cudaMemsetAsync(out->tag2col.p, 0xff, sizeof(int) * ((size_t)c->max_tag + 2), c->stream));
  k_coltag<<<ceil_div(c->nall, 256), 256, 0, c->stream>>>(c->tag.p, c->col_of_atom.p, c->nall, out->coltag.p, out->tag2col.p); ++c->launches;
  out->len_s.ensure(ns + 1); out->len_r.ensure(nh + 1);
  if (ns) { k_rows_len<<<ceil_div(ns, 256), 256, 0, c->stream>>>(A.row_len.p, h->send_idx.p, ns, out->len_s.p); ++c->launches; }
  NCCL_CHECK(g_nccl.GroupStart());
  for (int p = 0; p < R; ++p) {
    if (h->send_count[p]) NCCL_CHECK(g_nccl.Send(out->len_s.p + h->send_off[p], h->send_count[p], ncclInt, p, h->comm, c->stream));
    if (h->recv_count[p]) NCCL_CHECK(g_nccl.Recv(out->len_r.p + h->recv_off[p], h->recv_count[p], ncclInt, p, h->comm, c->stream));
  }
  NCCL_CHECK(g_nccl.GroupEnd());
  std::vector<int> ls(ns), lr(nh);
  if (ns) CUDA_CHECK(cudaMemcpyAsync(ls.data(), out->len_s.p, sizeof(int) * ns, cudaMemcpyDeviceToHost, c->stream));
  if (nh) CUDA_CHECK(cudaMemcpyAsync(lr.data(), out->len_r.p, sizeof(int) * nh, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  std::vector<long long> os(ns + 1, 0), orr(nh + 1, 0);
  for (int k = 0; k < ns; ++k) os[k + 1] = os[k] + ls[k];
  for (int k = 0; k < nh; ++k) orr[k + 1] = orr[k] + lr[k];
  const long long tot_s = os[ns], tot_r = orr[nh];
  out->off_s.ensure(ns + 1); out->off_r.ensure(nh + 1);
  CUDA_CHECK(cudaMemcpyAsync(out->off_s.p, os.data(), sizeof(long long) * (ns + 1), cudaMemcpyHostToDevice, c->stream));
  CUDA_CHECK(cudaMemcpyAsync(out->off_r.p, orr.data(), sizeof(long long) * (nh + 1), cudaMemcpyHostToDevice, c->stream));
  out->tag_s.ensure(tot_s + 1); out->val_s.ensure(tot_s + 1); out->tag_r.ensure(tot_r + 1); out->val_r.ensure(tot_r + 1); out->col_r.ensure(tot_r + 1); out->col_r2.ensure(tot_r + 1); out->val_r2.ensure(tot_r + 1);
  if (ns) { k_rows_pack<<<ceil_div((long long)ns * 32, 256), 256, 0, c->stream>>>(A.slice_off.p, A.col.p, A.val.p, h->send_idx.p, out->off_s.p, out->len_s.p, ns, out->coltag.p, out->tag_s.p, out->val_s.p); ++c->launches; }
  NCCL_CHECK(g_nccl.GroupStart());
  for (int p = 0; p < R; ++p) {
    const long long s0 = os[h->send_off[p]], s1 = os[h->send_off[p + 1]], r0 = orr[h->recv_off[p]], r1 = orr[h->recv_off[p + 1]];
    if (s1 > s0) { NCCL_CHECK(g_nccl.Send(out->tag_s.p + s0, (size_t)(s1 - s0), ncclInt, p, h->comm, c->stream)); NCCL_CHECK(g_nccl.Send(out->val_s.p + s0, (size_t)(s1 - s0), ncclDouble, p, h->comm, c->stream)); }
    if (r1 > r0) { NCCL_CHECK(g_nccl.Recv(out->tag_r.p + r0, (size_t)(r1 - r0), ncclInt, p, h->comm, c->stream)); NCCL_CHECK(g_nccl.Recv(out->val_r.p + r0, (size_t)(r1 - r0), ncclDouble, p, h->comm, c->stream)); }
  }
  NCCL_CHECK(g_nccl.GroupEnd());
  out->total = tot_r;
  if (tot_r > 0) {
    k_rows_map<<<ceil_div(tot_r, 256), 256, 0, c->stream>>>(out->tag_r.p, tot_r, out->tag2col.p, c->max_tag, nl, slot_ref, out->col_r.p); ++c->launches;
    size_t tb = 0;
    cub::DeviceSegmentedSort::SortPairs(nullptr, tb, out->col_r.p, out->col_r2.p, out->val_r.p, out->val_r2.p, tot_r, nh, out->off_r.p, out->off_r.p + 1, c->stream);
    h->cubtmp.ensure(tb); tb = h->cubtmp.cap;
    cub::DeviceSegmentedSort::SortPairs(h->cubtmp.p, tb, out->col_r.p, out->col_r2.p, out->val_r.p, out->val_r2.p, tot_r, nh, out->off_r.p, out->off_r.p + 1, c->stream); ++c->launches;
  }
  (void)nl;
}

// additive Schwarz, combine mode Add: the extended solution at the halo rows goes back to the owners and is added to their rows
// (peer after peer, in rank order: a row requested by several peers is summed in a fixed order)
void halo_export_add(Ctx *c, const double *zext_halo, double *z) {
  Halo *h = get(c); const int R = c->nranks;
  if (h->nhalo == 0 && h->nsend == 0) return;
  h->sendbuf.ensure((size_t)h->nsend * 9 + 8);
  NCCL_CHECK(g_nccl.GroupStart());
  for (int p = 0; p < R; ++p) {
    if (h->recv_count[p]) NCCL_CHECK(g_nccl.Send(zext_halo + h->recv_off[p], h->recv_count[p], ncclDouble, p, h->comm, c->stream));
    if (h->send_count[p]) NCCL_CHECK(g_nccl.Recv(h->sendbuf.p + h->send_off[p], h->send_count[p], ncclDouble, p, h->comm, c->stream));
  }
  NCCL_CHECK(g_nccl.GroupEnd());
  for (int p = 0; p < R; ++p) if (h->send_count[p]) {
    k_add_rows_from<<<ceil_div(h->send_count[p], 256), 256, 0, c->stream>>>(z, h->send_idx.p, h->sendbuf.p, h->send_off[p], h->send_off[p + 1]); ++c->launches;
  }
}

bool halo_fault(Ctx *c) {
  if (!c->halo || !c->halo->fault) return false;
  int f = 0; cudaMemcpy(&f, c->halo->fault, sizeof(int), cudaMemcpyDeviceToHost); return f != 0;
}

void halo_destroy(Ctx *c) {
  if (!c->halo) return; Halo *h = c->halo;
  ipc_close(h->mopened); ipc_close(h->hopened);
  if (h->mbox) cudaFree(h->mbox); if (h->hbox) cudaFree(h->hbox); if (h->fault) cudaFree(h->fault); if (h->d_tab) cudaFree(h->d_tab);
  if (h->inited && h->comm) g_nccl.CommDestroy(h->comm);
  h->d_plan.release(); h->row_sp.release(); h->row_sd.release(); h->row_cur.release(); h->cubtmp.release();
  h->mytags.release(); h->alltags.release(); h->gown.release(); h->gidx.release(); h->req.release(); h->pw.release(); h->gcol_saved.release(); h->h_small.release();
  h->send_idx.release(); h->itmp.release(); h->itmp2.release(); h->owner_tab.release(); h->sendbuf.release(); h->fieldbuf.release();
  delete h; c->halo = nullptr;
}

}  // namespace isph

extern "C" {

int isph_nccl_unique_id(void *id128) {
  if (!id128 || !isph::g_nccl.load()) return ISPH_FAILURE;
  ncclUniqueId id; if (isph::g_nccl.GetUniqueId(&id) != ncclSuccess) return ISPH_FAILURE;
  memset(id128, 0, 128); memcpy(id128, &id, sizeof(id) < 128 ? sizeof(id) : 128); return ISPH_SUCCESS;
}

// The device planner halo_setup runs (sort + number the distinct remote (owner, tag) pairs on the GPU), with the signature of
// isph_halo_plan_host so that a one-GPU test can compare the two on identical inputs.  Host arrays in, host arrays out.
int isph_halo_plan_device(isph_ctx *ctx, int nranks, int rank, int nlocal, int nghost, const int *ghost_tag, const int *ghost_owner, const int *ghost_owner_idx,
                          int *ghost_col, int *recv_count, int *request_idx, int *nhalo_out) {
  if (!ctx) return ISPH_FAILURE; isph::Ctx *c = reinterpret_cast<isph::Ctx *>(ctx);
  try {
    using namespace isph;
    CUDA_CHECK(cudaSetDevice(c->device));
    ISPH_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks && nlocal >= 0 && nghost >= 0, "bad plan arguments");
    const int save_r = c->nranks, save_k = c->rank; c->nranks = nranks; c->rank = rank;
    PlanWork pw; DevBuf<int> tag, own, idx, col, req, cnt;
    tag.ensure((size_t)nlocal + nghost + 1); own.ensure(nghost + 1); idx.ensure(nghost + 1); col.ensure(nghost + 1); req.ensure(nghost + 1); cnt.ensure(nranks + 1);
    CUDA_CHECK(cudaMemcpy(tag.p + nlocal, ghost_tag, sizeof(int) * nghost, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(own.p, ghost_owner, sizeof(int) * nghost, cudaMemcpyHostToDevice)); CUDA_CHECK(cudaMemcpy(idx.p, ghost_owner_idx, sizeof(int) * nghost, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemset(col.p, 0xff, sizeof(int) * (nghost + 1)));
    std::string err;
    try { plan_device(c, &pw, tag.p, own.p, idx.p, nlocal, nghost, col.p, req.p, cnt.p); CUDA_CHECK(cudaStreamSynchronize(c->stream)); } catch (const std::exception &e) { err = e.what(); }
    c->nranks = save_r; c->rank = save_k;
    std::vector<int> hc(nranks + 1, 0);
    if (err.empty()) {
      CUDA_CHECK(cudaMemcpy(hc.data(), cnt.p, sizeof(int) * (nranks + 1), cudaMemcpyDeviceToHost));
      int nh = 0; for (int p = 0; p < nranks; ++p) { recv_count[p] = hc[p]; nh += hc[p]; }
      *nhalo_out = nh;
      CUDA_CHECK(cudaMemcpy(ghost_col, col.p, sizeof(int) * nghost, cudaMemcpyDeviceToHost)); CUDA_CHECK(cudaMemcpy(request_idx, req.p, sizeof(int) * nh, cudaMemcpyDeviceToHost));
    }
    pw.release(); tag.release(); own.release(); idx.release(); col.release(); req.release(); cnt.release();
    if (!err.empty()) throw std::runtime_error(err);
    if (hc[nranks] != 0) { c->err = "halo plan: a ghost atom's tag is owned by no rank"; return ISPH_FAILURE; }
  } catch (const std::exception &e) { c->err = e.what(); cudaGetLastError(); return ISPH_FAILURE; }
  return ISPH_SUCCESS;
}

// Pure host code (no CUDA): halo columns for the ghosts owned by other ranks.  Distinct (owner, tag) pairs are numbered
// in (owner, tag) order, so each peer's block is contiguous and ordered identically on both sides.
int isph_halo_plan_host(int nranks, int rank, int nlocal, int nghost, const int *ghost_tag, const int *ghost_owner, const int *ghost_owner_idx,
                        int *ghost_col, int *recv_count, int *request_idx, int *nhalo_out) {
  std::vector<std::tuple<int, int, int, int>> rem;      // owner, tag, owner idx, ghost
  for (int p = 0; p < nranks; ++p) recv_count[p] = 0;
  for (int g = 0; g < nghost; ++g) {
    if (ghost_owner[g] == rank) { ghost_col[g] = ghost_owner_idx[g]; continue; }
    if (ghost_owner[g] < 0 || ghost_owner[g] >= nranks) return ISPH_FAILURE;
    rem.emplace_back(ghost_owner[g], ghost_tag[g], ghost_owner_idx[g], g);
  }
  std::sort(rem.begin(), rem.end());
  int slot = -1, po = -1, pt = -1;
  for (auto &e : rem) {
    if (std::get<0>(e) != po || std::get<1>(e) != pt) { ++slot; po = std::get<0>(e); pt = std::get<1>(e); request_idx[slot] = std::get<2>(e); ++recv_count[po]; }
    ghost_col[std::get<3>(e)] = nlocal + slot;
  }
  *nhalo_out = slot + 1;
  return ISPH_SUCCESS;
}

}  // extern "C"
