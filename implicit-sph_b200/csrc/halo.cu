// Multi-GPU plumbing: one process per GPU, rows partitioned along LAMMPS subdomains (Epetra_Map of owned tags,
// pair_isph.cpp:1258-1259).  Replaces what Epetra does through MPI underneath the reference's calls:
//   * column map / importer  -> halo plan: every ghost atom whose tag is owned by another rank gets a halo column
//     (distinct remote tags, grouped by owner so that each peer's data lands contiguously behind the owned rows);
//   * Epetra_Import in CrsMatrix::Multiply -> pack kernel + grouped ncclSend/ncclRecv straight into the halo part of x;
//   * Epetra_MpiComm::SumAll -> ncclAllReduce(double, sum) on the device scalars of the Krylov kernels;
//   * comm->forward_comm_pair (owner -> ghost field copy) -> the same plan with ncomp doubles per particle.
// NCCL is loaded with dlopen so that the library also loads on hosts without it (single-GPU use, CPU-side ABI tests).
// The plan itself (isph_halo_plan_host) is pure host code without any CUDA call: tests/test_multirank_cpu.py drives it
// with a gloo transport at world_size 2.
#include "isph_internal.h"
#include <nccl.h>
#include <dlfcn.h>
#include <algorithm>
#include <tuple>

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

struct Halo {
  ncclComm_t comm = nullptr; bool inited = false;
  int nhalo = 0, nsend = 0;
  std::vector<int> recv_count, recv_off, send_count, send_off;
  DevBuf<int> send_idx, itmp, itmp2; DevBuf<long long> owner_tab; DevBuf<double> sendbuf, fieldbuf;
};

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

static Halo *get(Ctx *c) {
  if (!c->halo) c->halo = new Halo();
  Halo *h = c->halo;
  if (!h->inited) {
    ISPH_REQUIRE(g_nccl.load(), "nranks > 1: libnccl.so.2 could not be loaded");
    ISPH_REQUIRE(c->have_nccl_id, "nranks > 1: no NCCL unique id");
    ncclUniqueId id; memcpy(&id, c->nccl_id, sizeof(id) < 128 ? sizeof(id) : 128);
    NCCL_CHECK(g_nccl.CommInitRank(&h->comm, c->nranks, id, c->rank));
    h->inited = true;
  }
  return h;
}

static void exchange(Ctx *c, Halo *h, const double *sendbuf, double *recv_base, int nc) {
  NCCL_CHECK(g_nccl.GroupStart());
  for (int p = 0; p < c->nranks; ++p) {
    if (h->send_count[p]) NCCL_CHECK(g_nccl.Send(sendbuf + (size_t)h->send_off[p] * nc, (size_t)h->send_count[p] * nc, ncclDouble, p, h->comm, c->stream));
    if (h->recv_count[p]) NCCL_CHECK(g_nccl.Recv(recv_base + (size_t)h->recv_off[p] * nc, (size_t)h->recv_count[p] * nc, ncclDouble, p, h->comm, c->stream));
  }
  NCCL_CHECK(g_nccl.GroupEnd());
}

void halo_setup(Ctx *c) {
  Halo *h = get(c); const int R = c->nranks, nl = c->nlocal, ng = c->nghost;
  c->tic("haloSetup");
  // owned-tag directory: allgather (padded to the largest rank) -> tag -> (owner rank, owner-local index)
  h->itmp.ensure((size_t)R + 8); h->itmp2.ensure((size_t)R * R + 8);
  CUDA_CHECK(cudaMemcpyAsync(h->itmp2.p, &nl, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  NCCL_CHECK(g_nccl.AllGather(h->itmp2.p, h->itmp.p, 1, ncclInt, h->comm, c->stream));
  std::vector<int> nloc_all(R);
  CUDA_CHECK(cudaMemcpyAsync(nloc_all.data(), h->itmp.p, sizeof(int) * R, cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream));
  const int maxn = *std::max_element(nloc_all.begin(), nloc_all.end());
  DevBuf<int> mytags, alltags; mytags.ensure(maxn); alltags.ensure((size_t)R * maxn);
  CUDA_CHECK(cudaMemsetAsync(mytags.p, 0, sizeof(int) * maxn, c->stream));
  CUDA_CHECK(cudaMemcpyAsync(mytags.p, c->tag.p, sizeof(int) * nl, cudaMemcpyDeviceToDevice, c->stream));
  NCCL_CHECK(g_nccl.AllGather(mytags.p, alltags.p, maxn, ncclInt, h->comm, c->stream));
  h->owner_tab.ensure((size_t)c->max_tag + 1);
  CUDA_CHECK(cudaMemsetAsync(h->owner_tab.p, 0xff, sizeof(long long) * ((size_t)c->max_tag + 1), c->stream));
  k_owner_tab<<<ceil_div((long long)R * maxn, 256), 256, 0, c->stream>>>(alltags.p, h->itmp.p, maxn, R, c->max_tag, h->owner_tab.p); ++c->launches;
  DevBuf<int> gown, gidx; gown.ensure(ng + 1); gidx.ensure(ng + 1);
  k_ghost_owner<<<ceil_div(ng, 256), 256, 0, c->stream>>>(c->tag.p, c->col_of_atom.p, h->owner_tab.p, nl, ng, c->rank, gown.p, gidx.p); ++c->launches;
  std::vector<int> owner(ng), idx(ng), gcol(ng), request(ng);
  CUDA_CHECK(cudaMemcpyAsync(owner.data(), gown.p, sizeof(int) * ng, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaMemcpyAsync(idx.data(), gidx.p, sizeof(int) * ng, cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream));
  mytags.release(); alltags.release(); gown.release(); gidx.release();
  h->recv_count.assign(R, 0);
  const int rc = isph_halo_plan_host(R, c->rank, nl, ng, c->h_tag.data() + nl, owner.data(), idx.data(), gcol.data(), h->recv_count.data(), request.data(), &h->nhalo);
  ISPH_REQUIRE(rc == ISPH_SUCCESS, "halo plan: a ghost atom's tag is owned by no rank");
  CUDA_CHECK(cudaMemcpyAsync(c->col_of_atom.p + nl, gcol.data(), sizeof(int) * ng, cudaMemcpyHostToDevice, c->stream));
  // who needs what from me: allgather the request-count matrix, then swap the index lists
  CUDA_CHECK(cudaMemcpyAsync(h->itmp.p, h->recv_count.data(), sizeof(int) * R, cudaMemcpyHostToDevice, c->stream));
  NCCL_CHECK(g_nccl.AllGather(h->itmp.p, h->itmp2.p, R, ncclInt, h->comm, c->stream));
  std::vector<int> M((size_t)R * R);
  CUDA_CHECK(cudaMemcpyAsync(M.data(), h->itmp2.p, sizeof(int) * R * R, cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream));
  h->send_count.assign(R, 0); h->send_off.assign(R + 1, 0); h->recv_off.assign(R + 1, 0);
  for (int p = 0; p < R; ++p) { h->send_count[p] = M[(size_t)p * R + c->rank]; h->send_off[p + 1] = h->send_off[p] + h->send_count[p]; h->recv_off[p + 1] = h->recv_off[p] + h->recv_count[p]; }
  h->nsend = h->send_off[R];
  DevBuf<int> req; req.ensure(h->nhalo + 1); h->send_idx.ensure(h->nsend + 1);
  CUDA_CHECK(cudaMemcpyAsync(req.p, request.data(), sizeof(int) * h->nhalo, cudaMemcpyHostToDevice, c->stream));
  NCCL_CHECK(g_nccl.GroupStart());
  for (int p = 0; p < R; ++p) {
    if (h->recv_count[p]) NCCL_CHECK(g_nccl.Send(req.p + h->recv_off[p], h->recv_count[p], ncclInt, p, h->comm, c->stream));
    if (h->send_count[p]) NCCL_CHECK(g_nccl.Recv(h->send_idx.p + h->send_off[p], h->send_count[p], ncclInt, p, h->comm, c->stream));
  }
  NCCL_CHECK(g_nccl.GroupEnd());
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  req.release();
  h->sendbuf.ensure((size_t)h->nsend * 9 + 8); h->fieldbuf.ensure((size_t)h->nhalo * 9 + 8);
  c->toc("haloSetup");
}

int halo_ncols(Ctx *c) { return c->nlocal + (c->halo ? c->halo->nhalo : 0); }

// x[nlocal .. nlocal+nhalo) <- owners' values (per vector)
void halo_exchange(Ctx *c, double *x, int nvec, int ldx) {
  Halo *h = get(c); if (h->nhalo == 0 && h->nsend == 0) return;
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
  NCCL_CHECK(g_nccl.AllReduce(buf, buf, count, ncclDouble, ncclSum, h->comm, c->stream));
}

void halo_destroy(Ctx *c) {
  if (!c->halo) return; Halo *h = c->halo;
  if (h->inited && h->comm) g_nccl.CommDestroy(h->comm);
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
