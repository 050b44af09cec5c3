// Graph build: LAMMPS full neighbor list -> sparsity pattern, directly in the SELL-32 layout the SpMV and the
// assembly kernels use.  Replaces FunctorOuterGraph (functor_graph.h:38-99: row i = {tag[j] : rsq < cutsq} U {tag[i]},
// FillComplete sorts and merges duplicates) + Epetra_CrsGraph/CrsMatrix construction (pair_isph.cpp:1258-1270).
//
// One CTA per slice of 32 rows; each of its 4 warps takes 8 rows: the warp evaluates the reference's distance test
// (same operation order, no FMA contraction: this TU is compiled with -fmad=false), sorts the surviving
// (column, atom) keys with a shared-memory bitonic network, and the CTA then writes the slice column-major so that
// every later pass over the matrix (assembly, SpMV, scaling) is perfectly coalesced with one thread per row.
#include "isph_internal.h"
#include <algorithm>
#include <numeric>

namespace isph {

// ---------------------------------------------------------------------------------------------------------------
__global__ void k_fill_int(int *p, int v, long long n) { long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; if (i < n) p[i] = v; }
__global__ void k_tag2own(const int *tag, int nlocal, int *tag2own) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < nlocal) tag2own[tag[i]] = i; }
__global__ void k_col_of_atom(const int *tag, const int *tag2own, int nlocal, int nall, int *col_of_atom, int *missing) {
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= nall) return;
  if (i < nlocal) { col_of_atom[i] = i; return; }
  int o = tag2own[tag[i]];
  col_of_atom[i] = o;                       // -1: owned by another rank -> halo slot assigned by halo_setup
  if (o < 0) atomicAdd(missing, 1);
}
__global__ void k_kind(const int *type, int nall, const PairTab *T, int *kind) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < nall) kind[i] = T->kind[type[i]]; }

void build_column_map(Ctx *c) {
  const int B = 256;
  c->tag2own.ensure((size_t)c->max_tag + 1); c->col_of_atom.ensure(c->nall); c->kind.ensure(c->nall); c->flag.ensure(16);
  k_fill_int<<<ceil_div(c->max_tag + 1, B), B, 0, c->stream>>>(c->tag2own.p, -1, c->max_tag + 1);
  k_tag2own<<<ceil_div(c->nlocal, B), B, 0, c->stream>>>(c->tag.p, c->nlocal, c->tag2own.p);
  CUDA_CHECK(cudaMemsetAsync(c->flag.p, 0, 16 * sizeof(int), c->stream));
  k_col_of_atom<<<ceil_div(c->nall, B), B, 0, c->stream>>>(c->tag.p, c->tag2own.p, c->nlocal, c->nall, c->col_of_atom.p, c->flag.p);
  k_kind<<<ceil_div(c->nall, B), B, 0, c->stream>>>(c->type.p, c->nall, c->d_tab.p, c->kind.p);
  c->launches += 4;
  int missing = 0;
  CUDA_CHECK(cudaMemcpyAsync(&missing, c->flag.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  if (c->nranks > 1) halo_setup(c);          // assigns halo columns to ghosts owned by other ranks
  else ISPH_REQUIRE(missing == 0, "isph_atoms_set: a ghost atom carries a tag that no owned atom has (single-rank run)");
}

// ---------------------------------------------------------------------------------------------------------------
struct GraphArgs {
  int n, dim; const int *ilist, *neigh, *type, *col_of_atom; const long long *noff; const double *x; const PairTab *T;
  const long long *slice_off; int *slice_len, *row_len, *diag_k, *col, *atom; int *counters;   // counters[0]=ndup, [1]=nnz(lo), ...
  unsigned long long *nnz;
};

template <int CAP> __global__ void __launch_bounds__(128) k_graph_slice(GraphArgs a) {
  extern __shared__ unsigned long long skey[];          // [32][CAP+1]
  __shared__ int s_len[32], s_selfcol[32], s_dup[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, slice = blockIdx.x;
  const unsigned long long INVALID = ~0ull;
  constexpr int E = CAP / 32;
  for (int rr = 0; rr < 8; ++rr) {
    unsigned long long keys[E];
    const int r = warp * 8 + rr, row = slice * 32 + r;
    unsigned long long *kb = skey + (size_t)r * (CAP + 1);
    int selfcol = row < a.n ? row : 0;
    if (row < a.n) {
      const int i = a.ilist[row], itype = a.type[i];
      const double xi0 = a.x[3 * (size_t)i], xi1 = a.x[3 * (size_t)i + 1], xi2 = a.x[3 * (size_t)i + 2];
      const long long beg = a.noff[row]; const int cnt = (int)(a.noff[row + 1] - beg);
      selfcol = a.col_of_atom[i];
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int t = e * 32 + lane;
        unsigned long long key = INVALID;
        if (t < cnt) {
          const int j = a.neigh[beg + t] & ISPH_NEIGHMASK, jtype = a.type[j];
          // functor_graph.h:76-84: rsq accumulated component by component, `rsq < cutsq[itype][jtype]`
          double d = xi0 - a.x[3 * (size_t)j]; double rsq = 0.0 + d * d;
          d = xi1 - a.x[3 * (size_t)j + 1]; rsq += d * d;
          if (a.dim == 3) { d = xi2 - a.x[3 * (size_t)j + 2]; rsq += d * d; }
          if (rsq < a.T->cutsq[itype][jtype]) key = ((unsigned long long)(unsigned)a.col_of_atom[j] << 32) | (unsigned)j;
        } else if (t == cnt) {
          key = ((unsigned long long)(unsigned)selfcol << 32) | (unsigned)i;      // self connectivity, functor_graph.h:87
        }
        keys[e] = key;
      }
    } else {
#pragma unroll
      for (int e = 0; e < E; ++e) keys[e] = INVALID;
    }
    // bitonic sort over the index space idx = e*32 + lane, entirely in registers: partners at distance < 32 are reached
    // with shuffles, larger distances are register-to-register in the same lane (the shared-memory version of this
    // network kept the L1/shared pipe at 92 %: profiles/r01_asm_ncu_full.txt)
#pragma unroll
    for (int k = 2; k <= CAP; k <<= 1) {
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) {
        if (j >= 32) {
          const int je = j >> 5;
#pragma unroll
          for (int e = 0; e < E; ++e) {
            if ((e & je) == 0) {
              const int idx = e * 32 + lane; const bool asc = ((idx & k) == 0);
              const unsigned long long u = keys[e], v = keys[e | je];
              if ((u > v) == asc) { keys[e] = v; keys[e | je] = u; }
            }
          }
        } else {
#pragma unroll
          for (int e = 0; e < E; ++e) {
            const int idx = e * 32 + lane; const bool asc = ((idx & k) == 0), lower = ((lane & j) == 0);
            const unsigned long long u = keys[e], v = __shfl_xor_sync(0xffffffffu, u, j);
            const bool take_min = (lower == asc);
            keys[e] = take_min ? (u < v ? u : v) : (u > v ? u : v);
          }
        }
      }
    }
#pragma unroll
    for (int e = 0; e < E; ++e) kb[e * 32 + lane] = keys[e];
    __syncwarp();
    int valid = 0, dup = 0;
    for (int t = lane; t < CAP; t += 32) {
      const unsigned long long u = kb[t];
      if (u != INVALID) { ++valid; if (t > 0 && (kb[t - 1] >> 32) == (u >> 32)) ++dup; }
    }
    for (int o = 16; o > 0; o >>= 1) { valid += __shfl_xor_sync(0xffffffffu, valid, o); dup += __shfl_xor_sync(0xffffffffu, dup, o); }
    if (lane == 0) {
      s_len[r] = valid; s_selfcol[r] = selfcol; s_dup[r] = dup;
      if (row < a.n) a.row_len[row] = valid;
    }
  }
  __syncthreads();
  int slen = 0, tot = 0, totdup = 0;
  for (int r = 0; r < 32; ++r) { slen = max(slen, s_len[r]); tot += s_len[r]; totdup += s_dup[r]; }
  if (threadIdx.x == 0) { a.slice_len[slice] = slen; if (totdup) atomicAdd(&a.counters[0], totdup); atomicAdd(a.nnz, (unsigned long long)(tot - totdup)); }
  const long long base = a.slice_off[slice];
  const int r = lane, row = slice * 32 + r;
  const unsigned long long *kb = skey + (size_t)r * (CAP + 1);
  const int mylen = s_len[r], selfcol = s_selfcol[r];
  for (int k = warp; k < slen; k += 4) {
    int cc = selfcol, at = -1;
    if (k < mylen) {
      const unsigned long long u = kb[k]; cc = (int)(u >> 32); at = (int)(u & 0xffffffffu);
      if (cc == selfcol && (k == 0 || (int)(kb[k - 1] >> 32) != selfcol)) a.diag_k[row] = k;
    }
    a.col[base + 32ll * k + r] = cc; a.atom[base + 32ll * k + r] = at;
  }
}

__global__ void k_zero_padding(const long long *slice_off, const int *slice_len, const int *row_len, int n, int nslices, double *val) {
  // every in-slice entry is zero-initialised (new Epetra_CrsMatrix(Copy, graph), pair_isph.cpp:1266); padding stays 0 forever
  const int row = blockIdx.x * blockDim.x + threadIdx.x, s = row >> 5, lane = row & 31; if (s >= nslices) return;
  const long long base = slice_off[s] + lane; const int slen = slice_len[s];
  for (int k = 0; k < slen; ++k) val[base + 32ll * k] = 0.0;
}

void graph_build(Ctx *c) {
  ISPH_REQUIRE(c->have_pair && c->have_atoms && c->have_neigh, "isph_graph_build: pair_coeff, atoms and neighbors must be set first");
  ISPH_REQUIRE(c->inum == c->nlocal, "isph_graph_build: the row map is the set of owned atoms (inum must equal nlocal, pair_isph.cpp:1258)");
  Matrix &A = c->A; const int n = c->nlocal;
  A.n = n; A.nslices = (n + 31) / 32; A.external = false;
  // slice capacities from the neighbor-list upper bound jnum+1 (the reference's static profile, functor_graph.h:46-52)
  A.slice_off.ensure(A.nslices + 1);
  if (c->neigh_on_device) {                                       // list built on the device: offsets never visit the host
    A.h_slice_off.clear();
    A.total = slice_offsets_device(c, n, A.nslices, A.slice_off.p);
  } else {
    A.h_slice_off.assign(A.nslices + 1, 0);
    for (int s = 0; s < A.nslices; ++s) {
      long long cap = 0; const int r1 = std::min(n, (s + 1) * 32);
      for (int r = s * 32; r < r1; ++r) cap = std::max(cap, c->h_noff[r + 1] - c->h_noff[r] + 1);
      A.h_slice_off[s + 1] = A.h_slice_off[s] + 32 * cap;
    }
    A.total = A.h_slice_off[A.nslices];
    CUDA_CHECK(cudaMemcpyAsync(A.slice_off.p, A.h_slice_off.data(), sizeof(long long) * (A.nslices + 1), cudaMemcpyHostToDevice, c->stream));
  }
  A.slice_len.ensure(A.nslices); A.row_len.ensure(n); A.diag_k.ensure(n);
  A.col.ensure(A.total); A.atom.ensure(A.total); A.val.ensure(A.total); A.diagonal.ensure(n); A.sld.ensure(n);
  c->flag.ensure(16);
  CUDA_CHECK(cudaMemsetAsync(c->flag.p, 0, 16 * sizeof(int), c->stream));
  CUDA_CHECK(cudaMemsetAsync(A.diag_k.p, 0xff, sizeof(int) * n, c->stream));
  GraphArgs a{n, c->tab.dim, c->ilist.p, c->neigh.p, c->type.p, c->col_of_atom.p, c->noff.p, c->x.p, c->d_tab.p,
              A.slice_off.p, A.slice_len.p, A.row_len.p, A.diag_k.p, A.col.p, A.atom.p, c->flag.p, (unsigned long long *)(c->flag.p + 2)};
  const int need = c->max_jnum + 1;
  auto launch = [&](auto capc) {
    constexpr int CAP = decltype(capc)::value;
    const size_t sm = (size_t)32 * (CAP + 1) * sizeof(unsigned long long);
    CUDA_CHECK(cudaFuncSetAttribute(k_graph_slice<CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_graph_slice<CAP><<<A.nslices, 128, sm, c->stream>>>(a);
  };
  if (need <= 32) launch(std::integral_constant<int, 32>());
  else if (need <= 64) launch(std::integral_constant<int, 64>());
  else if (need <= 128) launch(std::integral_constant<int, 128>());
  else if (need <= 256) launch(std::integral_constant<int, 256>());
  else if (need <= 512) launch(std::integral_constant<int, 512>());
  else ISPH_REQUIRE(false, "isph_graph_build: more than 511 neighbors in one row is not supported");
  CUDA_CHECK(cudaGetLastError());
  k_zero_padding<<<ceil_div(A.nslices * 32, 128), 128, 0, c->stream>>>(A.slice_off.p, A.slice_len.p, A.row_len.p, n, A.nslices, A.val.p);
  CUDA_CHECK(cudaMemsetAsync(A.diagonal.p, 0, sizeof(double) * n, c->stream));     // new Epetra_Vector(map) is zero-filled, pair_isph.cpp:1268-1269
  CUDA_CHECK(cudaMemsetAsync(A.sld.p, 0, sizeof(double) * n, c->stream));
  c->launches += 2;
  int h[4];
  CUDA_CHECK(cudaMemcpyAsync(h, c->flag.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  A.ndup = h[0]; unsigned long long nnz; memcpy(&nnz, &h[2], 8); A.nnz = (long long)nnz;
  A.ncols = c->nranks > 1 ? halo_ncols(c) : n;
  A.max_row = -1; A.is_filled = 0; A.built = true;
  solver_prepare_vectors(c);
}

// ---------------------------------------------------------------------------------------------------------------
// canonical export (tests / inspection, off the hot path): rows in nodal-map order, columns = ascending global tag,
// duplicate columns merged by summation (what Epetra's FillComplete + SumIntoGlobalValues leave behind)
void graph_export(Ctx *c, int *rowptr, int *col_tags, double *val) {
  Matrix &A = c->A; ISPH_REQUIRE(A.built, "matrix not built");
  std::vector<int> col(A.total), rlen(A.n); std::vector<double> v; if (val) v.resize(A.total);
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  if (A.h_slice_off.empty()) { A.h_slice_off.resize(A.nslices + 1); CUDA_CHECK(cudaMemcpy(A.h_slice_off.data(), A.slice_off.p, sizeof(long long) * (A.nslices + 1), cudaMemcpyDeviceToHost)); }
  CUDA_CHECK(cudaMemcpy(col.data(), A.col.p, sizeof(int) * A.total, cudaMemcpyDeviceToHost));
  CUDA_CHECK(cudaMemcpy(rlen.data(), A.row_len.p, sizeof(int) * A.n, cudaMemcpyDeviceToHost));
  if (val) CUDA_CHECK(cudaMemcpy(v.data(), A.val.p, sizeof(double) * A.total, cudaMemcpyDeviceToHost));
  std::vector<int> coltag;   // local column id -> global tag
  if (A.external) { coltag.resize(A.ncols); std::iota(coltag.begin(), coltag.end(), 0); }
  else {
    coltag.assign(A.ncols, 0);
    for (int i = 0; i < c->nlocal; ++i) coltag[i] = c->h_tag[i];
    if (c->nranks > 1) { std::vector<int> ca(c->nall); CUDA_CHECK(cudaMemcpy(ca.data(), c->col_of_atom.p, sizeof(int) * c->nall, cudaMemcpyDeviceToHost));
      for (int a = c->nlocal; a < c->nall; ++a) if (ca[a] >= c->nlocal) coltag[ca[a]] = c->h_tag[a]; }
  }
  std::vector<std::pair<int, double>> row; long long out = 0; int mx = 0;
  if (rowptr) rowptr[0] = 0;
  for (int r = 0; r < A.n; ++r) {
    const long long base = A.h_slice_off[r >> 5] + (r & 31);
    row.clear();
    for (int k = 0; k < rlen[r]; ++k) row.emplace_back(coltag[col[base + 32ll * k]], val ? v[base + 32ll * k] : 0.0);
    std::stable_sort(row.begin(), row.end(), [](const std::pair<int, double> &p, const std::pair<int, double> &q) { return p.first < q.first; });
    int cnt = 0;
    for (size_t k = 0; k < row.size(); ++k) {
      if (k > 0 && row[k].first == row[k - 1].first) { if (val) val[out - 1] += row[k].second; continue; }
      if (col_tags) col_tags[out] = row[k].first; if (val) val[out] = row[k].second; ++out; ++cnt;
    }
    if (rowptr) rowptr[r + 1] = (int)out; mx = std::max(mx, cnt);
  }
  A.max_row = mx;
}

// external matrix -> SELL (second API client / Krylov tests): host conversion, off the hot path
void matrix_from_csr(Ctx *c, int n, const int *rowptr, const int *col, const double *val) {
  Matrix &A = c->A; A.n = n; A.ncols = n; A.nslices = (n + 31) / 32; A.external = true; A.ndup = 0; A.nnz = rowptr[n];
  A.h_slice_off.assign(A.nslices + 1, 0); std::vector<int> slen(A.nslices, 0), rlen(n), dk(n, -1);
  for (int r = 0; r < n; ++r) { rlen[r] = rowptr[r + 1] - rowptr[r]; slen[r >> 5] = std::max(slen[r >> 5], rlen[r]); }
  for (int s = 0; s < A.nslices; ++s) A.h_slice_off[s + 1] = A.h_slice_off[s] + 32ll * slen[s];
  A.total = A.h_slice_off[A.nslices];
  std::vector<int> hc(A.total, 0), ha(A.total, -1); std::vector<double> hv(A.total, 0.0);
  std::vector<std::pair<int, double>> row;
  for (int r = 0; r < n; ++r) {
    const long long base = A.h_slice_off[r >> 5] + (r & 31);
    row.clear(); for (int p = rowptr[r]; p < rowptr[r + 1]; ++p) row.emplace_back(col[p], val[p]);
    std::sort(row.begin(), row.end(), [](const std::pair<int, double> &p, const std::pair<int, double> &q) { return p.first < q.first; });
    for (int k = 0; k < slen[r >> 5]; ++k) {
      if (k < rlen[r]) { hc[base + 32ll * k] = row[k].first; hv[base + 32ll * k] = row[k].second; if (row[k].first == r && dk[r] < 0) dk[r] = k; }
      else hc[base + 32ll * k] = r;
    }
  }
  A.slice_off.ensure(A.nslices + 1); A.slice_len.ensure(A.nslices); A.row_len.ensure(n); A.diag_k.ensure(n);
  A.col.ensure(A.total); A.atom.ensure(A.total); A.val.ensure(A.total); A.diagonal.ensure(n); A.sld.ensure(n);
  CUDA_CHECK(cudaMemcpy(A.slice_off.p, A.h_slice_off.data(), sizeof(long long) * (A.nslices + 1), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(A.slice_len.p, slen.data(), sizeof(int) * A.nslices, cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(A.row_len.p, rlen.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(A.diag_k.p, dk.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(A.col.p, hc.data(), sizeof(int) * A.total, cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(A.atom.p, ha.data(), sizeof(int) * A.total, cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(A.val.p, hv.data(), sizeof(double) * A.total, cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemset(A.diagonal.p, 0, sizeof(double) * n)); CUDA_CHECK(cudaMemset(A.sld.p, 0, sizeof(double) * n));
  c->nlocal = n; A.is_filled = 1; A.built = true; A.max_row = *std::max_element(rlen.begin(), rlen.end());
  solver_prepare_vectors(c);
}

// ---------------------------------------------------------------------------------------------------------------
// Epetra_CrsMatrix value operations on the SELL layout (one thread per row, coalesced across the slice)
#define SELL_ROW_PROLOGUE \
  const int row = blockIdx.x * blockDim.x + threadIdx.x; if (row >= n) return; \
  const long long base = slice_off[row >> 5] + (row & 31); const int rlen = row_len[row];

__global__ void k_put_scalar(const long long *slice_off, const int *row_len, int n, double *val, double a) {
  SELL_ROW_PROLOGUE for (int k = 0; k < rlen; ++k) val[base + 32ll * k] = a;
}
__global__ void k_scale(const long long *slice_off, const int *row_len, int n, double *val, double a, const double *s, int recip) {
  SELL_ROW_PROLOGUE const double f = s ? (recip ? 1.0 / s[row] : s[row]) : a;
  for (int k = 0; k < rlen; ++k) val[base + 32ll * k] *= f;
}
__global__ void k_extract_diag(const long long *slice_off, const int *row_len, const int *diag_k, int n, const double *val, double *d) {
  SELL_ROW_PROLOGUE (void)rlen; const int k = diag_k[row]; d[row] = k < 0 ? 0.0 : val[base + 32ll * k];
}
__global__ void k_replace_diag(const long long *slice_off, const int *row_len, const int *diag_k, int n, double *val, const double *d) {
  SELL_ROW_PROLOGUE (void)rlen; const int k = diag_k[row]; if (k >= 0) val[base + 32ll * k] = d[row];
}
// after a SumInto-style assembly: fold duplicate columns (two periodic images of one particle in a row) into the first
__global__ void k_merge_dup(const long long *slice_off, const int *row_len, int n, const int *col, double *val) {
  SELL_ROW_PROLOGUE
  int first = 0;
  for (int k = 1; k < rlen; ++k) {
    if (col[base + 32ll * k] == col[base + 32ll * first]) { val[base + 32ll * first] += val[base + 32ll * k]; val[base + 32ll * k] = 0.0; }
    else first = k;
  }
}

void matrix_put_scalar(Ctx *c, double a) { Matrix &A = c->A; k_put_scalar<<<ceil_div(A.n, 128), 128, 0, c->stream>>>(A.slice_off.p, A.row_len.p, A.n, A.val.p, a); ++c->launches; }
void matrix_scale(Ctx *c, double a) { Matrix &A = c->A; k_scale<<<ceil_div(A.n, 128), 128, 0, c->stream>>>(A.slice_off.p, A.row_len.p, A.n, A.val.p, a, nullptr, 0); ++c->launches; }
void matrix_left_scale_dev(Ctx *c, const double *s, bool recip) { Matrix &A = c->A; k_scale<<<ceil_div(A.n, 128), 128, 0, c->stream>>>(A.slice_off.p, A.row_len.p, A.n, A.val.p, 1.0, s, recip ? 1 : 0); ++c->launches; }
void matrix_extract_diag_dev(Ctx *c, double *d) { Matrix &A = c->A; k_extract_diag<<<ceil_div(A.n, 128), 128, 0, c->stream>>>(A.slice_off.p, A.row_len.p, A.diag_k.p, A.n, A.val.p, d); ++c->launches; }
void matrix_replace_diag_dev(Ctx *c, const double *d) { Matrix &A = c->A; k_replace_diag<<<ceil_div(A.n, 128), 128, 0, c->stream>>>(A.slice_off.p, A.row_len.p, A.diag_k.p, A.n, A.val.p, d); ++c->launches; }
void matrix_merge_duplicates(Ctx *c) { Matrix &A = c->A; if (A.ndup == 0) return; k_merge_dup<<<ceil_div(A.n, 128), 128, 0, c->stream>>>(A.slice_off.p, A.row_len.p, A.n, A.col.p, A.val.p); ++c->launches; }

}  // namespace isph
