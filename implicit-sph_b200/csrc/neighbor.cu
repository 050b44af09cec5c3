// Device-side construction of the full neighbor list (first half): what LAMMPS' binned neighbor build hands to
// PairISPH (full list requested in init_style, pair_isph.cpp:1887-1894; rebuilt every step, skin 0, sph-script/taylor-green-vortex-2d.lmp:
// 33,38-39) — for every OWNED atom i all atoms j != i (owned or ghost) with |x_i - x_j|^2 <= cutneigh^2.  The list is a SUPERSET of
// what the functors use: their own `rsq < cutsq[itype][jtype]` test (functor_graph.h:76-84) stays the sole decider, so the graph
// is bit-exact whatever the list order is; the order only fixes the summation order of the pre-computation functors, exactly as
// LAMMPS' own (bin-dependent) order does for the reference.
// Why on the device: the list is the bulk of what crosses PCIe every step (4.56 GB of the 4.56 GB for the 8M-particle problem,
// ~90 ms of the end-to-end step); positions, types and tags are 0.2 GB.
// Algorithm: uniform cells of edge cutneigh over the bounding box of all atoms; atoms sorted by cell (stable radix sort: inside a
// cell the atom order is the caller's); count pass -> exclusive scan -> fill pass over the 3^dim cell stencil.  List order of a row:
// stencil cells in (dz, dy, dx) ascending order, atoms of a cell in ascending atom index.
#include "isph_internal.h"
#include <cub/cub.cuh>

namespace isph {

struct CellGrid { double lo[3], inv; int nc[3]; int dim; };

__global__ void __launch_bounds__(256) k_bbox(const double *x, int nall, double *partials) {       // per-block min / max of the 3 coordinates
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (int i = blockIdx.x * 256 + threadIdx.x; i < nall; i += gridDim.x * 256)
    for (int k = 0; k < 3; ++k) { const double v = x[3 * (size_t)i + k]; mn[k] = fmin(mn[k], v); mx[k] = fmax(mx[k], v); }
  __shared__ double sm[6][8];
  for (int k = 0; k < 3; ++k) for (int o = 16; o > 0; o >>= 1) { mn[k] = fmin(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o)); mx[k] = fmax(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o)); }
  if ((threadIdx.x & 31) == 0) for (int k = 0; k < 3; ++k) { sm[k][threadIdx.x >> 5] = mn[k]; sm[3 + k][threadIdx.x >> 5] = mx[k]; }
  __syncthreads();
  if (threadIdx.x < 6) { double v = sm[threadIdx.x][0]; for (int w = 1; w < 8; ++w) v = threadIdx.x < 3 ? fmin(v, sm[threadIdx.x][w]) : fmax(v, sm[threadIdx.x][w]); partials[6 * blockIdx.x + threadIdx.x] = v; }
}
__device__ __forceinline__ int cell_coord(double v, double lo, double inv, int nc) { int c = (int)floor((v - lo) * inv); return c < 0 ? 0 : (c >= nc ? nc - 1 : c); }
__global__ void k_cell_of_atom(const double *x, int nall, CellGrid g, int *cell, int *atom) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= nall) return;
  const int cx = cell_coord(x[3 * (size_t)i], g.lo[0], g.inv, g.nc[0]), cy = cell_coord(x[3 * (size_t)i + 1], g.lo[1], g.inv, g.nc[1]);
  const int cz = g.dim == 3 ? cell_coord(x[3 * (size_t)i + 2], g.lo[2], g.inv, g.nc[2]) : 0;
  cell[i] = (cz * g.nc[1] + cy) * g.nc[0] + cx; atom[i] = i;
}
__global__ void k_cell_bounds(const int *cell_sorted, int nall, int *cell_start, int *cell_end) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x; if (k >= nall) return;
  const int c = cell_sorted[k];
  if (k == 0 || cell_sorted[k - 1] != c) cell_start[c] = k;
  if (k == nall - 1 || cell_sorted[k + 1] != c) cell_end[c] = k + 1;
}
// one thread per owned atom; FILL = false: count, FILL = true: write the row
template <int DIM, bool FILL> __global__ void __launch_bounds__(128)
k_neigh_rows(const double *__restrict__ x, int nlocal, CellGrid g, const int *__restrict__ cell_start, const int *__restrict__ cell_end, const int *__restrict__ atom_sorted,
             double cutneighsq, int *count, const long long *noff, int *neigh) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= nlocal) return;
  const double xi0 = x[3 * (size_t)i], xi1 = x[3 * (size_t)i + 1], xi2 = x[3 * (size_t)i + 2];
  const int cx = cell_coord(xi0, g.lo[0], g.inv, g.nc[0]), cy = cell_coord(xi1, g.lo[1], g.inv, g.nc[1]), cz = DIM == 3 ? cell_coord(xi2, g.lo[2], g.inv, g.nc[2]) : 0;
  int cnt = 0; long long o = FILL ? noff[i] : 0;
  for (int dz = (DIM == 3 ? -1 : 0); dz <= (DIM == 3 ? 1 : 0); ++dz) {
    const int z = cz + dz; if (z < 0 || z >= g.nc[2]) continue;
    for (int dy = -1; dy <= 1; ++dy) {
      const int y = cy + dy; if (y < 0 || y >= g.nc[1]) continue;
      for (int dx = -1; dx <= 1; ++dx) {
        const int xx = cx + dx; if (xx < 0 || xx >= g.nc[0]) continue;
        const int c = (z * g.nc[1] + y) * g.nc[0] + xx;
        for (int k = cell_start[c]; k < cell_end[c]; ++k) {
          const int j = atom_sorted[k]; if (j == i) continue;
          double d = xi0 - x[3 * (size_t)j]; double rsq = d * d;
          d = xi1 - x[3 * (size_t)j + 1]; rsq += d * d;
          if (DIM == 3) { d = xi2 - x[3 * (size_t)j + 2]; rsq += d * d; }
          if (rsq <= cutneighsq) { if (FILL) neigh[o + cnt] = j; ++cnt; }
        }
      }
    }
  }
  if (!FILL) count[i] = cnt;
}
__global__ void k_count_to_ll(const int *count, int n, long long *out) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i <= n) out[i] = i < n ? count[i] : 0; }
__global__ void k_iota(int *p, int n) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = i; }

struct NeighWork { DevBuf<int> cell, cell2, atom, atom2, cstart, cend, count; DevBuf<long long> cnt64; DevBuf<double> part; DevBuf<char> tmp; PinBuf<double> h; };

void neighbors_build(Ctx *c, double cutneigh) {
  ISPH_REQUIRE(c->have_pair && c->have_atoms, "isph_neighbors_build: pair_coeff and atoms must be set first");
  if (!c->nwork) c->nwork = new NeighWork();
  NeighWork &W = *c->nwork; const int nall = c->nall, nl = c->nlocal, dim = c->tab.dim;
  if (cutneigh <= 0.0) { for (int a = 1; a <= c->tab.ntypes; ++a) for (int b = 1; b <= c->tab.ntypes; ++b) cutneigh = std::max(cutneigh, c->tab.cut[a][b]); }
  c->tic("buildNeighbors");
  // bounding box of all atoms (owned + ghost)
  const int nb = 256; W.part.ensure(6 * nb); W.h.ensure(6 * nb + 16);
  k_bbox<<<nb, 256, 0, c->stream>>>(c->x.p, nall, W.part.p); ++c->launches;
  CUDA_CHECK(cudaMemcpyAsync(W.h.p, W.part.p, sizeof(double) * 6 * nb, cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream));
  CellGrid g; g.dim = dim; g.inv = 1.0 / cutneigh;
  long long ncell = 1;
  for (int k = 0; k < 3; ++k) {
    double mn = 1e300, mx = -1e300; for (int b = 0; b < nb; ++b) { mn = std::min(mn, W.h.p[6 * b + k]); mx = std::max(mx, W.h.p[6 * b + 3 + k]); }
    g.lo[k] = mn; g.nc[k] = (k < dim) ? std::max(1, (int)std::floor((mx - mn) * g.inv) + 1) : 1; ncell *= g.nc[k];
  }
  ISPH_REQUIRE(ncell < (1ll << 30), "isph_neighbors_build: too many cells");
  W.cell.ensure(nall); W.cell2.ensure(nall); W.atom.ensure(nall); W.atom2.ensure(nall); W.cstart.ensure(ncell + 1); W.cend.ensure(ncell + 1); W.count.ensure(nl + 1); W.cnt64.ensure(nl + 2);
  k_cell_of_atom<<<ceil_div(nall, 256), 256, 0, c->stream>>>(c->x.p, nall, g, W.cell.p, W.atom.p); ++c->launches;
  int bits = 1; while ((1ll << bits) < ncell) ++bits;
  size_t tb = 0; cub::DeviceRadixSort::SortPairs(nullptr, tb, W.cell.p, W.cell2.p, W.atom.p, W.atom2.p, nall, 0, bits, c->stream); W.tmp.ensure(tb);
  tb = W.tmp.cap; cub::DeviceRadixSort::SortPairs(W.tmp.p, tb, W.cell.p, W.cell2.p, W.atom.p, W.atom2.p, nall, 0, bits, c->stream); ++c->launches;
  CUDA_CHECK(cudaMemsetAsync(W.cstart.p, 0, sizeof(int) * (ncell + 1), c->stream)); CUDA_CHECK(cudaMemsetAsync(W.cend.p, 0, sizeof(int) * (ncell + 1), c->stream));
  k_cell_bounds<<<ceil_div(nall, 256), 256, 0, c->stream>>>(W.cell2.p, nall, W.cstart.p, W.cend.p); ++c->launches;
  const double cnsq = cutneigh * cutneigh;
  if (dim == 2) k_neigh_rows<2, false><<<ceil_div(nl, 128), 128, 0, c->stream>>>(c->x.p, nl, g, W.cstart.p, W.cend.p, W.atom2.p, cnsq, W.count.p, nullptr, nullptr);
  else k_neigh_rows<3, false><<<ceil_div(nl, 128), 128, 0, c->stream>>>(c->x.p, nl, g, W.cstart.p, W.cend.p, W.atom2.p, cnsq, W.count.p, nullptr, nullptr);
  ++c->launches;
  c->noff.ensure(nl + 1);
  k_count_to_ll<<<ceil_div(nl + 1, 256), 256, 0, c->stream>>>(W.count.p, nl, W.cnt64.p); ++c->launches;
  tb = 0; cub::DeviceScan::ExclusiveSum(nullptr, tb, W.cnt64.p, c->noff.p, nl + 1, c->stream); W.tmp.ensure(tb);
  tb = W.tmp.cap; cub::DeviceScan::ExclusiveSum(W.tmp.p, tb, W.cnt64.p, c->noff.p, nl + 1, c->stream); ++c->launches;
  int *d_max = c->flag.p + 7; tb = 0; cub::DeviceReduce::Max(nullptr, tb, W.count.p, d_max, nl, c->stream); W.tmp.ensure(tb);
  tb = W.tmp.cap; cub::DeviceReduce::Max(W.tmp.p, tb, W.count.p, d_max, nl, c->stream); ++c->launches;
  long long tot = 0; int mj = 0;
  CUDA_CHECK(cudaMemcpyAsync(&tot, c->noff.p + nl, sizeof(long long), cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaMemcpyAsync(&mj, d_max, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  c->neigh.ensure(tot > 0 ? tot : 1);
  if (dim == 2) k_neigh_rows<2, true><<<ceil_div(nl, 128), 128, 0, c->stream>>>(c->x.p, nl, g, W.cstart.p, W.cend.p, W.atom2.p, cnsq, nullptr, c->noff.p, c->neigh.p);
  else k_neigh_rows<3, true><<<ceil_div(nl, 128), 128, 0, c->stream>>>(c->x.p, nl, g, W.cstart.p, W.cend.p, W.atom2.p, cnsq, nullptr, c->noff.p, c->neigh.p);
  ++c->launches;
  c->ilist.ensure(nl); k_iota<<<ceil_div(nl, 256), 256, 0, c->stream>>>(c->ilist.p, nl); ++c->launches;
  c->inum = nl; c->max_jnum = mj; c->nneigh = tot; c->have_neigh = true; c->neigh_on_device = true; c->h_noff.clear(); c->A.built = false;
  c->toc("buildNeighbors");
}

void neighbors_destroy(Ctx *c) {
  if (!c->nwork) return; NeighWork &W = *c->nwork;
  W.cell.release(); W.cell2.release(); W.atom.release(); W.atom2.release(); W.cstart.release(); W.cend.release(); W.count.release(); W.cnt64.release(); W.part.release(); W.tmp.release(); W.h.release();
  delete c->nwork; c->nwork = nullptr;
}

// slice capacities of the SELL matrix from a device-resident list: cap(s) = max over the slice's rows of jnum + 1
__global__ void k_slice_cap(const long long *noff, int n, int nslices, long long *cap32) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x; if (s > nslices) return;
  long long m = 0;
  if (s < nslices) { const int r1 = min(n, (s + 1) * 32); for (int r = s * 32; r < r1; ++r) m = max(m, noff[r + 1] - noff[r] + 1); }
  cap32[s] = 32 * m;
}
long long slice_offsets_device(Ctx *c, int n, int nslices, long long *d_slice_off) {
  if (!c->nwork) c->nwork = new NeighWork();
  NeighWork &W = *c->nwork; W.cnt64.ensure(nslices + 2);
  k_slice_cap<<<ceil_div(nslices + 1, 256), 256, 0, c->stream>>>(c->noff.p, n, nslices, W.cnt64.p); ++c->launches;
  size_t tb = 0; cub::DeviceScan::ExclusiveSum(nullptr, tb, W.cnt64.p, d_slice_off, nslices + 1, c->stream); W.tmp.ensure(tb);
  tb = W.tmp.cap; cub::DeviceScan::ExclusiveSum(W.tmp.p, tb, W.cnt64.p, d_slice_off, nslices + 1, c->stream); ++c->launches;
  long long tot = 0; CUDA_CHECK(cudaMemcpyAsync(&tot, d_slice_off + nslices, sizeof(long long), cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream));
  return tot;
}

}  // namespace isph
