// Preconditioners behind PrecondWrapper_Ifpack (precond_ifpack.h:14-85): create() = Ifpack::Create(type, A, overlap)
// -> SetParameters -> Initialize -> Compute ; the Krylov solver calls ApplyInverse ; free() drops the object.
// Ifpack is third-party code not vendored with the reference; the semantics implemented here are restated in
// oracle/krylov_oracle.cpp (Jacobi = Ifpack_PointRelaxation, Chebyshev = Ifpack_Chebyshev, ILU = Ifpack_ILU level 0 on
// the rank-local block with overlap 0).  ILU(0) lives in ilu.cu.
#include "isph_internal.h"

namespace isph {

static const int VB = 256;
static int vgrid(int n) { int g = ceil_div(n, VB); return g < 592 ? (g < 1 ? 1 : g) : 592; }

void ilu_create(Ctx *c);                      // ilu.cu
void ilu_free(Ctx *c);
void ilu_apply(Ctx *c, const double *r, double *z);

__global__ void k_invdiag(const double *d, double *inv, double min_diag, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  double v = d[i]; if (fabs(v) < min_diag) v = min_diag;          // "min diagonal value" of Ifpack_PointRelaxation / Ifpack_Chebyshev
  inv[i] = v != 0.0 ? 1.0 / v : 0.0;
}
__global__ void __launch_bounds__(VB) k_jacobi_first(const double *r, const double *invdiag, double damping, double *z, int n) {
  for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) z[i] = damping * invdiag[i] * r[i];
}
__global__ void __launch_bounds__(VB) k_jacobi_sweep(const double *r, const double *Az, const double *invdiag, double damping, double *z, int n) {
  for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) z[i] += damping * invdiag[i] * (r[i] - Az[i]);
}
// Ifpack_Chebyshev::ApplyInverse, zero starting solution: W = invDiag * X / theta ; Y = W
__global__ void __launch_bounds__(VB) k_cheb_first(const double *r, const double *invdiag, double oneOverTheta, double *W, double *z, int n) {
  for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) { const double w = invdiag[i] * r[i] * oneOverTheta; W[i] = w; z[i] = w; }
}
// W = dtemp1 * W + dtemp2 * invDiag * (X - V) ; Y += W
__global__ void __launch_bounds__(VB) k_cheb_step(const double *r, const double *V, const double *invdiag, double dtemp1, double dtemp2, double *W, double *z, int n) {
  for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) { double w = W[i] * dtemp1; w += dtemp2 * invdiag[i] * (r[i] - V[i]); W[i] = w; z[i] += w; }
}
__global__ void __launch_bounds__(VB) k_hash_vec(double *y, const int *tag, int n, int salt) {
  for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) {
    unsigned long long z = (unsigned long long)(tag ? tag[i] : i + 1) + 0x9E3779B97F4A7C15ULL * (unsigned long long)(salt + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; z = z ^ (z >> 31);
    y[i] = 2.0 * ((double)(z >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
  }
}
__global__ void __launch_bounds__(VB) k_mul_inplace(double *y, const double *d, int n) { for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) y[i] *= d[i]; }
__global__ void __launch_bounds__(VB) k_scaled_copy(double *x, const double *y, double s, int n) { for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) x[i] = y[i] * s; }
// block partial sums of a.b, a.a, b.b (3 values) -> combined on the host (setup path only)
__global__ void __launch_bounds__(VB) k_dot3(const double *a, const double *b, int n, double *partials) {
  double s0 = 0, s1 = 0, s2 = 0;
  for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) { s0 += a[i] * b[i]; s1 += a[i] * a[i]; s2 += b[i] * b[i]; }
  __shared__ double sm[3][VB / 32];
  for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = s0; sm[1][threadIdx.x >> 5] = s1; sm[2][threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x < 3) { double s = 0; for (int w = 0; w < VB / 32; ++w) s += sm[threadIdx.x][w]; partials[blockIdx.x * 3 + threadIdx.x] = s; }
}

static void dot3(Ctx *c, const double *a, const double *b, int n, double out[3]) {
  const int g = vgrid(n); c->red.ensure((size_t)592 * 64);
  k_dot3<<<g, VB, 0, c->stream>>>(a, b, n, c->red.p); ++c->launches;
  std::vector<double> h((size_t)g * 3);
  CUDA_CHECK(cudaMemcpyAsync(h.data(), c->red.p, sizeof(double) * g * 3, cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream));
  out[0] = out[1] = out[2] = 0.0;
  for (int b_ = 0; b_ < g; ++b_) for (int k = 0; k < 3; ++k) out[k] += h[(size_t)b_ * 3 + k];
  if (c->nranks > 1) { c->hbuf.ensure(8192); CUDA_CHECK(cudaMemcpyAsync(c->hbuf.p + 4000, out, 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    halo_allreduce(c, c->hbuf.p + 4000, 3); CUDA_CHECK(cudaMemcpyAsync(out, c->hbuf.p + 4000, 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream)); }
}

void precond_create(Ctx *c) {
  Matrix &A = c->A; ISPH_REQUIRE(A.built, ">> A is null");                      // precond_ifpack.h:52
  const int n = A.n, ld = c->ld; const std::string &t = c->pp.type;
  c->tic("precondCreate");
  // with one rank and one block there is nothing to overlap with: Ifpack's "Overlap Level" (default 1, precond_ifpack.h:37) is a no-op
  // across ranks "Overlap Level" 1 (the reference's own default, precond_ifpack.h:37) is the additive-Schwarz ILU of ilu.cu (rows of the
  // halo particles imported from their owners, combine mode Add); not provided: more than one level, overlap between sub-blocks of a
  // rank, overlap for the relaxation / Chebyshev types
  ISPH_REQUIRE(t == "ML" || c->pp.overlap == 0 || (c->nranks == 1 && !c->have_blocks) || (c->pp.overlap == 1 && !c->have_blocks && t == "ILU"),
               "Overlap Level: 0, or 1 with Precond Type ILU and one block per rank (more levels, overlapping sub-blocks and overlapping relaxation are not implemented)");
  if (c->prec_parent) {                                                          // block solve: prec->create(dim) builds ONE preconditioner from the scalar matrix the wrapper was given (prec->setMatrix(A.crs))
    c->prec_kind = 5; precond_create(c->prec_parent); c->prec_ready = true; c->toc("precondCreate"); return;
  }
  if (t == "none") c->prec_kind = 0;
  else if (t == "point relaxation" || t == "point relaxation stand-alone" || t == "Jacobi") { ISPH_REQUIRE(c->pp.relax_type == "Jacobi", "relaxation: type must be Jacobi"); c->prec_kind = 1; }
  else if (t == "Chebyshev") c->prec_kind = 2;
  else if (t == "ILU") { ISPH_REQUIRE(c->pp.fill >= 0 && c->pp.fill <= 255, "fact: level-of-fill must be in 0..255"); c->prec_kind = 3; }
  else if (t == "ML") c->prec_kind = 4;                                           // Precond Package = ML (pair_isph.cpp:325-329): amg.cu
  else ISPH_REQUIRE(false, "Precond Type not supported: " + t);
  if (c->prec_kind == 1 || c->prec_kind == 2) {
    c->invdiag.ensure(ld); c->cv.ensure(ld); c->cw.ensure(ld);
    matrix_extract_diag_dev(c, c->cv.p);
    k_invdiag<<<ceil_div(n, 256), 256, 0, c->stream>>>(c->cv.p, c->invdiag.p, c->pp.min_diag, n); ++c->launches;
  }
  if (c->prec_kind == 2) {
    ISPH_REQUIRE(c->pp.cheb_degree >= 1, "chebyshev: degree must be >= 1");
    double lmax = c->pp.cheb_lmax;
    if (lmax <= 0.0) {                                                           // Ifpack_Chebyshev::PowerMethod on D^-1 A
      const int g = vgrid(n); double *x = c->cw.p, *y = c->cv.p; double d[3];
      k_hash_vec<<<g, VB, 0, c->stream>>>(x, A.external ? nullptr : c->tag.p, n, 7); ++c->launches;
      dot3(c, x, x, n, d);
      k_scaled_copy<<<g, VB, 0, c->stream>>>(x, x, 1.0 / sqrt(d[1]), n); ++c->launches;
      for (int it = 0; it < c->pp.cheb_eig_iters; ++it) {
        spmv(c, x, y, 1, ld, ld);
        k_mul_inplace<<<g, VB, 0, c->stream>>>(y, c->invdiag.p, n); ++c->launches;
        dot3(c, x, y, n, d);                                                    // d[0] = y.x, d[1] = x.x, d[2] = y.y
        lmax = d[0] / d[1];
        k_scaled_copy<<<g, VB, 0, c->stream>>>(x, y, 1.0 / sqrt(d[2]), n); ++c->launches;
      }
    }
    c->last_lmax = lmax;
  }
  if (c->prec_kind == 3) ilu_create(c);
  if (c->prec_kind == 4) { amg_create(c); c->last_lmax = 0.0; }
  c->prec_ready = true;
  c->toc("precondCreate");
}

void precond_free(Ctx *c) { if (c->prec_kind == 5 && c->prec_parent) precond_free(c->prec_parent); if (c->prec_kind == 3) ilu_free(c); if (c->prec_kind == 4) amg_free(c); c->prec_ready = false; }

// z = M^-1 r  (Ifpack_Preconditioner::ApplyInverse as wrapped by Belos::EpetraPrecOp, solver_lin_belos.h:155)
void precond_apply(Ctx *c, const double *r, double *z) {
  ISPH_REQUIRE(c->prec_ready, "preconditioner not created");
  const int n = c->A.n, ld = c->ld, g = vgrid(n);
  switch (c->prec_kind) {
  case 0: CUDA_CHECK(cudaMemcpyAsync(z, r, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream)); break;
  case 1:
    k_jacobi_first<<<g, VB, 0, c->stream>>>(r, c->invdiag.p, c->pp.damping, z, n); ++c->launches;
    for (int s = 1; s < c->pp.sweeps; ++s) { spmv(c, z, c->cv.p, 1, ld, ld); k_jacobi_sweep<<<g, VB, 0, c->stream>>>(r, c->cv.p, c->invdiag.p, c->pp.damping, z, n); ++c->launches; }
    break;
  case 2: {
    const double lmax = c->last_lmax, alpha = lmax / c->pp.cheb_ratio, beta = 1.1 * lmax, delta = 2.0 / (beta - alpha), theta = 0.5 * (beta + alpha), s1 = theta * delta;
    k_cheb_first<<<g, VB, 0, c->stream>>>(r, c->invdiag.p, 1.0 / theta, c->cw.p, z, n); ++c->launches;
    double rhok = 1.0 / s1;
    for (int deg = 0; deg < c->pp.cheb_degree - 1; ++deg) {
      spmv(c, z, c->cv.p, 1, ld, ld);
      const double rhokp1 = 1.0 / (2.0 * s1 - rhok), dtemp1 = rhokp1 * rhok, dtemp2 = 2.0 * rhokp1 * delta; rhok = rhokp1;
      k_cheb_step<<<g, VB, 0, c->stream>>>(r, c->cv.p, c->invdiag.p, dtemp1, dtemp2, c->cw.p, z, n); ++c->launches;
    }
    break; }
  case 3: ilu_apply(c, r, z); break;
  case 4: amg_apply(c, r, z); break;
  case 5: { const int nb = c->prec_parent->A.n; for (int k = 0; k < c->prec_dim; ++k) precond_apply(c->prec_parent, r + (size_t)k * nb, z + (size_t)k * nb); break; }   // the same operator on every diagonal block
  }
}

}  // namespace isph
