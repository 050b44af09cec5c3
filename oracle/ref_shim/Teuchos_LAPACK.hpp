// TEST INFRASTRUCTURE ONLY (oracle): Teuchos::LAPACK<int,double>::GESV stand-in (the image has no LAPACK).
// Column-major LU with partial pivoting (row interchanges), right-looking, then forward/back substitution;
// same mathematical algorithm as DGETF2+DGETRS.  Used by utils_reference.cpp:403 for the 3x3 / 6x6
// Laplacian-correction systems.  POSV / GELSS are never reached on the oracle's path and fail loudly.
#pragma once
#include <cmath>
#include <utility>
namespace Teuchos {
  template <class O, class S> class LAPACK {
  public:
    void GESV(int n, int nrhs, S *A, int lda, int *ipiv, S *B, int ldb, int *info) const {
      *info = 0;
      for (int k = 0; k < n; ++k) {
        int p = k; S mx = std::fabs(A[k + k * lda]);
        for (int i = k + 1; i < n; ++i) if (std::fabs(A[i + k * lda]) > mx) { mx = std::fabs(A[i + k * lda]); p = i; }
        ipiv[k] = p + 1;
        if (mx == S(0)) { if (!*info) *info = k + 1; continue; }
        if (p != k) for (int j = 0; j < n; ++j) std::swap(A[k + j * lda], A[p + j * lda]);
        S r = S(1) / A[k + k * lda];
        for (int i = k + 1; i < n; ++i) A[i + k * lda] *= r;
        for (int j = k + 1; j < n; ++j) { S t = A[k + j * lda]; for (int i = k + 1; i < n; ++i) A[i + j * lda] -= A[i + k * lda] * t; }
      }
      if (*info) return;
      for (int c = 0; c < nrhs; ++c) {
        S *b = B + c * ldb;
        for (int k = 0; k < n; ++k) { int p = ipiv[k] - 1; if (p != k) std::swap(b[k], b[p]); }
        for (int k = 0; k < n; ++k) for (int i = k + 1; i < n; ++i) b[i] -= A[i + k * lda] * b[k];
        for (int k = n - 1; k >= 0; --k) { b[k] /= A[k + k * lda]; for (int i = 0; i < k; ++i) b[i] -= A[i + k * lda] * b[k]; }
      }
    }
    void POSV(char, int, int, S *, int, S *, int, int *info) const { *info = -999; }
    void GELSS(int, int, int, S *, int, S *, int, S *, S, int *, S *, int, int *info) const { *info = -999; }
  };
}
