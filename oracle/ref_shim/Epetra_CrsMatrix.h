// TEST INFRASTRUCTURE ONLY (oracle).  Stand-in for the handful of Epetra classes that the
// reference's header-only assembly functors touch (SURVEY.md §8a "Epetra matrix ops used by the
// above").  This is NOT Trilinos and not product code: it exists so that the reference's own
// functor headers under /root/reference/IMPLICIT-SPH can be compiled, unmodified, into
// oracle/_ref/libisph_ref.so and serve as the parity checker for graph/values/RHS.
// Semantics kept: rows addressed by global id (atom tag), duplicate column insertions merged at
// FillComplete, SumInto accumulates, canonical in-row order = ascending global id.
#pragma once
#include <vector>
#include <map>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <unordered_map>

enum Epetra_DataAccess { Copy, View };

struct Epetra_Comm { int MyPID() const { return 0; } int NumProc() const { return 1; } };
typedef Epetra_Comm Epetra_MpiComm;
typedef Epetra_Comm Epetra_SerialComm;

class Epetra_SerialDenseVector {
  std::vector<double> own_; double *p_; int n_;
public:
  Epetra_SerialDenseVector() : p_(nullptr), n_(0) {}
  Epetra_SerialDenseVector(int n) : own_(n, 0.0), p_(nullptr), n_(n) { p_ = own_.data(); }
  Epetra_SerialDenseVector(Epetra_DataAccess cv, double *v, int n) : p_(v), n_(n) {
    if (cv == Copy) { own_.assign(v, v + n); p_ = own_.data(); }
  }
  int Size(int n) { own_.assign(n, 0.0); p_ = own_.data(); n_ = n; return 0; }
  int Resize(int n) { own_.resize(n, 0.0); p_ = own_.data(); n_ = n; return 0; }
  int Length() const { return n_; }
  double *Values() { return p_; }
  const double *Values() const { return p_; }
  double &operator[](int i) { return p_[i]; }
  const double &operator[](int i) const { return p_[i]; }
  int Scale(double a) { for (int i = 0; i < n_; ++i) p_[i] *= a; return 0; }
};

class Epetra_IntSerialDenseVector {
  std::vector<int> own_; int *p_; int n_;
public:
  Epetra_IntSerialDenseVector() : p_(nullptr), n_(0) {}
  Epetra_IntSerialDenseVector(int n) : own_(n, 0), p_(nullptr), n_(n) { p_ = own_.data(); }
  Epetra_IntSerialDenseVector(Epetra_DataAccess cv, int *v, int n) : p_(v), n_(n) {
    if (cv == Copy) { own_.assign(v, v + n); p_ = own_.data(); }
  }
  int Size(int n) { own_.assign(n, 0); p_ = own_.data(); n_ = n; return 0; }
  int Length() const { return n_; }
  int *Values() { return p_; }
  int &operator[](int i) { return p_[i]; }
  int InfNorm() const { int m = 0; for (int i = 0; i < n_; ++i) m = std::max(m, std::abs(p_[i])); return m; }
};

class Epetra_Map {
  std::vector<int> gid_; std::unordered_map<int, int> lid_;
public:
  Epetra_Map(int /*nglobal*/, int nlocal, const int *gids, int /*base*/, const Epetra_Comm &) {
    gid_.assign(gids, gids + nlocal);
    for (int i = 0; i < nlocal; ++i) lid_[gids[i]] = i;
  }
  int NumMyElements() const { return (int)gid_.size(); }
  int GID(int lid) const { return gid_[lid]; }
  int LID(int gid) const { auto it = lid_.find(gid); return it == lid_.end() ? -1 : it->second; }
};
typedef Epetra_Map Epetra_BlockMap;

class Epetra_MultiVector {
protected:
  const Epetra_Map *map_; std::vector<double> own_; double *p_; int lda_, nvec_;
public:
  Epetra_MultiVector(const Epetra_Map &m, int nvec, bool = true)
    : map_(&m), own_((size_t)m.NumMyElements() * nvec, 0.0), lda_(m.NumMyElements()), nvec_(nvec) { p_ = own_.data(); }
  Epetra_MultiVector(Epetra_DataAccess, const Epetra_Map &m, double *v, int lda, int nvec)
    : map_(&m), p_(v), lda_(lda), nvec_(nvec) {}
  int MyLength() const { return map_->NumMyElements(); }
  int NumVectors() const { return nvec_; }
  int Stride() const { return lda_; }
  double *Values() const { return p_; }
  double *col(int k) const { return p_ + (size_t)k * lda_; }
  int PutScalar(double a) { for (int k = 0; k < nvec_; ++k) for (int i = 0; i < MyLength(); ++i) col(k)[i] = a; return 0; }
  int Scale(double a) { for (int k = 0; k < nvec_; ++k) for (int i = 0; i < MyLength(); ++i) col(k)[i] *= a; return 0; }
  const Epetra_Map &Map() const { return *map_; }
};

class Epetra_Vector : public Epetra_MultiVector {
public:
  Epetra_Vector(const Epetra_Map &m, bool z = true) : Epetra_MultiVector(m, 1, z) {}
  Epetra_Vector(Epetra_DataAccess cv, const Epetra_Map &m, double *v) : Epetra_MultiVector(cv, m, v, m.NumMyElements(), 1) {}
  double &operator[](int i) { return p_[i]; }
  const double &operator[](int i) const { return p_[i]; }
  int Reciprocal(const Epetra_Vector &a) { for (int i = 0; i < MyLength(); ++i) p_[i] = 1.0 / a.p_[i]; return 0; }
};

class Epetra_CrsGraph {
  friend class Epetra_CrsMatrix;
  const Epetra_Map *map_;
  std::vector<std::vector<int>> rows_;   // global column ids per local row
  bool filled_;
public:
  std::vector<int> rowptr, col;          // after FillComplete: canonical CSR (cols = sorted global ids)
  Epetra_CrsGraph(Epetra_DataAccess, const Epetra_Map &m, const int *rowsizes, bool /*static profile*/)
    : map_(&m), rows_(m.NumMyElements()), filled_(false) {
    for (int i = 0; i < m.NumMyElements(); ++i) rows_[i].reserve(rowsizes[i]);
  }
  int InsertGlobalIndices(int grow, int n, int *idx) {
    int l = map_->LID(grow); if (l < 0) return -1;
    rows_[l].insert(rows_[l].end(), idx, idx + n); return 0;
  }
  int FillComplete() {
    if (filled_) return 0;
    int n = (int)rows_.size(); rowptr.assign(n + 1, 0);
    for (int i = 0; i < n; ++i) {
      auto &r = rows_[i]; std::sort(r.begin(), r.end()); r.erase(std::unique(r.begin(), r.end()), r.end());
      rowptr[i + 1] = rowptr[i] + (int)r.size();
    }
    col.resize(rowptr[n]);
    for (int i = 0; i < n; ++i) std::copy(rows_[i].begin(), rows_[i].end(), col.begin() + rowptr[i]);
    rows_.clear(); rows_.shrink_to_fit(); filled_ = true; return 0;
  }
  int OptimizeStorage() { return 0; }
  int MaxNumIndices() const { int m = 0; for (size_t i = 0; i + 1 < rowptr.size(); ++i) m = std::max(m, rowptr[i + 1] - rowptr[i]); return m; }
  const Epetra_Map &RowMap() const { return *map_; }
};

class Epetra_CrsMatrix {
  const Epetra_CrsGraph *g_; bool filled_;
  int find(int l, int gcol) const {
    const int *b = g_->col.data() + g_->rowptr[l], *e = g_->col.data() + g_->rowptr[l + 1];
    const int *it = std::lower_bound(b, e, gcol);
    return (it != e && *it == gcol) ? (int)(it - g_->col.data()) : -1;
  }
public:
  std::vector<double> val;
  Epetra_CrsMatrix(Epetra_DataAccess, const Epetra_CrsGraph &g) : g_(&g), filled_(false), val(g.col.size(), 0.0) {}
  int FillComplete() { filled_ = true; return 0; }
  int OptimizeStorage() { return 0; }
  bool Filled() const { return filled_; }
  int PutScalar(double a) { std::fill(val.begin(), val.end(), a); return 0; }
  int SumIntoGlobalValues(int grow, int n, const double *v, const int *idx) {
    int l = g_->map_->LID(grow); if (l < 0) return -1; int err = 0;
    for (int k = 0; k < n; ++k) { int p = find(l, idx[k]); if (p < 0) err = 2; else val[p] += v[k]; }
    return err;
  }
  int ReplaceGlobalValues(int grow, int n, const double *v, const int *idx) {
    int l = g_->map_->LID(grow); if (l < 0) return -1; int err = 0;
    for (int k = 0; k < n; ++k) { int p = find(l, idx[k]); if (p < 0) err = 2; else val[p] = v[k]; }
    return err;
  }
  int ExtractDiagonalCopy(Epetra_Vector &d) const {
    for (int l = 0; l < g_->map_->NumMyElements(); ++l) { int p = find(l, g_->map_->GID(l)); d[l] = p < 0 ? 0.0 : val[p]; }
    return 0;
  }
  int ReplaceDiagonalValues(const Epetra_Vector &d) {
    for (int l = 0; l < g_->map_->NumMyElements(); ++l) { int p = find(l, g_->map_->GID(l)); if (p >= 0) val[p] = d[l]; }
    return 0;
  }
  int LeftScale(const Epetra_Vector &s) {
    for (int l = 0; l < g_->map_->NumMyElements(); ++l)
      for (int p = g_->rowptr[l]; p < g_->rowptr[l + 1]; ++p) val[p] *= s[l];
    return 0;
  }
  int Scale(double a) { for (auto &v : val) v *= a; return 0; }
  // single-process stand-in: every column id is owned locally
  int Multiply(bool /*trans*/, const Epetra_MultiVector &X, Epetra_MultiVector &Y) const {
    const Epetra_Map &m = *g_->map_; int n = m.NumMyElements();
    for (int k = 0; k < X.NumVectors(); ++k) {
      const double *x = X.col(k); double *y = Y.col(k);
      for (int l = 0; l < n; ++l) {
        double s = 0.0;
        for (int p = g_->rowptr[l]; p < g_->rowptr[l + 1]; ++p) s += val[p] * x[m.LID(g_->col[p])];
        y[l] = s;
      }
    }
    return 0;
  }
  int Apply(const Epetra_MultiVector &X, Epetra_MultiVector &Y) const { return Multiply(false, X, Y); }
  int ExtractGlobalRowView(int grow, int &n, double *&values) {
    int l = g_->map_->LID(grow); if (l < 0) return -1;
    n = g_->rowptr[l + 1] - g_->rowptr[l]; values = val.data() + g_->rowptr[l]; return 0;
  }
  const Epetra_Map &RowMap() const { return *g_->map_; }
  const Epetra_CrsGraph &Graph() const { return *g_; }
  // local-index read access used by the Epetra-typed adapter (include/solver_lin_b200_epetra.h): single-process stand-in, every
  // column id is owned locally, local column id = LID of the global id
  int NumMyRows() const { return g_->map_->NumMyElements(); }
  int MaxNumEntries() const { return g_->MaxNumIndices(); }
  int ExtractMyRowCopy(int lrow, int len, int &n, double *values, int *indices) const {
    if (lrow < 0 || lrow >= NumMyRows()) return -1;
    n = g_->rowptr[lrow + 1] - g_->rowptr[lrow]; if (n > len) return -2;
    for (int k = 0; k < n; ++k) { values[k] = val[g_->rowptr[lrow] + k]; indices[k] = g_->map_->LID(g_->col[g_->rowptr[lrow] + k]); }
    return 0;
  }
};
