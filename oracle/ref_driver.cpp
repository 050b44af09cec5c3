// TEST INFRASTRUCTURE ONLY (oracle).  Drives the reference's OWN header-only functors
// (/root/reference/IMPLICIT-SPH/functor_*.h, included at build time, never copied) through a mock
// `PairIsph` that supplies exactly the members they duck-type on.  Output:
// oracle/_ref/libisph_ref.so exporting oracle_api.h.  The only logic restated here (because it lives in
// pair_isph*.cpp, which needs LAMMPS) is marked "restated from".
#include <vector>
#include <string>
#include <stdexcept>
#include <cstring>
#include <cmath>
#include <cfloat>
#include <iostream>
#include <iomanip>
#include <map>
#include <unordered_map>

#define FLERR __FILE__, __LINE__
#define NEIGHMASK 0x3FFFFFFF

#include "Epetra_CrsMatrix.h"
#include "utils.h"
namespace LAMMPS_NS { Utils util; }

#include "kernel.h"
#include "kernel_cubic.h"
#include "kernel_quintic.h"
#include "kernel_wendland.h"
#include "pair_for.h"
#include "filter.h"
#include "functor.h"
#include "functor_graph.h"
#include "functor_volume.h"
#include "functor_normal.h"
#include "functor_gradient_correction.h"
#include "functor_laplacian_correction.h"
#include "functor_gradient.h"
#include "functor_gradient_operator.h"
#include "functor_divergence.h"
#include "functor_laplacian.h"
#include "functor_laplacian_matrix.h"
#include "functor_boundary_morris_holmes.h"
#include "functor_gradient_dot_operator_matrix.h"
#include "functor_incomp_navier_stokes_helmholtz.h"
#include "functor_incomp_navier_stokes_poisson.h"
#include "functor_poisson_boltzmann_jacobian.h"
#include "functor_poisson_boltzmann_f.h"
#include "functor_applied_electric_potential.h"
#include "functor_solute_transport.h"
#include "functor_correct_velocity.h"
#include "functor_correct_pressure.h"
#include "functor_advance_time_begin.h"
#include "functor_advance_time_end.h"
#include "functor_boundary_navier_slip.h"
#include "functor_boundary_dirichlet.h"

#include "oracle_api.h"

using namespace LAMMPS_NS;

namespace {

struct MockPair;
struct MockAtom { double *vfrac, **x, *density, *viscosity, *pressure, **v, **f, *eps, *psi, *psi0, *sigma, *phi; int *type, *tag, *part; int nlocal, nghost; };
struct MockList { int inum, *ilist, *numneigh, **firstneigh; };
struct MockDomain { int dimension; };
struct MockError { void all(const char *, int, const char *msg) { throw std::runtime_error(msg); } };
struct MockComm { int me; MockPair *owner; void forward_comm_pair(MockPair *p); };
struct MockBlk { Epetra_CrsMatrix *operator()(int, int) const { return nullptr; } };

struct MockPair {
  // pair_isph.h:86-137
  enum CommType { Vfrac, NormalVector, Velocity, Pressure, Vstar, DeltaP, Psi, WorkScalar, WorkVector, TempScalar, TempVector, MaxCommType = 12 };
  enum ParticleKind { NoParticle = 0, FluidWithNormal = 1, Fluid = 99, SolidWithNormal = 4, Solid = 12, Boundary = 16,
                      BufferDirichlet = 32, BufferNeumann = 64, All = 127, MaxParticleKind = 128 };
  enum SingularPoisson { NotSingularPoisson = 0, NullSpace = 1, PinZero = 2, DoubleDiag = 3 };

  MockAtom *atom; MockList *list; MockDomain *domain; MockError *error; MockComm *comm;
  double **cutsq, **h; KernelFunction *kernel;
  double **Gc, **Lc, Gi[9], Li[6];
  double **normal, *pnd, **vstar, *dp, morris_safe_coeff;
  struct { Epetra_CrsMatrix *crs; Epetra_Vector *diagonal, *scaled_laplace_diagonal; int is_filled; } A;
  MockBlk A_blk;
  Epetra_CrsGraph *tags_in_cut; Epetra_Map *nodalmap;
  struct { int singular_poisson; bool is_incremental_pressure_used; double g[3]; } ns;
  struct { bool is_linearized; double ezcb, psiref, gamma; } pb;
  CommType comm_variable; int comm_forward;
  std::vector<int> kind_of_type, fixed_of_type;   // pinfo[0] (kind), pinfo[1] (fixed), pair_isph.cpp:160-167
  std::vector<int> owner_of_ghost;   // ghost atom -> owned atom with the same tag (single-process periodic images)

  int getParticleKind(int itype) const { return kind_of_type[itype]; }
  bool isParticleFixed(int itype) const { return fixed_of_type[itype] != 0; }

  // restated from pair_isph.cpp:493-520 (PairISPH::modifySingularMatrix)
  void modifySingularMatrix(const int row, double &diag, double &b) {
    switch (ns.singular_poisson) {
    case NotSingularPoisson: case NullSpace: break;
    case PinZero: { int ncol; double *values; A.crs->ExtractGlobalRowView(row, ncol, values);
                    memset(values, 0, sizeof(double) * ncol); diag = -1.0; b = 0.0; break; }
    case DoubleDiag: diag *= 1.5; break;
    }
  }
};

// stand-in for LAMMPS Comm::forward_comm_pair + PairISPH::pack/unpack_forward_comm (pair_isph.cpp:1924-2074):
// owned value -> every ghost copy of the same particle.
void MockComm::forward_comm_pair(MockPair *p) {
  const int nl = p->atom->nlocal, ng = p->atom->nghost;
  for (int g = 0; g < ng; ++g) {
    const int o = p->owner_of_ghost[g], a = nl + g;
    if (o < 0) continue;
    switch (p->comm_variable) {
    case MockPair::Vfrac: p->atom->vfrac[a] = p->atom->vfrac[o]; break;
    case MockPair::NormalVector: for (int k = 0; k < 3; ++k) p->normal[a][k] = p->normal[o][k]; p->pnd[a] = p->pnd[o]; break;
    case MockPair::Vstar: for (int k = 0; k < 3; ++k) p->vstar[a][k] = p->vstar[o][k]; break;
    case MockPair::DeltaP: p->dp[a] = p->dp[o]; break;
    case MockPair::Psi: p->atom->psi[a] = p->atom->psi[o]; break;
    default: break;
    }
  }
}

template <class T> struct Arr2 {   // contiguous [n][m] with row pointers (LAMMPS memory->create layout)
  std::vector<T> d; std::vector<T *> r; int m;
  void init(int n, int m_) { m = m_; d.assign((size_t)n * m_, T(0)); r.resize(n); for (int i = 0; i < n; ++i) r[i] = d.data() + (size_t)i * m_; }
  T **ptr() { return r.data(); }
};

typedef MockPair P;
}  // namespace

struct orc_problem {
  int dim, nlocal, nghost, nall, ntypes;
  Arr2<double> x, v, f, vstar, normal, Gc, Lc, cutsq, h;
  std::vector<double> vfrac, density, viscosity, pressure, eps, psi, psi0, sigma, phi, pnd, work, work3, dpv;
  std::vector<int> type, tag, ilist, numneigh, neigh; std::vector<int *> firstneigh;
  MockAtom atom; MockList list; MockDomain domain; MockError error; MockComm comm; MockPair pair;
  KernelFunction *kernel;
  std::string err;
  ~orc_problem() {
    delete pair.A.crs; delete pair.A.diagonal; delete pair.A.scaled_laplace_diagonal;
    delete pair.tags_in_cut; delete pair.nodalmap; delete kernel;
  }
};

extern "C" {

const char *orc_name(void) { return "reference functors (IMPLICIT-SPH/functor_*.h) + stand-in Epetra"; }

int orc_field_ncomp(int f) {
  static const int nc[ORC_F_COUNT] = {1, 9, 6, 3, 1, 1, 1, 1, 3, 3, 3, 1, 1, 1, 1, 1, 1};
  return (f >= 0 && f < ORC_F_COUNT) ? nc[f] : -1;
}

orc_problem *orc_create(int dim, int nlocal, int nghost, const double *x, const int *type, const int *tag,
                        int inum, const int *ilist, const long long *noff, const int *neigh,
                        int ntypes, const int *kind_of_type, double h_one, double h_min, double cut_over_h,
                        int kernel_id, double morris_safe) {
  orc_problem *q = new orc_problem();
  q->dim = dim; q->nlocal = nlocal; q->nghost = nghost; q->nall = nlocal + nghost; q->ntypes = ntypes;
  const int nall = q->nall;
  q->x.init(nall, 3); q->v.init(nall, 3); q->f.init(nall, 3); q->vstar.init(nall, 3); q->normal.init(nall, 3);
  q->Gc.init(nall, 9); q->Lc.init(nall, 6);
  memcpy(q->x.d.data(), x, sizeof(double) * 3 * nall);
  q->vfrac.assign(nall, 0.0); q->density.assign(nall, 1.0); q->viscosity.assign(nall, 0.0); q->pressure.assign(nall, 0.0);
  q->eps.assign(nall, 1.0); q->psi.assign(nall, 0.0); q->psi0.assign(nall, 0.0); q->sigma.assign(nall, 1.0); q->phi.assign(nall, 0.0); q->dpv.assign(nall, 0.0); q->pnd.assign(nall, 0.0); q->work.assign(nall, 0.0); q->work3.assign((size_t)nall * 3, 0.0);
  q->type.assign(type, type + nall); q->tag.assign(tag, tag + nall);
  q->ilist.assign(ilist, ilist + inum);
  q->neigh.assign(neigh, neigh + noff[inum]);
  q->numneigh.assign(nall, 0); q->firstneigh.assign(nall, nullptr);
  for (int ii = 0; ii < inum; ++ii) { int i = ilist[ii]; q->numneigh[i] = (int)(noff[ii + 1] - noff[ii]); q->firstneigh[i] = q->neigh.data() + noff[ii]; }

  MockPair &p = q->pair;
  q->atom = MockAtom{q->vfrac.data(), q->x.ptr(), q->density.data(), q->viscosity.data(), q->pressure.data(), q->v.ptr(), q->f.ptr(),
                     q->eps.data(), q->psi.data(), q->psi0.data(), q->sigma.data(), q->phi.data(), q->type.data(), q->tag.data(), nullptr, nlocal, nghost};
  q->list = MockList{inum, q->ilist.data(), q->numneigh.data(), q->firstneigh.data()};
  q->domain.dimension = dim; q->comm.me = 0; q->comm.owner = &p;
  p.atom = &q->atom; p.list = &q->list; p.domain = &q->domain; p.error = &q->error; p.comm = &q->comm;
  p.kind_of_type.assign(kind_of_type, kind_of_type + ntypes + 1);
  p.fixed_of_type.assign(ntypes + 1, 0);

  // restated from pair_isph_corrected.cpp:1289-1337 (PairISPH_Corrected::coeff): kernel choice, one cutoff for all
  // type pairs, h = h_one for equal kinds else h_min, MorrisSafeCoeff.
  switch (kernel_id) {
  case ORC_KERNEL_CUBIC: q->kernel = new KernelFuncCubic(dim); break;
  case ORC_KERNEL_QUINTIC: q->kernel = new KernelFuncQuintic(dim); break;
  default: q->kernel = new KernelFuncWendland(dim); break;
  }
  p.kernel = q->kernel;
  const double cut_one = h_one * cut_over_h, cut_one_sq = cut_one * cut_one;
  q->cutsq.init(ntypes + 1, ntypes + 1); q->h.init(ntypes + 1, ntypes + 1);
  for (int i = 1; i <= ntypes; ++i) for (int j = 1; j <= ntypes; ++j) {
    q->cutsq.ptr()[i][j] = cut_one_sq;
    q->h.ptr()[i][j] = (p.getParticleKind(i) == p.getParticleKind(j)) ? h_one : h_min;
  }
  p.cutsq = q->cutsq.ptr(); p.h = q->h.ptr(); p.morris_safe_coeff = morris_safe;
  p.Gc = q->Gc.ptr(); p.Lc = q->Lc.ptr(); p.normal = q->normal.ptr(); p.pnd = q->pnd.data(); p.vstar = q->vstar.ptr(); p.dp = q->dpv.data();
  // identity correction operators, restated from pair_isph_corrected.cpp:342-346,363-366
  memset(p.Gi, 0, sizeof(p.Gi)); memset(p.Li, 0, sizeof(p.Li));
  for (int k2 = 0; k2 < dim; ++k2) for (int k1 = 0; k1 < dim; ++k1) VIEW2(p.Gi, dim, k1, k2) = (k1 == k2);
  for (int k2 = 0, op = 0; k2 < dim; ++k2) for (int k1 = 0; k1 < (k2 + 1); ++k1, ++op) p.Li[op] = (k1 == k2);
  p.A.crs = nullptr; p.A.diagonal = nullptr; p.A.scaled_laplace_diagonal = nullptr; p.A.is_filled = 0;
  p.tags_in_cut = nullptr;
  p.ns.singular_poisson = P::NullSpace; p.ns.is_incremental_pressure_used = true; p.ns.g[0] = p.ns.g[1] = p.ns.g[2] = 0.0;
  p.pb.is_linearized = false; p.pb.ezcb = 0.5; p.pb.psiref = 1.0; p.pb.gamma = 0.0;
  // nodal map, restated from pair_isph.cpp:1258-1259
  p.nodalmap = new Epetra_Map(-1, nlocal, q->tag.data(), 1, Epetra_MpiComm());
  p.owner_of_ghost.assign(nghost, -1);
  for (int g = 0; g < nghost; ++g) p.owner_of_ghost[g] = p.nodalmap->LID(q->tag[nlocal + g]);
  return q;
}

void orc_destroy(orc_problem *p) { delete p; }

static double *field_ptr(orc_problem *q, int f) {
  switch (f) {
  case ORC_F_VFRAC: return q->vfrac.data(); case ORC_F_GC: return q->Gc.d.data(); case ORC_F_LC: return q->Lc.d.data();
  case ORC_F_NORMAL: return q->normal.d.data(); case ORC_F_PND: return q->pnd.data(); case ORC_F_DENSITY: return q->density.data();
  case ORC_F_VISCOSITY: return q->viscosity.data(); case ORC_F_PRESSURE: return q->pressure.data(); case ORC_F_VELOCITY: return q->v.d.data();
  case ORC_F_VSTAR: return q->vstar.d.data(); case ORC_F_FORCE: return q->f.d.data(); case ORC_F_EPS: return q->eps.data(); case ORC_F_PSI: return q->psi.data(); case ORC_F_DP: return q->dpv.data(); case ORC_F_PSI0: return q->psi0.data(); case ORC_F_SIGMA: return q->sigma.data(); case ORC_F_PHI: return q->phi.data();
  }
  return nullptr;
}
int orc_set_field(orc_problem *q, int f, const double *d) { double *p = field_ptr(q, f); if (!p) return -1; memcpy(p, d, sizeof(double) * q->nall * orc_field_ncomp(f)); return 0; }
int orc_get_field(orc_problem *q, int f, double *d) { double *p = field_ptr(q, f); if (!p) return -1; memcpy(d, p, sizeof(double) * q->nall * orc_field_ncomp(f)); return 0; }

#define ORC_TRY(...) try { __VA_ARGS__; return 0; } catch (std::exception & e) { q->err = e.what(); fprintf(stderr, "oracle/_ref: %s\n", e.what()); return -1; } catch (int) { return -1; }

int orc_compute_volumes(orc_problem *q) { ORC_TRY({ Corrected::FunctorOuterVolume<P> f(&q->pair); PairFor(f, f.getNumberOfWork()); }) }
int orc_compute_gradient_correction(orc_problem *q) { ORC_TRY({ Corrected::FunctorOuterGradientCorrection<P> f(&q->pair); PairFor(f, f.getNumberOfWork()); }) }
int orc_compute_laplacian_correction(orc_problem *q) { ORC_TRY({ Corrected::FunctorOuterLaplacianCorrection<P> f(&q->pair); PairFor(f, f.getNumberOfWork()); }) }

int orc_compute_normals(orc_problem *q) {
  // restated from pair_isph_corrected.cpp:379-427 (computeNormals, boundary_particle == Solid, no `part`, no bd_coord)
  ORC_TRY({
    MockPair &p = q->pair;
    double orient[P::MaxParticleKind] = {};
    orient[P::Fluid] = -1.0; orient[P::BufferDirichlet] = -1.0; orient[P::BufferNeumann] = -1.0; orient[P::Solid] = 1.0; orient[P::Boundary] = 1.0;
    { FilterBinary filter; filter.setPairYes(P::Fluid, P::Solid);
      Corrected::FunctorOuterNormal<P> f(&p, NULL, &orient[0], p.normal, p.pnd, NULL); f.setFilter(&filter); PairFor(f, f.getNumberOfWork()); }
    { FilterBinary filter; filter.setPairYes(P::Solid, P::Fluid);
      Corrected::FunctorOuterNormal<P> f(&p, NULL, &orient[0], p.normal, p.pnd, NULL); f.setFilter(&filter); PairFor(f, f.getNumberOfWork()); }
    p.comm_variable = P::NormalVector; p.comm_forward = 4; p.comm->forward_comm_pair(&p);
  })
}

long long orc_graph(orc_problem *q) {
  try {
    MockPair &p = q->pair;
    delete p.A.crs; delete p.A.diagonal; delete p.A.scaled_laplace_diagonal; delete p.tags_in_cut;
    p.A.crs = nullptr; p.tags_in_cut = nullptr;
    FunctorOuterGraph<P> f(&p, p.nodalmap, &p.tags_in_cut);
    PairFor(f, f.getNumberOfWork());
    // restated from pair_isph.cpp:1266-1270
    p.A.crs = new Epetra_CrsMatrix(Copy, *p.tags_in_cut); p.A.crs->FillComplete();
    p.A.diagonal = new Epetra_Vector(*p.nodalmap); p.A.scaled_laplace_diagonal = new Epetra_Vector(*p.nodalmap); p.A.is_filled = 0;
    return (long long)p.tags_in_cut->col.size();
  } catch (std::exception &e) { q->err = e.what(); return -1; }
}
int orc_graph_get(orc_problem *q, int *rowptr, int *col) {
  Epetra_CrsGraph *g = q->pair.tags_in_cut; if (!g) return -1;
  memcpy(rowptr, g->rowptr.data(), sizeof(int) * g->rowptr.size()); memcpy(col, g->col.data(), sizeof(int) * g->col.size()); return 0;
}
int orc_graph_max_row(orc_problem *q) { return q->pair.tags_in_cut ? q->pair.tags_in_cut->MaxNumIndices() : -1; }

int orc_ns_poisson(orc_problem *q, double dt, int anti, int singular, int mh, double *b) {
  ORC_TRY({
    MockPair &p = q->pair; p.ns.singular_poisson = singular;
    std::vector<double> bb(q->nall, 0.0);
    using namespace Corrected;
    // functor bindings as in pair_isph_corrected.cpp:169-178,210-219 ; call as in :969-1015
    if (anti && !mh) { FunctorOuterIncompNavierStokesPoisson<P, FunctorOuterLaplacianMatrixAntiSymmetric, FunctorOuterDivergenceAntiSymmetric, FunctorOuterGradientOperator>
        f(&p, dt, p.normal, p.atom->density, p.vstar, p.atom->pressure, bb.data(), q->work.data()); PairFor(f, f.getNumberOfWork()); }
    else if (anti && mh) { FunctorOuterIncompNavierStokesPoisson<P, FunctorOuterLaplacianMatrixAntiSymmetric, FunctorOuterDivergenceAntiSymmetric_MorrisHolmes, FunctorOuterGradientOperator>
        f(&p, dt, p.normal, p.atom->density, p.vstar, p.atom->pressure, bb.data(), q->work.data()); PairFor(f, f.getNumberOfWork()); }
    else if (!anti && !mh) { FunctorOuterIncompNavierStokesPoisson<P, FunctorOuterLaplacianMatrixSymmetric, FunctorOuterDivergenceSymmetric, FunctorOuterGradientOperator>
        f(&p, dt, p.normal, p.atom->density, p.vstar, p.atom->pressure, bb.data(), q->work.data()); PairFor(f, f.getNumberOfWork()); }
    else { FunctorOuterIncompNavierStokesPoisson<P, FunctorOuterLaplacianMatrixSymmetric, FunctorOuterDivergenceSymmetric_MorrisHolmes, FunctorOuterGradientOperator>
        f(&p, dt, p.normal, p.atom->density, p.vstar, p.atom->pressure, bb.data(), q->work.data()); PairFor(f, f.getNumberOfWork()); }
    memcpy(b, bb.data(), sizeof(double) * q->nlocal);
  })
}

int orc_ns_helmholtz(orc_problem *q, double dt, double theta, int anti, int mh, int incp, const double *g, double *b) {
  ORC_TRY({
    MockPair &p = q->pair; p.ns.is_incremental_pressure_used = incp != 0; for (int k = 0; k < 3; ++k) p.ns.g[k] = g ? g[k] : 0.0;
    std::fill(q->work3.begin(), q->work3.end(), 0.0);   // clearCommArray(WorkVector), pair_isph_corrected.cpp:873
    using namespace Corrected;
    // bindings: pair_isph_corrected.cpp:154-161,195-202 ; call :868-915
    if (anti && !mh) { FunctorOuterIncompNavierStokesHelmholtz<P, FunctorOuterLaplacianMatrixAntiSymmetric, FunctorOuterGradientAntiSymmetric>
        f(&p, dt, theta, p.atom->viscosity, p.atom->density, p.atom->pressure, p.atom->f, b, q->nlocal, q->work3.data()); PairFor(f, f.getNumberOfWork()); }
    else if (anti && mh) { FunctorOuterIncompNavierStokesHelmholtz<P, FunctorOuterLaplacianMatrixAntiSymmetric_MorrisHolmes, FunctorOuterGradientAntiSymmetric>
        f(&p, dt, theta, p.atom->viscosity, p.atom->density, p.atom->pressure, p.atom->f, b, q->nlocal, q->work3.data()); PairFor(f, f.getNumberOfWork()); }
    else if (!anti && !mh) { FunctorOuterIncompNavierStokesHelmholtz<P, FunctorOuterLaplacianMatrixSymmetric, FunctorOuterGradientSymmetric>
        f(&p, dt, theta, p.atom->viscosity, p.atom->density, p.atom->pressure, p.atom->f, b, q->nlocal, q->work3.data()); PairFor(f, f.getNumberOfWork()); }
    else { FunctorOuterIncompNavierStokesHelmholtz<P, FunctorOuterLaplacianMatrixSymmetric_MorrisHolmes, FunctorOuterGradientSymmetric>
        f(&p, dt, theta, p.atom->viscosity, p.atom->density, p.atom->pressure, p.atom->f, b, q->nlocal, q->work3.data()); PairFor(f, f.getNumberOfWork()); }
  })
}

int orc_pb_jacobian(orc_problem *q, int mh, int linearized, double ezcb, double psiref, double gamma) {
  ORC_TRY({
    MockPair &p = q->pair; p.pb.is_linearized = linearized != 0; p.pb.ezcb = ezcb; p.pb.psiref = psiref; p.pb.gamma = gamma;
    using namespace Corrected;
    // bindings: pair_isph_corrected.cpp:110-115 ; call :489-523
    if (!mh) { FunctorOuterPoissonBoltzmannJacobian<P, FunctorOuterLaplacianMatrixSymmetric> f(&p, p.atom->psi, p.atom->eps); PairFor(f, f.getNumberOfWork()); }
    else { FunctorOuterPoissonBoltzmannJacobian<P, FunctorOuterLaplacianMatrixSymmetric_MorrisHolmes> f(&p, p.atom->psi, p.atom->eps); PairFor(f, f.getNumberOfWork()); }
  })
}

int orc_scalar_gradient(orc_problem *q, int field, int anti, int mh, int f0, int f1, double *grad) {
  ORC_TRY({
    MockPair &p = q->pair; using namespace Corrected;
    double *fld = field_ptr(q, field); if (!fld || orc_field_ncomp(field) != 1) throw std::runtime_error("scalar field expected");
    if (field == ORC_F_PSI) { p.comm_variable = MockPair::Psi; p.comm_forward = 1; p.comm->forward_comm_pair(&p); }
    else for (int g = 0; g < q->nghost; ++g) { const int o = p.owner_of_ghost[g]; if (o >= 0) fld[q->nlocal + g] = fld[o]; }
    Arr2<double> out; out.init(q->nall, 3);
    FilterBinary filter; filter.setPairYes(f0, f1);
    // bindings: functor_gradient.h:15-16, functor_boundary_morris_holmes.h ; call pair_isph_corrected.cpp:537-549
    if (mh) { FunctorOuterGradient_MorrisHolmes<P> f(&p, fld, 1.0, out.ptr()); f.setFilter(&filter); PairFor(f, f.getNumberOfWork()); }
    else if (anti) { FunctorOuterGradient<P, true> f(&p, fld, 1.0, out.ptr()); f.setFilter(&filter); PairFor(f, f.getNumberOfWork()); }
    else { FunctorOuterGradient<P, false> f(&p, fld, 1.0, out.ptr()); f.setFilter(&filter); PairFor(f, f.getNumberOfWork()); }
    memcpy(grad, out.d.data(), sizeof(double) * 3 * q->nlocal);
  })
}
int orc_applied_electric_potential(orc_problem *q, double *b) {
  ORC_TRY({
    MockPair &p = q->pair; using namespace Corrected;
    // binding: pair_isph_corrected.cpp:142-144 ; call :604-611 (normal is passed but unused by the functor)
    FunctorOuterAppliedElectricPotential<P, FunctorOuterLaplacianMatrixSymmetric> f(&p, p.atom->sigma, p.atom->phi, p.normal, b);
    PairFor(f, f.getNumberOfWork());
  })
}
int orc_solute_transport(orc_problem *q, double dt, double theta, double dcoeff, double *b) {
  ORC_TRY({
    MockPair &p = q->pair; using namespace Corrected;
    std::fill(q->work.begin(), q->work.end(), 0.0);                                // clearCommArray(WorkScalar), pair_isph_corrected.cpp:850
    // binding: pair_isph_corrected.cpp:136-139 ; call :852-857
    FunctorOuterSoluteTransport<P, FunctorOuterLaplacianMatrixSymmetric, FunctorOuterGradientOperator> f(&p, p.atom->v, false, dt, theta, dcoeff, b, q->work.data());
    PairFor(f, f.getNumberOfWork());
  })
}

int orc_pb_residual(orc_problem *q, int mh, int linearized, double ezcb, double psiref, double gamma, const double *extra_f, double *fout) {
  ORC_TRY({
    MockPair &p = q->pair; p.pb.is_linearized = linearized != 0; p.pb.ezcb = ezcb; p.pb.psiref = psiref; p.pb.gamma = gamma;
    using namespace Corrected;
    // restated from pair_isph_corrected.cpp:446-450: psi is communicated to the ghosts before the functor runs
    p.comm_variable = MockPair::Psi; p.comm_forward = 1; p.comm->forward_comm_pair(&p);
    Epetra_Vector fv(View, *p.nodalmap, fout);
    // bindings: pair_isph_corrected.cpp:98-103 ; call :452-468
    if (!mh) { FunctorOuterPoissonBoltzmannF<P, FunctorOuterLaplacian> f(&p, p.atom->psi, p.atom->psi0, p.atom->eps, &fv); PairFor(f, f.getNumberOfWork()); }
    else { FunctorOuterPoissonBoltzmannF<P, FunctorOuterLaplacian_MorrisHolmes> f(&p, p.atom->psi, p.atom->psi0, p.atom->eps, &fv); PairFor(f, f.getNumberOfWork()); }
    // restated from functor_poisson_boltzmann_extra_f.h:76-90 (the expression itself is evaluated by the caller)
    if (extra_f) for (int ii = 0; ii < q->list.inum; ++ii) { const int i = q->list.ilist[ii]; if (!(p.getParticleKind(p.atom->type[i]) & MockPair::Solid)) fout[i] += extra_f[i]; }
  })
}

int orc_ns_correct(orc_problem *q, double dt, int anti, int incp, const double *dp_owned) {
  ORC_TRY({
    MockPair &p = q->pair; p.ns.is_incremental_pressure_used = incp != 0;
    memcpy(p.dp, dp_owned, sizeof(double) * q->nlocal);
    p.comm_variable = P::DeltaP; p.comm_forward = 1; p.comm->forward_comm_pair(&p);                  // pair_isph.cpp:1017-1019
    if (incp) {                                                                                       // restated from pair_isph.cpp:422-464 (computeZeroMeanPressure)
      int nloc = 0; double mysum = 0.0;
      for (int ii = 0; ii < q->list.inum; ++ii) { const int i = q->list.ilist[ii]; const int ikind = p.getParticleKind(q->type[i]);
        if (ikind == P::Solid) p.dp[i] = 0.0; else { mysum += p.dp[i]; ++nloc; } }
      const double mean_val = mysum / nloc;
      for (int i = 0; i < q->nall; ++i) p.dp[i] -= mean_val * (p.getParticleKind(q->type[i]) != P::Solid);
    }
    using namespace Corrected;                                                                        // bindings: pair_isph_corrected.cpp:180-182,221-223
    if (anti) { FunctorOuterCorrectVelocity<P, FunctorOuterGradientAntiSymmetric> f(&p, dt, p.atom->density, p.dp, p.vstar); PairFor(f, f.getNumberOfWork()); }
    else { FunctorOuterCorrectVelocity<P, FunctorOuterGradientSymmetric> f(&p, dt, p.atom->density, p.dp, p.vstar); PairFor(f, f.getNumberOfWork()); }
    { FunctorOuterCorrectPressure<P> f(&p, p.atom->pressure, p.dp, p.atom->nghost); PairFor(f, f.getNumberOfWork()); }   // pair_isph_corrected.cpp:1039-1052
  })
}
int orc_set_fixed(orc_problem *q, const int *fixed_of_type) { q->pair.fixed_of_type.assign(fixed_of_type, fixed_of_type + q->ntypes + 1); return 0; }
int orc_get_x(orc_problem *q, double *x) { memcpy(x, q->x.d.data(), sizeof(double) * 3 * q->nall); return 0; }
int orc_advance_time(orc_problem *q, double dt, int anti) {
  ORC_TRY({
    MockPair &p = q->pair; using namespace Corrected;
    // bindings pair_isph_corrected.cpp:185-187,226-228 ; call :1183-1194 (PairISPH_Corrected::advanceTime, ns enabled, no ALE)
    if (anti) { FunctorOuterAdvanceTimeBegin<P, FunctorOuterGradientAntiSymmetric> f(&p, dt, p.atom->v, p.atom->pressure); PairFor(f, f.getNumberOfWork()); }
    else { FunctorOuterAdvanceTimeBegin<P, FunctorOuterGradientSymmetric> f(&p, dt, p.atom->v, p.atom->pressure); PairFor(f, f.getNumberOfWork()); }
    { FunctorOuterAdvanceTimeEnd<P> f(&p, dt, p.atom->v, p.atom->pressure, p.atom->nghost); PairFor(f, f.getNumberOfWork()); }
  })
}
int orc_boundary_navier_slip(orc_problem *q, double beta) {
  ORC_TRY({
    MockPair &p = q->pair; using namespace Corrected;
    // call pair_isph_corrected.cpp:921-926 (after the Helmholtz functor; SumInto A.crs)
    if (beta != 0.0) { FunctorOuterBoundaryNavierSlip<P> f(&p, beta, p.normal, p.atom->density); PairFor(f, f.getNumberOfWork()); }
  })
}
int orc_boundary_dirichlet(orc_problem *q, double *b, int lda) {
  ORC_TRY({
    MockPair &p = q->pair; using namespace Corrected;
    FunctorOuterBoundaryDirichlet<P> f(&p, p.normal, b, lda); PairFor(f, f.getNumberOfWork());       // call pair_isph_corrected.cpp:928-932
  })
}
int orc_invalidate_matrix(orc_problem *q) { q->pair.A.is_filled = 0; return 0; }
int orc_matrix_get(orc_problem *q, double *val) { if (!q->pair.A.crs) return -1; memcpy(val, q->pair.A.crs->val.data(), sizeof(double) * q->pair.A.crs->val.size()); return 0; }
int orc_diag_get(orc_problem *q, double *d, double *s) {
  if (!q->pair.A.diagonal) return -1;
  for (int i = 0; i < q->nlocal; ++i) { if (d) d[i] = (*q->pair.A.diagonal)[i]; if (s) s[i] = (*q->pair.A.scaled_laplace_diagonal)[i]; }
  return 0;
}
int orc_spmv(orc_problem *q, const double *x, double *y, int nvec) {
  if (!q->pair.A.crs) return -1;
  Epetra_MultiVector X(View, *q->pair.nodalmap, const_cast<double *>(x), q->nlocal, nvec), Y(View, *q->pair.nodalmap, y, q->nlocal, nvec);
  return q->pair.A.crs->Multiply(false, X, Y);
}

}  // extern "C"
