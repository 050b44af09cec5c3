"""TEST INFRASTRUCTURE ONLY — ctypes front-end to the CPU oracles (see oracle_api.h).

`Oracle(kind="port")` loads oracle/libisph_oracle.so (our restatement), `kind="ref"` loads
oracle/_ref/libisph_ref.so (the reference's own functor headers).  Both export the same symbols.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libisph_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libisph_ref.so")

FLUID, SOLID, BOUNDARY, BUFFER_DIRICHLET, BUFFER_NEUMANN, ALL = 99, 12, 16, 32, 64, 127
NOT_SINGULAR, NULLSPACE, PINZERO, DOUBLEDIAG = 0, 1, 2, 3
WENDLAND, CUBIC, QUINTIC = 0, 1, 2
F_VFRAC, F_GC, F_LC, F_NORMAL, F_PND, F_DENSITY, F_VISCOSITY, F_PRESSURE, F_VELOCITY, F_VSTAR, F_FORCE, F_EPS, F_PSI, F_DP, F_PSI0, F_SIGMA, F_PHI = range(17)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_longlong)


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def build(which=("port", "ref")):
    """(Re)build the oracle libraries with oracle/Makefile; `ref` is skipped when /root/reference is absent."""
    for w in which:
        subprocess.run(["make", "-s", "-C", HERE, w], check=True)


def have_ref():
    return os.path.exists(REF_SO)


_libs = {}


def _load(kind):
    if kind in _libs:
        return _libs[kind]
    path = PORT_SO if kind == "port" else REF_SO
    if not os.path.exists(path):
        build((kind,))
    L = C.CDLL(path)
    L.orc_create.restype = C.c_void_p
    L.orc_create.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _ip, _ip, C.c_int, _ip, _lp, _ip, C.c_int, _ip,
                             C.c_double, C.c_double, C.c_double, C.c_int, C.c_double]
    L.orc_destroy.argtypes = [C.c_void_p]
    L.orc_name.restype = C.c_char_p
    L.orc_graph.restype = C.c_longlong
    for f in ("orc_set_field", "orc_get_field"):
        getattr(L, f).argtypes = [C.c_void_p, C.c_int, _dp]
    for f in ("orc_compute_volumes", "orc_compute_gradient_correction", "orc_compute_laplacian_correction",
              "orc_compute_normals", "orc_graph", "orc_graph_max_row", "orc_invalidate_matrix"):
        getattr(L, f).argtypes = [C.c_void_p]
    L.orc_graph_get.argtypes = [C.c_void_p, _ip, _ip]
    L.orc_ns_poisson.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int, _dp]
    L.orc_ns_helmholtz.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, _dp, _dp]
    L.orc_pb_jacobian.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
    L.orc_ns_correct.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, _dp]
    L.orc_applied_electric_potential.argtypes = [C.c_void_p, _dp]
    L.orc_scalar_gradient.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp]
    L.orc_solute_transport.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, _dp]
    L.orc_pb_residual.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, _dp, _dp]
    L.orc_matrix_get.argtypes = [C.c_void_p, _dp]
    L.orc_set_fixed.argtypes = [C.c_void_p, _ip]
    L.orc_get_x.argtypes = [C.c_void_p, _dp]
    L.orc_advance_time.argtypes = [C.c_void_p, C.c_double, C.c_int]
    L.orc_boundary_navier_slip.argtypes = [C.c_void_p, C.c_double]
    L.orc_boundary_dirichlet.argtypes = [C.c_void_p, _dp, C.c_int]
    L.orc_diag_get.argtypes = [C.c_void_p, _dp, _dp]
    L.orc_spmv.argtypes = [C.c_void_p, _dp, _dp, C.c_int]
    _libs[kind] = L
    return L


class Oracle:
    """One particle configuration (= one LAMMPS time step's worth of inputs) on the CPU oracle."""

    def __init__(self, P, kinds=(0, FLUID), h_over_dx=1.5, h=None, h_min=None, cut_over_h=2.0, kernel=WENDLAND,
                 morris_safe=0.43301, kind="port"):
        self.L = _load(kind)
        self.kind = kind
        self.P = P
        self.dim = P["dim"]; self.nlocal = P["nlocal"]; self.nghost = P["nghost"]; self.nall = self.nlocal + self.nghost
        h = P["dx"] * h_over_dx if h is None else h
        h_min = h if h_min is None else h_min
        self.h = h
        kinds = np.asarray(kinds, dtype=np.int32)
        x = np.ascontiguousarray(P["x"], dtype=np.float64)
        self._keep = (x, P["type"], P["tag"], P["ilist"], P["noff"], P["neigh"], kinds)
        self.p = self.L.orc_create(self.dim, self.nlocal, self.nghost, _d(x), _i(P["type"]), _i(P["tag"]),
                                   len(P["ilist"]), _i(P["ilist"]), P["noff"].ctypes.data_as(_lp), _i(P["neigh"]),
                                   len(kinds) - 1, _i(kinds), h, h_min, cut_over_h, kernel, morris_safe)
        self.nnz = None

    def close(self):
        if self.p:
            self.L.orc_destroy(self.p); self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"oracle[{self.kind}] {what} failed rc={rc}")

    def set_field(self, f, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.size == self.nall * self.L.orc_field_ncomp(f), (a.shape, f)
        self._ck(self.L.orc_set_field(self.p, f, _d(a)), "set_field")

    def get_field(self, f):
        nc = self.L.orc_field_ncomp(f)
        a = np.empty((self.nall, nc) if nc > 1 else (self.nall,), dtype=np.float64)
        self._ck(self.L.orc_get_field(self.p, f, _d(a)), "get_field")
        return a

    def compute_pre(self, normals=False):
        self._ck(self.L.orc_compute_volumes(self.p), "volumes")
        self._ck(self.L.orc_compute_gradient_correction(self.p), "gradient_correction")
        self._ck(self.L.orc_compute_laplacian_correction(self.p), "laplacian_correction")
        if normals:
            self._ck(self.L.orc_compute_normals(self.p), "normals")

    def graph(self):
        nnz = self.L.orc_graph(self.p)
        if nnz < 0:
            raise RuntimeError("oracle graph failed")
        self.nnz = int(nnz)
        rowptr = np.empty(self.nlocal + 1, dtype=np.int32); col = np.empty(self.nnz, dtype=np.int32)
        self._ck(self.L.orc_graph_get(self.p, _i(rowptr), _i(col)), "graph_get")
        return rowptr, col

    def ns_poisson(self, dt, anti=True, singular=NULLSPACE, morris_holmes=False):
        b = np.zeros(self.nlocal)
        self._ck(self.L.orc_ns_poisson(self.p, dt, int(anti), singular, int(morris_holmes), _d(b)), "ns_poisson")
        return b

    def ns_helmholtz(self, dt, theta, b, anti=True, morris_holmes=False, incremental_pressure=True, g=(0.0, 0.0, 0.0)):
        b = np.asfortranarray(np.array(b, dtype=np.float64).reshape(self.nlocal, self.dim, order="F"))
        gg = np.asarray(g, dtype=np.float64)
        self._ck(self.L.orc_ns_helmholtz(self.p, dt, theta, int(anti), int(morris_holmes), int(incremental_pressure), _d(gg), _d(b)), "ns_helmholtz")
        return b

    def pb_jacobian(self, morris_holmes=False, linearized=False, ezcb=0.5, psiref=1.0, gamma=0.0):
        self._ck(self.L.orc_pb_jacobian(self.p, int(morris_holmes), int(linearized), ezcb, psiref, gamma), "pb_jacobian")

    def scalar_gradient(self, field, anti=False, morris_holmes=False, filter_i=FLUID, filter_j=127):
        g = np.zeros((self.nlocal, 3)); self._ck(self.L.orc_scalar_gradient(self.p, field, int(anti), int(morris_holmes), filter_i, filter_j, _d(g)), "scalar_gradient"); return g

    def applied_electric_potential(self):
        b = np.zeros(self.nlocal); self._ck(self.L.orc_applied_electric_potential(self.p, _d(b)), "applied_electric_potential"); return b

    def solute_transport(self, dt, theta, dcoeff, conc):
        b = np.array(conc, dtype=np.float64)[:self.nlocal].copy(); self._ck(self.L.orc_solute_transport(self.p, dt, theta, dcoeff, _d(b)), "solute_transport"); return b

    def pb_residual(self, morris_holmes=False, linearized=False, ezcb=0.5, psiref=1.0, gamma=0.0, extra_f=None):
        f = np.zeros(self.nlocal); ex = None if extra_f is None else np.ascontiguousarray(extra_f, dtype=np.float64)
        self._ck(self.L.orc_pb_residual(self.p, int(morris_holmes), int(linearized), ezcb, psiref, gamma,
                                        None if ex is None else _d(ex), _d(f)), "pb_residual")
        return f

    def ns_correct(self, dt, dp, anti=True, incremental_pressure=True):
        dp = np.ascontiguousarray(dp, dtype=np.float64); assert dp.size == self.nlocal
        self._ck(self.L.orc_ns_correct(self.p, dt, int(anti), int(incremental_pressure), _d(dp)), "ns_correct")

    def set_fixed(self, fixed_of_type):
        a = np.ascontiguousarray(fixed_of_type, dtype=np.int32); self._ck(self.L.orc_set_fixed(self.p, _i(a)), "set_fixed")

    def get_x(self):
        x = np.empty((self.nall, 3)); self._ck(self.L.orc_get_x(self.p, _d(x)), "get_x"); return x

    def advance_time(self, dt, anti=True):
        self._ck(self.L.orc_advance_time(self.p, dt, int(anti)), "advance_time")

    def boundary_navier_slip(self, beta):
        self._ck(self.L.orc_boundary_navier_slip(self.p, beta), "boundary_navier_slip")

    def boundary_dirichlet(self, b):
        b = np.asfortranarray(np.array(b, dtype=np.float64).reshape(self.nlocal, self.dim, order="F"))
        self._ck(self.L.orc_boundary_dirichlet(self.p, _d(b), self.nlocal), "boundary_dirichlet"); return b

    def invalidate_matrix(self):
        self.L.orc_invalidate_matrix(self.p)

    def matrix(self):
        v = np.empty(self.nnz)
        self._ck(self.L.orc_matrix_get(self.p, _d(v)), "matrix_get")
        return v

    def diagonals(self):
        d = np.empty(self.nlocal); s = np.empty(self.nlocal)
        self._ck(self.L.orc_diag_get(self.p, _d(d), _d(s)), "diag_get")
        return d, s

    def spmv(self, x):
        x = np.asfortranarray(np.asarray(x, dtype=np.float64).reshape(self.nlocal, -1, order="F"))
        y = np.zeros_like(x, order="F")
        self._ck(self.L.orc_spmv(self.p, _d(x), _d(y), x.shape[1]), "spmv")
        return y


# ---- Krylov / preconditioner oracle (port only; Trilinos is not available: "parity unpinned", krylov_oracle.cpp) ----
class KrylovParams(C.Structure):
    _fields_ = [("solver", C.c_int), ("flexible", C.c_int), ("num_blocks", C.c_int), ("max_iters", C.c_int), ("max_restarts", C.c_int),
                ("tol", C.c_double), ("precond", C.c_int), ("jacobi_sweeps", C.c_int), ("jacobi_damping", C.c_double), ("min_diag", C.c_double),
                ("cheb_degree", C.c_int), ("cheb_ratio", C.c_double), ("cheb_lambda_max", C.c_double), ("cheb_eig_iters", C.c_int),
                ("row_gid", _ip), ("ilu_fill", C.c_int), ("overlap", C.c_int),
                ("amg_max_levels", C.c_int), ("amg_threshold", C.c_double), ("amg_smoother", C.c_int), ("amg_pre", C.c_int), ("amg_post", C.c_int),
                ("amg_level_sweeps", C.c_int), ("amg_coarse_sweeps", C.c_int), ("amg_alpha", C.c_double), ("amg_coarse_alpha", C.c_double),
                ("amg_eig_iters", C.c_int), ("amg_max_coarse", C.c_int), ("amg_scale", C.c_double), ("amg_damping", C.c_double), ("amg_level_alpha", C.c_double), ("amg_level_scale", C.c_double), ("amg_coarse_direct", C.c_int)]


SOLVER_GMRES, SOLVER_CG = 0, 1
PREC_NONE, PREC_JACOBI, PREC_CHEBYSHEV, PREC_ILU0, PREC_AMG = 0, 1, 2, 3, 4


def krylov_params(**kw):
    L = _load("port")
    p = KrylovParams(); L.orc_krylov_default_params(C.byref(p))
    keep = []
    for k, v in kw.items():
        if k == "row_gid":
            a = np.ascontiguousarray(v, dtype=np.int32); keep.append(a); p.row_gid = _i(a)
        else:
            setattr(p, k, v)
    p._keep = keep
    return p


def krylov_solve(rowptr, col, val, b, x0=None, params=None, blocks=None, null_mask=None, use_null=False, history=False):
    """col: local row index per entry (-1 = column outside this process).  Returns x, info."""
    L = _load("port")
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32); col = np.ascontiguousarray(col, dtype=np.int32); val = np.ascontiguousarray(val, dtype=np.float64)
    n = len(rowptr) - 1
    b = np.array(b, dtype=np.float64); x = np.zeros(n) if x0 is None else np.array(x0, dtype=np.float64)
    p = params if params is not None else krylov_params()
    it = C.c_int(); rr = C.c_double()
    hist = np.zeros(p.max_iters + 8) if history else None
    bl = None if blocks is None else np.ascontiguousarray(blocks, dtype=np.int32)
    nm = None if null_mask is None else np.ascontiguousarray(null_mask, dtype=np.int32)
    rc = L.orc_krylov_solve(n, _i(rowptr), _i(col), _d(val), C.byref(p), None if bl is None else _i(bl), None if nm is None else _i(nm),
                            int(use_null), _d(b), _d(x), C.byref(it), C.byref(rr), None if hist is None else _d(hist), 0 if hist is None else len(hist))
    return x, dict(iters=it.value, relres=rr.value, converged=(rc == 0), history=None if hist is None else hist[:it.value + 1], b=b)


def krylov_solve_block(dim, stacked, prec, b, x0=None, params=None):
    """solveBlockProblem: `stacked` = scipy CSR of the dim x dim block operator (dim * nb rows), `prec` = scipy CSR of the scalar nb x nb matrix
    the block-diagonal preconditioner is built from; b, x = stacked vectors."""
    L = _load("port"); nb = prec.shape[0]
    rp = np.ascontiguousarray(stacked.indptr, dtype=np.int32); ci = np.ascontiguousarray(stacked.indices, dtype=np.int32); va = np.ascontiguousarray(stacked.data, dtype=np.float64)
    prp = np.ascontiguousarray(prec.indptr, dtype=np.int32); pci = np.ascontiguousarray(prec.indices, dtype=np.int32); pva = np.ascontiguousarray(prec.data, dtype=np.float64)
    b = np.array(b, dtype=np.float64); x = np.zeros(nb * dim) if x0 is None else np.array(x0, dtype=np.float64)
    p = params if params is not None else krylov_params(); it = C.c_int(); rr = C.c_double()
    rc = L.orc_krylov_solve_block(nb, int(dim), _i(rp), _i(ci), _d(va), _i(prp), _i(pci), _d(pva), C.byref(p), _d(b), _d(x), C.byref(it), C.byref(rr))
    return x, dict(iters=it.value, relres=rr.value, converged=(rc == 0))


def precond_apply(rowptr, col, val, r, params, blocks=None):
    L = _load("port")
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32); col = np.ascontiguousarray(col, dtype=np.int32); val = np.ascontiguousarray(val, dtype=np.float64)
    n = len(rowptr) - 1; r = np.ascontiguousarray(r, dtype=np.float64); z = np.zeros(n); lm = C.c_double()
    bl = None if blocks is None else np.ascontiguousarray(blocks, dtype=np.int32)
    L.orc_precond_apply(n, _i(rowptr), _i(col), _d(val), C.byref(params), None if bl is None else _i(bl), _d(r), _d(z), C.byref(lm))
    return z, lm.value


def amg_hierarchy(rowptr, col, val, params, blocks=None, cap_rows=0, cap_nnz=0):
    """Hierarchy of the multilevel stand-in for ML (amg_oracle.h): sizes per level, the aggregates of the finest level and (when it fits
    the capacities) the level-1 Galerkin operator."""
    L = _load("port")
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32); col = np.ascontiguousarray(col, dtype=np.int32); val = np.ascontiguousarray(val, dtype=np.float64)
    n = len(rowptr) - 1; rows = np.zeros(16, dtype=np.int32); nnz = np.zeros(16, dtype=np.int64); lmax = np.zeros(16); agg = np.zeros(n, dtype=np.int32)
    bl = None if blocks is None else np.ascontiguousarray(blocks, dtype=np.int32)
    crp = np.zeros(cap_rows + 1, dtype=np.int32); cci = np.zeros(max(cap_nnz, 1), dtype=np.int32); cva = np.zeros(max(cap_nnz, 1))
    L.orc_amg_hierarchy.restype = C.c_int
    nl = L.orc_amg_hierarchy(n, _i(rowptr), _i(col), _d(val), C.byref(params), None if bl is None else _i(bl), _i(rows), nnz.ctypes.data_as(C.POINTER(C.c_longlong)), _d(lmax), _i(agg),
                             int(cap_rows), C.c_longlong(int(cap_nnz)), _i(crp) if cap_rows else None, _i(cci) if cap_rows else None, _d(cva) if cap_rows else None)
    out = dict(levels=nl, rows=rows[:nl].copy(), nnz=nnz[:nl].copy(), lmax=lmax[:nl].copy(), agg=agg)
    if cap_rows and nl > 1 and rows[1] <= cap_rows and nnz[1] <= cap_nnz:
        out["coarse"] = (crp[:rows[1] + 1].copy(), cci[:nnz[1]].copy(), cva[:nnz[1]].copy())
    return out


def set_num_threads(n):
    """OpenMP threads of the port's row loops (returns the count in effect)."""
    return int(_load("port").orc_set_num_threads(int(n)))


def tags_to_local(col_tags, tags_owned):
    """canonical graph columns (global tags) -> local row indices (-1 when the tag is not owned here)"""
    lut = -np.ones(int(max(col_tags.max(), tags_owned.max())) + 2, dtype=np.int64)
    lut[tags_owned] = np.arange(len(tags_owned))
    return lut[col_tags].astype(np.int32)
