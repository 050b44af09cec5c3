"""implicit-sph_b200 — B200-native linear-solve hot path of implicit-sph behind a C ABI.

The product is `libisph_b200.so` (csrc/, hand-written sm_100a CUDA + host C++, include/isph_b200.h).  This Python
module is only the ctypes binding used by tests/ and bench.py; there is no CPU fallback: creating a context without
a CUDA device fails loudly.  Import with `importlib.import_module("implicit-sph_b200")`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libisph_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "isph_b200.h")

KIND_FLUID, KIND_SOLID, KIND_BOUNDARY, KIND_BUFFER_DIRICHLET, KIND_BUFFER_NEUMANN, KIND_ALL = 99, 12, 16, 32, 64, 127
NOT_SINGULAR, NULLSPACE, PINZERO, DOUBLEDIAG = 0, 1, 2, 3
WENDLAND, CUBIC, QUINTIC = 0, 1, 2
INIT_RANDOM, INIT_ZERO, INIT_VALUE = 0, 1, 2
F_VFRAC, F_GC, F_LC, F_NORMAL, F_PND, F_DENSITY, F_VISCOSITY, F_PRESSURE, F_VELOCITY, F_VSTAR, F_FORCE, F_EPS, F_PSI, F_DP, F_PSI0, F_SIGMA, F_PHI = range(17)
FIELD_NCOMP = (1, 9, 6, 3, 1, 1, 1, 1, 3, 3, 3, 1, 1, 1, 1, 1, 1)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_longlong)


def build(force: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libisph_b200.so (nvcc cross-compiles without a GPU)."""
    args = ["make", "-s", "-C", os.path.join(HERE, "csrc"), "-j8"]
    if force:
        subprocess.run(args + ["clean"], check=True)
    subprocess.run(args, check=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run __graft_entry__.build() (there is no fallback path)")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        L.isph_last_error.restype = C.c_char_p
        L.isph_version.restype = C.c_char_p
        L.isph_graph_nnz.restype = C.c_longlong
        L.isph_neighbors_count.restype = C.c_longlong
        L.isph_kernel_launches.restype = C.c_longlong
        L.isph_solver_second_passes.restype = C.c_longlong
        L.isph_timer_ms.restype = C.c_double
        L.isph_timer_ms.argtypes = [C.c_void_p, C.c_char_p]
        _lib = L
    return _lib


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


class IsphError(RuntimeError):
    pass


class Context:
    """One GPU context = one LAMMPS rank's `PairISPH` solver state (matrix A, SolverLin, PrecondWrapper)."""

    def __init__(self, device=0, nranks=1, rank=0, nccl_id: bytes | None = None):
        self.L = lib()
        h = C.c_void_p()
        idbuf = C.create_string_buffer(nccl_id, 128) if nccl_id is not None else None
        rc = self.L.isph_ctx_create(C.byref(h), device, nranks, rank, idbuf)
        if rc != 0 or not h:
            raise IsphError("isph_ctx_create failed: no usable CUDA device or bad arguments (no CPU fallback exists)")
        self.h = h
        self.nlocal = self.nall = self.dim = 0
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.L.isph_ctx_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise IsphError(self.L.isph_last_error(self.h).decode())

    def call(self, name, *args):
        self._ck(getattr(self.L, name)(self.h, *args))

    # ---- inputs
    def set_stream(self, cuda_stream: int):
        self.call("isph_set_stream", C.c_void_p(cuda_stream))

    def synchronize(self):
        self.call("isph_synchronize")

    def pair_coeff(self, dim, kinds=(0, KIND_FLUID), h=1.0, h_min=None, cut_over_h=2.0, kernel=WENDLAND, morris_safe=0.43301):
        kinds = np.asarray(kinds, dtype=np.int32)
        self.dim = dim
        self.call("isph_pair_coeff", dim, len(kinds) - 1, _i(kinds), C.c_double(h), C.c_double(h if h_min is None else h_min),
                  C.c_double(cut_over_h), kernel, C.c_double(morris_safe))

    def atoms_set(self, nlocal, nghost, x, type_, tag):
        x = np.ascontiguousarray(x, dtype=np.float64); type_ = np.ascontiguousarray(type_, dtype=np.int32); tag = np.ascontiguousarray(tag, dtype=np.int32)
        assert x.shape == (nlocal + nghost, 3)
        self.nlocal, self.nall = nlocal, nlocal + nghost
        self.call("isph_atoms_set", nlocal, nghost, _d(x), _i(type_), _i(tag))

    def neighbors_set_packed(self, ilist, noff, neigh):
        ilist = np.ascontiguousarray(ilist, dtype=np.int32); noff = np.ascontiguousarray(noff, dtype=np.int64); neigh = np.ascontiguousarray(neigh, dtype=np.int32)
        self.call("isph_neighbors_set_packed", len(ilist), _i(ilist), noff.ctypes.data_as(_lp), _i(neigh))

    def neighbors_set(self, ilist, numneigh, firstneigh_rows):
        """LAMMPS layout: firstneigh_rows[i] is an int32 array for atom i (kept alive for the duration of the call)."""
        ilist = np.ascontiguousarray(ilist, dtype=np.int32); numneigh = np.ascontiguousarray(numneigh, dtype=np.int32)
        ptrs = (_ip * len(firstneigh_rows))(*[r.ctypes.data_as(_ip) for r in firstneigh_rows])
        self.call("isph_neighbors_set", len(ilist), _i(ilist), _i(numneigh), ptrs)

    def neighbors_build(self, cutneigh=0.0):
        """full neighbor list built on the device from the atoms already set (cutneigh <= 0: the pair cutoff)"""
        self.call("isph_neighbors_build", C.c_double(cutneigh))

    def neighbors_get(self):
        n = int(self.L.isph_neighbors_count(self.h)); noff = np.empty(self.nlocal + 1, dtype=np.int64); neigh = np.empty(max(n, 1), dtype=np.int32)
        self.call("isph_neighbors_get", noff.ctypes.data_as(_lp), _i(neigh)); return noff, neigh[:n]

    def set_particles(self, P, kinds=(0, KIND_FLUID), h_over_dx=1.5, h=None, h_min=None, cut_over_h=2.0, kernel=WENDLAND, morris_safe=0.43301):
        """Convenience: everything `lattice.make_brick` produced, in the call order LAMMPS would use."""
        h = P["dx"] * h_over_dx if h is None else h
        self.pair_coeff(P["dim"], kinds, h, h_min, cut_over_h, kernel, morris_safe)
        self.atoms_set(P["nlocal"], P["nghost"], P["x"], P["type"], P["tag"])
        self.neighbors_set_packed(P["ilist"], P["noff"], P["neigh"])

    def field_set(self, f, a):
        a = np.ascontiguousarray(a, dtype=np.float64); assert a.size == self.nall * FIELD_NCOMP[f], (a.shape, f)
        self.call("isph_field_set", f, _d(a))

    def field_get(self, f):
        nc = FIELD_NCOMP[f]; a = np.empty((self.nall, nc) if nc > 1 else (self.nall,), dtype=np.float64)
        self.call("isph_field_get", f, _d(a)); return a

    # ---- pre-computation / graph / matrix
    def compute_pre(self, normals=False):
        self.call("isph_compute_volumes"); self.call("isph_compute_gradient_correction"); self.call("isph_compute_laplacian_correction")
        if normals:
            self.call("isph_compute_normals")

    def graph_build(self):
        self.call("isph_graph_build")

    @property
    def nnz(self):
        return int(self.L.isph_graph_nnz(self.h))

    def graph_get(self):
        rowptr = np.empty(self.nlocal + 1, dtype=np.int32); col = np.empty(self.nnz, dtype=np.int32)
        self.call("isph_graph_get", _i(rowptr), _i(col)); return rowptr, col

    def matrix_get(self):
        v = np.empty(self.nnz); self.call("isph_matrix_get", _d(v)); return v

    def matrix_set_csr(self, rowptr, col, val):
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int32); col = np.ascontiguousarray(col, dtype=np.int32); val = np.ascontiguousarray(val, dtype=np.float64)
        self.nlocal = len(rowptr) - 1; self.nall = self.nlocal
        self.call("isph_matrix_set_csr", self.nlocal, _i(rowptr), _i(col), _d(val))

    def matrix_multiply(self, x):
        x = np.asfortranarray(np.asarray(x, dtype=np.float64).reshape(self.nlocal, -1, order="F")); y = np.zeros_like(x, order="F")
        self.call("isph_matrix_multiply", _d(x), _d(y), self.nlocal, x.shape[1]); return y

    def diagonals_get(self):
        d = np.empty(self.nlocal); s = np.empty(self.nlocal); self.call("isph_diagonals_get", _d(d), _d(s)); return d, s

    def ns_poisson(self, dt, anti=True, singular=NULLSPACE, morris_holmes=False):
        self.call("isph_ns_poisson", C.c_double(dt), int(anti), singular, int(morris_holmes))

    def ns_helmholtz(self, dt, theta, anti=True, morris_holmes=False, incremental_pressure=True, g=(0.0, 0.0, 0.0)):
        gg = np.asarray(g, dtype=np.float64)
        self.call("isph_ns_helmholtz", C.c_double(dt), C.c_double(theta), int(anti), int(morris_holmes), int(incremental_pressure), _d(gg))

    def applied_electric_potential(self):
        self.call("isph_applied_electric_potential")

    def solute_transport(self, dt, theta, dcoeff):
        self.call("isph_solute_transport", C.c_double(dt), C.c_double(theta), C.c_double(dcoeff))

    def pb_residual(self, morris_holmes=False, linearized=False, ezcb=0.5, psiref=1.0, gamma=0.0, extra_f=None):
        f = np.zeros(self.nlocal); ex = None if extra_f is None else np.ascontiguousarray(extra_f, dtype=np.float64)
        self.call("isph_pb_residual", int(morris_holmes), int(linearized), C.c_double(ezcb), C.c_double(psiref), C.c_double(gamma), None if ex is None else _d(ex), _d(f))
        return f

    def pb_newton(self, morris_holmes=False, linearized=False, ezcb=0.5, psiref=1.0, gamma=0.0, extra_f=None, max_newton=100, tol_f=1e-8, tol_update=1e-5, use_prec=True):
        ex = None if extra_f is None else np.ascontiguousarray(extra_f, dtype=np.float64)
        it = C.c_int(); li = C.c_int(); nf = C.c_double(); cv = C.c_int()
        self.call("isph_pb_newton", int(morris_holmes), int(linearized), C.c_double(ezcb), C.c_double(psiref), C.c_double(gamma), None if ex is None else _d(ex),
                  int(max_newton), C.c_double(tol_f), C.c_double(tol_update), int(use_prec), C.byref(it), C.byref(li), C.byref(nf), C.byref(cv))
        return dict(newton_iters=it.value, linear_iters=li.value, normf=nf.value, converged=bool(cv.value))

    def pb_jacobian(self, morris_holmes=False, linearized=False, ezcb=0.5, psiref=1.0, gamma=0.0):
        self.call("isph_pb_jacobian", int(morris_holmes), int(linearized), C.c_double(ezcb), C.c_double(psiref), C.c_double(gamma))

    def ns_correct(self, dt, anti=True, incremental_pressure=True, dp=None):
        if dp is not None:
            dp = np.ascontiguousarray(dp, dtype=np.float64); assert dp.size == self.nlocal
        self.call("isph_ns_correct", C.c_double(dt), int(anti), int(incremental_pressure), _d(dp))

    def pair_fixed(self, fixed_of_type):
        a = np.ascontiguousarray(fixed_of_type, dtype=np.int32); self.call("isph_pair_fixed", _i(a))

    def advance_time(self, dt, anti=True):
        self.call("isph_advance_time", C.c_double(dt), int(anti))

    def atoms_get_x(self):
        x = np.empty((self.nall, 3)); self.call("isph_atoms_get_x", _d(x)); return x

    def boundary_navier_slip(self, beta):
        self.call("isph_boundary_navier_slip", C.c_double(beta))

    def boundary_dirichlet(self):
        self.call("isph_boundary_dirichlet")

    def matrix_invalidate(self):
        self.call("isph_matrix_invalidate")

    def graph_invalidate(self):
        self.call("isph_graph_invalidate")

    def profile_spmv(self, enable=True):
        self.call("isph_profile_spmv", int(enable))

    def profile_spmv_get(self):
        ms = C.c_double(); n = C.c_longlong(); self.call("isph_profile_spmv_get", C.byref(ms), C.byref(n)); return ms.value, n.value

    def profile_precond_get(self):
        ms = C.c_double(); n = C.c_longlong(); self.call("isph_profile_precond_get", C.byref(ms), C.byref(n)); return ms.value, n.value

    def precond_info(self):
        z = C.c_longlong(); a = C.c_int(); b = C.c_int(); m = C.c_int(); self.call("isph_precond_info", C.byref(z), C.byref(a), C.byref(b), C.byref(m))
        return dict(factor_nnz=z.value, levels_lower=a.value, levels_upper=b.value, max_row=m.value)

    def precond_ml_info(self):
        nl = C.c_int(); rows = (C.c_int * 16)(); nnz = (C.c_longlong * 16)(); lm = (C.c_double * 16)()
        self.call("isph_precond_ml_info", C.byref(nl), rows, nnz, lm, 16)
        L = lib(); L.isph_precond_ml_setup_ms.restype = C.c_double
        return dict(levels=nl.value, rows=list(rows[:nl.value]), nnz=list(nnz[:nl.value]), lambda_max=list(lm[:nl.value]),
                    setup_ms={k: L.isph_precond_ml_setup_ms(self.h, k.encode()) for k in ("aggregate", "galerkin", "eigen")})

    def precond_ml_aggregates(self):
        a = np.zeros(self.nlocal, dtype=np.int32); self.call("isph_precond_ml_aggregates", _i(a)); return a

    # ---- SolverLin mirror
    def create_solution(self, x=None, nvec=1):
        """x: Fortran-ordered (nlocal, nvec) float64 array that receives the solution (a View, like the reference), or None."""
        if x is not None:
            assert x.dtype == np.float64 and (x.ndim == 1 or x.flags.f_contiguous); self._keep.append(x)
            self.call("isph_solver_create_solution_multivector", _d(x), self.nlocal, nvec)
        else:
            self.call("isph_solver_create_solution_multivector", None, self.nlocal, nvec)

    def create_load(self, b=None, nvec=1):
        if b is not None:
            assert b.dtype == np.float64 and (b.ndim == 1 or b.flags.f_contiguous); self._keep.append(b)
            self.call("isph_solver_create_load_multivector", _d(b), self.nlocal, nvec)
        else:
            self.call("isph_solver_create_load_multivector", None, self.nlocal, nvec)

    def load_set(self, b):
        b = np.asfortranarray(np.asarray(b, dtype=np.float64).reshape(self.nlocal, -1, order="F")); self.call("isph_solver_load_set", _d(b), self.nlocal)

    def load_get(self, nvec=1):
        b = np.zeros((self.nlocal, nvec), order="F"); self.call("isph_solver_load_get", _d(b), self.nlocal); return b

    def solution_set(self, x):
        x = np.asfortranarray(np.asarray(x, dtype=np.float64).reshape(self.nlocal, -1, order="F")); self.call("isph_solver_solution_set", _d(x), self.nlocal)

    def solution_get(self, nvec=1):
        x = np.zeros((self.nlocal, nvec), order="F"); self.call("isph_solver_solution_get", _d(x), self.nlocal); return x

    def set_null_vector_mask(self, mask):
        if mask is None:
            self.call("isph_solver_set_null_vector_mask", None)
        else:
            m = np.ascontiguousarray(mask, dtype=np.int32); self.call("isph_solver_set_null_vector_mask", _i(m))

    def set_matrix_is_singular(self, flag):
        self.call("isph_solver_set_matrix_is_singular", int(flag))

    def set_initial_solution(self, init_type, val=0.0):
        self.call("isph_solver_set_initial_solution", init_type, C.c_double(val))

    def solver_param(self, name, v):
        n = name.encode()
        if isinstance(v, bool) or isinstance(v, (int, np.integer)):
            self.call("isph_solver_set_param_int", n, int(v))
        elif isinstance(v, float):
            self.call("isph_solver_set_param_double", n, C.c_double(v))
        else:
            self.call("isph_solver_set_param_str", n, str(v).encode())

    def precond_param(self, name, v):
        n = name.encode()
        if isinstance(v, bool) or isinstance(v, (int, np.integer)):
            self.call("isph_precond_set_param_int", n, int(v))
        elif isinstance(v, float):
            self.call("isph_precond_set_param_double", n, C.c_double(v))
        else:
            self.call("isph_precond_set_param_str", n, str(v).encode())

    def precond_set_blocks(self, blk):
        if blk is None:
            self.call("isph_precond_set_blocks", None)
        else:
            b = np.ascontiguousarray(blk, dtype=np.int32); self.call("isph_precond_set_blocks", _i(b))

    def precond_create(self):
        self.call("isph_precond_create")

    def precond_free(self):
        self.call("isph_precond_free")

    def precond_apply(self, r):
        r = np.ascontiguousarray(r, dtype=np.float64); z = np.empty_like(r); self.call("isph_precond_apply", _d(r), _d(z)); return z

    def block_matrix(self, dim, blocks, name="Block"):
        """blocks: {(i, j): scipy CSR} over the nodal map (createBlockMatrix / setBlock / setBlockEnd, solver_lin.cpp:78-138)"""
        self.call("isph_solver_create_block_matrix", int(dim), name.encode())
        for (i, j), A in blocks.items():
            rp = np.ascontiguousarray(A.indptr, dtype=np.int32); ci = np.ascontiguousarray(A.indices, dtype=np.int32); va = np.ascontiguousarray(A.data, dtype=np.float64)
            self.call("isph_solver_set_block_csr", int(i), int(j), A.shape[0], _i(rp), _i(ci), _d(va))
        self.call("isph_solver_set_block_end")

    def solve_block(self, use_prec=True, label="Block"):
        self.call("isph_solver_solve_block", int(use_prec), label.encode())
        it = C.c_int(); rr = C.c_double(); cv = C.c_int(); lm = C.c_double()
        self.call("isph_solver_stats", C.byref(it), C.byref(rr), C.byref(cv), C.byref(lm))
        return dict(iters=it.value, relres=rr.value, converged=bool(cv.value))

    def solve(self, use_prec=True, label="Poisson"):
        self.call("isph_solver_solve", int(use_prec), label.encode())
        it = C.c_int(); rr = C.c_double(); cv = C.c_int(); lm = C.c_double()
        self.call("isph_solver_stats", C.byref(it), C.byref(rr), C.byref(cv), C.byref(lm))
        return dict(iters=it.value, relres=rr.value, converged=bool(cv.value), lambda_max=lm.value,
                    second_passes=int(self.L.isph_solver_second_passes(self.h)))

    def timer_ms(self, name):
        return float(self.L.isph_timer_ms(self.h, name.encode()))

    def timer_reset(self):
        self.call("isph_timer_reset")

    def halo_counts(self):
        a = C.c_int(); b = C.c_int(); p = C.c_int(); self.call("isph_halo_counts", C.byref(a), C.byref(b), C.byref(p))
        return dict(nhalo=a.value, nsend=b.value, npeers=p.value)

    @property
    def launches(self):
        return int(self.L.isph_kernel_launches(self.h))

    def measure_fp64_peak(self):
        t = C.c_double(); self.call("isph_measure_fp64_peak", C.byref(t)); return t.value

    def bench_spmv(self, reps=20):
        ms = C.c_double(); self.call("isph_bench_spmv", reps, C.byref(ms)); return ms.value


def declared_symbols():
    """Every function include/isph_b200.h declares (used by the CPU-side symbol test)."""
    import re
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(isph_[a-z0-9_]+)\s*\(", txt)))
