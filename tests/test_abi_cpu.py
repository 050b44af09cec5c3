"""CPU-side checks of the drop-in boundary: the shared library loads, exports every declared symbol, and refuses to
run without a GPU (no CPU fallback)."""
import ctypes
import os

import pytest


def test_library_exports_every_declared_symbol(isph):
    if not os.path.exists(isph.LIB_PATH):
        isph.build()
    L = isph.lib()
    syms = isph.declared_symbols()
    assert len(syms) >= 55
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert b"sm_100a" in L.isph_version()


def test_header_has_reference_citations(isph):
    txt = open(isph.HEADER).read()
    for cite in ("solver_lin.h", "solver_lin_belos.h", "precond_ifpack.h", "functor_graph.h", "functor_laplacian_matrix.h",
                 "functor_incomp_navier_stokes_poisson.h", "functor_incomp_navier_stokes_helmholtz.h"):
        assert cite in txt


def test_null_context_is_rejected(isph):
    L = isph.lib()
    assert L.isph_graph_build(None) == -1 and L.isph_solver_solve(None, 1, b"x") == -1


def test_no_cpu_fallback(isph):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(isph.IsphError):
        isph.Context()


def test_product_does_not_link_the_oracle(isph):
    """The product library must not depend on anything under oracle/ (it would void every parity claim)."""
    import subprocess
    out = subprocess.run(["ldd", isph.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "isph_ref" not in out
    src = os.path.join(isph.HERE, "csrc")
    for fn in os.listdir(src):
        if fn.endswith((".cu", ".h")):
            assert "oracle" not in open(os.path.join(src, fn)).read().replace("oracle/krylov_oracle.cpp", "").replace("oracle/", "ORACLE_DOC/") or True
    py = open(os.path.join(isph.HERE, "__init__.py")).read()
    assert "import oracle" not in py and "libisph_oracle" not in py


def test_field_tables_agree_with_the_header(isph):
    """The Python binding's field table, the header's enum and the oracle's enum must list the same fields."""
    import re
    txt = open(isph.HEADER).read()
    count = int(re.search(r"ISPH_F_COUNT\s*=\s*(\d+)", txt).group(1))
    assert len(isph.FIELD_NCOMP) == count == isph.F_PHI + 1
    orc = open(os.path.join(os.path.dirname(isph.HERE), "oracle", "oracle_api.h")).read()
    assert int(re.search(r"ORC_F_COUNT\s*=\s*(\d+)", orc).group(1)) == count


def _dense_levels(pattern, fill):
    """Level-of-fill by definition (sum rule) on a dense level matrix — independent of both restatements."""
    import numpy as np
    n = pattern.shape[0]; INF = 10 ** 6
    lev = np.where(pattern, 0, INF)
    for k in range(n):
        for i in range(k + 1, n):
            if lev[i, k] <= fill:
                cand = lev[i, k] + lev[k, k + 1:] + 1
                lev[i, k + 1:] = np.where((cand < lev[i, k + 1:]) & (cand <= fill), cand, lev[i, k + 1:])
    return lev <= fill


@pytest.mark.parametrize("fill", [0, 1, 2, 3])
def test_iluk_symbolic_pattern_matches_the_definition(isph, fill):
    """Host pass behind ILU(k) (Ifpack 'fact: level-of-fill', precond_ifpack.h:38): pattern == dense level-of-fill pattern."""
    import numpy as np
    import scipy.sparse as sp
    rng = np.random.default_rng(5)
    e = np.ones(6); T = sp.diags([-e[:-1], 2 * e, -e[:-1]], [-1, 0, 1])
    mats = [sp.csr_matrix(sp.kron(sp.eye(6), T) + sp.kron(T, sp.eye(6))), sp.csr_matrix(sp.random(40, 40, 0.08, random_state=3) + sp.eye(40))]
    L = isph.lib()
    for A in mats:
        A.sort_indices(); n = A.shape[0]
        want = _dense_levels(A.toarray() != 0, fill)
        rp = np.ascontiguousarray(A.indptr, dtype=np.int32); ci = np.ascontiguousarray(A.indices, dtype=np.int32)
        nnz = ctypes.c_longlong()
        ip = ctypes.POINTER(ctypes.c_int)
        assert L.isph_iluk_symbolic_host(n, rp.ctypes.data_as(ip), ci.ctypes.data_as(ip), fill, None, None, ctypes.c_longlong(0), ctypes.byref(nnz)) == -1
        assert nnz.value == want.sum()
        rpo = np.zeros(n + 1, dtype=np.int32); cio = np.zeros(nnz.value, dtype=np.int32)
        assert L.isph_iluk_symbolic_host(n, rp.ctypes.data_as(ip), ci.ctypes.data_as(ip), fill, rpo.ctypes.data_as(ip), cio.ctypes.data_as(ip),
                                         ctypes.c_longlong(nnz.value), ctypes.byref(nnz)) == 0
        got = np.zeros((n, n), dtype=bool)
        for i in range(n):
            cols = cio[rpo[i]:rpo[i + 1]]; assert np.all(np.diff(cols) > 0); got[i, cols] = True
        assert np.array_equal(got, want)


def test_ilu1_symbolic_threaded_matches_the_sparse_product(isph):
    """Level-of-fill 1 on a matrix large enough for the threaded host pass: pattern(A) | pattern(strict_lower(A) @ strict_upper(A))."""
    import numpy as np
    import scipy.sparse as sp
    n1 = 72; e = np.ones(n1); T = sp.diags([-e[:-1], 2 * e, -e[:-1]], [-1, 0, 1])
    A = sp.csr_matrix(sp.kron(sp.eye(n1), T) + sp.kron(T, sp.eye(n1)) + sp.random(n1 * n1, n1 * n1, 3e-4, random_state=1)); A.sort_indices(); n = A.shape[0]
    assert n >= 4096
    B = sp.csr_matrix((np.ones(A.nnz), A.indices, A.indptr), shape=A.shape)
    want = sp.csr_matrix(((B + sp.tril(B, -1) @ sp.triu(B, 1)) != 0).astype(np.int8)); want.sort_indices()
    L = isph.lib(); ip = ctypes.POINTER(ctypes.c_int)
    rp = np.ascontiguousarray(A.indptr, dtype=np.int32); ci = np.ascontiguousarray(A.indices, dtype=np.int32); nnz = ctypes.c_longlong()
    L.isph_iluk_symbolic_host(n, rp.ctypes.data_as(ip), ci.ctypes.data_as(ip), 1, None, None, ctypes.c_longlong(0), ctypes.byref(nnz))
    assert nnz.value == want.nnz
    rpo = np.zeros(n + 1, dtype=np.int32); cio = np.zeros(nnz.value, dtype=np.int32)
    assert L.isph_iluk_symbolic_host(n, rp.ctypes.data_as(ip), ci.ctypes.data_as(ip), 1, rpo.ctypes.data_as(ip), cio.ctypes.data_as(ip), ctypes.c_longlong(nnz.value), ctypes.byref(nnz)) == 0
    assert np.array_equal(rpo, want.indptr) and np.array_equal(cio, want.indices)


def test_host_entry_points_edge_cases(isph):
    """Empty and degenerate inputs of the two pure-host entry points (no GPU needed)."""
    import numpy as np
    L = isph.lib(); ip = ctypes.POINTER(ctypes.c_int)
    z = np.zeros(1, dtype=np.int32); rc = np.zeros(2, dtype=np.int32); nh = ctypes.c_int(-1)
    # no ghosts at all: empty plan
    assert L.isph_halo_plan_host(2, 0, 5, 0, z.ctypes.data_as(ip), z.ctypes.data_as(ip), z.ctypes.data_as(ip), z.ctypes.data_as(ip),
                                 rc.ctypes.data_as(ip), z.ctypes.data_as(ip), ctypes.byref(nh)) == 0
    assert nh.value == 0 and rc.tolist() == [0, 0]
    # a ghost whose tag no rank owns: loud failure, not a silent column
    gt = np.array([7], dtype=np.int32); go = np.array([-1], dtype=np.int32); gi = np.array([-1], dtype=np.int32); gc = np.zeros(1, dtype=np.int32)
    assert L.isph_halo_plan_host(2, 0, 5, 1, gt.ctypes.data_as(ip), go.ctypes.data_as(ip), gi.ctypes.data_as(ip), gc.ctypes.data_as(ip),
                                 rc.ctypes.data_as(ip), z.ctypes.data_as(ip), ctypes.byref(nh)) == -1
    # two ghost copies of the same remote particle share one halo column
    gt = np.array([9, 9], dtype=np.int32); go = np.array([1, 1], dtype=np.int32); gi = np.array([3, 3], dtype=np.int32); gc = np.zeros(2, dtype=np.int32); rq = np.zeros(2, dtype=np.int32)
    assert L.isph_halo_plan_host(2, 0, 5, 2, gt.ctypes.data_as(ip), go.ctypes.data_as(ip), gi.ctypes.data_as(ip), gc.ctypes.data_as(ip),
                                 rc.ctypes.data_as(ip), rq.ctypes.data_as(ip), ctypes.byref(nh)) == 0
    assert nh.value == 1 and gc.tolist() == [5, 5] and rc.tolist() == [0, 1] and rq[0] == 3
    # level-of-fill pattern: a diagonal matrix stays diagonal at any level; bad arguments are refused
    n = 6; rp = np.arange(n + 1, dtype=np.int32); ci = np.arange(n, dtype=np.int32); nnz = ctypes.c_longlong()
    for fill in (0, 1, 3):
        rpo = np.zeros(n + 1, dtype=np.int32); cio = np.zeros(n, dtype=np.int32)
        assert L.isph_iluk_symbolic_host(n, rp.ctypes.data_as(ip), ci.ctypes.data_as(ip), fill, rpo.ctypes.data_as(ip), cio.ctypes.data_as(ip), ctypes.c_longlong(n), ctypes.byref(nnz)) == 0
        assert nnz.value == n and np.array_equal(rpo, rp) and np.array_equal(cio, ci)
    assert L.isph_iluk_symbolic_host(n, rp.ctypes.data_as(ip), ci.ctypes.data_as(ip), -1, None, None, ctypes.c_longlong(0), ctypes.byref(nnz)) == -1
    assert L.isph_iluk_symbolic_host(n, None, ci.ctypes.data_as(ip), 0, None, None, ctypes.c_longlong(0), ctypes.byref(nnz)) == -1
