"""GPU parity of the multilevel preconditioner that stands in for PrecondWrapper_ML (precond_ml.h:17-172; implicit-sph_b200/csrc/amg.cu)
against its sequential restatement (oracle/amg_oracle.h).  PARITY UNPINNED against ML itself by construction (ML is un-vendored,
un-pinned third-party code, and the algorithm is the data-parallel member of ML's option space, not its default): what these tests pin
is that the device builds the SAME hierarchy as the restatement (aggregates identical, level sizes identical, Galerkin operators and
V-cycle to rounding) and that the preconditioned solves agree (iteration counts +-2, solutions <= 1e-8 with both solves driven to 1e-13).
"""
import importlib

import numpy as np
import pytest
import scipy.sparse as sp

import oracle as O
from problems import make_case
from test_gpu_krylov import lap2d, check, TIGHT, TIGHT_ORACLE, _case_with_oracle

isph = importlib.import_module("implicit-sph_b200")
pytestmark = pytest.mark.gpu

ML = {"max levels": "amg_max_levels", "aggregation: threshold": "amg_threshold", "smoother: pre sweeps": "amg_pre", "smoother: post sweeps": "amg_post",
      "smoother: sweeps (coarse levels)": "amg_level_sweeps", "coarse: sweeps": "amg_coarse_sweeps", "coarse: max size": "amg_max_coarse",
      "smoother: Chebyshev alpha": "amg_alpha", "coarse correction scale": "amg_scale", "eigen-analysis: iterations": "amg_eig_iters",
      "smoother: Chebyshev alpha (coarse levels)": "amg_level_alpha", "coarse correction scale (coarse levels)": "amg_level_scale"}


def ml_configure(c, flexible=True, smoother="Chebyshev", **ml):
    c.solver_param("Solver Type", "Block GMRES"); c.solver_param("Flexible Gmres", bool(flexible))
    c.precond_param("Precond Package", "ML"); c.precond_param("smoother: type", smoother); c.precond_param("coarse: type", ml.pop("coarse: type", smoother))
    for k, v in ml.items():
        c.precond_param(k, v)


def oracle_params(smoother="Chebyshev", **ml):
    direct = ml.pop("coarse: type", "").startswith("Amesos")
    kw = {ML[k]: v for k, v in ml.items()}
    return dict(precond=O.PREC_AMG, amg_smoother=1 if smoother == "Jacobi" else 0, amg_coarse_direct=int(direct), **kw)


@pytest.mark.parametrize("smoother,ml", [("Chebyshev", {}), ("Jacobi", {"smoother: pre sweeps": 2, "smoother: post sweeps": 2}),
                                          ("Chebyshev", {"aggregation: threshold": 0.2, "coarse: max size": 20, "max levels": 4}),
                                          ("Chebyshev", {"coarse: type": "Amesos-KLU"}),
                                          ("Chebyshev", {"smoother: post sweeps": 2, "smoother: Chebyshev alpha": 10.0, "coarse correction scale": 2.0,
                                                         "smoother: Chebyshev alpha (coarse levels)": 4.0, "coarse correction scale (coarse levels)": 1.5})])      # the reference's default coarse solver (precond_ml.h:55): explicit inverse of the coarsest operator
@pytest.mark.parametrize("flex", [True, False])
def test_ml_standin_external_matrix(smoother, ml, flex):
    """second API client's situation (fix_qeq_reax hands over a CSR matrix): 5-point operator, 6400 rows -> three levels; flexible and
    standard right-preconditioned GMRES (the V-cycle is a fixed linear operator, so both are legitimate)"""
    A = lap2d(80, 0.002, 0.2); n = A.shape[0]; b = np.random.default_rng(0).standard_normal(n)
    ml = dict({"aggregation: threshold": 0.1}, **ml)
    okw = oracle_params(smoother, **dict(ml)); okw["flexible"] = int(flex)
    h = O.amg_hierarchy(A.indptr, A.indices, A.data, O.krylov_params(**okw))
    xo, info = O.krylov_solve(A.indptr, A.indices, A.data, b, params=O.krylov_params(**okw))
    xj, infoj = O.krylov_solve(A.indptr, A.indices, A.data, b, params=O.krylov_params(precond=O.PREC_JACOBI))
    c = isph.Context(); c.matrix_set_csr(A.indptr, A.indices, A.data)
    x = np.zeros(n); c.create_solution(x, 1); c.create_load(None, 1); c.load_set(b)
    ml_configure(c, flex, smoother, **dict(ml)); c.set_initial_solution(isph.INIT_ZERO)
    st = c.solve(True, "ext-ml"); hi = c.precond_ml_info(); agg = c.precond_ml_aggregates()
    assert hi["levels"] == h["levels"] >= 3 and list(hi["rows"]) == list(h["rows"]) and list(hi["nnz"][1:]) == list(h["nnz"][1:]), (hi, h)
    assert np.array_equal(agg, h["agg"])                                       # the same aggregates, numbered the same way
    if smoother == "Chebyshev":                                                # (no eigenvalue estimate is made for the Jacobi smoother)
        assert np.allclose(hi["lambda_max"], h["lmax"], rtol=1e-10)
    check(st, info, x, xo, sol_tol=1e-6)
    assert info["iters"] * 3 <= infoj["iters"]                                 # ... and it is a multilevel method: far fewer iterations than Jacobi
    # z = M^-1 r through the ABI against the restatement, entry by entry
    r = np.random.default_rng(1).standard_normal(n)
    c.precond_create(); z = c.precond_apply(r); c.precond_free()
    zo, _ = O.precond_apply(A.indptr, A.indices, A.data, r, O.krylov_params(**okw))
    assert np.abs(z - zo).max() <= 1e-11 * np.abs(zo).max()
    # the 1e-8 solution bar proper: both solves driven to 1e-13
    xo2, info2 = O.krylov_solve(A.indptr, A.indices, A.data, b, params=O.krylov_params(**okw, **TIGHT_ORACLE))
    for k, v in TIGHT.items():
        c.solver_param(k, v)
    x[:] = 0.0; c.set_initial_solution(isph.INIT_ZERO); st2 = c.solve(True, "ext-ml-tight"); c.close()
    assert st2["converged"] and info2["converged"] and np.linalg.norm(x - xo2) / np.linalg.norm(xo2) <= 1e-8


def _sph_ml(name, tight=False, blocks=None, **ml):
    import harness
    P, F, ref = _case_with_oracle(name); cs = P["case"]; nl = P["nlocal"]
    col = O.tags_to_local(ref["col"], P["tag"][:nl]); b = ref["b_poisson"].copy(); mask = np.ones(nl, dtype=np.int32)
    okw = oracle_params(**ml); okw.update(TIGHT_ORACLE if tight else {})
    prm = O.krylov_params(row_gid=P["tag"][:nl], **okw)
    xo, info = O.krylov_solve(ref["rowptr"], col, ref["A_poisson"], b, params=prm, null_mask=mask, use_null=True, blocks=blocks)
    h = O.amg_hierarchy(ref["rowptr"], col, ref["A_poisson"], prm, blocks=blocks)
    c = harness.cuda_context(P, F)
    c.compute_pre(); c.graph_build(); c.create_load(None, 1); c.ns_poisson(cs["dt"])
    x = np.zeros(nl); c.create_solution(x, 1)
    c.set_null_vector_mask(mask); c.set_matrix_is_singular(True); c.set_initial_solution(isph.INIT_ZERO)
    ml_configure(c, **ml)
    for k, v in (TIGHT if tight else {}).items():
        c.solver_param(k, v)
    if blocks is not None:
        c.precond_set_blocks(blocks)
    st = c.solve(True, "Poisson-ML"); hi = c.precond_ml_info(); agg = c.precond_ml_aggregates(); c.close()
    return st, info, x, xo, hi, h, agg


@pytest.mark.parametrize("name,ml", [("jitter3d", {}), ("lattice3d", {}), ("jitter2d", {"coarse: max size": 30}), ("cloud3d", {}), ("cloud3d_50k", {}), ("tgv128", {})])
def test_ml_standin_sph_pressure_poisson(name, ml):
    """The reference's per-step Poisson solve with its default preconditioner package (pair_isph.cpp:325-329): singular operator, null-space
    projection in the Krylov operator (solver_lin_belos.h:138-219), lattices (all ties between equal-strength neighbours are broken by the
    hashed tags, not by rounding) and ragged clouds."""
    st, info, x, xo, hi, h, agg = _sph_ml(name, **ml)
    assert hi["levels"] == h["levels"] >= 2 and list(hi["rows"]) == list(h["rows"]), (hi, h)
    assert np.array_equal(agg, h["agg"])
    check(st, info, x, xo, sol_tol=1e-5)
    st, info, x, xo, hi, h, agg = _sph_ml(name, tight=True, **ml)
    err = np.linalg.norm(x - xo) / np.linalg.norm(xo)
    assert st["converged"] and info["converged"] and err <= 1e-8, (st, info, err)


def test_ml_standin_aggregates_stay_inside_a_block():
    """'Uncoupled' aggregation: aggregates never cross an MPI rank.  With block_of_row = the bricks a 4-rank CPU run would own, no aggregate
    holds rows of two bricks, on the device and in the restatement alike (the multi-GPU form of this is tests/multi_gpu_check.py)."""
    P, F = make_case("jitter2d"); N = P["nglobal"][0]; g = P["gidx"][:P["nlocal"]]
    blocks = ((g % N) >= N // 2).astype(np.int32) + 2 * ((g // N) >= N // 2).astype(np.int32)
    st, info, x, xo, hi, h, agg = _sph_ml("jitter2d", blocks=blocks, **{"coarse: max size": 30})
    assert np.array_equal(agg, h["agg"]) and abs(st["iters"] - info["iters"]) <= 2
    for a in np.unique(agg[agg >= 0]):
        assert len(np.unique(blocks[agg == a])) == 1


def test_ml_parameter_list_is_checked():
    """precond_ml.h:44-58 sets symmetric Gauss-Seidel: refused by name (not silently replaced); Amesos-KLU is provided (direct coarse solve)"""
    A = lap2d(20, 0.1); c = isph.Context(); c.matrix_set_csr(A.indptr, A.indices, A.data)
    x = np.zeros(A.shape[0]); c.create_solution(x, 1); c.create_load(None, 1); c.load_set(np.ones(A.shape[0]))
    c.precond_param("Precond Package", "ML"); c.precond_param("smoother: type", "symmetric Gauss-Seidel")
    with pytest.raises(isph.IsphError, match="smoother: type"):
        c.solve(True, "x")
    c.precond_param("smoother: type", "Chebyshev"); c.precond_param("coarse: type", "Gauss-Seidel")
    with pytest.raises(isph.IsphError, match="coarse: type"):
        c.solve(True, "x")
    c.precond_param("coarse: type", "Chebyshev"); c.precond_param("aggregation: damping factor", 1.333)
    with pytest.raises(isph.IsphError, match="damping factor"):
        c.solve(True, "x")
    with pytest.raises(isph.IsphError):
        c.precond_param("Precond Package", "Hypre")
    c.close()


# ---- solveBlockProblem (solver_lin_belos.h:53-128) through the ABI ----------------------------------------------------------------------
@pytest.mark.parametrize("prec,name", [(O.PREC_JACOBI, "point relaxation"), (O.PREC_ILU0, "ILU"), (O.PREC_AMG, "ML"), (None, None)])
def test_solve_block_problem_matches_the_stacked_oracle(prec, name):
    """The 3 x 3 block operator of SolverLin::setBlock (Thyra blocked operator over one nodal map) as one stacked matrix, x / b as n x 3
    multivectors, ONE preconditioner built from the scalar matrix and applied to every diagonal block — against the oracle's stacked solve."""
    from test_krylov_oracle_cpu import block_system
    dim = 3; blocks, S, A0 = block_system(); n = A0.shape[0]; rng = np.random.default_rng(0)
    b = np.asfortranarray(rng.standard_normal((n, dim)))
    prm = O.krylov_params(precond=prec if prec is not None else O.PREC_NONE, amg_threshold=0.1, amg_max_coarse=20)
    xo, info = O.krylov_solve_block(dim, S, A0, b.reshape(-1, order="F"), params=prm)
    c = isph.Context(); c.matrix_set_csr(A0.indptr, A0.indices, A0.data)               # li_solver->setMatrix(A.crs); prec->setMatrix(A.crs)
    x = np.asfortranarray(np.zeros((n, dim))); c.create_solution(x, dim); c.create_load(None, dim); c.load_set(b)
    c.solver_param("Solver Type", "Block GMRES"); c.solver_param("Flexible Gmres", True)
    if prec is not None:
        if name == "ML":
            ml_configure(c, **{"aggregation: threshold": 0.1, "coarse: max size": 20})
        else:
            c.precond_param("Precond Package", "Ifpack"); c.precond_param("Precond Type", name); c.precond_param("Overlap Level", 0); c.precond_param("fact: level-of-fill", 0)
    c.block_matrix(dim, blocks, "Block 3x3"); c.set_matrix_is_singular(False); c.set_initial_solution(isph.INIT_ZERO)
    st = c.solve_block(prec is not None, "Block 3x3 Helmholtz")
    assert st["converged"] and info["converged"] and abs(st["iters"] - info["iters"]) <= 2, (st, info)
    assert np.linalg.norm(x.reshape(-1, order="F") - xo) / np.linalg.norm(xo) <= 1e-7
    assert np.linalg.norm(b.reshape(-1, order="F") - S @ x.reshape(-1, order="F")) / np.linalg.norm(b) <= 2e-8
    # error behaviour of the reference (solver_lin_belos.h:58-61): rhs columns != dim, singular problems
    c.set_matrix_is_singular(True)
    with pytest.raises(isph.IsphError, match="singular"):
        c.solve_block(False, "x")
    c.set_matrix_is_singular(False); c.create_load(None, 1); c.create_solution(None, 1)
    with pytest.raises(isph.IsphError, match="dimension of rhs"):
        c.solve_block(False, "x")
    c.close()


@pytest.mark.parametrize("n,m,reach_min,cols_min", [(600, 40, 48, 0), (4000, 75, 48, 384)])
def test_ml_standin_rows_that_reach_many_aggregates(n, m, reach_min, cols_min):
    """small aggregates under a wide stencil: only the two ring neighbours of a row are strong (aggregates of 3-5 rows) while every row has
    2m weak entries spread over the whole matrix, so a row reaches > 48 aggregates — the wide variant of the row compression takes over —
    and (second case) a coarse row holds > 384 columns — the wide variant of the merge kernel; hierarchy and iteration count still equal
    the restatement's"""
    rng = np.random.default_rng(7)
    r = np.repeat(np.arange(n), m); cidx = rng.integers(0, n, size=len(r)); keep = r != cidx
    W = sp.coo_matrix((-np.ones(keep.sum()), (r[keep], cidx[keep])), shape=(n, n)).tocsr(); W.data[:] = -1.0; W = W + W.T; W.data[:] = -1.0
    i = np.arange(n); R = sp.coo_matrix((-5.0 * np.ones(2 * n), (np.r_[i, i], np.r_[(i + 1) % n, (i - 1) % n])), shape=(n, n)).tocsr()
    off = sp.csr_matrix(W + R); A = sp.csr_matrix(off + sp.diags(-np.asarray(off.sum(1)).ravel() + 0.5)); A.sum_duplicates(); A.sort_indices()
    ml = {"aggregation: threshold": 0.03, "coarse: max size": 40}; okw = oracle_params(**ml)
    h = O.amg_hierarchy(A.indptr, A.indices, A.data, O.krylov_params(**okw), cap_rows=n, cap_nnz=n * n // 4)
    reach = max(len(set(h["agg"][A.indices[A.indptr[q]:A.indptr[q + 1]]])) for q in range(n)); assert reach > reach_min, reach
    assert np.diff(h["coarse"][0]).max() > cols_min
    b = rng.standard_normal(n); xo, info = O.krylov_solve(A.indptr, A.indices, A.data, b, params=O.krylov_params(**okw))
    c = isph.Context(); c.matrix_set_csr(A.indptr, A.indices, A.data)
    x = np.zeros(n); c.create_solution(x, 1); c.create_load(None, 1); c.load_set(b)
    ml_configure(c, **ml); c.set_initial_solution(isph.INIT_ZERO)
    st = c.solve(True, "wide"); hi = c.precond_ml_info(); agg = c.precond_ml_aggregates(); c.close()
    assert hi["levels"] == h["levels"] >= 2 and list(hi["rows"]) == list(h["rows"]) and np.array_equal(agg, h["agg"]), (hi, h)
    check(st, info, x, xo, sol_tol=1e-6)


@pytest.mark.parametrize("N", [32, 64])
def test_known_answer_poisson_boltzmann_table_with_the_ml_standin(N):
    """The reference's recorded err.psi.norm2 (sph-script/conv-poisson-boltzmann-harmonic-2d-rev390.txt) was produced by a run with the
    reference's default preconditioner package — ML — under GMRES.  The same table row from the CUDA path with the multilevel stand-in in the
    Jacobian solves: a different preconditioner must not move the converged answer (and the Newton iteration count stays what it is with ILU)."""
    from test_oracle_cpu import PB_TABLE, pb_harmonic_problem
    lat = importlib.import_module("implicit-sph_b200.lattice")
    P, s, ex = pb_harmonic_problem(lat, N); nl = P["nlocal"]
    c = isph.Context(); c.set_particles(P)
    c.field_set(isph.F_EPS, np.ones(len(s))); c.field_set(isph.F_PSI0, np.zeros(len(s))); c.field_set(isph.F_PSI, np.zeros(len(s)))
    c.compute_pre(); c.graph_build(); c.create_solution(None, 1); c.create_load(None, 1)
    ml_configure(c, **{"coarse: max size": 40}); c.solver_param("Convergence Tolerance", 1e-12); c.solver_param("Maximum Iterations", 2000)
    st = c.pb_newton(extra_f=ex, tol_f=1e-12, tol_update=1e-6); h = c.precond_ml_info()
    psi = c.field_get(isph.F_PSI)[:nl]; c.close()
    err = np.sqrt(np.mean((psi - s[:nl]) ** 2))
    assert st["converged"] and st["newton_iters"] < 10 and h["levels"] >= 2, (st, h)
    assert abs(err - PB_TABLE[N]) <= 1e-10 * PB_TABLE[N], (err, PB_TABLE[N])


def test_reference_benchmark_ml_parameter_file_chebyshev_variant():
    """The parameter list of the reference's own scaling benchmark (bench-script/hopper/tgv/4096/ml.xml) with the smoother switched to the
    Chebyshev lines that file keeps commented (:15-18) is accepted key by key and solves the singular pressure Poisson problem (coarse: type
    Amesos-KLU replaced by the smoother, as PrecondWrapper_ML::setNullVector does) with the iteration count of the restatement."""
    import harness
    ml_xml = [("ML output", 10), ("max levels", 10), ("increasing or decreasing", "increasing"), ("aggregation: type", "Uncoupled"),
              ("smoother: pre or post", "both"), ("smoother: type", "Chebyshev"), ("smoother: sweeps", 4), ("coarse: type", "Amesos-KLU")]
    P, F, ref = _case_with_oracle("jitter3d"); cs = P["case"]; nl = P["nlocal"]
    col = O.tags_to_local(ref["col"], P["tag"][:nl]); b = ref["b_poisson"].copy(); mask = np.ones(nl, dtype=np.int32)
    prm = O.krylov_params(precond=O.PREC_AMG, amg_max_levels=10, amg_pre=4, amg_post=4, amg_coarse_direct=0, row_gid=P["tag"][:nl])
    xo, info = O.krylov_solve(ref["rowptr"], col, ref["A_poisson"], b, params=prm, null_mask=mask, use_null=True)
    c = harness.cuda_context(P, F)
    c.compute_pre(); c.graph_build(); c.create_load(None, 1); c.ns_poisson(cs["dt"])
    x = np.zeros(nl); c.create_solution(x, 1)
    c.set_null_vector_mask(mask); c.set_matrix_is_singular(True); c.set_initial_solution(isph.INIT_ZERO)
    c.solver_param("Solver Type", "Block GMRES"); c.precond_param("Precond Package", "ML")
    for k, v in ml_xml:
        c.precond_param(k, v)
    st = c.solve(True, "Poisson"); c.close()
    check(st, info, x, xo, sol_tol=1e-5)
    with pytest.raises(isph.IsphError):                                        # ... and the Gauss-Seidel lines of the same file are refused by name
        c2 = isph.Context(); c2.precond_param("smoother: type", "ML Gauss-Seidel"); A = lap2d(10, 0.1); c2.matrix_set_csr(A.indptr, A.indices, A.data)
        c2.precond_param("Precond Package", "ML"); xx = np.zeros(100); c2.create_solution(xx, 1); c2.create_load(None, 1); c2.load_set(np.ones(100)); c2.solve(True, "x")


def test_ml_standin_under_block_cg_needs_a_symmetric_cycle():
    """Block CG (solver_lin_belos.h:181-182) with the multilevel preconditioner: a V-cycle with equal pre- and post-smoothing is a symmetric
    operator (Chebyshev polynomial in D^-1 A on both sides of a Galerkin correction) and PCG matches the restatement; the default asymmetric
    cycle (1 + 2 sweeps, chosen for flexible GMRES) is refused under CG instead of silently breaking the recurrence."""
    A = lap2d(64, 0.001); n = A.shape[0]; b = np.random.default_rng(4).standard_normal(n)
    ml = {"aggregation: threshold": 0.1, "smoother: pre sweeps": 2, "smoother: post sweeps": 2, "coarse correction scale": 1.0}
    xo, info = O.krylov_solve(A.indptr, A.indices, A.data, b, params=O.krylov_params(solver=O.SOLVER_CG, **oracle_params(**ml)))
    c = isph.Context(); c.matrix_set_csr(A.indptr, A.indices, A.data)
    x = np.zeros(n); c.create_solution(x, 1); c.create_load(None, 1); c.load_set(b)
    ml_configure(c, **ml); c.solver_param("Solver Type", "Block CG"); c.set_initial_solution(isph.INIT_ZERO)
    st = c.solve(True, "cg-ml")
    check(st, info, x, xo, sol_tol=1e-6)
    c.precond_param("smoother: post sweeps", 3); c.set_initial_solution(isph.INIT_ZERO)
    with pytest.raises(isph.IsphError, match="symmetric preconditioner"):
        c.solve(True, "cg-ml")
    c.close()


@pytest.mark.parametrize("seed", range(10))
def test_ml_standin_random_geometric_operators(seed):
    """seeded sweep over random particle-like operators (kernel-weighted graph Laplacians of random point clouds in 2-D / 3-D with random
    positive shifts, some identity rows, varying thresholds and coarse sizes): the device hierarchy equals the restatement's (aggregates,
    level sizes) and the solves agree — corner cases of the aggregation (isolated rows, leftovers, tiny aggregates, early stop) included."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(1000 + seed); dim = 2 + seed % 2; n = int(rng.integers(300, 3000))
    X = rng.random((n, dim)); rad = (6.0 if dim == 2 else 9.0) ** (1.0 / dim) * (1.0 / n) ** (1.0 / dim) * (1.3 + 0.5 * rng.random())
    pairs = cKDTree(X).query_pairs(rad, output_type="ndarray"); d = np.linalg.norm(X[pairs[:, 0]] - X[pairs[:, 1]], axis=1)
    w = (1.0 - d / rad) ** 3 + 1e-3
    W = sp.coo_matrix((np.r_[w, w], (np.r_[pairs[:, 0], pairs[:, 1]], np.r_[pairs[:, 1], pairs[:, 0]])), shape=(n, n)).tocsr()
    A = sp.lil_matrix(sp.diags(np.asarray(W.sum(1)).ravel() * (1.0 + 0.01 * rng.random(n)) + 1e-3) - W)
    for r in rng.choice(n, size=n // 50, replace=False):                       # identity rows (solid particles): no strong connection
        A.rows[r] = [int(r)]; A.data[r] = [1.0]
    A = sp.csr_matrix(A); A.sort_indices()
    ml = {"aggregation: threshold": float(rng.choice([0.0, 0.02, 0.1, 0.25])), "coarse: max size": int(rng.choice([10, 40, 128])), "max levels": int(rng.choice([2, 3, 5]))}
    okw = oracle_params(**dict(ml)); b = rng.standard_normal(n)
    h = O.amg_hierarchy(A.indptr, A.indices, A.data, O.krylov_params(**okw)); xo, info = O.krylov_solve(A.indptr, A.indices, A.data, b, params=O.krylov_params(**okw))
    c = isph.Context(); c.matrix_set_csr(A.indptr, A.indices, A.data)
    x = np.zeros(n); c.create_solution(x, 1); c.create_load(None, 1); c.load_set(b)
    ml_configure(c, **dict(ml)); c.set_initial_solution(isph.INIT_ZERO)
    st = c.solve(True, "fuzz"); hi = c.precond_ml_info(); agg = c.precond_ml_aggregates(); c.close()
    assert hi["levels"] == h["levels"] and list(hi["rows"]) == list(h["rows"]) and list(hi["nnz"][1:]) == list(h["nnz"][1:]), (ml, hi, h)
    assert np.array_equal(agg, h["agg"]), ml
    assert st["converged"] == info["converged"] and abs(st["iters"] - info["iters"]) <= 2, (ml, st, info)
    if info["converged"]:
        assert np.linalg.norm(x - xo) / np.linalg.norm(xo) <= 1e-5


def test_ml_standin_device_matches_the_golden_fixture():
    """the committed fixture of the restatement (tests/golden/amg/amg_restatement.npz) reproduced by the device path alone, without running the
    oracle: aggregates and level sizes exactly, one V-cycle to 1e-11, the iteration count exactly"""
    import importlib.util, os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden_amg", os.path.join(here, "golden", "amg", "make_golden_amg.py")); m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    A, b, r = m.problem(); want = np.load(os.path.join(here, "golden", "amg", "amg_restatement.npz")); n = A.shape[0]
    c = isph.Context(); c.matrix_set_csr(A.indptr, A.indices, A.data)
    x = np.zeros(n); c.create_solution(x, 1); c.create_load(None, 1); c.load_set(b)
    ml_configure(c, **{"aggregation: threshold": 0.1, "coarse: max size": 20}); c.set_initial_solution(isph.INIT_ZERO)
    st = c.solve(True, "golden"); hi = c.precond_ml_info(); agg = c.precond_ml_aggregates()
    c.precond_create(); z = c.precond_apply(r); c.precond_free(); c.close()
    assert np.array_equal(agg, want["agg"]) and list(hi["rows"]) == list(want["rows"]) and list(hi["nnz"][1:]) == list(want["nnz"][1:])
    assert st["iters"] == int(want["iters"]) and np.abs(z - want["z"]).max() <= 1e-11 * np.abs(want["z"]).max()
    assert np.linalg.norm(x - want["x"]) <= 1e-6 * np.linalg.norm(want["x"])


def test_ml_standin_with_one_level_is_the_coarse_solver():
    """'max levels' 1: the hierarchy is the finest level alone and the preconditioner is the coarse solver on it ('coarse: sweeps' Chebyshev
    steps over the 'coarse: Chebyshev alpha' interval) — the same on the device and in the restatement"""
    A = lap2d(40, 0.05, 0.2); n = A.shape[0]; b = np.random.default_rng(2).standard_normal(n)
    ml = {"max levels": 1}; okw = oracle_params(**ml)
    xo, info = O.krylov_solve(A.indptr, A.indices, A.data, b, params=O.krylov_params(**okw))
    c = isph.Context(); c.matrix_set_csr(A.indptr, A.indices, A.data)
    x = np.zeros(n); c.create_solution(x, 1); c.create_load(None, 1); c.load_set(b)
    ml_configure(c, **ml); c.set_initial_solution(isph.INIT_ZERO)
    st = c.solve(True, "one-level"); hi = c.precond_ml_info()
    r = np.random.default_rng(3).standard_normal(n); c.precond_create(); z = c.precond_apply(r); c.precond_free(); c.close()
    zo, _ = O.precond_apply(A.indptr, A.indices, A.data, r, O.krylov_params(**okw))
    assert hi["levels"] == 1 and np.abs(z - zo).max() <= 1e-12 * np.abs(zo).max()
    check(st, info, x, xo, sol_tol=1e-7)
