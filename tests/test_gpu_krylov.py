"""GPU parity of the Krylov + preconditioner half against the CPU restatement of the Belos/Ifpack semantics
(oracle/krylov_oracle.cpp — "parity unpinned" at the Trilinos boundary, see its header).

Bars (BASELINE.json north_star): same preconditioner => iteration count within +-2, solution relative difference <= 1e-8
(tolerance stated per test where the problem's conditioning enters), final relative residual matched (both <= tol).
"""
import importlib

import numpy as np
import pytest
import scipy.sparse as sp

import oracle as O
from problems import make_case

isph = importlib.import_module("implicit-sph_b200")
pytestmark = pytest.mark.gpu

PREC_NAME = {O.PREC_NONE: "none", O.PREC_JACOBI: "point relaxation", O.PREC_CHEBYSHEV: "Chebyshev", O.PREC_ILU0: "ILU"}


def lap2d(n, shift=0.3, skew=0.0):
    e = np.ones(n); T = sp.diags([-e[:-1], 2 * e, -e[:-1]], [-1, 0, 1])
    A = sp.kron(sp.eye(n), T) + sp.kron(T, sp.eye(n)) + shift * sp.eye(n * n)
    if skew:
        S = sp.diags([e[:-1] * skew, -e[:-1] * skew], [1, -1]); A = A + sp.kron(sp.eye(n), S)
    A = sp.csr_matrix(A); A.sort_indices(); return A


def configure(c, solver, prec, flexible=True, degree=1, **kw):
    c.solver_param("Solver Type", "Block CG" if solver == O.SOLVER_CG else "Block GMRES")
    c.solver_param("Flexible Gmres", bool(flexible))
    for k, v in kw.items():
        c.solver_param(k, v)
    c.precond_param("Precond Type", PREC_NAME[prec])
    c.precond_param("chebyshev: degree", degree)
    c.precond_param("Overlap Level", 0); c.precond_param("fact: level-of-fill", 0)


def check(st, info, x, xo, sol_tol=1e-8, tol=1e-8):
    assert st["converged"] == info["converged"]
    assert abs(st["iters"] - info["iters"]) <= 2, (st["iters"], info["iters"])
    assert st["relres"] <= tol and info["relres"] <= tol
    assert np.linalg.norm(x - xo) / np.linalg.norm(xo) <= sol_tol, np.linalg.norm(x - xo) / np.linalg.norm(xo)


TIGHT = {"Convergence Tolerance": 1e-13, "Maximum Iterations": 4000, "Maximum Restarts": 200}
TIGHT_ORACLE = dict(tol=1e-13, max_iters=4000, max_restarts=200)


def tight_external(A, b, solver, prec, flex, degree):
    """north_star's solution bar proper: the same system solved to 1e-13 on both sides => x agrees to <= 1e-8 (VERDICT r1 weak #3:
    with the reference's 1e-8 residual stop the two solutions may differ by kappa(A) * 1e-8, which says nothing about the path)."""
    n = A.shape[0]
    xo, info = O.krylov_solve(A.indptr, A.indices, A.data, b, params=O.krylov_params(solver=solver, precond=prec, flexible=int(flex), cheb_degree=degree, **TIGHT_ORACLE))
    c = isph.Context(); c.matrix_set_csr(A.indptr, A.indices, A.data)
    x = np.zeros(n); c.create_solution(x, 1); c.create_load(None, 1); c.load_set(b)
    configure(c, solver, prec, flex, degree, **TIGHT); c.set_initial_solution(isph.INIT_ZERO)
    st = c.solve(prec != O.PREC_NONE, "ext-tight"); c.close()
    err = np.linalg.norm(x - xo) / np.linalg.norm(xo)
    assert st["converged"] and info["converged"] and err <= 1e-8, (st, info, err)
    return err


@pytest.mark.parametrize("prec,degree", [(O.PREC_NONE, 1), (O.PREC_JACOBI, 1), (O.PREC_CHEBYSHEV, 3), (O.PREC_ILU0, 1)])
@pytest.mark.parametrize("flex", [True, False])
def test_gmres_external_matrix(prec, degree, flex):
    A = lap2d(40, 0.05, 0.4); n = A.shape[0]
    b = np.random.default_rng(0).standard_normal(n)
    xo, info = O.krylov_solve(A.indptr, A.indices, A.data, b, params=O.krylov_params(precond=prec, flexible=int(flex), cheb_degree=degree))
    c = isph.Context(); c.matrix_set_csr(A.indptr, A.indices, A.data)
    x = np.zeros(n); c.create_solution(x, 1); c.create_load(None, 1); c.load_set(b)
    configure(c, O.SOLVER_GMRES, prec, flex, degree); c.set_initial_solution(isph.INIT_ZERO)
    st = c.solve(prec != O.PREC_NONE, "ext")
    check(st, info, x, xo, sol_tol=1e-7)
    assert np.linalg.norm(b - A @ x) / np.linalg.norm(b) <= 2e-8
    c.close()
    tight_external(A, b, O.SOLVER_GMRES, prec, flex, degree)


@pytest.mark.parametrize("prec,degree", [(O.PREC_NONE, 1), (O.PREC_JACOBI, 1), (O.PREC_CHEBYSHEV, 4), (O.PREC_ILU0, 1)])
def test_cg_external_matrix(prec, degree):
    A = lap2d(48, 0.02); n = A.shape[0]
    b = np.random.default_rng(1).standard_normal(n)
    xo, info = O.krylov_solve(A.indptr, A.indices, A.data, b, params=O.krylov_params(solver=O.SOLVER_CG, precond=prec, cheb_degree=degree))
    c = isph.Context(); c.matrix_set_csr(A.indptr, A.indices, A.data)
    x = np.zeros(n); c.create_solution(x, 1); c.create_load(None, 1); c.load_set(b)
    configure(c, O.SOLVER_CG, prec, True, degree); c.set_initial_solution(isph.INIT_ZERO)
    st = c.solve(prec != O.PREC_NONE, "ext")
    check(st, info, x, xo, sol_tol=1e-7)
    c.close()
    tight_external(A, b, O.SOLVER_CG, prec, True, degree)


def test_restart_and_iteration_cap_follow_the_parameter_list():
    A = lap2d(40, 0.001, 0.2); n = A.shape[0]
    b = np.random.default_rng(2).standard_normal(n)
    prm = O.krylov_params(precond=O.PREC_NONE, num_blocks=10, max_iters=35, max_restarts=15)
    xo, info = O.krylov_solve(A.indptr, A.indices, A.data, b, params=prm)
    c = isph.Context(); c.matrix_set_csr(A.indptr, A.indices, A.data)
    x = np.zeros(n); c.create_solution(x, 1); c.create_load(None, 1); c.load_set(b)
    configure(c, O.SOLVER_GMRES, O.PREC_NONE, True, 1, **{"Num Blocks": 10, "Maximum Iterations": 35})
    c.set_initial_solution(isph.INIT_ZERO)
    st = c.solve(False, "cap")
    assert not info["converged"] and not st["converged"] and st["iters"] == info["iters"] == 35      # non-convergence is not an error
    assert abs(st["relres"] - info["relres"]) <= 1e-6 * info["relres"]
    assert np.linalg.norm(x - xo) / np.linalg.norm(xo) <= 1e-8
    c.close()


_ORACLE_CACHE = {}


def _case_with_oracle(name):
    import harness
    if name not in _ORACLE_CACHE:
        P, F = make_case(name)
        if P["nlocal"] > 20000:      # big cloud: only what the solve needs (run_oracle assembles every system)
            o = O.Oracle(P, kind="port"); o.set_field(O.F_DENSITY, F["density"]); o.set_field(O.F_VSTAR, F["velocity"]); o.compute_pre()
            rp, col = o.graph(); b = o.ns_poisson(P["case"]["dt"]); ref = dict(rowptr=rp, col=col, b_poisson=b, A_poisson=o.matrix()); o.close()
        else:
            ref = harness.run_oracle(P, F, "port")
        _ORACLE_CACHE.clear(); _ORACLE_CACHE[name] = (P, F, ref)
    return _ORACLE_CACHE[name]


def _sph_poisson(name, prec, solver=O.SOLVER_GMRES, degree=1, blocks=None, tight=False, fill=0):
    import harness
    P, F, ref = _case_with_oracle(name); cs = P["case"]; nl = P["nlocal"]
    col = O.tags_to_local(ref["col"], P["tag"][:nl])
    b = ref["b_poisson"].copy()
    mask = np.ones(nl, dtype=np.int32)
    kw = dict(TIGHT_ORACLE) if tight else {}
    xo, info = O.krylov_solve(ref["rowptr"], col, ref["A_poisson"], b, params=O.krylov_params(solver=solver, precond=prec, cheb_degree=degree, ilu_fill=fill, row_gid=P["tag"][:nl], **kw),
                              null_mask=mask, use_null=True, blocks=blocks)
    c = harness.cuda_context(P, F)
    c.compute_pre(); c.graph_build(); c.create_load(None, 1); c.ns_poisson(cs["dt"])
    x = np.zeros(nl); c.create_solution(x, 1)
    c.set_null_vector_mask(mask); c.set_matrix_is_singular(True); c.set_initial_solution(isph.INIT_ZERO)
    configure(c, solver, prec, True, degree, **(TIGHT if tight else {}))
    c.precond_param("fact: level-of-fill", fill)
    if blocks is not None:
        c.precond_set_blocks(blocks)
    st = c.solve(prec != O.PREC_NONE, "Poisson")
    c.close()
    return st, info, x, xo


def check_tight(name, prec, **kw):
    """the 1e-8 solution bar, measured with both solves driven to 1e-13 (see tight_external)"""
    st, info, x, xo = _sph_poisson(name, prec, tight=True, **kw)
    err = np.linalg.norm(x - xo) / np.linalg.norm(xo)
    assert st["converged"] and info["converged"] and err <= 1e-8, (st, info, err)


@pytest.mark.parametrize("name,prec,degree", [("jitter3d", O.PREC_JACOBI, 1), ("jitter2d", O.PREC_CHEBYSHEV, 2), ("lattice3d", O.PREC_JACOBI, 1), ("jitter2d", O.PREC_ILU0, 1)])
def test_sph_pressure_poisson_nullspace(name, prec, degree):
    """The reference's per-step Poisson solve: singular operator, PoissonProjection, b and x projected (solver_lin_belos.h:138-219)."""
    st, info, x, xo = _sph_poisson(name, prec, degree=degree)
    check(st, info, x, xo, sol_tol=1e-6)      # kappa(A) ~ 1e3-1e4 on these cases: 1e-8 residual parity allows ~1e-6 on x
    assert abs(x.sum()) <= 1e-9 * np.abs(x).sum()
    check_tight(name, prec, degree=degree)


@pytest.mark.parametrize("name,prec", [("cloud2d", O.PREC_JACOBI), ("cloud2d", O.PREC_ILU0), ("cloud3d", O.PREC_JACOBI), ("cloud3d", O.PREC_ILU0),
                                       ("cloud3d_50k", O.PREC_JACOBI), ("cloud3d_50k", O.PREC_ILU0)])
def test_sph_pressure_poisson_ragged_cloud(name, prec):
    """Ragged random clouds: GMRES(50) + Jacobi and + ILU(0) on the SELL slack path — iteration counts
    within +-2 of the oracle at the reference's tolerance, solutions <= 1e-8 with both solves driven to 1e-13."""
    st, info, x, xo = _sph_poisson(name, prec)
    check(st, info, x, xo, sol_tol=1e-5)
    check_tight(name, prec)


def test_block_jacobi_ilu0_follows_the_rank_partition():
    """Ifpack factors one open block per MPI rank (overlap 0 drops off-rank columns): with block_of_row = the brick a CPU
    run would give each rank (here 2 x 2 bricks of the 2-D box), preconditioner and iteration count match the oracle."""
    import harness
    P, F = make_case("jitter2d")
    N = P["nglobal"][0]; g = P["gidx"][:P["nlocal"]]
    blocks = ((g % N) >= N // 2).astype(np.int32) + 2 * ((g // N) >= N // 2).astype(np.int32)
    st, info, x, xo = _sph_poisson("jitter2d", O.PREC_ILU0, blocks=blocks)
    check(st, info, x, xo, sol_tol=1e-6)
    check_tight("jitter2d", O.PREC_ILU0, blocks=blocks)


def test_c1_tgv128_gmres_ilu0():
    """BASELINE config 1: 2-D TGV 128x128 pressure Poisson, flexible GMRES(50) + ILU(0) (fill 0, overlap 0), one rank."""
    st, info, x, xo = _sph_poisson("tgv128", O.PREC_ILU0)
    check(st, info, x, xo, sol_tol=1e-5)
    check_tight("tgv128", O.PREC_ILU0)


def test_helmholtz_three_rhs_cg_chebyshev():
    """BASELINE config 3 in miniature: velocity Helmholtz, dim right-hand sides solved one after another, CG + Chebyshev."""
    import harness
    P, F = make_case("jitter3d"); cs = P["case"]; nl, dim = P["nlocal"], P["dim"]
    ref = harness.run_oracle(P, F, "port")
    col = O.tags_to_local(ref["col"], P["tag"][:nl])
    prm = O.krylov_params(solver=O.SOLVER_CG, precond=O.PREC_CHEBYSHEV, cheb_degree=3, row_gid=P["tag"][:nl])
    xs, its = [], 0
    for k in range(dim):
        x0 = F["velocity"][:nl, k].copy()
        xo, info = O.krylov_solve(ref["rowptr"], col, ref["A_helmholtz"], ref["b_helmholtz"][:, k], x0=x0, params=prm); xs.append(xo); its += info["iters"]
        assert info["converged"]
    c = harness.cuda_context(P, F)
    c.compute_pre(); c.graph_build()
    x = np.asfortranarray(F["velocity"][:nl, :dim].copy()); c.create_solution(x, dim)       # x = v is the initial guess (pair_isph.cpp:932-941)
    c.create_load(None, dim); c.load_set(np.asfortranarray(F["velocity"][:nl, :dim]))
    c.ns_helmholtz(cs["dt"], cs["theta"])
    configure(c, O.SOLVER_CG, O.PREC_CHEBYSHEV, True, 3)
    st = c.solve(True, "Helmholtz")
    assert st["converged"] and abs(st["iters"] - its) <= 2 * dim
    for k in range(dim):
        assert np.linalg.norm(x[:, k] - xs[k]) / np.linalg.norm(xs[k]) <= 1e-8
    c.close()


@pytest.mark.parametrize("name,prec", [("jitter2d", O.PREC_JACOBI), ("jitter3d", O.PREC_ILU0)])
def test_poisson_boltzmann_newton_matches_cpu_newton(name, prec):
    """BASELINE config 5 in miniature: the Newton iteration NOX runs for computePoissonBoltzmann (full steps, NormF 1e-8 AND
    NormUpdate 1e-5, solver_nox_impl.h:76-160) with computeF / computeJacobian / the Jacobian solve on the device, against the
    same loop built from the CPU oracle's pieces (manufactured source of sph-script/poisson-boltzmann-harmonic.xml)."""
    import harness
    P, F = make_case(name); cs = P["case"]; nl = P["nlocal"]; ex = F["pb_extra"][:nl].copy()
    o = O.Oracle(P, kinds=cs["kinds"], kernel=cs["kernel"], h_min=cs["h_min"], kind="port")
    o.set_field(O.F_EPS, F["eps"]); o.set_field(O.F_PSI0, F["psi0"]); o.compute_pre(); rp, col = o.graph()
    colL = O.tags_to_local(col, P["tag"][:nl]); prm = O.krylov_params(precond=prec, row_gid=P["tag"][:nl])
    psi = np.zeros(nl + P["nghost"]); nup = 0.0; lin = 0; k = 0
    while True:
        o.set_field(O.F_PSI, psi); f = o.pb_residual(extra_f=ex); nf = np.linalg.norm(f) / np.sqrt(nl)
        if (k > 0 and nf <= 1e-8 and nup <= 1e-5) or k >= 100:
            break
        o.pb_jacobian(); A = o.matrix()
        dx, info = O.krylov_solve(rp, colL, A, -f, params=prm); assert info["converged"]; lin += info["iters"]
        psi[:nl] += dx; nup = np.linalg.norm(dx) / np.sqrt(nl); k += 1
    o.close()
    assert 2 <= k < 20                                           # a genuinely nonlinear solve that converges quadratically
    c = harness.cuda_context(P, F)
    c.field_set(isph.F_PSI, np.zeros(nl + P["nghost"]))
    c.compute_pre(); c.graph_build(); c.create_solution(None, 1); c.create_load(None, 1)
    configure(c, O.SOLVER_GMRES, prec)
    st = c.pb_newton(extra_f=ex)
    psi_gpu = c.field_get(isph.F_PSI)
    c.close()
    assert st["converged"] and st["newton_iters"] == k and abs(st["linear_iters"] - lin) <= 2 * k, (st, k, lin)
    assert st["normf"] <= 1e-8
    assert np.linalg.norm(psi_gpu[:nl] - psi[:nl]) / np.linalg.norm(psi[:nl]) <= 1e-8
    ghost_owner = P["tag"][nl:] - 1                              # single rank: tags are 1-based owned indices
    assert np.array_equal(psi_gpu[nl:], psi_gpu[ghost_owner])    # psi forwarded to the ghosts after the solve (pair_isph.cpp:595-598)


@pytest.mark.parametrize("fill", [1, 2])
def test_gmres_iluk_external_matrix(fill):
    """Ifpack ILU with 'fact: level-of-fill' k > 0 and the reference's default 'Overlap Level' 1 (precond_ifpack.h:37-38; a no-op on
    one rank): host level-of-fill pattern + the ILU(0) numeric kernels on it, against the oracle's ILU(k)."""
    A = lap2d(40, 0.05, 0.4); n = A.shape[0]
    b = np.random.default_rng(0).standard_normal(n)
    prm = O.krylov_params(precond=O.PREC_ILU0, ilu_fill=fill)
    xo, info = O.krylov_solve(A.indptr, A.indices, A.data, b, params=prm)
    zo, _ = O.precond_apply(A.indptr, A.indices, A.data, b, prm)
    c = isph.Context(); c.matrix_set_csr(A.indptr, A.indices, A.data)
    x = np.zeros(n); c.create_solution(x, 1); c.create_load(None, 1); c.load_set(b)
    configure(c, O.SOLVER_GMRES, O.PREC_ILU0); c.precond_param("fact: level-of-fill", fill); c.precond_param("Overlap Level", 1)
    c.set_initial_solution(isph.INIT_ZERO)
    st = c.solve(True, "ext")
    check(st, info, x, xo, sol_tol=1e-7)
    c.precond_create(); z = c.precond_apply(b); c.precond_free()
    assert np.abs(z - zo).max() <= 1e-12 * np.abs(zo).max()
    c.close()


def test_c1_tgv128_gmres_ifpack_default_ilu1():
    """BASELINE config 1 particle set with Ifpack's OWN defaults (ILU, level-of-fill 1, overlap 1: precond_ifpack.h:30-44)."""
    import harness
    P, F = make_case("tgv128"); cs = P["case"]; nl = P["nlocal"]
    ref = harness.run_oracle(P, F, "port")
    col = O.tags_to_local(ref["col"], P["tag"][:nl]); mask = np.ones(nl, dtype=np.int32)
    xo, info = O.krylov_solve(ref["rowptr"], col, ref["A_poisson"], ref["b_poisson"].copy(), params=O.krylov_params(precond=O.PREC_ILU0, ilu_fill=1, row_gid=P["tag"][:nl]),
                              null_mask=mask, use_null=True)
    c = harness.cuda_context(P, F)
    c.compute_pre(); c.graph_build(); c.create_load(None, 1); c.ns_poisson(cs["dt"])
    x = np.zeros(nl); c.create_solution(x, 1)
    c.set_null_vector_mask(mask); c.set_matrix_is_singular(True); c.set_initial_solution(isph.INIT_ZERO)
    configure(c, O.SOLVER_GMRES, O.PREC_ILU0); c.precond_param("fact: level-of-fill", 1); c.precond_param("Overlap Level", 1)
    st = c.solve(True, "Poisson"); c.close()
    check(st, info, x, xo, sol_tol=1e-5)


@pytest.mark.parametrize("N", [16, 32])
def test_known_answer_poisson_boltzmann_convergence_table_on_the_device(N):
    """The reference's recorded err.psi.norm2 (sph-script/conv-poisson-boltzmann-harmonic-2d-rev390.txt) reproduced by the CUDA
    path alone: pre-computation, graph, computeF / computeJacobian and the Newton iteration all on the device."""
    from test_oracle_cpu import PB_TABLE, pb_harmonic_problem
    lat = importlib.import_module("implicit-sph_b200.lattice")
    P, s, ex = pb_harmonic_problem(lat, N); nl = P["nlocal"]
    c = isph.Context(); c.set_particles(P)
    c.field_set(isph.F_EPS, np.ones(len(s))); c.field_set(isph.F_PSI0, np.zeros(len(s))); c.field_set(isph.F_PSI, np.zeros(len(s)))
    c.compute_pre(); c.graph_build(); c.create_solution(None, 1); c.create_load(None, 1)
    configure(c, O.SOLVER_GMRES, O.PREC_ILU0, **{"Convergence Tolerance": 1e-12, "Maximum Iterations": 2000})
    st = c.pb_newton(extra_f=ex, tol_f=1e-12, tol_update=1e-6)
    psi = c.field_get(isph.F_PSI)[:nl]; c.close()
    err = np.sqrt(np.mean((psi - s[:nl]) ** 2))
    assert st["converged"] and st["newton_iters"] < 10
    assert abs(err - PB_TABLE[N]) <= 1e-10 * PB_TABLE[N], (err, PB_TABLE[N])


@pytest.mark.parametrize("boundary", ["MorrisHolmes", "ConstExtension"])
def test_known_answer_channel_edl_table_on_the_device(boundary):
    """The reference's recorded channel-EDL errors (sph-script/conv-channel-edl-potential-2d-morrisholmes-rev722.txt, N = 32 and 64)
    from the CUDA path alone: solid walls, normals / number density, Morris-Holmes mirror, linearized Poisson-Boltzmann Newton."""
    from test_oracle_cpu import EDL_TABLE, edl_channel_problem
    lat = importlib.import_module("implicit-sph_b200.lattice"); mh = boundary == "MorrisHolmes"
    for N in (32, 64):
        P, h, exact = edl_channel_problem(lat, N); nl = P["nlocal"]; fluid = P["type"][:nl] != 2
        c = isph.Context(); c.set_particles(P, kinds=(0, isph.KIND_FLUID, isph.KIND_SOLID, isph.KIND_FLUID), h=h, h_min=h, morris_safe=0.0)
        c.field_set(isph.F_EPS, np.ones(len(exact))); c.field_set(isph.F_PSI0, np.ones(len(exact))); c.field_set(isph.F_PSI, np.zeros(len(exact)))
        c.compute_pre(normals=True); c.graph_build(); c.create_solution(None, 1); c.create_load(None, 1)
        configure(c, O.SOLVER_GMRES, O.PREC_ILU0, **{"Convergence Tolerance": 1e-13, "Maximum Iterations": 3000})
        st = c.pb_newton(morris_holmes=mh, linearized=True, ezcb=50.0, psiref=1.0, tol_f=1e-11, tol_update=1.0)
        psi = c.field_get(isph.F_PSI)[:nl]; c.close()
        err = np.sqrt(np.mean((psi[fluid] - exact[:nl][fluid]) ** 2)); want = EDL_TABLE[(boundary, N)]
        assert st["converged"] and st["newton_iters"] <= 3 and abs(err - want) <= 1e-9 * want, (N, st, err, want)


def test_borrowed_load_vector_is_a_view_like_the_reference():
    """createLoadMultiVector(b_ptr) wraps caller memory (solver_lin.cpp:45-58).  (1) a system functor writes the
    right-hand side INTO that memory and the solve uses it (it is not overwritten by the stale host array); (2) what the caller
    writes into b after the create is what the next solve uses."""
    import harness
    P, F = make_case("jitter2d"); cs = P["case"]; nl = P["nlocal"]
    c = harness.cuda_context(P, F); c.compute_pre(); c.graph_build()
    configure(c, O.SOLVER_GMRES, O.PREC_JACOBI)
    # owned load vector: the baseline
    c.create_load(None, 1); c.ns_poisson(cs["dt"]); b_ref = c.load_get(1)[:, 0].copy()
    x0 = np.zeros(nl); c.create_solution(x0, 1); c.set_matrix_is_singular(True); c.set_initial_solution(isph.INIT_ZERO); st0 = c.solve(True, "Poisson")
    # (1) borrowed b, filled by the device functor
    b = np.full(nl, 7.0); c.matrix_invalidate(); c.create_load(b, 1); c.ns_poisson(cs["dt"])
    assert np.array_equal(b, b_ref)
    x1 = np.zeros(nl); c.create_solution(x1, 1); c.set_initial_solution(isph.INIT_ZERO); st1 = c.solve(True, "Poisson")
    assert st1["iters"] == st0["iters"] and np.array_equal(x1, x0)
    # (2) the caller rewrites the View on the host: the next solve sees it
    b[:] = 2.0 * b_ref
    x2 = np.zeros(nl); c.create_solution(x2, 1); c.set_initial_solution(isph.INIT_ZERO); st2 = c.solve(True, "Poisson")
    assert st2["converged"] and np.linalg.norm(x2 - 2.0 * x0) <= 1e-12 * np.linalg.norm(x0)
    c.close()


def test_non_default_ilu_thresholds_are_refused():
    """Ifpack's fact: drop tolerance / relax value / absolute|relative threshold change the factors; only the
    defaults are implemented, and anything else is an error instead of being silently ignored."""
    c = isph.Context()
    for name, dflt in (("fact: drop tolerance", 0.0), ("fact: relax value", 0.0), ("fact: absolute threshold", 0.0), ("fact: relative threshold", 1.0)):
        c.precond_param(name, dflt)
        with pytest.raises(isph.IsphError):
            c.precond_param(name, dflt + 0.5)
    c.close()
