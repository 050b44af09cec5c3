"""Multi-GPU parity check of the NCCL path (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py

Every rank owns one brick of a periodic jittered lattice; graph + assembly + three solves run distributed (halo import and
all-reduces over NVLink peer memory, or NCCL with ISPH_NO_P2P=1) and are compared on rank 0 with the CPU oracle on the
GLOBAL problem: graph bit-exact, values <= 1e-12, iteration counts +-2, solutions <= 1e-6 (kappa ~ 1e3).
  1. pressure Poisson, NullSpace, flexible GMRES(50) + Jacobi                       (BASELINE configs[1])
  2. the same system with block-Jacobi ILU(0), one open block per rank (= Ifpack overlap 0 on an MPI run)   (configs[3])
  2b. ILU(0) with Overlap Level 1 and combine mode Add — Ifpack's defaults as the reference sets them (precond_ifpack.h:35-43)
  2c. the multilevel stand-in for ML, the reference's default package (aggregates inside each rank, coarse levels replicated)
  3. velocity Helmholtz, 3 right-hand sides one after another, CG + Chebyshev(2)     (configs[2]; SpMM with a 3-vector import)
  4. Poisson-Boltzmann Newton iteration (computeF / computeJacobian / GMRES + Jacobi Jacobian solves)      (configs[4])
tests/test_gpu_multi.py wraps this for pytest when >= 2 GPUs are visible.
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist
    isph = importlib.import_module("implicit-sph_b200"); lat = importlib.import_module("implicit-sph_b200.lattice")
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(lr)
    dist.init_process_group("gloo", init_method="env://")
    idt = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (isph.C.c_ubyte * 128)(); assert isph.lib().isph_nccl_unique_id(buf) == 0
        idt = torch.tensor(list(buf), dtype=torch.uint8)
    dist.broadcast(idt, 0)
    res = run_check(isph, lat, torch, dist, rank, world, lr, bytes(idt.tolist()))
    dist.destroy_process_group()
    sys.exit(0 if res["pass"] else 1)


def run_check(isph, lat, torch, dist, rank, world, lr, nccl_id, quiet=False):
    """The comparison itself (collective; every rank calls it).  Returns {"pass": bool, ...measured errors...} on every rank.
    bench.py calls this before its timed region at n_gpus > 1 and prints the result as the line's "parity" block."""
    say = (lambda *a: None) if quiet else print
    dim = 3; grid = lat.brick_grid(world, dim); per = (10, 8, 8); nglobal = tuple(per[k] * grid[k] for k in range(dim))
    lo, nloc = lat.brick_of_rank(rank, grid, nglobal)
    dx = 2 * np.pi / nglobal[0]
    P = lat.make_brick(dim, nglobal, dx, lo=lo, nloc=nloc, rs2=12, jitter=0.04)
    nl = P["nlocal"]; xw = P["xw"]
    v = lat.tgv_velocity(xw)
    for k in range(dim):
        v[:, k] += 0.05 * (2.0 * lat._hash01(P["gidx"] + 1, 100 + k) - 1.0)
    c = isph.Context(lr, world, rank, nccl_id)
    c.set_particles(P)
    nu = 0.1 + 0.01 * np.cos(xw[:, 1]); pr = np.sin(xw[:, 0]) * np.cos(xw[:, 1])
    c.field_set(isph.F_VSTAR, v); c.field_set(isph.F_VELOCITY, v); c.field_set(isph.F_VISCOSITY, nu); c.field_set(isph.F_PRESSURE, pr)
    c.compute_pre(); c.graph_build()
    c.create_load(None, 1); dt = 0.05; c.ns_poisson(dt)
    rp, col = c.graph_get(); A = c.matrix_get(); b = c.load_get(1)[:, 0]; vf = c.field_get(isph.F_VFRAC)[:nl]
    x = np.zeros(nl); c.create_solution(x, 1)
    c.set_matrix_is_singular(True); c.set_initial_solution(isph.INIT_ZERO); c.precond_param("Precond Type", "point relaxation")
    st = c.solve(True, "Poisson")
    # 2. block-Jacobi ILU(0): default block = this rank's rows (off-rank columns dropped)
    x2 = np.zeros(nl); c.create_solution(x2, 1); c.set_initial_solution(isph.INIT_ZERO)
    c.precond_param("Precond Type", "ILU"); c.precond_param("Overlap Level", 0); c.precond_param("fact: level-of-fill", 0)
    st2 = c.solve(True, "PoissonILU")
    # 2b. the reference's own Ifpack default across ranks: ILU with Overlap Level 1, combine mode Add (precond_ifpack.h:35-43)
    x2b = np.zeros(nl); c.create_solution(x2b, 1); c.set_initial_solution(isph.INIT_ZERO)
    c.precond_param("Overlap Level", 1)
    st2b = c.solve(True, "PoissonILUoverlap1")
    c.precond_param("Overlap Level", 0)
    # 2c. the reference's DEFAULT preconditioner package (ML, pair_isph.cpp:325-329): the multilevel stand-in, aggregates inside each rank
    x2c = np.zeros(nl); c.create_solution(x2c, 1); c.set_initial_solution(isph.INIT_ZERO)
    c.precond_param("Precond Package", "ML"); c.precond_param("coarse: max size", 20)
    st2c = c.solve(True, "PoissonML"); agg2c = c.precond_ml_aggregates(); ml2c = c.precond_ml_info()
    c.precond_param("Precond Package", "Ifpack")
    # 3. Helmholtz, dim right-hand sides, CG + Chebyshev(2)
    theta = 0.5
    c.matrix_invalidate(); c.create_load(None, dim); c.load_set(np.asfortranarray(v[:nl, :dim])); c.ns_helmholtz(dt, theta)
    Ah = c.matrix_get(); bh = c.load_get(dim)
    x3 = np.asfortranarray(np.zeros((nl, dim))); c.create_solution(x3, dim); c.set_matrix_is_singular(False); c.set_initial_solution(isph.INIT_ZERO)
    c.solver_param("Solver Type", "Block CG"); c.precond_param("Precond Type", "Chebyshev"); c.precond_param("chebyshev: degree", 2)
    st3 = c.solve(True, "Helmholtz")
    # 4. Poisson-Boltzmann: Newton iteration from psi = 0 with the manufactured source
    s_ = np.sin(xw[:, 0]) * np.cos(xw[:, 1]); ex = (-2.0 * s_ - np.sinh(s_))[:nl].copy()
    c.matrix_invalidate(); c.field_set(isph.F_EPS, 1.0 + 0.2 * np.cos(xw[:, 0])); c.field_set(isph.F_PSI0, 0.3 + 0.0 * s_); c.field_set(isph.F_PSI, 0.0 * s_)
    c.create_solution(None, 1); c.create_load(None, 1); c.set_matrix_is_singular(False)
    c.solver_param("Solver Type", "Block GMRES"); c.precond_param("Precond Type", "point relaxation")
    st4 = c.pb_newton(extra_f=ex); psi4 = c.field_get(isph.F_PSI)[:nl].copy()
    mine = dict(tag=P["tag"][:nl].copy(), rp=rp, col=col, A=A, b=b, x=x, vf=vf, st=st, x2=x2, st2=st2, x2b=x2b, st2b=st2b, x2c=x2c, st2c=st2c, agg2c=agg2c, ml2c=ml2c, Ah=Ah, bh=bh, x3=x3, st3=st3, st4=st4, psi4=psi4)
    allr = [None] * world
    dist.gather_object(mine, allr if rank == 0 else None, 0)
    ok = True; rep = {}
    if rank == 0:
        import oracle as O
        from problems import relerr
        G = lat.make_brick(dim, nglobal, dx, rs2=12, jitter=0.04)
        vg = lat.tgv_velocity(G["xw"])
        for k in range(dim):
            vg[:, k] += 0.05 * (2.0 * lat._hash01(G["gidx"] + 1, 100 + k) - 1.0)
        xg = G["xw"]
        o = O.Oracle(G, kind="port"); o.set_field(O.F_VSTAR, vg); o.set_field(O.F_VELOCITY, vg)
        o.set_field(O.F_VISCOSITY, 0.1 + 0.01 * np.cos(xg[:, 1])); o.set_field(O.F_PRESSURE, np.sin(xg[:, 0]) * np.cos(xg[:, 1])); o.compute_pre(); grp, gcol = o.graph(); gb = o.ns_poisson(dt); gA = o.matrix(); gvf = o.get_field(O.F_VFRAC)
        n = G["nlocal"]
        xo, info = O.krylov_solve(grp, O.tags_to_local(gcol, G["tag"][:n]), gA, gb.copy(), params=O.krylov_params(precond=O.PREC_JACOBI), null_mask=np.ones(n, dtype=np.int32), use_null=True)
        row_of_tag = -np.ones(n + 2, dtype=np.int64); row_of_tag[G["tag"][:n]] = np.arange(n)
        xd = np.zeros(n); worst = 0.0
        for d in allr:
            for li, t in enumerate(d["tag"]):
                gi = row_of_tag[t]; sl = slice(grp[gi], grp[gi + 1]); ll = slice(d["rp"][li], d["rp"][li + 1])
                assert np.array_equal(gcol[sl], d["col"][ll]), "graph differs"
                worst = max(worst, relerr(gA[sl], d["A"][ll]))
            gi = row_of_tag[d["tag"]]
            worst = max(worst, relerr(gvf[gi], d["vf"]), float(np.abs(gb[gi] - d["b"]).max() / np.abs(gb).max()))
            xd[gi] = d["x"]
        its = allr[0]["st"]["iters"]
        xerr = np.linalg.norm(xd - xo) / np.linalg.norm(xo)
        say(f"multi_gpu_check world={world} rows={n}: values/b/vfrac max err {worst:.2e}; iters gpu {its} vs oracle {info['iters']}; x rel diff {xerr:.2e}; converged {allr[0]['st']['converged']}")
        ok = worst <= 1e-12 and abs(its - info["iters"]) <= 2 and xerr <= 1e-6 and allr[0]["st"]["converged"] and all(d["st"]["iters"] == its for d in allr)
        # 2. ILU(0), one block per rank
        colL = O.tags_to_local(gcol, G["tag"][:n]); blocks = np.zeros(n, dtype=np.int32); x2d = np.zeros(n)
        for r_, d in enumerate(allr):
            gi = row_of_tag[d["tag"]]; blocks[gi] = r_; x2d[gi] = d["x2"]
        x2o, info2 = O.krylov_solve(grp, colL, gA, gb.copy(), params=O.krylov_params(precond=O.PREC_ILU0, row_gid=G["tag"][:n]), null_mask=np.ones(n, dtype=np.int32), use_null=True, blocks=blocks)
        its2 = allr[0]["st2"]["iters"]; x2err = np.linalg.norm(x2d - x2o) / np.linalg.norm(x2o)
        say(f"  block-Jacobi ILU(0): iters gpu {its2} vs oracle {info2['iters']}; x rel diff {x2err:.2e}; converged {allr[0]['st2']['converged']}")
        ok = ok and abs(its2 - info2["iters"]) <= 2 and x2err <= 1e-6 and allr[0]["st2"]["converged"]
        # 2b. Overlap Level 1: every rank's block extended by the rows of its halo columns, additive Schwarz with combine mode Add
        x2bd = np.zeros(n)
        for d in allr:
            x2bd[row_of_tag[d["tag"]]] = d["x2b"]
        x2bo, info2b = O.krylov_solve(grp, colL, gA, gb.copy(), params=O.krylov_params(precond=O.PREC_ILU0, overlap=1, row_gid=G["tag"][:n]), null_mask=np.ones(n, dtype=np.int32), use_null=True, blocks=blocks)
        its2b = allr[0]["st2b"]["iters"]; x2berr = np.linalg.norm(x2bd - x2bo) / np.linalg.norm(x2bo)
        say(f"  ILU(0), Overlap Level 1 (Add): iters gpu {its2b} vs oracle {info2b['iters']}; x rel diff {x2berr:.2e}; converged {allr[0]['st2b']['converged']}")
        ok = ok and abs(its2b - info2b["iters"]) <= 2 and x2berr <= 1e-6 and allr[0]["st2b"]["converged"]
        # 2c. multilevel stand-in for ML: the same aggregates as the restatement on the global problem (as a partition: the coarse numbering is
        # rank-major on the GPUs, row-major in the oracle), the same level sizes, the same iteration count
        x2cd = np.zeros(n); aggd = -np.ones(n, dtype=np.int64)
        for d in allr:
            gi = row_of_tag[d["tag"]]; x2cd[gi] = d["x2c"]; aggd[gi] = d["agg2c"]
        prm2c = O.krylov_params(precond=O.PREC_AMG, amg_max_coarse=20, row_gid=G["tag"][:n])
        x2co, info2c = O.krylov_solve(grp, colL, gA, gb.copy(), params=prm2c, null_mask=np.ones(n, dtype=np.int32), use_null=True, blocks=blocks)
        hier = O.amg_hierarchy(grp, colL, gA, prm2c, blocks=blocks)
        same_part = len(set(zip(aggd.tolist(), hier["agg"].tolist()))) == len(set(aggd.tolist())) == len(set(hier["agg"].tolist()))
        same_lev = all(list(d["ml2c"]["rows"][1:]) == list(hier["rows"][1:]) for d in allr) and sum(d["ml2c"]["rows"][0] for d in allr) == n
        its2c = allr[0]["st2c"]["iters"]; x2cerr = np.linalg.norm(x2cd - x2co) / np.linalg.norm(x2co)
        say(f"  ML stand-in (levels {hier['levels']}, rows {list(hier['rows'])}): same aggregates {same_part}, same level sizes {same_lev}; iters gpu {its2c} vs oracle {info2c['iters']}; x rel diff {x2cerr:.2e}; converged {allr[0]['st2c']['converged']}")
        ok = ok and same_part and same_lev and hier["levels"] >= 2 and abs(its2c - info2c["iters"]) <= 2 and x2cerr <= 1e-6 and allr[0]["st2c"]["converged"]
        # 3. Helmholtz, CG + Chebyshev(2), dim right-hand sides
        o.invalidate_matrix(); gbh = o.ns_helmholtz(dt, 0.5, np.asfortranarray(vg[:n, :dim])); gAh = o.matrix()
        x3d = np.zeros((n, dim)); worst3 = 0.0
        for d in allr:
            gi = row_of_tag[d["tag"]]; x3d[gi] = d["x3"]
            worst3 = max(worst3, float(np.abs(gbh[gi] - d["bh"]).max() / np.abs(gbh).max()))
            for li, gr in enumerate(gi):
                worst3 = max(worst3, relerr(gAh[grp[gr]:grp[gr + 1]], d["Ah"][d["rp"][li]:d["rp"][li + 1]]))
        prm = O.krylov_params(solver=O.SOLVER_CG, precond=O.PREC_CHEBYSHEV, cheb_degree=2, row_gid=G["tag"][:n]); its3o = 0; x3err = 0.0
        for k in range(dim):
            xk, ik = O.krylov_solve(grp, colL, gAh, gbh[:, k], params=prm); its3o += ik["iters"]
            x3err = max(x3err, np.linalg.norm(x3d[:, k] - xk) / np.linalg.norm(xk))
        its3 = allr[0]["st3"]["iters"]
        say(f"  Helmholtz CG+Chebyshev(2) x{dim}: values/b max err {worst3:.2e}; iters gpu {its3} vs oracle {its3o}; x rel diff {x3err:.2e}; converged {allr[0]['st3']['converged']}")
        ok = ok and worst3 <= 1e-12 and abs(its3 - its3o) <= 2 * dim and x3err <= 1e-7 and allr[0]["st3"]["converged"]
        # 4. Poisson-Boltzmann Newton on the global problem, built from the oracle's pieces (same stopping rule)
        sg = np.sin(xg[:, 0]) * np.cos(xg[:, 1]); exg = (-2.0 * sg - np.sinh(sg))[:n].copy()
        o.set_field(O.F_EPS, 1.0 + 0.2 * np.cos(xg[:, 0])); o.set_field(O.F_PSI0, 0.3 + 0.0 * sg)
        psi = np.zeros(len(xg)); nup = 0.0; lin = 0; kn = 0; prm4 = O.krylov_params(precond=O.PREC_JACOBI, row_gid=G["tag"][:n])
        while True:
            o.set_field(O.F_PSI, psi); fv = o.pb_residual(extra_f=exg); nf = np.linalg.norm(fv) / np.sqrt(n)
            if (kn > 0 and nf <= 1e-8 and nup <= 1e-5) or kn >= 100:
                break
            o.invalidate_matrix() if kn == 0 else None
            o.pb_jacobian(); Aj = o.matrix()
            dxk, ik = O.krylov_solve(grp, colL, Aj, -fv, params=prm4); lin += ik["iters"]
            psi[:n] += dxk; nup = np.linalg.norm(dxk) / np.sqrt(n); kn += 1
        psid = np.zeros(n)
        for d in allr:
            psid[row_of_tag[d["tag"]]] = d["psi4"]
        s4 = allr[0]["st4"]; p4err = np.linalg.norm(psid - psi[:n]) / np.linalg.norm(psi[:n])
        say(f"  Poisson-Boltzmann Newton: newton its gpu {s4['newton_iters']} vs oracle {kn}; linear its {s4['linear_iters']} vs {lin}; psi rel diff {p4err:.2e}; ||F|| {s4['normf']:.1e}; converged {s4['converged']}")
        ok = ok and s4["converged"] and s4["newton_iters"] == kn and abs(s4["linear_iters"] - lin) <= 2 * kn and p4err <= 1e-8
        say("MULTI_GPU_CHECK", "PASS" if ok else "FAIL")
        rep = dict(rows=int(n), graph="bit-exact", values_max_rel_err=float(worst), gmres_jacobi=dict(iters=int(its), oracle_iters=int(info["iters"]), x_rel_diff=float(xerr)),
                   gmres_block_ilu0=dict(iters=int(its2), oracle_iters=int(info2["iters"]), x_rel_diff=float(x2err)),
                   gmres_ilu0_overlap1=dict(iters=int(its2b), oracle_iters=int(info2b["iters"]), x_rel_diff=float(x2berr)),
                   gmres_ml_standin=dict(iters=int(its2c), oracle_iters=int(info2c["iters"]), x_rel_diff=float(x2cerr), same_aggregates=bool(same_part), levels=int(hier["levels"]), rows=[int(v) for v in hier["rows"]]),
                   helmholtz_cg_chebyshev=dict(iters=int(its3), oracle_iters=int(its3o), x_rel_diff=float(x3err), values_max_rel_err=float(worst3)),
                   pb_newton=dict(newton_iters=int(s4["newton_iters"]), oracle_newton_iters=int(kn), linear_iters=int(s4["linear_iters"]), oracle_linear_iters=int(lin), psi_rel_diff=float(p4err)))
    c.close()
    box = [dict(rep, **{"pass": bool(ok), "world": world, "against": "CPU oracle (port) on the global problem, tests/multi_gpu_check.py"})]
    dist.broadcast_object_list(box, 0)
    return box[0]


if __name__ == "__main__":
    main()
