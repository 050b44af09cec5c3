"""Multi-GPU parity check of the NCCL path (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py

Every rank owns one brick of a periodic jittered lattice; graph + assembly + (NullSpace) GMRES/Jacobi run distributed
(halo exchange with ncclSend/ncclRecv, reductions with ncclAllReduce) and are compared on rank 0 with the CPU oracle on
the GLOBAL problem: graph bit-exact, values <= 1e-12, iteration count +-2, solution <= 1e-6 (kappa ~ 1e3).
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
    dim = 3; grid = lat.brick_grid(world, dim); per = (10, 8, 8); nglobal = tuple(per[k] * grid[k] for k in range(dim))
    lo, nloc = lat.brick_of_rank(rank, grid, nglobal)
    dx = 2 * np.pi / nglobal[0]
    P = lat.make_brick(dim, nglobal, dx, lo=lo, nloc=nloc, rs2=12, jitter=0.04)
    nl = P["nlocal"]; xw = P["xw"]
    v = lat.tgv_velocity(xw)
    for k in range(dim):
        v[:, k] += 0.05 * (2.0 * lat._hash01(P["gidx"] + 1, 100 + k) - 1.0)
    c = isph.Context(lr, world, rank, bytes(idt.tolist()))
    c.set_particles(P)
    c.field_set(isph.F_VSTAR, v)
    c.compute_pre(); c.graph_build()
    c.create_load(None, 1); dt = 0.05; c.ns_poisson(dt)
    rp, col = c.graph_get(); A = c.matrix_get(); b = c.load_get(1)[:, 0]; vf = c.field_get(isph.F_VFRAC)[:nl]
    x = np.zeros(nl); c.create_solution(x, 1)
    c.set_matrix_is_singular(True); c.set_initial_solution(isph.INIT_ZERO); c.precond_param("Precond Type", "point relaxation")
    st = c.solve(True, "Poisson")
    mine = dict(tag=P["tag"][:nl].copy(), rp=rp, col=col, A=A, b=b, x=x, vf=vf, st=st)
    allr = [None] * world
    dist.gather_object(mine, allr if rank == 0 else None, 0)
    ok = True
    if rank == 0:
        import oracle as O
        from problems import relerr
        G = lat.make_brick(dim, nglobal, dx, rs2=12, jitter=0.04)
        vg = lat.tgv_velocity(G["xw"])
        for k in range(dim):
            vg[:, k] += 0.05 * (2.0 * lat._hash01(G["gidx"] + 1, 100 + k) - 1.0)
        o = O.Oracle(G, kind="port"); o.set_field(O.F_VSTAR, vg); o.compute_pre(); grp, gcol = o.graph(); gb = o.ns_poisson(dt); gA = o.matrix(); gvf = o.get_field(O.F_VFRAC)
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
        print(f"multi_gpu_check world={world} rows={n}: values/b/vfrac max err {worst:.2e}; iters gpu {its} vs oracle {info['iters']}; x rel diff {xerr:.2e}; converged {allr[0]['st']['converged']}")
        ok = worst <= 1e-12 and abs(its - info["iters"]) <= 2 and xerr <= 1e-6 and allr[0]["st"]["converged"] and all(d["st"]["iters"] == its for d in allr)
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL")
    c.close()
    flag = torch.tensor([1 if ok else 0]); dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
