"""World-size-2 test of the multi-rank host logic on CPU (gloo): brick partition of the lattice, the halo planner that
the NCCL path uses (isph_halo_plan_host, pure host code inside libisph_b200.so), owner->ghost field forwarding and the
halo-exchanged SpMV — checked against the single-rank CPU oracle on the global problem.  The transport here is gloo
send/recv; on the GPUs the same plan drives ncclSend/ncclRecv (implicit-sph_b200/csrc/halo.cu)."""
import ctypes as C
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, dim, nglobal, jitter, rs2, out):
    import torch.distributed as dist
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        isph = importlib.import_module("implicit-sph_b200"); lat = importlib.import_module("implicit-sph_b200.lattice")
        import oracle as O
        L = isph.lib()
        ip = C.POINTER(C.c_int)
        grid = lat.brick_grid(world, dim); lo, nloc = lat.brick_of_rank(rank, grid, nglobal)
        dx = 2 * np.pi / nglobal[0]
        P = lat.make_brick(dim, nglobal, dx, lo=lo, nloc=nloc, rs2=rs2, jitter=jitter)
        nl, ng = P["nlocal"], P["nghost"]
        # ---- owned-tag directory (what halo.cu builds with ncclAllGather)
        tags = [None] * world
        dist.all_gather_object(tags, P["tag"][:nl].copy())
        owner_of = {};
        mt = max(int(t.max()) for t in tags) + 1
        own_rank = -np.ones(mt, dtype=np.int32); own_idx = -np.ones(mt, dtype=np.int32)
        for r, t in enumerate(tags):
            own_rank[t] = r; own_idx[t] = np.arange(len(t), dtype=np.int32)
        gt = np.ascontiguousarray(P["tag"][nl:], dtype=np.int32)
        g_owner = np.ascontiguousarray(own_rank[gt]); g_idx = np.ascontiguousarray(own_idx[gt])
        gcol = np.zeros(ng, dtype=np.int32); rcount = np.zeros(world, dtype=np.int32); req = np.zeros(max(ng, 1), dtype=np.int32); nh = C.c_int()
        rc = L.isph_halo_plan_host(world, rank, nl, ng, gt.ctypes.data_as(ip), g_owner.ctypes.data_as(ip), g_idx.ctypes.data_as(ip),
                                   gcol.ctypes.data_as(ip), rcount.ctypes.data_as(ip), req.ctypes.data_as(ip), C.byref(nh))
        assert rc == 0
        nhalo = nh.value; roff = np.concatenate([[0], np.cumsum(rcount)])
        assert roff[-1] == nhalo and rcount[rank] == 0
        # every remote ghost maps to a slot whose request points at the right owner-local particle
        rem = g_owner != rank
        assert np.all(gcol[~rem] == g_idx[~rem]) and np.all(gcol[rem] >= nl)
        assert np.all(req[gcol[rem] - nl] == g_idx[rem])
        # ---- swap request lists (halo.cu: grouped ncclSend/ncclRecv of ints)
        send_lists = [None] * world
        reqs = [torch.from_numpy(req[roff[p]:roff[p + 1]].copy()) for p in range(world)]
        cnts = [None] * world
        dist.all_gather_object(cnts, rcount.tolist())
        ops = []
        for p in range(world):
            if p == rank:
                continue
            send_lists[p] = torch.zeros(cnts[p][rank], dtype=torch.int32)
            if rcount[p]:
                ops.append(dist.isend(reqs[p], p))
            if cnts[p][rank]:
                ops.append(dist.irecv(send_lists[p], p))
        for o in ops:
            o.wait()

        def forward(values, nc=1):
            """owner -> halo slots of an [nl, nc] array; returns [nhalo, nc]"""
            values = np.ascontiguousarray(values.reshape(nl, nc)); halo = np.zeros((nhalo, nc)); ops = []; bufs = []
            for p in range(world):
                if p == rank:
                    continue
                if cnts[p][rank]:
                    sb = torch.from_numpy(values[send_lists[p].numpy()].copy()); bufs.append(sb); ops.append(dist.isend(sb, p))
                if rcount[p]:
                    rb = torch.zeros((int(rcount[p]), nc), dtype=torch.float64); bufs.append((p, rb)); ops.append(dist.irecv(rb, p))
            for o in ops:
                o.wait()
            for b in bufs:
                if isinstance(b, tuple):
                    halo[roff[b[0]]:roff[b[0] + 1]] = b[1].numpy()
            return halo

        # ---- per-rank assembly with the CPU oracle; vfrac of remote ghosts comes through the plan (forward_comm_pair)
        xw = P["xw"]; v = lat.tgv_velocity(xw); v[:, 0] += 0.05 * np.sin(xw[:, 0])
        o = O.Oracle(P, kind="port"); o.set_field(O.F_VSTAR, v)
        o.compute_pre()
        vf = o.get_field(O.F_VFRAC)
        halo_vf = forward(vf[:nl])
        vf[nl:][rem] = halo_vf[gcol[rem] - nl, 0]
        o.set_field(O.F_VFRAC, vf)
        rp, col = o.graph(); b = o.ns_poisson(0.05); A = o.matrix()
        # ---- halo-exchanged SpMV
        xg = np.random.default_rng(7).standard_normal(int(np.prod(nglobal)) + 1)       # indexed by tag
        x_own = xg[P["tag"][:nl]]
        x_cols = np.concatenate([x_own, forward(x_own)[:, 0]])
        tag2col = -np.ones(mt, dtype=np.int64); tag2col[P["tag"][:nl]] = np.arange(nl)
        remote_tags = gt[rem]; tag2col[remote_tags] = gcol[rem]
        y = np.add.reduceat(A * x_cols[tag2col[col]], rp[:-1])
        out[rank] = dict(tag=P["tag"][:nl].copy(), rp=rp, col=col, A=A, b=b, y=y, vf=vf[:nl].copy(), nhalo=nhalo)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dim,nglobal,jitter,rs2", [(3, (12, 8, 8), 0.04, 12), (2, (24, 16), 0.0, 9)])
def test_two_rank_partition_matches_global_oracle(dim, nglobal, jitter, rs2, oracle_mod, lattice):
    import torch.multiprocessing as mp
    mgr = mp.Manager(); out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, dim, nglobal, jitter, rs2, out), nprocs=2, join=True)
    # global single-rank oracle
    O = oracle_mod
    dx = 2 * np.pi / nglobal[0]
    P = lattice.make_brick(dim, nglobal, dx, rs2=rs2, jitter=jitter)
    xw = P["xw"]; v = lattice.tgv_velocity(xw); v[:, 0] += 0.05 * np.sin(xw[:, 0])
    o = O.Oracle(P, kind="port"); o.set_field(O.F_VSTAR, v); o.compute_pre(); rp, col = o.graph(); b = o.ns_poisson(0.05); A = o.matrix()
    vf = o.get_field(O.F_VFRAC)
    xg = np.random.default_rng(7).standard_normal(int(np.prod(nglobal)) + 1)
    yg = np.add.reduceat(A * xg[col], rp[:-1])
    row_of_tag = {int(t): i for i, t in enumerate(P["tag"][:P["nlocal"]])}
    seen = 0
    for r in range(2):
        d = out[r]; assert d["nhalo"] > 0
        for li, t in enumerate(d["tag"]):
            gi = row_of_tag[int(t)]; seen += 1
            cg = col[rp[gi]:rp[gi + 1]]; cl = d["col"][d["rp"][li]:d["rp"][li + 1]]
            assert np.array_equal(cg, cl)                                                   # graph: bit-exact, any partition
            assert np.array_equal(A[rp[gi]:rp[gi + 1]], d["A"][d["rp"][li]:d["rp"][li + 1]])  # values: same arithmetic per row
            assert d["b"][li] == b[gi] and d["vf"][li] == vf[gi]
            assert abs(d["y"][li] - yg[gi]) <= 1e-13 * max(1.0, abs(yg[gi]))
    assert seen == P["nlocal"]


@pytest.mark.parametrize("world", [4, 8])
def test_halo_plan_for_four_and_eight_bricks(world, lattice):
    """The host planner on the 2x2x1 / 2x2x2 brick layouts bench.py uses at 4 / 8 GPUs (in-process: the planner is a pure function):
    every remote ghost gets a halo column, slots are grouped by owner in (owner, tag) order, and the request list of rank r for
    peer p names exactly the owner-local particles whose tags r's slots carry."""
    isph = importlib.import_module("implicit-sph_b200"); L = isph.lib(); ip = C.POINTER(C.c_int)
    dim = 3; grid = lattice.brick_grid(world, dim); ng_ = tuple(8 * g for g in grid); dx = 2 * np.pi / ng_[0]
    bricks = []
    for r in range(world):
        lo, nloc = lattice.brick_of_rank(r, grid, ng_)
        bricks.append(lattice.make_brick(dim, ng_, dx, lo=lo, nloc=nloc, rs2=12, jitter=0.03))
    mt = max(int(P["tag"].max()) for P in bricks) + 1
    own_rank = -np.ones(mt, dtype=np.int32); own_idx = -np.ones(mt, dtype=np.int32)
    for r, P in enumerate(bricks):
        t = P["tag"][:P["nlocal"]]; assert np.all(own_rank[t] < 0); own_rank[t] = r; own_idx[t] = np.arange(len(t), dtype=np.int32)
    assert np.count_nonzero(own_rank >= 0) == int(np.prod(ng_))                      # every particle owned exactly once
    recv = np.zeros((world, world), dtype=np.int64)
    for r, P in enumerate(bricks):
        nl, ngh = P["nlocal"], P["nghost"]
        gt = np.ascontiguousarray(P["tag"][nl:], dtype=np.int32); g_owner = np.ascontiguousarray(own_rank[gt]); g_idx = np.ascontiguousarray(own_idx[gt])
        gcol = np.zeros(ngh, dtype=np.int32); rcount = np.zeros(world, dtype=np.int32); req = np.zeros(max(ngh, 1), dtype=np.int32); nh = C.c_int()
        assert L.isph_halo_plan_host(world, r, nl, ngh, gt.ctypes.data_as(ip), g_owner.ctypes.data_as(ip), g_idx.ctypes.data_as(ip),
                                     gcol.ctypes.data_as(ip), rcount.ctypes.data_as(ip), req.ctypes.data_as(ip), C.byref(nh)) == 0
        nhalo = nh.value; roff = np.concatenate([[0], np.cumsum(rcount)]); recv[r] = rcount
        rem = g_owner != r
        assert roff[-1] == nhalo == len(np.unique(gt[rem])) and rcount[r] == 0       # one slot per distinct remote tag
        assert np.all(gcol[~rem] == g_idx[~rem]) and np.all((gcol[rem] >= nl) & (gcol[rem] < nl + nhalo))
        slot_tag = np.zeros(nhalo, dtype=np.int64); slot_tag[gcol[rem] - nl] = gt[rem]
        for p in range(world):
            st = slot_tag[roff[p]:roff[p + 1]]
            assert np.all(own_rank[st] == p) and np.all(np.diff(st) > 0)              # grouped by owner, ascending tag inside a group
            assert np.array_equal(bricks[p]["tag"][req[roff[p]:roff[p + 1]]], st)     # the request names the right particles on the owner
    assert np.array_equal(recv > 0, (recv > 0).T)                                    # exchanges are pairwise in both directions


def test_halo_plan_host_beyond_eight_ranks(isph):
    """the planner itself has no 8-rank limit (the NVLink peer path has, and is skipped above it); slots are numbered
    in (owner, tag) order whatever the world size, repeated tags share a slot, ghosts owned here keep the owner-local index."""
    import ctypes as C
    rng = np.random.default_rng(12); R, rank, nl, ng = 12, 5, 40, 2500
    owner = rng.integers(0, R, ng).astype(np.int32); oidx = rng.integers(0, 200, ng).astype(np.int32)
    tag = (1 + owner.astype(np.int64) * 1000 + oidx).astype(np.int32)
    col = np.zeros(ng, dtype=np.int32); rc = np.zeros(R, dtype=np.int32); req = np.zeros(ng, dtype=np.int32); nh = C.c_int()
    assert isph.lib().isph_halo_plan_host(R, rank, nl, ng, isph._i(tag), isph._i(owner), isph._i(oidx), isph._i(col), isph._i(rc), isph._i(req), C.byref(nh)) == 0
    remote = owner != rank
    keys = sorted(set(zip(owner[remote].tolist(), tag[remote].tolist())))
    assert nh.value == len(keys) and rc.sum() == len(keys) and rc[rank] == 0
    slot = {k: i for i, k in enumerate(keys)}
    assert all(col[g] == (nl + slot[(owner[g], tag[g])] if remote[g] else oidx[g]) for g in range(ng))
    assert all(req[slot[k]] == k[1] - 1 - k[0] * 1000 for k in keys)
    assert [int(rc[p]) for p in range(R)] == [sum(1 for k in keys if k[0] == p) for p in range(R)]
