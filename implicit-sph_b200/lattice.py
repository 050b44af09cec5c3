"""Synthetic LAMMPS-shaped inputs for the linear-solve hot path (host-side helper for tests / bench).

Produces exactly what the drop-in boundary receives from LAMMPS (SURVEY.md §8b "Input data layouts"):
owned atoms 0..nlocal-1 followed by ghost atoms (periodic images / off-rank copies carrying the owner's
tag), row-major x[nall][3], int type/tag, and a *full* neighbor list (ilist, packed offsets, neigh).
The lattice follows the reference's scripts: `lattice sc|sq ${dx}` in a periodic box [0, N*dx)^d
(IMPLICIT-SPH/sph-script/taylor-green-vortex-2d.lmp:69-74, taylor-green-vortex-3d.lmp), one brick of it
per rank (LAMMPS spatial decomposition, pair_isph.cpp:1258-1259).

The neighbor list is a superset of the cut sphere (all lattice offsets with |o|^2 <= rs2), the same role a
LAMMPS list with a small skin plays: the functors' own `rsq < cutsq` test stays the sole decider
(functor_graph.h:76-84; SURVEY.md §7 "Bit-exact graph on lattices").
"""
from __future__ import annotations

import numpy as np


def _hash01(tag: np.ndarray, salt: int) -> np.ndarray:
    """Deterministic per-tag uniform(0,1) (splitmix64): ghosts get bit-identical jitter to their owner."""
    with np.errstate(over="ignore"):
        z = tag.astype(np.uint64) + np.array([0x9E3779B97F4A7C15], dtype=np.uint64) * np.array([salt + 1], dtype=np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def stencil_offsets(dim: int, rs2: int) -> np.ndarray:
    r = int(np.floor(np.sqrt(rs2)))
    rng = np.arange(-r, r + 1)
    if dim == 2:
        oy, ox = np.meshgrid(rng, rng, indexing="ij")
        off = np.stack([ox.ravel(), oy.ravel(), np.zeros(ox.size, dtype=np.int64)], axis=1)
    else:
        oz, oy, ox = np.meshgrid(rng, rng, rng, indexing="ij")
        off = np.stack([ox.ravel(), oy.ravel(), oz.ravel()], axis=1)
    n2 = (off ** 2).sum(axis=1)
    return off[(n2 <= rs2) & (n2 > 0)]


def make_brick(dim, nglobal, dx, lo=None, nloc=None, rs2=9, jitter=0.0, seed=42, origin=0.0,
               type_fn=None, tag_perm=None):
    """One rank's particle set.

    nglobal : (Nx,Ny[,Nz]) global periodic lattice;  lo/nloc : this rank's brick (default: whole box)
    jitter  : displacement amplitude in units of dx (uniform in [-jitter, jitter] per component)
    type_fn : callable(gx, gy, gz) -> int type array (default all 1)
    returns dict with x,type,tag,nlocal,nghost,ilist,noff,neigh,gidx (global lattice index per atom)
    """
    nglobal = tuple(int(v) for v in nglobal) + ((1,) if dim == 2 else ())
    lo = (0, 0, 0) if lo is None else tuple(lo) + ((0,) if len(lo) == 2 else ())
    nloc = nglobal if nloc is None else tuple(nloc) + ((1,) if len(nloc) == 2 else ())
    g = int(np.floor(np.sqrt(rs2)))
    gz = 0 if dim == 2 else g
    ext = (nloc[0] + 2 * g, nloc[1] + 2 * g, nloc[2] + 2 * gz)

    # extended-lattice coordinates (x fastest, like LAMMPS create_atoms)
    ez, ey, ex = np.meshgrid(np.arange(-gz, nloc[2] + gz), np.arange(-g, nloc[1] + g), np.arange(-g, nloc[0] + g), indexing="ij")
    ex = ex.ravel(); ey = ey.ravel(); ez = ez.ravel()
    owned = (ex >= 0) & (ex < nloc[0]) & (ey >= 0) & (ey < nloc[1]) & (ez >= 0) & (ez < nloc[2])
    nlocal = int(owned.sum()); next_ = ex.size; nghost = next_ - nlocal
    ext2atom = np.empty(next_, dtype=np.int32)
    ext2atom[owned] = np.arange(nlocal, dtype=np.int32)
    ext2atom[~owned] = np.arange(nlocal, next_, dtype=np.int32)
    order = np.argsort(ext2atom, kind="stable")            # atom -> extended index
    ax = ex[order] + lo[0]; ay = ey[order] + lo[1]; az = ez[order] + lo[2]   # unwrapped global lattice coords
    wx = np.mod(ax, nglobal[0]); wy = np.mod(ay, nglobal[1]); wz = np.mod(az, nglobal[2])
    gidx = (wz.astype(np.int64) * nglobal[1] + wy) * nglobal[0] + wx
    tag = (gidx + 1).astype(np.int32) if tag_perm is None else tag_perm[gidx].astype(np.int32)

    # positions: owner's coordinate (lattice + per-tag jitter) shifted by whole box lengths for images
    L = np.array([nglobal[0] * dx, nglobal[1] * dx, nglobal[2] * dx])
    xw = np.empty((next_, 3)); xw[:, 0] = (wx + origin) * dx; xw[:, 1] = (wy + origin) * dx; xw[:, 2] = (wz + (origin if dim == 3 else 0.0)) * dx
    if jitter > 0.0:
        for k in range(dim):
            xw[:, k] += (2.0 * _hash01(gidx + 1, seed * 3 + k) - 1.0) * (jitter * dx)
    x = xw.copy()
    x[:, 0] += ((ax - wx) // nglobal[0]) * L[0]; x[:, 1] += ((ay - wy) // nglobal[1]) * L[1]
    if dim == 3:
        x[:, 2] += ((az - wz) // nglobal[2]) * L[2]

    typ = np.ones(next_, dtype=np.int32) if type_fn is None else np.asarray(type_fn(wx, wy, wz), dtype=np.int32)

    # full neighbor list of the owned atoms: stencil gather on the extended lattice
    off = stencil_offsets(dim, rs2)
    lin = (off[:, 2] * ext[1] + off[:, 1]) * ext[0] + off[:, 0]
    own_ext = order[:nlocal].astype(np.int64)
    ns = lin.size
    neigh = np.empty((nlocal, ns), dtype=np.int32)
    step = max(1, (1 << 24) // ns)
    for s in range(0, nlocal, step):
        neigh[s:s + step] = ext2atom[own_ext[s:s + step, None] + lin[None, :]]
    noff = np.arange(nlocal + 1, dtype=np.int64) * ns
    return dict(dim=dim, nlocal=nlocal, nghost=nghost, x=np.ascontiguousarray(x), xw=xw, type=typ, tag=tag,
                ilist=np.arange(nlocal, dtype=np.int32), noff=noff, neigh=neigh.ravel(), gidx=gidx,
                dx=dx, nglobal=nglobal[:dim], lo=lo[:dim], nloc=nloc[:dim])


def brick_grid(n_ranks: int, dim: int = 3):
    """Processor grid for n_ranks (1,2,4,8 -> 1x1x1, 2x1x1, 2x2x1, 2x2x2), LAMMPS-style bricks."""
    grid = [1, 1, 1]
    k = 0
    r = n_ranks
    while r > 1:
        assert r % 2 == 0, "power-of-two rank counts only"
        grid[k % dim] *= 2; r //= 2; k += 1
    return tuple(grid[:dim])


def brick_of_rank(rank: int, grid, nglobal):
    dim = len(grid)
    c = []; r = rank
    for k in range(dim):
        c.append(r % grid[k]); r //= grid[k]
    nloc = tuple(nglobal[k] // grid[k] for k in range(dim))
    lo = tuple(c[k] * nloc[k] for k in range(dim))
    return lo, nloc


def tgv_velocity(xw: np.ndarray, U: float = 0.1) -> np.ndarray:
    """Taylor-Green initial velocity, taylor-green-vortex-2d.lmp:135-142 (same expression used in 3-D, w=0)."""
    v = np.zeros_like(xw)
    v[:, 0] = U * np.sin(xw[:, 0]) * np.cos(xw[:, 1])
    v[:, 1] = -U * np.cos(xw[:, 0]) * np.sin(xw[:, 1])
    return v


def make_cloud(dim, n, box, reach, seed=7, min_sep=0.0, type_fn=None):
    """A non-lattice particle set in the same LAMMPS shape: `n` owned particles uniformly random in the periodic box
    [0, box)^dim (rejected below `min_sep` of an earlier one), ghosts = every periodic image within `reach` of the box, and a
    full neighbor list holding, for each owned particle, all atoms (owned or ghost) within `reach` in an order unrelated to
    position — rows have ragged lengths.  `reach` must be >= the pair cutoff; the functors' own rsq < cutsq test decides."""
    rng = np.random.default_rng(seed)
    pts = []
    while len(pts) < n:
        p = rng.uniform(0.0, box, size=3); p[dim:] = 0.0
        if min_sep > 0.0 and pts:
            d = np.abs(np.asarray(pts) - p); d = np.minimum(d, box - d); d[:, dim:] = 0.0
            if (np.sqrt((d ** 2).sum(axis=1)) < min_sep).any():
                continue
        pts.append(p)
    own = np.asarray(pts)
    shifts = [s for s in np.ndindex(*(3,) * dim)]
    gx, gt = [], []
    for s in shifts:
        sh = np.zeros(3); sh[:dim] = (np.asarray(s) - 1) * box
        if not sh.any():
            continue
        img = own + sh
        near = np.all((img[:, :dim] > -reach) & (img[:, :dim] < box + reach), axis=1)
        gx.append(img[near]); gt.append(np.nonzero(near)[0])
    ghost = np.concatenate(gx) if gx else np.zeros((0, 3)); gtag = np.concatenate(gt) if gt else np.zeros(0, dtype=np.int64)
    x = np.concatenate([own, ghost]); tag = np.concatenate([np.arange(n), gtag]).astype(np.int32) + 1
    xw = np.concatenate([own, own[gtag]])
    neigh, noff = [], [0]
    for i in range(n):
        d = x - own[i]; r2 = (d ** 2).sum(axis=1)
        j = np.nonzero((r2 < reach * reach) & (np.arange(len(x)) != i))[0]
        neigh.append(rng.permutation(j)); noff.append(noff[-1] + len(j))
    typ = np.ones(len(x), dtype=np.int32) if type_fn is None else np.asarray(type_fn(xw), dtype=np.int32)
    return dict(dim=dim, nlocal=n, nghost=len(ghost), x=np.ascontiguousarray(x), xw=xw, type=typ, tag=tag,
                ilist=np.arange(n, dtype=np.int32), noff=np.asarray(noff, dtype=np.int64), neigh=np.concatenate(neigh).astype(np.int32),
                gidx=(tag - 1).astype(np.int64), dx=box / n ** (1.0 / dim), nglobal=(n,), lo=(0,) * dim, nloc=(n,))


def make_cloud_kd(dim, n, box, reach, seed=11, min_sep=0.0, type_fn=None):
    """`make_cloud` for tens of thousands of particles: the same construction (uniformly random owned particles in the
    periodic box, thinned below `min_sep`; ghosts = every periodic image within `reach` of the box; full neighbor list of all
    atoms within `reach`, each row in a random order) with k-d trees instead of the O(n^2) scans."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(seed)
    own = np.zeros((0, 3))
    while len(own) < n:
        cand = np.zeros((int(1.5 * (n - len(own))) + 16, 3)); cand[:, :dim] = rng.uniform(0.0, box, size=(len(cand), dim))
        pts = np.concatenate([own, cand])
        if min_sep > 0.0:
            t = cKDTree(pts[:, :dim], boxsize=box)
            drop = np.zeros(len(pts), dtype=bool)
            for a, b in sorted(t.query_pairs(min_sep)):          # keep the earlier particle of every close pair
                if not drop[a]:
                    drop[b] = True
            pts = pts[~drop]
        own = pts[:n]
    gx, gt = [], []
    for s in np.ndindex(*(3,) * dim):
        sh = np.zeros(3); sh[:dim] = (np.asarray(s) - 1) * box
        if not sh.any():
            continue
        img = own + sh
        near = np.all((img[:, :dim] > -reach) & (img[:, :dim] < box + reach), axis=1)
        gx.append(img[near]); gt.append(np.nonzero(near)[0])
    ghost = np.concatenate(gx); gtag = np.concatenate(gt)
    x = np.concatenate([own, ghost]); tag = np.concatenate([np.arange(n), gtag]).astype(np.int32) + 1
    xw = np.concatenate([own, own[gtag]])
    tree = cKDTree(x[:, :dim])
    rows = tree.query_ball_point(own[:, :dim], reach * (1.0 - 1e-12))
    neigh, noff = [], np.zeros(n + 1, dtype=np.int64)
    for i, r in enumerate(rows):
        j = np.asarray(r, dtype=np.int32); j = j[j != i]
        neigh.append(rng.permutation(j)); noff[i + 1] = noff[i] + len(j)
    typ = np.ones(len(x), dtype=np.int32) if type_fn is None else np.asarray(type_fn(xw), dtype=np.int32)
    return dict(dim=dim, nlocal=n, nghost=len(ghost), x=np.ascontiguousarray(x), xw=xw, type=typ, tag=tag,
                ilist=np.arange(n, dtype=np.int32), noff=noff, neigh=np.concatenate(neigh).astype(np.int32),
                gidx=(tag - 1).astype(np.int64), dx=box / n ** (1.0 / dim), nglobal=(n,), lo=(0,) * dim, nloc=(n,))
