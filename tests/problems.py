"""Shared problem definitions for the parity tests (small enough for the CPU oracle to finish in seconds).

Each case = one particle configuration + the per-particle fields a LAMMPS step would hold.  The same dict drives the
CPU oracle (oracle/oracle.py) and the CUDA path (implicit-sph_b200.Context) so that both see identical bytes.
"""
import importlib

import numpy as np

lat = importlib.import_module("implicit-sph_b200.lattice")
FLUID, SOLID = 99, 12

# name: dim, N, jitter(dx), rs2, kinds, solid slab (rows of y below this index are type 2), kernel, h_min factor
CASES = {
    "lattice2d":   dict(dim=2, N=24, jitter=0.0,  rs2=9,  kinds=(0, FLUID)),
    "jitter2d":    dict(dim=2, N=24, jitter=0.05, rs2=12, kinds=(0, FLUID)),
    "lattice3d":   dict(dim=3, N=10, jitter=0.0,  rs2=9,  kinds=(0, FLUID)),
    "jitter3d":    dict(dim=3, N=10, jitter=0.04, rs2=12, kinds=(0, FLUID)),
    "solid2d":     dict(dim=2, N=28, jitter=0.03, rs2=12, kinds=(0, FLUID, SOLID), slab=4),
    "solid3d":     dict(dim=3, N=10, jitter=0.03, rs2=12, kinds=(0, FLUID, SOLID), slab=3),
    "quintic2d":   dict(dim=2, N=20, jitter=0.04, rs2=12, kinds=(0, FLUID), kernel=2),
    "cubic3d":     dict(dim=3, N=9,  jitter=0.03, rs2=12, kinds=(0, FLUID), kernel=1),
    "buffer2d":    dict(dim=2, N=24, jitter=0.04, rs2=12, kinds=(0, FLUID, 32), slab=3),                 # type 2 = BufferDirichlet rows (applied potential / solute transport branches)
    "tgv128":      dict(dim=2, N=128, jitter=0.0, rs2=9,  kinds=(0, FLUID), origin=0.5),     # BASELINE C1 particle set
}


# ragged inputs: uniformly random particle clouds (rows of very different lengths, neighbor lists in an order unrelated to position)
CLOUDS = {
    "cloud2d":  dict(dim=2, n=500),
    "cloud3d":  dict(dim=3, n=700),
    "cloud3d_50k": dict(dim=3, n=50000, kd=True),
}


def make_case(name):
    if name in CLOUDS:
        return make_cloud_case(name)
    c = dict(CASES[name]); dim, N = c["dim"], c["N"]
    dx = 2.0 * np.pi / N
    slab = c.get("slab")
    type_fn = (lambda wx, wy, wz: np.where(wy < slab, 2, 1)) if slab else None
    P = lat.make_brick(dim, (N,) * dim, dx, rs2=c["rs2"], jitter=c["jitter"], type_fn=type_fn, origin=c.get("origin", 0.0))
    return P, fields_for(P, c, name, dx, slab)


def make_cloud_case(name):
    c = dict(CLOUDS[name]); dim, n = c["dim"], c["n"]; box = 2.0 * np.pi
    dx = box / n ** (1.0 / dim); h = 1.5 * dx
    gen = lat.make_cloud_kd if c.get("kd") else lat.make_cloud
    P = gen(dim, n, box, reach=2 * h * 1.05, min_sep=0.45 * dx)
    c["kinds"] = (0, FLUID)
    return P, fields_for(P, c, name, dx, None)


def fields_for(P, c, name, dx, slab):
    dim = P["dim"]
    xw = P["xw"]
    F = {}
    F["density"] = 1.0 + 0.1 * np.sin(xw[:, 0])
    F["viscosity"] = 0.1 + 0.01 * np.cos(xw[:, 1])
    F["pressure"] = np.sin(xw[:, 0]) * np.cos(xw[:, 1])
    v = lat.tgv_velocity(xw); v[:, 0] += 0.05 * np.sin(xw[:, 0]); v[:, 1] += 0.02 * np.cos(2 * xw[:, 1])
    if dim == 3:
        v[:, 2] = 0.03 * np.sin(xw[:, 2])
    F["velocity"] = v
    F["force"] = 0.01 * np.stack([np.cos(xw[:, 0]), np.sin(xw[:, 1]), np.zeros(len(xw))], axis=1)
    F["eps"] = 1.0 + 0.2 * np.cos(xw[:, 0])
    F["psi"] = np.sin(xw[:, 0]) * np.cos(xw[:, 1])
    F["sigma"] = 1.0 + 0.3 * np.sin(xw[:, 0])                                     # electric conductivity
    F["phi"] = np.cos(xw[:, 1]) + 0.1 * np.sin(2 * xw[:, 0])                      # applied potential (buffer rows: Dirichlet data)
    F["conc"] = 0.5 + 0.2 * np.sin(xw[:, 0]) * np.cos(xw[:, 1])                   # solute concentration c^n
    F["psi0"] = 0.3 + 0.1 * np.cos(xw[:, 0])                                     # prescribed potential of solid / boundary particles
    F["pb_extra"] = -2.0 * np.sin(xw[:, 0]) * np.cos(xw[:, 1]) - np.sinh(np.sin(xw[:, 0]) * np.cos(xw[:, 1]))   # poisson-boltzmann-harmonic.xml:14-30
    has_solid = bool(slab) and SOLID in c["kinds"]
    P["case"] = dict(name=name, kinds=c["kinds"], kernel=c.get("kernel", 0), has_solid=has_solid, dt=0.05 * dx / 0.1, theta=0.5,
                     h_min=(0.8 * 1.5 * dx) if has_solid else None)
    return F


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64).ravel(); b = np.asarray(b, dtype=np.float64).ravel()
    d = np.abs(a - b); s = np.maximum(np.abs(a), np.abs(b)); m = s > 0
    return float((d[m] / s[m]).max()) if m.any() else 0.0


def scaled_err(a, b):
    """max |a-b| / max|b| (for vectors whose individual entries may legitimately cancel to ~0)"""
    a = np.asarray(a, dtype=np.float64).ravel(); b = np.asarray(b, dtype=np.float64).ravel()
    s = np.abs(b).max()
    return float(np.abs(a - b).max() / s) if s > 0 else float(np.abs(a - b).max())


def mixed_err(a, b, rowptr, abs_frac=1e-15):
    """max over entries of |a-b| / (|b| + (abs_frac / 1e-12) * rowmax|b|): the value that has to stay <= 1e-12 for the mixed bar
    |a-b| <= 1e-12 |b| + abs_frac * (largest entry of the row).  Used on random clouds, where a handful of entries per million are
    the difference of two nearly equal terms (|entry| ~ 1e-12 of the row) and the REFERENCE ITSELF does not determine them to 1e-12:
    reordering a neighbor list (which LAMMPS does not keep stable across runs) moves them by up to 1e-10 (tests/test_oracle_cpu.py::
    test_cloud_entries_are_only_defined_up_to_the_neighbor_order)."""
    a = np.asarray(a, dtype=np.float64).ravel(); b = np.asarray(b, dtype=np.float64).ravel()
    rowmax = np.maximum.reduceat(np.abs(b), rowptr[:-1]); rm = np.repeat(rowmax, np.diff(rowptr))
    den = np.abs(b) + (abs_frac / 1e-12) * rm
    m = den > 0
    return float((np.abs(a - b)[m] / den[m]).max()) if m.any() else 0.0
