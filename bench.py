#!/usr/bin/env python
"""bench.py — implicit-sph linear-solve hot path on B200 (see DESIGN.md "Measurement").

One *step* = what `PairISPH::compute` does per time step for the pressure Poisson problem (pair_isph.cpp:1241-1380,
:988-1026): pre-computation (volumes / gradient & Laplacian corrections), nodal map + graph, operator assembly + RHS
(computePoisson), flexible GMRES(50) + preconditioner solve (solvePoisson) — everything rebuilt from scratch, as the
reference does.  Metric: rows assembled-and-solved per second (whole job), with `ms_per_step` = the absolute Poisson
step time BASELINE.json asks for and `roofline` = the SpMV kernel's achieved HBM bandwidth inside the solve.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload p8m|p8m_ml|c2|c2_ml|c1|c3|c4|c4s|c4s_ml|c5|c5_ml|c2j] [--n LATTICE]

Default workload = the configuration BASELINE.json's metric and target are quoted on: the 3-D 8M-particle (200^3) pressure
Poisson GMRES solve — it fits one B200 (17 GB), and the same global problem is split over N GPUs (strong scaling).  The line
also carries, at every N, secondary blocks for BASELINE configs[2] (`configs2_c3`: 8M Helmholtz, 3 RHS, CG + Chebyshev),
configs[3] (`configs3_c4`: 8M corrected-operator Poisson, GMRES + block-Jacobi ILU(0) on 4x4x4 bricks of 50^3; `configs3_c4_ml`: the same
problem with the multilevel stand-in for ML), `p8m_ml` (the headline problem with the ML stand-in) and configs[4]
(`configs4_c5`: 4M Poisson-Boltzmann Newton), at N=1 also configs[1] (`configs1_c2`, 1M particles), and at N>1 a `parity`
block: the multi-GPU path checked against the CPU oracle on a small global problem before the timed region (the run fails on
a mismatch).  `--impl reference` times the CPU path (oracle port, OpenMP over all host cores) on a bounded sample of the same
workload, with the same `config`.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: dim, lattice edge, jitter, rs2, precond, solver, description
    "c2": dict(dim=3, n=100, jitter=0.0, rs2=9, prec="point relaxation", solver="Block GMRES",
               desc="BASELINE configs[1]: 3-D pressure Poisson, 1M-particle periodic cubic lattice, Wendland h=1.5dx cut=2h, NullSpace, flexible GMRES(50)+Jacobi"),
    "c1": dict(dim=2, n=128, jitter=0.0, rs2=9, prec="ILU", solver="Block GMRES", origin=0.5,
               desc="BASELINE configs[0]: 2-D TGV 128x128 pressure Poisson, flexible GMRES(50)+ILU(0)"),
    "c2j": dict(dim=3, n=100, jitter=0.05, rs2=12, prec="point relaxation", solver="Block GMRES",
                desc="configs[1] particle count with positions jittered by 0.05 dx (no pair on the cutoff)"),
    # BASELINE configs[3] operator family per GPU: corrected (Gc/Lc) operators, jittered particles, block-Jacobi ILU(0) with
    # 2x2x2 Ifpack-rank-equivalent bricks per GPU (= 4x4x4 bricks of 50^3 at 8 GPUs)
    "c4": dict(dim=3, n=100, jitter=0.04, rs2=12, prec="ILU", solver="Block GMRES", anti=False, blocks=2,
               desc="BASELINE configs[3] per GPU: corrected-operator (Gc/Lc) pressure Poisson, 1M particles per GPU, GMRES(50) + block-Jacobi ILU(0), 8 bricks of 50^3 per GPU"),
    # BASELINE configs[3] as BASELINE.md §3 defines it: the 8M-row problem (strong scaling), block-Jacobi ILU(0) on the Ifpack-rank
    # partition 4x4x4 bricks of 50^3 (64 / N blocks per GPU), corrected (Gc/Lc) operators, jittered particles
    "c4s": dict(dim=3, n=200, jitter=0.04, rs2=12, prec="ILU", solver="Block GMRES", anti=False, strong=True, global_blocks=4,
                desc="BASELINE configs[3]: 3-D corrected-operator (Gc/Lc) pressure Poisson, 8M particles (200^3), GMRES(50) + block-Jacobi ILU(0) on 4x4x4 Ifpack-rank-equivalent bricks of 50^3, fixed global size (strong scaling)"),
    # BASELINE configs[2]: velocity Helmholtz (I - theta dt nu lap) v* = rhs, dim right-hand sides one after another, CG + Chebyshev
    "c3": dict(dim=3, n=200, jitter=0.0, rs2=9, prec="Chebyshev", solver="Block CG", strong=True, system="helmholtz", theta=0.5,
               desc="BASELINE configs[2]: 3-D velocity Helmholtz (functor_incomp_navier_stokes_helmholtz), 8M particles, 3 right-hand sides, CG + Chebyshev(1), x0 = v^n, fixed global size (strong scaling)"),
    # BASELINE configs[4]: Poisson-Boltzmann Newton-Krylov (full-step Newton, NormF 1e-8 AND NormUpdate 1e-5), manufactured source
    "c5": dict(dim=3, n=160, jitter=0.0, rs2=9, prec="point relaxation", solver="Block GMRES", strong=True, system="pb",
               desc="BASELINE configs[4]: 3-D Poisson-Boltzmann electrostatics, 4M particles (160^3), Newton iteration with device-resident computeF / computeJacobian and GMRES(50)+Jacobi Jacobian solves, manufactured source of poisson-boltzmann-harmonic.xml, psi0 = 0, fixed global size (strong scaling)"),
    # the headline problems with the reference's DEFAULT preconditioner package (ML, pair_isph.cpp:325-329): the multilevel stand-in of csrc/amg.cu
    "p8m_ml": dict(dim=3, n=200, jitter=0.0, rs2=9, prec="ML", solver="Block GMRES", strong=True,
                   desc="north_star problem (3-D 8M-particle pressure Poisson) with the reference's default preconditioner package: flexible GMRES(50) + the multilevel stand-in for ML (MIS aggregation, Chebyshev smoothers, V-cycle), fixed global size (strong scaling)"),
    "c2_ml": dict(dim=3, n=100, jitter=0.0, rs2=9, prec="ML", solver="Block GMRES",
                  desc="BASELINE configs[1] problem (1M particles) with flexible GMRES(50) + the multilevel stand-in for ML"),
    "c4s_ml": dict(dim=3, n=200, jitter=0.04, rs2=12, prec="ML", solver="Block GMRES", anti=False, strong=True,
                   desc="BASELINE configs[3] problem (8M particles, corrected Gc/Lc operators, jittered lattice) with GMRES(50) + the multilevel stand-in for ML instead of block-Jacobi ILU(0)"),
    "c5_ml": dict(dim=3, n=160, jitter=0.0, rs2=9, prec="ML", solver="Block GMRES", strong=True, system="pb",
                  desc="BASELINE configs[4] problem (4M-particle Poisson-Boltzmann Newton) with GMRES(50) + the multilevel stand-in for ML in the Jacobian solves"),
    # north_star target: fixed 8M-particle problem split over the GPUs (strong scaling)
    "p8m": dict(dim=3, n=200, jitter=0.0, rs2=9, prec="point relaxation", solver="Block GMRES", strong=True,
                desc="north_star target: 3-D 8M-particle (200^3) pressure Poisson, flexible GMRES(50)+Jacobi, fixed global size (strong scaling)"),
}


def workload_config(wname, w, world, lat):
    """The declarative description of the workload: identical on the B200 arm and on the reference arm (the driver compares the
    two `config` objects).  Everything a run MEASURES (iterations, residual, nnz) goes into `result`, not here."""
    dim = w["dim"]; grid = lat.brick_grid(world, dim)
    nglobal = (w["n"],) * dim if w.get("strong") else tuple(w["n"] * g for g in grid)
    rows = int(np.prod(nglobal)); per = rows // world
    nnz_row = 93 if dim == 3 else 25                         # ideal stored entries per row
    return dict(workload=wname, description=w["desc"], rows=rows, rows_per_gpu=per, bricks="x".join(map(str, grid)), lattice="x".join(map(str, nglobal)),
                system=w.get("system", "poisson"), precond=w["prec"], solver=w["solver"], restart=50, tolerance=1e-8,
                operators="antisymmetric" if w.get("anti", True) else "corrected (Gc/Lc)", jitter_dx=w["jitter"],
                l2="inputs larger than L2: matrix stream ~%.0f MB per SpMV per GPU, Krylov basis ~%.0f MB per GPU (L2 = 126 MB)" % ((12.0 * nnz_row + 20.0) * per / 1e6, 8e-6 * per * 101))


_PROBLEMS = {}


def make_problem(w, n, lat, lo=None, nloc=None, nglobal=None):
    dim = w["dim"]; ng = nglobal if nglobal is not None else (n,) * dim
    dx = 2.0 * np.pi / ng[0]
    key = (dim, tuple(ng), None if lo is None else tuple(lo), None if nloc is None else tuple(nloc), w["rs2"], w["jitter"], w.get("origin", 0.0))
    if key not in _PROBLEMS:                                  # p8m and c3 share one 200^3 lattice: generated once (one big lattice is kept at a time)
        if int(np.prod(nloc if nloc is not None else ng)) >= 500000:
            for k in [k for k, v in _PROBLEMS.items() if v["nlocal"] >= 500000]:
                del _PROBLEMS[k]
        _PROBLEMS[key] = lat.make_brick(dim, ng, dx, lo=lo, nloc=nloc, rs2=w["rs2"], jitter=w["jitter"], origin=w.get("origin", 0.0))
    P = _PROBLEMS[key]
    xw = P["xw"]
    v = lat.tgv_velocity(xw)
    # v* of a real step is not the analytic vortex: add a broadband (per-particle, tag-hashed => identical on every rank
    # that holds a copy) perturbation so that the RHS -div(v*) excites the whole spectrum.  On the unperturbed periodic
    # lattice the operator is translation invariant and the smooth vortex is (nearly) an eigenvector: GMRES would stop
    # after 2-3 iterations and the benchmark would not exercise the Krylov loop at all.
    for k in range(dim):
        v[:, k] += 0.05 * (2.0 * lat._hash01(P["gidx"] + 1, 100 + k) - 1.0)
    F = dict(vstar=v, density=np.ones(len(xw)), viscosity=np.full(len(xw), 0.1))      # rho = 1, nu = 0.1 (taylor-green-vortex-2d.lmp:138-140)
    dt = (0.05 * dx / 0.1) if dim == 2 else (0.1 * 1.5 * dx / 0.1)                    # taylor-green-vortex-{2d,3d}.lmp dt
    return P, F, dt


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False); self.p = None; self.gpu = gpu

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            t = [s.strip() for s in line.split(",")]
            if len(t) < 7:
                continue
            try:
                sm.append(float(t[0])); mx.append(float(t[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), t[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        busy = [s for s in sm if s > 0.5 * max(sm)] or sm
        return dict(sm_mhz=float(np.median(busy)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))


def nvlink_counters(gpu):
    """Cumulative NVLink data counters of one GPU in bytes (nvidia-smi nvlink -gt d: per-link 'Data Tx/Rx: N KiB'), or None."""
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(gpu)], capture_output=True, text=True, timeout=20).stdout
    except Exception:
        return None
    tx = rx = 0; seen = False
    for line in out.splitlines():
        t = line.replace(":", " ").split()
        if "Tx" in t and "KiB" in t:
            tx += int(t[t.index("KiB") - 1]) * 1024; seen = True
        elif "Rx" in t and "KiB" in t:
            rx += int(t[t.index("KiB") - 1]) * 1024; seen = True
    return dict(tx=tx, rx=rx) if seen else None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------------------------------
# GMRES iterations of the FULL-SIZE workloads (measured by the B200 arm, BENCH_r01 / profiles/; on every size where both ran
# the CPU oracle needs the same count, e.g. 188 = 188 on the 1M-particle configs[1]).  The CPU arm's bounded sample pins its
# solve to this count, so that a sampled row costs what a row of the full workload costs.
FULL_ITERS = {"p8m": 452, "c2": 188}


def cpu_step(O, P, F, dt, w, pin_iters=None):
    """One step of the CPU path (oracle port): pre-computation, graph, assembly, Krylov solve.  pin_iters: run exactly that
    many Krylov iterations (Convergence Tolerance 0, Maximum Iterations = pin_iters) instead of stopping at 1e-8."""
    t0 = time.perf_counter()
    o = O.Oracle(P, kind="port")
    o.set_field(O.F_VSTAR, F["vstar"]); o.set_field(O.F_DENSITY, F["density"])
    o.compute_pre()
    rp, col = o.graph()
    b = o.ns_poisson(dt)
    A = o.matrix()
    t1 = time.perf_counter()
    nl = P["nlocal"]
    colL = O.tags_to_local(col, P["tag"][:nl])
    prec = {"point relaxation": O.PREC_JACOBI, "ILU": O.PREC_ILU0, "Chebyshev": O.PREC_CHEBYSHEV, "ML": O.PREC_AMG}[w["prec"]]
    kw = dict(tol=0.0, max_iters=int(pin_iters)) if pin_iters else {}
    x, info = O.krylov_solve(rp, colL, A, b, params=O.krylov_params(precond=prec, row_gid=P["tag"][:nl], **kw), null_mask=np.ones(nl, dtype=np.int32), use_null=True)
    t2 = time.perf_counter()
    o.close()
    return dict(assemble_s=t1 - t0, solve_s=t2 - t1, iters=info["iters"], converged=info["converged"], nnz=len(col))


def cpu_sample_edge(args, w):
    """Lattice edge of the CPU arm's bounded sample: ~10-20 s of CPU work per step on the bench box (16-32 host cores)."""
    return args.cpu_n or (min(80, w["n"]) if w["dim"] == 3 else w["n"])


def cpu_sample_text(w, n, rows, r, pinned):
    full = f"{w['n']}^{w['dim']}"
    how = (f"solve pinned to the full workload's {r['iters']} GMRES iterations (tolerance 0), so a sampled row costs what a row of the full workload costs"
           if pinned else f"{r['iters']} GMRES iterations to 1e-8")
    return (f"{w['dim']}-D {n}^{w['dim']} = {rows} rows of the same periodic lattice / physics / solver as the {full} workload; every step = "
            f"pre-computation + graph + assembly + solve, {how}")


def run_reference(args, w, lat):
    """CPU arm.  Runs on rank 0 only; all host threads (torchrun exports OMP_NUM_THREADS=1: overridden here)."""
    threads = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(threads)                 # unconditional, and before the OpenMP runtime loads
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    O.build(("port",))
    threads = O.set_num_threads(threads)
    budget = float(os.environ.get("ISPH_REF_BUDGET_S", "540"))   # the whole --steps K --warmup W run ends within a few minutes
    t_start = time.perf_counter()
    n = cpu_sample_edge(args, w)
    pin = FULL_ITERS.get(args.workload) if n != w["n"] else None
    P, F, dt = make_problem(w, n, lat)
    cpu_step(O, P, F, dt, w, pin)                                # one warm-up step (page faults, OpenMP pool)
    ts, last = [], None
    for k in range(args.steps):
        t = time.perf_counter(); last = cpu_step(O, P, F, dt, w, pin); ts.append(time.perf_counter() - t)
        if time.perf_counter() - t_start + 1.5 * ts[-1] > budget:
            break
    sec = float(np.mean(ts)); val = P["nlocal"] / sec / 1e6
    world = int(os.environ.get("WORLD_SIZE", "1"))
    sample = cpu_sample_text(w, n, P["nlocal"], last, pin is not None)
    line = dict(metric="sph_poisson_step_throughput", value=val, unit="Mrow/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=sec * 1e3,
                higher_is_better=True, scaling="strong" if w.get("strong") else "weak", vs_baseline=None, dtype="f64", data="synthetic", impl="reference",
                config=workload_config(args.workload, w, world, lat),
                cpu_baseline=dict(value=val, unit="Mrow/s", cores=threads, kind="port", sample=sample, sample_rows=P["nlocal"], sample_iters=last["iters"],
                                  steps_timed=len(ts), sample_ms_per_step=sec * 1e3,
                                  note="CPU restatement of the reference path (oracle port: reference functor algorithms + Belos/Ifpack semantics), OpenMP over rows; "
                                       "Trilinos/LAMMPS/MPI are not installable here (DESIGN.md §5)"),
                e2e=dict(value=val, unit="Mrow/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                assemble_ms=last["assemble_s"] * 1e3, solve_ms=last["solve_s"] * 1e3,
                full_workload_ms_per_step=sec * 1e3 * (float(w["n"]) ** w["dim"]) / P["nlocal"])
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
def krylov_bytes_per_solve(n, nnz, iters, second, m=50):
    """Algorithmic HBM bytes of one flexible GMRES(m)+Jacobi solve on one GPU: per Arnoldi step with basis
    size nv: SpMV 12 nnz + 20 n; pass-0 coefficients 8 n (nv+2); first update 8 n (nv+3); and, when the DGKS test asks for
    it (`second` of the `iters` steps), pass-1 coefficients — 8 n (nv+1) for nv <= 8, nothing extra for nv > 8 where they
    are formed in the same sweep as the first update — and the second update inside the closing sweep 8 n nv; closing
    sweep (normalise + Jacobi) 8 n 4.  The x update at the end of a cycle: 8 n (ncol + 2)."""
    tot = 0.0; frac2 = second / max(iters, 1)
    for it in range(iters):
        nv = it % m + 1
        tot += 12.0 * nnz + 20.0 * n + 8.0 * n * ((nv + 2) + (nv + 3) + 4) + frac2 * 8.0 * n * ((nv + 1 if nv <= 8 else 0) + nv)
    cycles = (iters + m - 1) // m
    tot += cycles * 8.0 * n * (min(iters, m) + 2)
    return tot


def measure(args, wname, steps, isph, lat, torch, dist, rank, world, local_rank, nccl_id, with_cpu, warmup=None):
    w = dict(WORKLOADS[wname]); warmup = max(args.warmup, 3) if warmup is None else max(warmup, 3)
    if args.n:
        w["n"] = args.n
    # ---- this rank's brick of the periodic lattice (weak: w['n']^dim rows per GPU; strong: the global lattice is split)
    dim = w["dim"]; grid = lat.brick_grid(world, dim)
    nglobal = (w["n"],) * dim if w.get("strong") else tuple(w["n"] * g for g in grid)
    lo, nloc = lat.brick_of_rank(rank, grid, nglobal)
    P, F, dt = make_problem(w, w["n"], lat, lo=lo, nloc=nloc, nglobal=nglobal)
    nl = P["nlocal"]

    # host inputs in pinned memory (what LAMMPS would hand over each step)
    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory(); return t, t.numpy()
    hx = pin(P["x"]); htype = pin(P["type"]); htag = pin(P["tag"]); hneigh = pin(P["neigh"]); hnoff = pin(P["noff"]); hil = pin(P["ilist"])
    hv = pin(F["vstar"]); hrho = pin(F["density"]); hsol = pin(np.zeros(nl))
    helm = w.get("system") == "helmholtz"; pbs = w.get("system") == "pb"
    if pbs:
        xw_ = P["xw"]; s_ = np.sin(xw_[:, 0]) * np.cos(xw_[:, 1])
        hex_ = pin((-2.0 * s_ - np.sinh(s_))[:nl]); hpsi = pin(np.zeros(len(xw_))); heps = pin(np.ones(len(xw_))); hsol = pin(np.zeros(len(xw_)))
    if helm:
        hnu = pin(F["viscosity"]); hvn = pin(np.asfortranarray(F["vstar"][:nl, :dim]).reshape(-1, order="F")); hsol = pin(np.zeros((nl, dim), order="F").reshape(-1, order="F"))

    c = isph.Context(local_rank, world, rank, nccl_id)
    stream = torch.cuda.Stream(); c.set_stream(stream.cuda_stream)
    c.pair_coeff(dim, (0, isph.KIND_FLUID), 1.5 * P["dx"])
    c.solver_param("Solver Type", w["solver"]); c.precond_param("Precond Type", w["prec"]); c.precond_param("Overlap Level", 0); c.precond_param("fact: level-of-fill", 0)
    anti = w.get("anti", True)
    blk = None
    if w.get("blocks"):          # Ifpack-rank-equivalent bricks inside this GPU's brick
        nb = w["blocks"]; li = P["gidx"][:nl]; ng = nglobal
        gx = li % ng[0] - lo[0]; gy = (li // ng[0]) % ng[1] - lo[1]; gz = (li // (ng[0] * ng[1])) - (lo[2] if dim == 3 else 0)
        blk = ((gx * nb) // nloc[0] + nb * ((gy * nb) // nloc[1]) + (nb * nb * ((gz * nb) // nloc[2]) if dim == 3 else 0)).astype(np.int32)
    if w.get("global_blocks"):   # the Ifpack-rank partition of the GLOBAL lattice (4x4x4 bricks of 50^3), whatever the GPU count
        nb = w["global_blocks"]; li = P["gidx"][:nl]; ng = nglobal
        bx = (li % ng[0]) * nb // ng[0]; by = ((li // ng[0]) % ng[1]) * nb // ng[1]; bz = (li // (ng[0] * ng[1])) * nb // ng[2]
        blk = (bx + nb * (by + nb * bz)).astype(np.int32)

    def upload(device_list=False):
        c.atoms_set(nl, P["nghost"], hx[1], htype[1], htag[1])
        if device_list:
            c.neighbors_build()                              # the list is built on the device from the atoms: nothing but atoms and fields cross PCIe
        else:
            c.neighbors_set_packed(hil[1], hnoff[1], hneigh[1])
        c.field_set(isph.F_VSTAR, hv[1]); c.field_set(isph.F_DENSITY, hrho[1])
        if helm:
            c.field_set(isph.F_VELOCITY, hv[1]); c.field_set(isph.F_VISCOSITY, hnu[1])
        if pbs:
            c.field_set(isph.F_EPS, heps[1]); c.field_set(isph.F_PSI0, hpsi[1])
    h2d_bytes = hx[1].nbytes + htype[1].nbytes + htag[1].nbytes + hneigh[1].nbytes + hnoff[1].nbytes + hil[1].nbytes + hv[1].nbytes + hrho[1].nbytes
    if pbs:
        h2d_bytes += heps[1].nbytes + 2 * hpsi[1].nbytes + hex_[1].nbytes   # eps, psi0, psi (initial guess), source
    if helm:
        h2d_bytes += hv[1].nbytes + hnu[1].nbytes + 2 * hvn[1].nbytes      # + load and initial guess (both = v^n) written every step
    d2h_bytes = hsol[1].nbytes
    solve_label = "Helmholtz" if helm else ("PoissonBoltzmannJacobian" if pbs else "Poisson")

    def pb_step():
        """pair_isph.cpp:572-600: Newton iteration for psi (initial guess written every step, result read back)"""
        c.graph_invalidate()
        c.compute_pre()
        c.graph_build()
        c.create_solution(None, 1); c.create_load(None, 1)
        c.field_set(isph.F_PSI, hpsi[1])
        c.set_matrix_is_singular(False)
        r = c.pb_newton(extra_f=hex_[1])
        c.call("isph_field_get", isph.F_PSI, isph._d(hsol[1]))
        c.matrix_invalidate()
        return dict(iters=r["linear_iters"], relres=r["normf"], converged=r["converged"], newton_iters=r["newton_iters"])

    def helmholtz_step():
        """pair_isph.cpp:924-982: x = b = v^n (transposed view), computeHelmholtz, solveHelmholtz (dim right-hand sides)"""
        c.graph_invalidate()
        c.compute_pre()
        c.graph_build()
        c.create_solution(None, dim); c.create_load(None, dim)
        c.call("isph_solver_load_set", isph._d(hvn[1]), nl); c.call("isph_solver_solution_set", isph._d(hvn[1]), nl)
        c.ns_helmholtz(dt, w["theta"], anti=anti)
        c.set_matrix_is_singular(False)
        st = c.solve(True, "Helmholtz")
        c.call("isph_solver_solution_get", isph._d(hsol[1]), nl)
        c.matrix_invalidate()
        return st

    def device_step():
        """inputs resident in HBM; everything else rebuilt (end-of-step delete, pair_isph.cpp:1351-1372)"""
        c.graph_invalidate()
        c.compute_pre()
        c.graph_build()
        c.create_solution(hsol[1], 1); c.create_load(None, 1)
        c.ns_poisson(dt, anti=anti)
        if blk is not None:
            c.precond_set_blocks(blk)
        c.set_matrix_is_singular(True); c.set_initial_solution(isph.INIT_ZERO)
        st = c.solve(True, "Poisson")
        c.matrix_invalidate()
        return st

    if helm:
        device_step = helmholtz_step
    if pbs:
        device_step = pb_step

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    upload()
    st = None
    for _ in range(warmup):
        st = device_step()
    # ---- timed region 1: device-resident inputs
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    c.timer_reset(); c.profile_spmv(True); c.profile_spmv_get()
    launches0 = c.launches
    nvl0 = nvlink_counters(local_rank) if (world > 1 and rank == 0) else None
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(steps):
            st = device_step()
        e1.record(stream)
    barrier()
    ms_dev = e0.elapsed_time(e1) / steps
    nvl1 = nvlink_counters(local_rank) if nvl0 is not None else None
    spmv_ms, spmv_cnt = c.profile_spmv_get(); prec_ms, prec_cnt = c.profile_precond_get(); c.profile_spmv(False)
    ilu = c.precond_info() if w["prec"] == "ILU" else None
    ml = c.precond_ml_info() if w["prec"] == "ML" else None
    launches = (c.launches - launches0) // steps
    timers = {k: c.timer_ms(k) / steps for k in ("computeVolumes", "computeGradientCorrection", "computeLaplacianCorrection", "computeGraph", "precondCreate", "solve" + solve_label) +
              (("computeFPoissonBoltzmann", "computeJacobianPoissonBoltzmann") if pbs else ("compute" + solve_label,))}
    if w["prec"] == "ILU":
        timers.update({k: c.timer_ms(k) / steps for k in ("iluPattern", "iluLevels", "iluFactor", "iluPermute")})
    c.timer_reset()
    clocks = sampler.stop() if rank == 0 else None
    # ---- timed region 2: end to end through the C ABI with host buffers (H2D of the step's inputs, D2H of the solution)
    barrier()
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(steps):
            upload(); st = device_step()
        e1.record(stream)
    barrier()
    ms_e2e = e0.elapsed_time(e1) / steps
    wall_e2e = (time.perf_counter() - t0) * 1e3 / steps
    e2e_timers = {k: c.timer_ms(k) / steps for k in ("h2dAtoms", "h2dNeighbors", "haloSetup")}
    ms_e2e = max(ms_e2e, wall_e2e)          # host-side packing/validation inside the ABI calls is part of the end-to-end cost
    # ---- timed region 3: end to end with the neighbor list built on the device (isph_neighbors_build) instead of uploaded
    upload(device_list=True); device_step()                   # one untimed step: the builder's work buffers are allocated here
    c.timer_reset(); barrier()
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(steps):
            upload(device_list=True); st_dl = device_step()
        e1.record(stream)
    barrier()
    ms_e2e_dl = max(e0.elapsed_time(e1) / steps, (time.perf_counter() - t0) * 1e3 / steps)
    dl_build_ms = c.timer_ms("buildNeighbors") / steps
    dl_same = (st_dl["iters"] == st["iters"])
    if world > 1:
        t = torch.tensor([ms_dev, ms_e2e, timers["solve" + solve_label], ms_e2e_dl], dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms_dev, ms_e2e, solve_ms, ms_e2e_dl = t.tolist()
    else:
        solve_ms = timers["solve" + solve_label]
    nnz = c.nnz
    fp64_peak = c.measure_fp64_peak() if rank == 0 else None
    halo = None
    if world > 1:
        try:
            halo = c.halo_counts()
        except Exception:                                    # reporting only: never let it take the benchmark line down
            halo = None
    if world > 1:
        t = torch.tensor([float(nnz), float(nl)], dtype=torch.float64); dist.all_reduce(t); nnz_g, rows_g = t.tolist()
    else:
        nnz_g, rows_g = float(nnz), float(nl)
    c.close()
    if rank != 0:
        return None

    peak, peak_src = measured_peak()
    spmv_bytes = 12.0 * nnz + 20.0 * nl                      # 12 B per nonzero + 20 B per row, this rank's launch
    avg = spmv_ms / max(spmv_cnt, 1)
    ach = spmv_bytes / (avg * 1e-3) / 1e9 if avg > 0 else 0.0
    traffic = None
    try:                                                     # dram bytes per SpMV launch from the committed ncu --set full capture of this workload (1 GPU)
        tj = json.load(open(os.path.join(ROOT, "profiles", "spmv_traffic.json")))
        if world == 1 and not args.n and wname in tj:
            traffic = float(tj[wname]["dram_bytes_per_launch"])
    except Exception:
        pass
    result = dict(iters=st["iters"], relres=st["relres"], converged=st["converged"], rows=int(rows_g), nnz=int(nnz_g), rows_this_gpu=nl, **({"newton_iters": st["newton_iters"]} if pbs else {}))
    line = dict(metric="sph_poisson_step_throughput", value=rows_g / (ms_dev * 1e-3) / 1e6, unit="Mrow/s", n_gpus=world, steps=steps, warmup=warmup,
                ms_per_step=ms_dev, higher_is_better=True, scaling="strong" if w.get("strong") else "weak", vs_baseline=None, dtype="f64", data="synthetic",
                config=workload_config(wname, w, world, lat), result=result,
                e2e=dict(value=rows_g / (ms_e2e * 1e-3) / 1e6, unit="Mrow/s", ms_per_step=ms_e2e, h2d_bytes_per_step=int(h2d_bytes), d2h_bytes_per_step=int(d2h_bytes),
                         upload_ms=e2e_timers),
                e2e_device_neighbors=dict(value=rows_g / (ms_e2e_dl * 1e-3) / 1e6, unit="Mrow/s", ms_per_step=ms_e2e_dl, build_neighbors_ms=dl_build_ms,
                                          h2d_bytes_per_step=int(h2d_bytes - hneigh[1].nbytes - hnoff[1].nbytes - hil[1].nbytes), d2h_bytes_per_step=int(d2h_bytes), same_iterations=bool(dl_same),
                                          note="same step with the full neighbor list built on the device from the uploaded atoms (isph_neighbors_build, SURVEY.md §8f.4) instead of "
                                               "uploaded from the host; `e2e` above is the reference-shaped hand-over (LAMMPS builds the list on the host)"),
                gpu_launches=int(launches), clocks=clocks,
                roofline=dict(bound="hbm", kernel="k_spmv_sell<1>", achieved=ach, peak=peak, unit="GB/s", frac=ach / peak, traffic=traffic, peak_source=peak_src,
                              traffic_source="dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture of this workload "
                                             "(profiles/spmv_traffic.json); not re-measured in this run" if traffic else None,
                              launches_timed=int(spmv_cnt), avg_launch_ms=avg, algorithmic_bytes_per_launch=spmv_bytes,
                              share_of_step=spmv_ms / steps / ms_dev,
                              note="per-launch time from CUDA events on the launching stream; at n_gpus > 1 it includes the NVLink import of the halo"),
                breakdown_ms=timers, ms_per_iter=solve_ms / max(st["iters"], 1))
    if fp64_peak:                                            # assembly against both rooflines (~160 flop per in-cut pair, ~1530 B per row in 3-D)
        key = "compute" + solve_label; t_asm = timers.get(key, 0.0)
        if t_asm > 0 and not pbs:
            fl = 160.0 * nnz if dim == 3 else 90.0 * nnz; by = 4.0 * float(P["noff"][-1]) + (44.0 + (0.0 if anti else 120.0)) * nl + 12.0 * nnz
            line["assembly_roofline"] = dict(kernel="k_laplacian_rows + system rows (" + key + ")", ms=t_asm, flops_per_launch=fl, bytes_per_launch=by,
                                             achieved_tflops=fl / (t_asm * 1e-3) / 1e12, fp64_peak_tflops=fp64_peak, frac_fp64=fl / (t_asm * 1e-3) / 1e12 / fp64_peak,
                                             achieved_gbs=by / (t_asm * 1e-3) / 1e9, frac_hbm=by / (t_asm * 1e-3) / 1e9 / peak,
                                             note="flop and byte counts are SURVEY.md §8(d)'s per-pair / per-row figures x this launch's pairs / rows (divisions and square roots "
                                                  "counted as one flop each: the FP64-pipe time they take is several times that); fp64 peak = FMA loop measured in this run (isph_measure_fp64_peak)")
        line["fp64_peak_tflops"] = fp64_peak
    if ilu is not None and prec_cnt:                         # triangular solves of the block-Jacobi ILU apply (12 nnz_factor + 32 n, and the level count)
        pav = prec_ms / prec_cnt; pb = 12.0 * ilu["factor_nnz"] + 32.0 * nl
        line["ilu_roofline"] = dict(bound="hbm", kernel="k_ilu_solve", achieved=pb / (pav * 1e-3) / 1e9, peak=peak, unit="GB/s", frac=pb / (pav * 1e-3) / 1e9 / peak,
                                    algorithmic_bytes_per_launch=pb, avg_launch_ms=pav, launches_timed=int(prec_cnt), share_of_step=prec_ms / steps / ms_dev,
                                    factor_nnz=ilu["factor_nnz"], levels_lower=ilu["levels_lower"], levels_upper=ilu["levels_upper"],
                                    us_per_level=1e3 * pav / max(ilu["levels_lower"] + ilu["levels_upper"], 1),
                                    note="latency-bound by the dependency levels of the forward + backward sweeps, not by bytes (DESIGN.md §3)")
    if ml is not None:                                       # hierarchy of the multilevel preconditioner (rebuilt every solve, like every preconditioner of the reference)
        line["ml_hierarchy"] = dict(levels=ml["levels"], rows=ml["rows"], nnz=ml["nnz"], lambda_max=ml["lambda_max"], setup_ms_last_solve=ml["setup_ms"],
                                    spmv_per_iteration=spmv_cnt / max(steps, 1) / max(st["iters"], 1),
                                    note="finest level = this rank's SELL matrix; coarser levels replicated on every GPU; 'parity unpinned' against ML by construction (DESIGN.md)")
    if halo is not None:                                     # NVLink side of the roofline (rank 0's brick): bytes that leave this GPU per operator apply
        spmv_per_s = spmv_cnt / max(steps, 1) / max(ms_dev * 1e-3, 1e-12)
        line["nvlink"] = dict(what="halo import of one SpMV on rank 0 (NVLink peer stores, 8 B per value) and the two all-reduces of an Arnoldi step (<= 53 doubles to each peer)",
                              send_values_per_spmv=halo["nsend"], recv_values_per_spmv=halo["nhalo"], peers=halo["npeers"],
                              bytes_sent_per_spmv=8 * halo["nsend"], bytes_sent_per_second=8.0 * halo["nsend"] * spmv_per_s,
                              peak_gbs_per_direction=900.0, frac_of_nvlink_peak=8.0 * halo["nsend"] * spmv_per_s / 900e9,
                              note="0.1 % of the local SpMV traffic: the exchange is latency-, not bandwidth-bound (DESIGN.md §6)")
        if nvl0 is None:
            line["nvlink"]["measured"] = None
            line["nvlink"]["measured_unavailable"] = ("`nvidia-smi nvlink -gt d` reports 'Data Tx: N/A' for every link on this driver (profiles/r02_nvlink_counters.txt) and ncu may "
                                                      "only profile single-GPU commands here, so the link bytes are the plan's arithmetic, not a counter reading")
        if nvl0 is not None and nvl1 is not None:            # MEASURED: this GPU's NVLink data counters around the timed region (nvidia-smi nvlink -gt d)
            dtx = nvl1["tx"] - nvl0["tx"]; drx = nvl1["rx"] - nvl0["rx"]
            line["nvlink"]["measured"] = dict(tx_bytes_per_step=dtx / steps, rx_bytes_per_step=drx / steps, tx_bytes_per_spmv=dtx / max(spmv_cnt, 1), rx_bytes_per_spmv=drx / max(spmv_cnt, 1),
                                              tx_gbs=dtx / max(steps * ms_dev * 1e-3, 1e-12) / 1e9, frac_of_nvlink_peak=dtx / max(steps * ms_dev * 1e-3, 1e-12) / 900e9,
                                              how="difference of the per-link Data Tx/Rx counters of rank 0's GPU (nvidia-smi nvlink -gt d) around the timed region: halo values, "
                                                  "peer all-reduce messages, sentinel re-arming is local; includes the per-step field forwards over NCCL")
    if w["prec"] == "point relaxation" and w["solver"] == "Block GMRES" and not pbs:
        kb = krylov_bytes_per_solve(nl, nnz, st["iters"], st.get("second_passes", st["iters"]))
        sg = kb / (solve_ms * 1e-3) / 1e9
        line["solve_roofline"] = dict(bound="hbm", what="whole GMRES(50)+Jacobi solve on one GPU of the job: SpMV + Gram-Schmidt sweeps + solution update (DESIGN.md §3), rank 0's rows",
                                      achieved=sg, peak=peak, unit="GB/s", frac=sg / peak, algorithmic_bytes_per_solve=kb, solve_ms=solve_ms,
                                      iters=st["iters"], second_pass_steps=st.get("second_passes"))
    if with_cpu:                                             # the same bounded sample the reference arm times (see run_reference), two steps
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as O
        O.build(("port",))
        threads = O.set_num_threads(os.cpu_count() or 1)
        ncpu = cpu_sample_edge(args, w); pin = FULL_ITERS.get(wname) if ncpu != w["n"] else None
        Pc, Fc, dtc = make_problem(w, ncpu, lat)
        cpu_step(O, Pc, Fc, dtc, w, pin)                     # warm
        t = time.perf_counter(); r = cpu_step(O, Pc, Fc, dtc, w, pin); sec = time.perf_counter() - t
        line["cpu_baseline"] = dict(value=Pc["nlocal"] / sec / 1e6, unit="Mrow/s", cores=threads, kind="port",
                                    sample=cpu_sample_text(w, ncpu, Pc["nlocal"], r, pin is not None) + f"; one step timed, {sec:.1f} s")
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1); ap.add_argument("--steps", type=int, default=5); ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"]); ap.add_argument("--workload", default="p8m", choices=sorted(WORKLOADS))
    ap.add_argument("--n", type=int, default=0, help="lattice edge override (per GPU)"); ap.add_argument("--cpu-n", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true"); ap.add_argument("--no-secondary", action="store_true"); ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    lat = importlib.import_module("implicit-sph_b200.lattice")
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            w = dict(WORKLOADS[args.workload])
            if args.n:
                w["n"] = args.n
            run_reference(args, w, lat)
        return

    import torch
    import torch.distributed as dist
    isph = importlib.import_module("implicit-sph_b200")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("gloo", init_method="env://")          # plumbing only: unique-id broadcast, barrier, max-over-ranks

    def fresh_id():                                          # one NCCL communicator per context: a new unique id for each
        if world == 1:
            return None
        idt = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (isph.C.c_ubyte * 128)(); assert isph.lib().isph_nccl_unique_id(buf) == 0, "NCCL unique id"
            idt = torch.tensor(list(buf), dtype=torch.uint8)
        dist.broadcast(idt, 0); return bytes(idt.tolist())

    parity = None
    if world > 1 and not args.no_parity:                     # multi-GPU parity against the CPU oracle BEFORE anything is timed (tests/ may use the oracle)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import multi_gpu_check
        parity = multi_gpu_check.run_check(isph, lat, torch, dist, rank, world, local_rank, fresh_id(), quiet=True)
        if not parity["pass"]:
            if rank == 0:
                print(json.dumps(dict(metric="sph_poisson_step_throughput", value=None, n_gpus=world, parity=parity, error="multi-GPU parity check failed: nothing was timed")))
            dist.destroy_process_group()
            raise SystemExit(3)

    line = measure(args, args.workload, args.steps, isph, lat, torch, dist, rank, world, local_rank, fresh_id(), with_cpu=(not args.no_cpu_baseline and world == 1))
    if parity is not None and rank == 0:
        line["parity"] = parity
    # the other BASELINE configs beside the headline line: same code path, measured in the same run, at this GPU count
    if args.workload == "p8m" and not args.n and not args.no_secondary:
        sec_steps = max(1, min(args.steps, 3))
        todo = ([("configs1_c2", "c2"), ("configs1_c2_ml", "c2_ml")] if world == 1 else []) + [("p8m_ml", "p8m_ml"), ("configs2_c3", "c3"), ("configs3_c4", "c4s"), ("configs3_c4_ml", "c4s_ml"), ("configs4_c5", "c5")]
        for key, wn in todo:
            sec = measure(args, wn, min(args.steps, 5) if wn == "c2" else sec_steps, isph, lat, torch, dist, rank, world, local_rank, fresh_id(), with_cpu=False, warmup=3)
            if rank != 0:
                continue
            blk = {k: sec[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "e2e", "gpu_launches", "breakdown_ms", "ms_per_iter", "result", "ilu_roofline", "nvlink", "ml_hierarchy") if k in sec}
            blk.update(workload=sec["config"]["description"], rows=sec["config"]["rows"], iters=sec["result"]["iters"], converged=sec["result"]["converged"],
                       spmv_roofline_frac=sec["roofline"]["frac"], spmv_gbs=sec["roofline"]["achieved"], spmv_bytes_per_launch=sec["roofline"]["algorithmic_bytes_per_launch"],
                       solve_roofline_frac=sec.get("solve_roofline", {}).get("frac"))
            line[key] = blk
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
