"""The reference-side view of the boundary: a C++ program using the SolverLin / PrecondWrapper adapter classes
(include/solver_lin_b200.h) in the call order of pair_isph.cpp:986-1026, with LAMMPS-layout neighbor pages."""
import os
import struct
import subprocess

import numpy as np
import pytest

from problems import make_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp):
    exe = os.path.join(tmp, "adapter_poisson")
    libdir = os.path.join(ROOT, "implicit-sph_b200")
    subprocess.run(["g++", "-O1", "-std=c++14", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "adapter_poisson.cpp"),
                    "-L", libdir, "-l:libisph_b200.so", f"-Wl,-rpath,{libdir}", "-o", exe], check=True)
    return exe


def test_adapter_compiles_against_the_abi(tmp_path):
    """CPU: the adapter header + ABI header are self-contained C++ and link against the shared library."""
    assert os.path.exists(_build(str(tmp_path)))


@pytest.mark.gpu
@pytest.mark.parametrize("package", ["Ifpack", "ML"])
def test_adapter_poisson_matches_oracle(tmp_path, package):
    import oracle as O
    P, F = make_case("jitter3d"); cs = P["case"]; nl, nall = P["nlocal"], P["nlocal"] + P["nghost"]
    rho = np.ones(nall); v = F["velocity"]
    fn = str(tmp_path / "in.bin"); out = str(tmp_path / "out.bin")
    nneigh = len(P["neigh"])
    with open(fn, "wb") as f:
        f.write(struct.pack("6i", P["dim"], nl, P["nghost"], nneigh & 0x7FFFFFFF, nneigh >> 31, 0))
        f.write(struct.pack("2d", 1.5 * P["dx"], cs["dt"]))
        for a in (P["x"], v, rho):
            f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
        for a in (P["type"], P["tag"], np.diff(P["noff"]).astype(np.int32), P["neigh"]):
            f.write(np.ascontiguousarray(a, dtype=np.int32).tobytes())
    exe = _build(str(tmp_path))
    r = subprocess.run([exe, fn, out, package], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    its = int([l for l in r.stdout.splitlines() if l.startswith("iterations")][0].split()[1])
    x = np.fromfile(out, dtype=np.float64)
    o = O.Oracle(P, kind="port"); o.set_field(O.F_VSTAR, v); o.set_field(O.F_DENSITY, rho); o.compute_pre(); rp, col = o.graph(); b = o.ns_poisson(cs["dt"]); A = o.matrix()
    xo, info = O.krylov_solve(rp, O.tags_to_local(col, P["tag"][:nl]), A, b, params=O.krylov_params(precond=O.PREC_JACOBI if package == "Ifpack" else O.PREC_AMG, row_gid=P["tag"][:nl]), null_mask=np.ones(nl, dtype=np.int32), use_null=True)
    assert abs(its - info["iters"]) <= 2 and np.linalg.norm(x - xo) / np.linalg.norm(xo) <= 1e-6
