"""CPU-side checks of bench.py's contract pieces that do not need a GPU: the reference arm's JSON line, the traffic model
behind `solve_roofline`, the brick decomposition used for N = 1, 2, 4, 8, and the loud failure without a CUDA device."""
import importlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-n", "16"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["dtype"] == "f64" and line["vs_baseline"] is None and line["scaling"] == "strong"
    assert line["config"]["workload"] == "p8m" and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    # same `config` as the B200 arm (the driver compares them): the FULL workload is named, the bounded sample is described in cpu_baseline
    bench = importlib.import_module("bench"); lat = importlib.import_module("implicit-sph_b200.lattice")
    assert line["config"] == json.loads(json.dumps(bench.workload_config("p8m", bench.WORKLOADS["p8m"], 1, lat))) and line["config"]["rows"] == 8000000
    cb = line["cpu_baseline"]
    assert cb["sample_rows"] == 16 ** 3 and cb["sample_iters"] == 452 and "452" in cb["sample"]      # the sample's solve is pinned to the full workload's iteration count
    assert cb["cores"] == (os.cpu_count() or 1)


def test_b200_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_solve_traffic_model():
    bench = importlib.import_module("bench")
    n, nnz = 1000, 100000
    one = bench.krylov_bytes_per_solve(n, nnz, 1, 1)             # nv = 1: SpMV + (3 + 4 + 4 + 2 + 1) n doubles + x update (1 + 2) n
    assert one == 12.0 * nnz + 20.0 * n + 8.0 * n * (3 + 4 + 4 + 2 + 1) + 8.0 * n * 3
    full = bench.krylov_bytes_per_solve(n, nnz, 50, 50)          # one full cycle: fused sweep for nv > 8
    want = sum(12.0 * nnz + 20.0 * n + 8.0 * n * ((nv + 2) + (nv + 3) + 4 + (nv + 1 if nv <= 8 else 0) + nv) for nv in range(1, 51)) + 8.0 * n * 52
    assert abs(full - want) <= 1e-9 * want
    assert bench.krylov_bytes_per_solve(n, nnz, 50, 0) < full    # no second passes => less traffic


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_bricks_tile_the_global_lattice(world):
    lat = importlib.import_module("implicit-sph_b200.lattice")
    dim = 3; ng = (8, 8, 8); grid = lat.brick_grid(world, dim)
    assert int(np.prod(grid)) == world
    seen = np.zeros(ng, dtype=np.int32)
    for r in range(world):
        lo, nloc = lat.brick_of_rank(r, grid, ng)
        seen[lo[0]:lo[0] + nloc[0], lo[1]:lo[1] + nloc[1], lo[2]:lo[2] + nloc[2]] += 1
    assert np.all(seen == 1)


def test_reference_arm_under_torchrun_prints_one_line_from_rank_0():
    """Driver contract for N > 1: launched with torchrun, rank 0 alone runs the CPU arm and prints the line; the other ranks exit 0."""
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29577",
                        os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--cpu-n", "12"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0]); assert line["impl"] == "reference" and line["n_gpus"] == 2
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm must still use every host core (the N > 1 reference arm timed out)
    assert line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    bench = importlib.import_module("bench"); lat = importlib.import_module("implicit-sph_b200.lattice")
    assert line["config"] == json.loads(json.dumps(bench.workload_config("p8m", bench.WORKLOADS["p8m"], 2, lat))) and line["config"]["bricks"] == "2x1x1"
