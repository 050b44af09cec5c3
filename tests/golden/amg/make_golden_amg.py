"""Regenerates tests/golden/amg/amg_restatement.npz: a self-consistency pin of the multilevel stand-in's sequential restatement
(oracle/amg_oracle.h).  NOT a reference pin — ML is not available (parity unpinned); the fixture freezes the algorithm's
discrete decisions (aggregates, level sizes) and numbers (lambda_max, one V-cycle, iteration count) so that a later change to the
restatement or to the device code (the GPU tests compare the two) cannot drift unnoticed.

    python tests/golden/amg/make_golden_amg.py
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402


def problem():
    n1 = 40; e = np.ones(n1); T = sp.diags([-e[:-1], 2 * e, -e[:-1]], [-1, 0, 1]); S = sp.diags([0.2 * e[:-1], -0.2 * e[:-1]], [1, -1])
    A = sp.csr_matrix(sp.kron(sp.eye(n1), T) + sp.kron(T, sp.eye(n1)) + 0.002 * sp.eye(n1 * n1) + sp.kron(sp.eye(n1), S)); A.sort_indices()
    rng = np.random.default_rng(11); return A, rng.standard_normal(A.shape[0]), rng.standard_normal(A.shape[0])


def compute():
    A, b, r = problem(); prm = O.krylov_params(precond=O.PREC_AMG, amg_threshold=0.1, amg_max_coarse=20)
    h = O.amg_hierarchy(A.indptr, A.indices, A.data, prm)
    z, _ = O.precond_apply(A.indptr, A.indices, A.data, r, prm)
    x, info = O.krylov_solve(A.indptr, A.indices, A.data, b, params=prm)
    return dict(agg=h["agg"], rows=h["rows"], nnz=h["nnz"], lmax=h["lmax"], z=z, iters=np.array(info["iters"]), x=x)


if __name__ == "__main__":
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "amg", "amg_restatement.npz"), **compute())
    print("written")
