#!/usr/bin/env python
"""Replays a matrix exported by a DROPS run (exportMTX / operator<< / out(), see drops_b200/matrix_io.py)
through the CUDA path and, with --check, through the oracle: SpMV bit-parity and solver parity.

  python tools/replay_matrix.py A.mtx [--solver pcg|cg|gmres|minres|bicgstab|gcr] [--pc jacobi|dummy]
         [--tol 1e-10] [--maxiter 1000] [--restart 100] [--rhs b.txt] [--check]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("matrix")
    ap.add_argument("--solver", default="pcg", choices=["pcg", "cg", "gmres", "minres", "bicgstab", "gcr"])
    ap.add_argument("--pc", default="jacobi", choices=["jacobi", "dummy"])
    ap.add_argument("--tol", type=float, default=1e-10)
    ap.add_argument("--maxiter", type=int, default=1000)
    ap.add_argument("--restart", type=int, default=100)
    ap.add_argument("--rhs", default=None, help="vector file in the format of out(os, VectorCL); default b = A*1")
    ap.add_argument("--check", action="store_true", help="also run the oracle (CPU) and compare")
    a = ap.parse_args()
    from drops_b200 import capi, matrix_io as mio
    n, m, rb, ci, va = mio.read_matrix(a.matrix)
    capi.init(int(os.environ.get("LOCAL_RANK", "0")))
    A = capi.Matrix(n, m, rb, ci, va)
    x1 = np.ones(m)
    if a.rhs:
        tok = open(a.rhs).read().split()
        b = np.array(tok[1:1 + int(tok[0])], dtype=np.float64)
    else:
        b = A.spmv(x1)
    pc = capi.PC_JACOBI if a.pc == "jacobi" else capi.PC_DUMMY
    x0 = np.zeros(n)
    run = {"pcg": lambda M: M.pcg(x0, b, a.maxiter, a.tol, True, pc),
           "cg": lambda M: M.cg(x0, b, a.maxiter, a.tol, True),
           "gmres": lambda M: M.gmres(x0, b, a.restart, a.maxiter, a.tol, True, pc),
           "minres": lambda M: M.minres(x0, b, a.maxiter, a.tol, True),
           "bicgstab": lambda M: M.bicgstab(x0, b, a.maxiter, a.tol, True, pc),
           "gcr": lambda M: M.gcr(x0, b, a.restart, a.maxiter, a.tol, True, pc)}[a.solver]
    t0 = time.perf_counter()
    out = run(A)
    dt = time.perf_counter() - t0
    rep = {"matrix": a.matrix, "rows": n, "cols": m, "nnz": int(len(va)), "format": A.format(), "solver": a.solver,
           "pc": a.pc, "gpu": {"converged": out[0], "iterations": out[1], "residual": out[2], "seconds": dt}}
    if a.check:
        from oracle import port
        P = port.PortMatrix(n, m, rb, ci, va)
        rep["spmv_bit_identical"] = bool(np.array_equal(A.spmv(x1), P.spmv(x1)))
        pk = 1 if a.pc == "jacobi" else 0
        ref = {"pcg": lambda: (P.pcg_jacobi if pk else P.pcg_dummy)(x0, b, a.maxiter, a.tol, True),
               "cg": lambda: P.cg(x0, b, a.maxiter, a.tol, True),
               "gmres": lambda: (P.gmres_jacobi if pk else P.gmres_dummy)(x0, b, a.restart, a.maxiter, a.tol, True),
               "minres": lambda: P.minres(x0, b, a.maxiter, a.tol, True),
               "bicgstab": lambda: P.bicgstab(x0, b, a.maxiter, a.tol, True, pk),
               "gcr": lambda: P.gcr(x0, b, a.restart, a.maxiter, a.tol, True, pk)}[a.solver]
        t0 = time.perf_counter()
        r = ref()
        rep["oracle"] = {"converged": r[0], "iterations": r[1], "residual": r[2], "seconds": time.perf_counter() - t0}
        rep["solution_rel_diff"] = float(np.linalg.norm(out[3] - r[3]) / max(np.linalg.norm(r[3]), 1e-300))
    print(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()
