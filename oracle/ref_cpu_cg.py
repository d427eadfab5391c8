#!/usr/bin/env python
"""CPU baseline with the REFERENCE itself (NGSolve built by oracle/build_reference.sh into oracle/_ref/ngs):
the TaskManager-parallel C++ CGSolver + JacobiPrecond + SparseMatrix<double>::MultAdd of /root/reference, timed on
this host's cores on a bounded sample of the bench workload.  Test/bench infrastructure only (bench.py's cpu_baseline and
--impl reference legs run it in a subprocess because `import ngsolve` must precede numpy/torch, SURVEY.md 8c pitfall 4).

    source oracle/_ref/ngs/env.sh && python oracle/ref_cpu_cg.py --m 33 --iters 150 [--threads T]

The system is the one bench.py solves on the GPU (same generator, host loop), injected with SparseMatrixd.CreateFromCOO
(linalg/python_linalg.cpp:144-152).  Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import ngsolve                      # noqa: E402  (before numpy)
from ngsolve import la, BitArray, CGSolver, TaskManager, SetNumThreads

import numpy as np                  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=33)
    ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--iters", type=int, default=150)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--threads", type=int, default=0)
    a = ap.parse_args()
    T = a.threads or (os.cpu_count() or 1)
    ngsolve.ngsglobals.msg_level = 0
    SetNumThreads(T)                # one thread count per process: SparseMatrix::balance is frozen at construction
    from ngsolve_b200 import workloads as W
    box = W.FemBox(a.m, order=a.order)
    rp, col, val, rhs = box.host_csr()
    free = box.dof_info()[2].astype(bool)
    n = box.ndof
    nnz = int(rp[-1])
    t0 = time.perf_counter()
    with TaskManager():
        rows = np.repeat(np.arange(n, dtype=np.int32), np.diff(rp.astype(np.int64)))
        mat = la.SparseMatrixd.CreateFromCOO(rows, col, val, n, n)
        del rows
        assert mat.nze == nnz
        fd = BitArray(n)
        fd.Clear()
        for i in np.flatnonzero(free):
            fd.Set(int(i))
        jac = mat.CreateSmoother(fd)
        f = mat.CreateColVector()
        u = mat.CreateColVector()
        f.FV().NumPy()[:] = rhs
        setup = time.perf_counter() - t0
        # SpMV alone (BaseMatrix::Mult = SetZero + MultAdd)
        y = mat.CreateColVector()
        for _ in range(3):
            mat.Mult(f, y)
        reps = 20
        t0 = time.perf_counter()
        for _ in range(reps):
            mat.Mult(f, y)
        t_spmv = (time.perf_counter() - t0) / reps
        # Jacobi-PCG, the C++ CGSolver (linalg/cg.cpp:503-633); precision 0 -> runs exactly maxsteps iterations
        inv = CGSolver(mat, jac, precision=0.0, maxsteps=a.warmup)
        u.data = inv * f
        inv = CGSolver(mat, jac, precision=0.0, maxsteps=a.iters)
        t0 = time.perf_counter()
        u.data = inv * f
        dt = time.perf_counter() - t0
        steps = inv.GetSteps()
    its = steps - 1 if steps > a.iters else steps
    b_spmv = nnz * 12 + n * 20
    out = dict(kind="reference", ngsolve_version=ngsolve.__version__, threads=T, m=a.m, ndof=n, nnz=nnz, iterations=int(its), steps=int(steps),
               seconds=dt, it_per_s=its / dt, spmv_ms=t_spmv * 1e3, spmv_gbs=b_spmv / t_spmv / 1e9, setup_s=setup,
               u_norm=float(np.linalg.norm(u.FV().NumPy())))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
