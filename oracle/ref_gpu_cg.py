#!/usr/bin/env python
"""Time the REFERENCE's own device path (ngscuda, built by oracle/build_reference_cuda.sh) on this GPU: cuSPARSE
DevSparseMatrix::Mult (ngscuda/cuda_linalg.cpp:244-276), NGSolve's C++ CGSolver loop on UnifiedVectors and the graph-
captured DevCGSolver (ngscuda/cuda_krylov.cpp:19-203) -- "that kernel to beat on the same box" (SURVEY.md 8a20).
Run under the reference environment (source oracle/_ref/ngs/env.sh).  One JSON line.  Test/measurement infrastructure."""
import argparse
import faulthandler
import json
import os
import sys
import time

faulthandler.enable()
import ngsolve
from ngsolve import *          # noqa: F401,F403
from netgen.csg import unit_cube

ap = argparse.ArgumentParser()
ap.add_argument("--maxh", type=float, default=0.03)
ap.add_argument("--nref", type=int, default=0)
ap.add_argument("--order", type=int, default=3)
ap.add_argument("--iters", type=int, default=100)
ap.add_argument("--full", action="store_true")
args = ap.parse_args()
ngsolve.ngsglobals.msg_level = 0
SetNumThreads(os.cpu_count())


def log(*a):
    print(*a, file=sys.stderr, flush=True)


out = {"ngsolve": ngsolve.__version__, "maxh": args.maxh, "nref": args.nref, "order": args.order}
with TaskManager():
    mesh = Mesh(unit_cube.GenerateMesh(maxh=args.maxh))
    for _ in range(args.nref):
        mesh.Refine()
    fes = H1(mesh, order=args.order, dirichlet=".*")
    u, v = fes.TnT()
    a = BilinearForm(grad(u) * grad(v) * dx).Assemble()
    f = LinearForm(1 * v * dx).Assemble()
    jac = a.mat.CreateSmoother(fes.FreeDofs())
    out.update(ndof=fes.ndof, nnz=a.mat.nze)
    log("assembled", fes.ndof)
    # ngsolve/ngscuda.py refuses to load when the core was configured without CUDA; the compiled module is complete
    import ngsolve._ngscuda as ngscuda
    log("ngscuda imported")
    adev = a.mat.CreateDeviceMatrix()
    log("adev", type(adev).__name__)
    jdev = jac.CreateDeviceMatrix()
    log("jdev", type(jdev).__name__)
    fdev = f.vec.CreateDeviceVector()
    log("fdev", type(fdev).__name__)
    out["types"] = [type(adev).__name__, type(jdev).__name__, type(fdev).__name__]
    xd = fdev.CreateVector(); yd = fdev.CreateVector()
    xd.data = fdev
    for _ in range(3):
        yd.data = adev * xd
    InnerProduct(yd, yd)                       # cublasDdot returns to the host: a device synchronisation
    log("spmv warm")
    reps = 20
    t1 = time.perf_counter()
    for _ in range(reps):
        yd.data = adev * xd
    InnerProduct(yd, yd)
    dt = (time.perf_counter() - t1) / reps
    b_alg = a.mat.nze * 12 + fes.ndof * 20
    out.update(spmv_ms=dt * 1e3, spmv_gbs_algorithmic=b_alg / dt / 1e9)
    log("spmv", dt)
    K = args.iters
    inv = CGSolver(adev, jdev, precision=1e-30, maxsteps=K, printrates=False)
    res = (inv * fdev).Evaluate()
    t1 = time.perf_counter()
    res = (inv * fdev).Evaluate()
    InnerProduct(res, res)
    dt = time.perf_counter() - t1
    out.update(cg_hostloop_it_per_s=(inv.GetSteps() - 1) / dt, cg_hostloop_steps=inv.GetSteps())
    log("hostloop cg", dt)
    try:
        # (mat, pre, adev_raw, cdev_raw): Mult only touches the last two (ngscuda/cuda_krylov.cpp:21-22)
        dinv = ngscuda.DevCGSolver(adev, jdev, adev, jdev, precision=1e-30, maxsteps=K)
        res = (dinv * fdev).Evaluate()
        t1 = time.perf_counter()
        res = (dinv * fdev).Evaluate()
        InnerProduct(res, res)
        dt = time.perf_counter() - t1
        out.update(devcg_it_per_s=K / dt, devcg_steps_reported=dinv.GetSteps())
        log("devcg", dt)
        if args.full:
            gfu = GridFunction(fes)
            cinv = CGSolver(a.mat, jac, precision=1e-8, maxsteps=20000, printrates=False)
            t1 = time.perf_counter()
            gfu.vec.data = cinv * f.vec
            out.update(cpu_full_steps=cinv.GetSteps(), cpu_full_s=time.perf_counter() - t1)
            dinv = ngscuda.DevCGSolver(adev, jdev, adev, jdev, precision=1e-8, maxsteps=20000)
            t1 = time.perf_counter()
            res = (dinv * fdev).Evaluate()
            InnerProduct(res, res)
            out.update(devcg_full_s=time.perf_counter() - t1)
            hinv = CGSolver(adev, jdev, precision=1e-8, maxsteps=20000, printrates=False)
            t1 = time.perf_counter()
            res2 = (hinv * fdev).Evaluate()
            InnerProduct(res2, res2)
            out.update(hostloop_full_s=time.perf_counter() - t1, hostloop_full_steps=hinv.GetSteps())
            d = gfu.vec.CreateVector()
            g2 = GridFunction(fes)
            g2.vec.data = res
            d.data = gfu.vec - g2.vec
            out.update(devcg_full_rel_diff=Norm(d) / Norm(gfu.vec))
    except Exception as e:          # noqa: BLE001
        out["devcg_error"] = repr(e)[:400]
print(json.dumps(out))
