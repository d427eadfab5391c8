#!/usr/bin/env python
"""Generate a netgen-meshed system with the REFERENCE build (oracle/_ref/ngs, source its env.sh first or let
tools/netgen_scale.py do it) and store it for the GPU measurements: SURVEY.md 8(d) inputs C1 / C3-fallback,
    Mesh(unit_cube.GenerateMesh(maxh)), mesh.Refine() x nref, H1(order, dirichlet=".*"), grad(u)*grad(v)*dx, rhs 1*v*dx.
Also times the reference's own CPU solve on it: C++ CGSolver + JacobiPrecond under TaskManager (it/s over --cpu-iters
iterations; with --cpu-full the whole solve to 1e-8 and its GetSteps()).
`import ngsolve` must come before numpy (SURVEY.md 8c pitfall 4)."""
import argparse
import hashlib
import json
import os
import time

import ngsolve
from ngsolve import *          # noqa: F401,F403
from netgen.csg import unit_cube

import numpy as np

ap = argparse.ArgumentParser()
ap.add_argument("--maxh", type=float, default=0.05)
ap.add_argument("--nref", type=int, default=2)
ap.add_argument("--order", type=int, default=3)
ap.add_argument("--threads", type=int, default=0)
ap.add_argument("--cpu-iters", type=int, default=20)
ap.add_argument("--cpu-full", action="store_true")
ap.add_argument("--gpu-reference", action="store_true", help="also time the reference's OWN device path (ngscuda: cuSPARSE DevSparseMatrix, "
                "cuBLAS UnifiedVector, DevCGSolver; built by oracle/build_reference_cuda.sh) on this GPU")
ap.add_argument("--gpu-iters", type=int, default=100)
ap.add_argument("--no-save", action="store_true")
ap.add_argument("--out", required=True)
args = ap.parse_args()
ngsolve.ngsglobals.msg_level = 0
T = args.threads or os.cpu_count()
SetNumThreads(T)
os.makedirs(args.out, exist_ok=True)
meta = {"ngsolve": ngsolve.__version__, "maxh": args.maxh, "nref": args.nref, "order": args.order, "threads": T}
t0 = time.perf_counter()
with TaskManager():
    mesh = Mesh(unit_cube.GenerateMesh(maxh=args.maxh))
    for _ in range(args.nref):
        mesh.Refine()
    meta.update(ne=mesh.ne, nv=mesh.nv, mesh_s=time.perf_counter() - t0)
    fes = H1(mesh, order=args.order, dirichlet=".*")
    u, v = fes.TnT()
    a = BilinearForm(grad(u) * grad(v) * dx).Assemble()
    f = LinearForm(1 * v * dx).Assemble()
    meta.update(ndof=fes.ndof, nnz=a.mat.nze, setup_s=time.perf_counter() - t0)
    val, col, rowptr = a.mat.CSR()
    rowptr = np.asarray(rowptr); col = np.asarray(col); val = np.asarray(val)
    fd = fes.FreeDofs()
    # BitArray -> numpy without a Python loop over 1e8 dofs: through a projector product
    ones = a.mat.CreateColVector(); res = a.mat.CreateColVector()
    ones.FV().NumPy()[:] = 1.0
    res.data = Projector(fd, True) * ones
    free = res.FV().NumPy() > 0.5
    bits = np.packbits(free, bitorder="little")
    if not args.no_save:
        np.save(os.path.join(args.out, "rowptr.npy"), rowptr)
        np.save(os.path.join(args.out, "col.npy"), col)
        np.save(os.path.join(args.out, "val.npy"), val)
        np.save(os.path.join(args.out, "f.npy"), f.vec.FV().NumPy())
        np.save(os.path.join(args.out, "freebits.npy"), bits)
    meta["sha256_rowptr"] = hashlib.sha256(rowptr.tobytes()).hexdigest()
    meta["sha256_col"] = hashlib.sha256(col.tobytes()).hexdigest()
    meta["saved_s"] = time.perf_counter() - t0
    # ---- the reference's CPU solve
    jac = a.mat.CreateSmoother(fd)
    if args.cpu_iters > 0:
        gfu = GridFunction(fes)
        inv = CGSolver(a.mat, jac, precision=1e-30, maxsteps=3, printrates=False)
        gfu.vec.data = inv * f.vec
        inv = CGSolver(a.mat, jac, precision=1e-30, maxsteps=args.cpu_iters, printrates=False)
        t1 = time.perf_counter()
        gfu.vec.data = inv * f.vec
        dt = time.perf_counter() - t1
        meta.update(cpu_iters=inv.GetSteps() - 1, cpu_it_per_s=(inv.GetSteps() - 1) / dt)
        xs = a.mat.CreateRowVector(); ys = a.mat.CreateColVector()
        xs.FV().NumPy()[:] = 1.0
        t1 = time.perf_counter()
        for _ in range(5):
            ys.data = a.mat * xs
        meta.update(cpu_spmv_ms=(time.perf_counter() - t1) / 5 * 1e3)
    if args.cpu_full:
        gfu = GridFunction(fes)
        inv = CGSolver(a.mat, jac, precision=1e-8, maxsteps=20000, printrates=False)
        t1 = time.perf_counter()
        gfu.vec.data = inv * f.vec
        meta.update(cpu_full_steps=inv.GetSteps(), cpu_full_s=time.perf_counter() - t1)
        np.save(os.path.join(args.out, "u_ref.npy"), gfu.vec.FV().NumPy())
    if args.gpu_reference:
        # the stock reference device layer (ngscuda/cuda_linalg.cpp:187-316 DevSparseMatrix = cuSPARSE SpMV with a generic-API
        # descriptor per call, ngscuda/unifiedvector.cpp cuBLAS vectors, ngscuda/cuda_krylov.cpp:19-203 DevCGSolver).  The
        # python wrapper ngsolve/ngscuda.py refuses to load when the core library was configured without CUDA, the
        # compiled module itself is complete: import it directly.
        gr = {}
        try:
            from ngsolve._ngscuda import DevCGSolver as RefDevCGSolver      # registers the device creators at import
            adev = a.mat.CreateDeviceMatrix()
            jdev = jac.CreateDeviceMatrix()
            fdev = f.vec.CreateDeviceVector()
            gr["types"] = [type(adev).__name__, type(jdev).__name__, type(fdev).__name__]
            xd = fdev.CreateVector(); yd = fdev.CreateVector()
            xd.data = fdev
            for _ in range(3):
                yd.data = adev * xd
            InnerProduct(yd, yd)                       # cublasDdot returns to the host: a device synchronisation
            reps = 20
            t1 = time.perf_counter()
            for _ in range(reps):
                yd.data = adev * xd
            InnerProduct(yd, yd)
            dt = (time.perf_counter() - t1) / reps
            b_alg = a.mat.nze * 12 + fes.ndof * 20
            gr.update(spmv_ms=dt * 1e3, spmv_gbs_algorithmic=b_alg / dt / 1e9)
            # NGSolve's C++ CGSolver loop driving the device objects op by op (docs/i-tutorials/unit-5.5-cuda/poisson_cuda.ipynb)
            K = args.gpu_iters
            inv = CGSolver(adev, jdev, precision=1e-30, maxsteps=K, printrates=False)
            res = (inv * fdev).Evaluate()
            t1 = time.perf_counter()
            res = (inv * fdev).Evaluate()
            InnerProduct(res, res)
            dt = time.perf_counter() - t1
            gr.update(cg_hostloop_it_per_s=(inv.GetSteps() - 1) / dt, cg_hostloop_steps=inv.GetSteps())
            # DevCGSolver: the graph-captured device CG
            dinv = RefDevCGSolver(adev, jdev, adev, jdev, precision=1e-30, maxsteps=K)
            res = (dinv * fdev).Evaluate()
            t1 = time.perf_counter()
            res = (dinv * fdev).Evaluate()
            InnerProduct(res, res)
            dt = time.perf_counter() - t1
            gr.update(devcg_it_per_s=(dinv.GetSteps() - 1) / dt if dinv.GetSteps() > 1 else K / dt, devcg_steps=dinv.GetSteps())
            if args.cpu_full:
                dinv = RefDevCGSolver(adev, jdev, adev, jdev, precision=1e-8, maxsteps=20000)
                t1 = time.perf_counter()
                res = (dinv * fdev).Evaluate()
                InnerProduct(res, res)
                gr.update(devcg_full_steps=dinv.GetSteps(), devcg_full_s=time.perf_counter() - t1)
        except Exception as e:                  # noqa: BLE001 -- report what the reference device path did
            gr["error"] = repr(e)[:500]
        meta["gpu_reference"] = gr
json.dump(meta, open(os.path.join(args.out, "meta.json"), "w"))
print(json.dumps(meta))
