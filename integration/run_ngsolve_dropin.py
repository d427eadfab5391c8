#!/usr/bin/env python
"""The drop-in, end to end: an UNCHANGED NGSolve solve script (netgen mesh, BilinearForm.Assemble, CreateSmoother,
CGSolver(mat, pre)) run once on NGSolve's own CPU path and once behind CreateDeviceMatrix()/CreateDeviceVector() with
the adapter module `_ngsb200` imported -- the same lines as docs/i-tutorials/unit-5.5-cuda/poisson_cuda.ipynb cells 5-6
with `import ngsolve.ngscuda` replaced by `import _ngsb200`.

Needs NGSolve (oracle/build_reference.sh), the adapter (integration/build_adapter.sh) and a GPU:
    source oracle/_ref/ngs/env.sh && python integration/run_ngsolve_dropin.py [--maxh 0.05] [--order 3]
Prints one JSON line (steps, timings, difference of the two solutions).
"""
import argparse
import json
import os
import sys
import time

import ngsolve
from ngsolve import *          # noqa: F401,F403
from netgen.csg import unit_cube

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_build"))

ap = argparse.ArgumentParser()
ap.add_argument("--maxh", type=float, default=0.05)
ap.add_argument("--order", type=int, default=3)
ap.add_argument("--threads", type=int, default=0)
ap.add_argument("--reorder", type=int, default=None, help="library option reorder (-1 automatic, 0 off, 1 always)")
ap.add_argument("--opt", action="append", default=[], help="library option name=value")
args = ap.parse_args()
ngsolve.ngsglobals.msg_level = 0
SetNumThreads(args.threads or os.cpu_count())

import _ngsb200               # registers the device creators (BaseMatrix::RegisterDeviceMatrixCreator, ...)
if args.reorder is not None:
    _ngsb200.SetOption("reorder", args.reorder)
for o in args.opt:
    _ngsb200.SetOption(o.split("=")[0], int(o.split("=")[1]))

out = {"ngsolve": ngsolve.__version__, "maxh": args.maxh, "order": args.order, "threads": args.threads or os.cpu_count()}
with TaskManager():
    t0 = time.perf_counter()
    mesh = Mesh(unit_cube.GenerateMesh(maxh=args.maxh))
    fes = H1(mesh, order=args.order, dirichlet=".*")
    u, v = fes.TnT()
    a = BilinearForm(grad(u) * grad(v) * dx).Assemble()
    f = LinearForm(1 * v * dx).Assemble()
    jac = a.mat.CreateSmoother(fes.FreeDofs())
    out.update(ndof=fes.ndof, nze=a.mat.nze, ne=mesh.ne, setup_s=time.perf_counter() - t0)

    # ---- NGSolve's own CPU path
    gfu = GridFunction(fes)
    inv = CGSolver(a.mat, jac, precision=1e-8, maxsteps=20000, printrates=False)
    t0 = time.perf_counter()
    gfu.vec.data = inv * f.vec
    out.update(cpu_steps=inv.GetSteps(), cpu_solve_s=time.perf_counter() - t0)

    # ---- the same script lines behind the device creators
    t0 = time.perf_counter()
    adev = a.mat.CreateDeviceMatrix()
    jdev = jac.CreateDeviceMatrix()
    fdev = f.vec.CreateDeviceVector()
    out.update(upload_s=time.perf_counter() - t0, types=[type(adev).__name__, type(jdev).__name__, type(fdev).__name__],
               is_host_object=[adev is a.mat, jdev is jac])
    # the unchanged script line.  With `_ngsb200` imported the factory hands back the same KrylovSpaceSolver interface whose
    # Mult is the library's device-resident loop (integration/ngsb200_ngla.cpp, B200CG)
    invdev = CGSolver(adev, jdev, precision=1e-8, maxsteps=20000, printrates=False)
    res = (invdev * fdev).Evaluate()
    t0 = time.perf_counter()
    res = (invdev * fdev).Evaluate()
    dt_dev = time.perf_counter() - t0
    reps = []
    for _ in range(4):
        t0 = time.perf_counter()
        res = (invdev * fdev).Evaluate()
        reps.append(time.perf_counter() - t0)
    out.update(dev_steps=invdev.GetSteps(), dev_solve_s=min([dt_dev] + reps), dev_solve_s_all=[dt_dev] + reps, dev_fused=_ngsb200.WasFused(invdev),
               dev_solver_type=type(invdev).__name__, reorder_info=list(_ngsb200.ReorderInfo(adev)))
    # a host matrix still gets the reference's own solver from the same factory
    out.update(host_factory_type=type(inv).__name__, host_factory_fused=_ngsb200.WasFused(inv))
    gfu2 = GridFunction(fes)
    gfu2.vec.data = res
    diff = gfu.vec.CreateVector()
    diff.data = gfu.vec - gfu2.vec
    out.update(rel_diff=Norm(diff) / Norm(gfu.vec))

    # ---- the reference's OWN python solvers (python/krylovspace.py:263-290, 988-1095), unmodified, on the device objects:
    # every `vec.data = expr`, InnerProduct and Norm goes through the adapter's BaseVector / BaseMatrix virtuals
    from ngsolve.krylovspace import CGSolver as PyCGSolver, GMResSolver as PyGMResSolver
    pinv = PyCGSolver(mat=adev, pre=jdev, tol=1e-8, maxiter=20000, printrates=False)
    t0 = time.perf_counter()
    pres = pinv.Solve(rhs=fdev)
    out.update(pycg_iterations=pinv.iterations, pycg_solve_s=time.perf_counter() - t0)
    gfu2 = GridFunction(fes)
    gfu2.vec.data = pres
    diff = gfu.vec.CreateVector()
    diff.data = gfu.vec - gfu2.vec
    out.update(pycg_rel_diff=Norm(diff) / Norm(gfu.vec))
    pinv_host = PyCGSolver(mat=a.mat, pre=jac, tol=1e-8, maxiter=20000, printrates=False)
    pinv_host.Solve(rhs=f.vec)
    out.update(pycg_host_iterations=pinv_host.iterations)
    pg = PyGMResSolver(mat=adev, pre=jdev, tol=1e-8, maxiter=400, printrates=False)
    gres = pg.Solve(rhs=fdev)
    pgh = PyGMResSolver(mat=a.mat, pre=jac, tol=1e-8, maxiter=400, printrates=False)
    gres_h = pgh.Solve(rhs=f.vec)
    gfu2.vec.data = gres
    diff.data = gres_h - gfu2.vec
    out.update(pygmres_iterations=pg.iterations, pygmres_host_iterations=pgh.iterations, pygmres_rel_diff=Norm(diff) / Norm(gres_h))

    # ---- host/device coherence of Range() views in both directions (ngscuda/unifiedvector.cpp:363-405)
    vv = fdev.CreateVector()
    vv.data = fdev
    half = fes.ndof // 2
    view = vv.Range(0, half)
    view.data = 2.0 * view                               # device write through the view ...
    h = vv.FV().NumPy()
    ref = f.vec.FV().NumPy()
    ok1 = bool(abs(h[:half] - 2 * ref[:half]).max() == 0 and abs(h[half:] - ref[half:]).max() == 0)   # ... seen by the parent's host side
    vv.FV().NumPy()[:] = 1.0                             # host write to the parent ...
    ok2 = bool(abs(Norm(view) ** 2 - half) < 1e-9 * half)   # ... seen by a device read through the view
    view.FV().NumPy()[:] = 3.0                           # host write through the view ...
    ok3 = bool(abs(InnerProduct(vv, vv) - (9.0 * half + (fes.ndof - half))) < 1e-9 * fes.ndof)   # ... seen by the parent on the device
    vv.data = 5.0 * vv                                   # device write to the parent ...
    ok4 = bool(abs(view.FV().NumPy() - 15.0).max() == 0)    # ... seen by the view's host side
    out.update(range_coherence=[ok1, ok2, ok3, ok4])

    # ---- fused device solver (same interface as ngscuda.DevCGSolver)
    fused = _ngsb200.DevCGSolver(adev, jdev, maxsteps=20000, precision=1e-8)
    res3 = (fused * fdev).Evaluate()
    t0 = time.perf_counter()
    res3 = (fused * fdev).Evaluate()
    out.update(fused_steps=fused.GetSteps(), fused_solve_s=time.perf_counter() - t0)
    gfu2.vec.data = res3
    diff.data = gfu.vec - gfu2.vec
    out.update(fused_rel_diff=Norm(diff) / Norm(gfu.vec))
    # ---- block Jacobi (the tutorial's second device example): BlockJacobiPrecond -> DevBlockJacobiMatrix
    blocks = fes.CreateSmoothingBlocks()
    bj = a.mat.CreateBlockSmoother(blocks)
    invb = CGSolver(a.mat, bj, precision=1e-8, maxsteps=20000, printrates=False)
    t0 = time.perf_counter()
    gfu.vec.data = invb * f.vec
    out.update(bj_cpu_steps=invb.GetSteps(), bj_cpu_solve_s=time.perf_counter() - t0)
    bjdev = bj.CreateDeviceMatrix()
    invbd = CGSolver(adev, bjdev, precision=1e-8, maxsteps=20000, printrates=False)
    res4 = (invbd * fdev).Evaluate()
    t0 = time.perf_counter()
    res4 = (invbd * fdev).Evaluate()
    out.update(bj_dev_steps=invbd.GetSteps(), bj_dev_solve_s=time.perf_counter() - t0, bj_type=type(bjdev).__name__,
               bj_fused=_ngsb200.WasFused(invbd))     # False: block Jacobi runs the inherited reference loop, op by op on the device
    gfu2.vec.data = res4
    diff.data = gfu.vec - gfu2.vec
    out.update(bj_rel_diff=Norm(diff) / Norm(gfu.vec))

    # ---- device-resident scalars (BaseScalar hooks, linalg/basevector.cpp:259-298): one CG-style step without host scalars
    try:
        sc, sc2 = fdev.CreateScalar(), fdev.CreateScalar()
        w = fdev.CreateVector()
        w.data = jdev * fdev
        fdev.InnerProduct(w, sc)                       # <f, C f> stays on the device
        host_ip = InnerProduct(fdev, w)
        t = fdev.CreateVector()
        t.data = fdev
        t.data += sc * w                               # t = f + <f, C f> * C f with the device scalar (BaseScalar * vector expression)
        t2 = fdev.CreateVector()
        t2.data = fdev + host_ip * w
        dd = t.CreateVector()
        dd.data = t - t2
        t.InnerProduct(t, sc2)
        out.update(scalar_type=type(sc).__name__, scalar_ip_rel_err=abs(sc.__float__() - host_ip) / abs(host_ip) if hasattr(sc, "__float__") else None,
                   scalar_axpy_rel_diff=Norm(dd) / Norm(t2), scalar_str=str(sc2))
    except Exception as e:                           # noqa: BLE001 -- report, keep the other results
        out["scalar_error"] = str(e)[:300]

    # ---- the other two entry kinds through the same registry: Mat<3,3,double> (elasticity, H1 dim=3) and Complex (Helmholtz, GMRES)
    def attempt(tag, fn):
        try:
            fn()
        except Exception as e:                       # noqa: BLE001 -- report, keep the other results
            out[tag + "_error"] = str(e)[:300]

    def elasticity():
        mesh2 = Mesh(unit_cube.GenerateMesh(maxh=0.12))
        fe = H1(mesh2, order=2, dim=3, dirichlet="back")
        uu, vv = fe.TnT()
        E, nu = 210.0, 0.2
        mu, lam = E / 2 / (1 + nu), E * nu / ((1 + nu) * (1 - 2 * nu))
        eps = lambda w: 0.5 * (grad(w) + grad(w).trans)
        aa = BilinearForm(2 * mu * InnerProduct(eps(uu), eps(vv)) * dx + lam * Trace(grad(uu)) * Trace(grad(vv)) * dx).Assemble()
        ff = LinearForm(CoefficientFunction((0, 0, -1)) * vv * dx).Assemble()
        jj = aa.mat.CreateSmoother(fe.FreeDofs())
        g1 = GridFunction(fe)
        ic = CGSolver(aa.mat, jj, precision=1e-8, maxsteps=20000, printrates=False)
        g1.vec.data = ic * ff.vec
        ad, jd, fd = aa.mat.CreateDeviceMatrix(), jj.CreateDeviceMatrix(), ff.vec.CreateDeviceVector()
        idv = CGSolver(ad, jd, precision=1e-8, maxsteps=20000, printrates=False)
        r = (idv * fd).Evaluate()
        out.update(b3_fused=_ngsb200.WasFused(idv))
        g2 = GridFunction(fe)
        g2.vec.data = r
        dd = g1.vec.CreateVector()
        dd.data = g1.vec - g2.vec
        out.update(b3_mat_type=type(aa.mat).__name__, b3_ndof=fe.ndof, b3_cpu_steps=ic.GetSteps(), b3_dev_steps=idv.GetSteps(),
                   b3_rel_diff=Norm(dd) / Norm(g1.vec), b3_is_host_object=ad is aa.mat)

    def helmholtz():
        mesh3 = Mesh(unit_cube.GenerateMesh(maxh=0.12))
        fe = H1(mesh3, order=3, complex=True)
        uu, vv = fe.TnT()
        om = 6.0
        aa = BilinearForm(grad(uu) * grad(vv) * dx - om * om * uu * vv * dx - 1j * om * uu * vv * ds).Assemble()
        ff = LinearForm(exp(-40 * ((x - 0.5) ** 2 + (y - 0.5) ** 2 + (z - 0.5) ** 2)) * vv * dx).Assemble()
        jj = aa.mat.CreateSmoother(fe.FreeDofs())
        g1 = GridFunction(fe)
        ic = GMRESSolver(aa.mat, jj, printrates=False, precision=1e-8, maxsteps=400)
        g1.vec.data = ic * ff.vec
        ad, jd, fd = aa.mat.CreateDeviceMatrix(), jj.CreateDeviceMatrix(), ff.vec.CreateDeviceVector()
        idv = GMRESSolver(ad, jd, printrates=False, precision=1e-8, maxsteps=400)
        r = (idv * fd).Evaluate()
        out.update(z_fused=_ngsb200.WasFused(idv))
        g2 = GridFunction(fe)
        g2.vec.data = r
        dd = g1.vec.CreateVector()
        dd.data = g1.vec - g2.vec
        out.update(z_mat_type=type(aa.mat).__name__, z_ndof=fe.ndof, z_cpu_gmres_steps=ic.GetSteps(), z_dev_gmres_steps=idv.GetSteps(),
                   z_rel_diff=Norm(dd) / Norm(g1.vec), z_is_host_object=ad is aa.mat)

    attempt("b3", elasticity)
    attempt("z", helmholtz)
print(json.dumps(out))
