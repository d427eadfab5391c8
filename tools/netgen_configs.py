#!/usr/bin/env python
"""BASELINE configs[1], [3], [4] on REAL netgen meshes at the named sizes, one GPU (run under the reference environment):

  c2  linear elasticity, H1 order 4, dim 3  -> SparseMatrix<Mat<3,3>>   (maxh 0.05 + 1x Refine: ~4.2 M block rows = 12.5 M dofs), Jacobi-CG
  c4  Maxwell curl-curl + mass, HCurl order 2 (maxh 0.05 + 2x Refine: ~20 M dofs, irregular rows), Jacobi-CG
  c5  complex Helmholtz, H1 order 4, impedance boundary term -i w u v ds, w = 10 (maxh 0.05 + 2x Refine: ~32 M dofs), Jacobi-GMRES

    source oracle/_ref/ngs/env.sh && python tools/netgen_configs.py c2 c4 c5 --out gpurun_out/r2_netgen_configs.jsonl

The reference assembles (and its CPU solver is timed for a few iterations on the same system); the library gets the matrix in
NGSolve's numbering straight from NGSolve's memory.  `import ngsolve` has to precede numpy."""
import argparse
import json
import os
import sys
import time

import ngsolve
from ngsolve import *          # noqa: F401,F403
from netgen.csg import unit_cube

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("configs", nargs="+", choices=["c2", "c4", "c5"])
ap.add_argument("--maxh", type=float, default=0.05)
ap.add_argument("--nref-scale", type=int, default=0, help="subtract this many refinements (smoke runs)")
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--gmres-steps", type=int, default=60)
ap.add_argument("--cpu-iters", type=int, default=3)
ap.add_argument("--out", default=None)
ap.add_argument("--dry", action="store_true", help="stop before the device (checks assembly and the CPU arm where there is no GPU)")
args = ap.parse_args()
ngsolve.ngsglobals.msg_level = 0
T = os.cpu_count()
SetNumThreads(T)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def build(cfg, mesh):
    if cfg == "c2":
        fes = H1(mesh, order=4, dim=3, dirichlet="back")
        u, v = fes.TnT()
        E, nu = 210.0, 0.2
        mu, lam = E / 2 / (1 + nu), E * nu / ((1 + nu) * (1 - 2 * nu))
        eps = lambda w: 0.5 * (grad(w) + grad(w).trans)          # noqa: E731
        a = BilinearForm((2 * mu * InnerProduct(eps(u), eps(v)) + lam * Trace(grad(u)) * Trace(grad(v))) * dx)
        f = LinearForm(CF((0, 0, -1)) * v * dx)
    elif cfg == "c4":
        fes = HCurl(mesh, order=2, dirichlet=".*")
        u, v = fes.TnT()
        a = BilinearForm((curl(u) * curl(v) + u * v) * dx)
        f = LinearForm(CF((1, 0.5, -0.25)) * v * dx)
    else:
        fes = H1(mesh, order=4, complex=True)
        u, v = fes.TnT()
        om = 10.0
        a = BilinearForm((grad(u) * grad(v) - om * om * u * v) * dx - 1j * om * u * v * ds)
        f = LinearForm(exp(-40 * ((x - 0.5) ** 2 + (y - 0.5) ** 2 + (z - 0.5) ** 2)) * v * dx)
    return fes, a, f


NREF = {"c2": 1, "c4": 2, "c5": 2}
from ngsolve_b200 import la     # noqa: E402  (ctypes only)
ctx = None
lines = []
for cfg in args.configs:
    out = {"config": cfg, "ngsolve": ngsolve.__version__, "threads": T, "maxh": args.maxh, "nref": max(0, NREF[cfg] - args.nref_scale)}
    t0 = time.perf_counter()
    with TaskManager():
        mesh = Mesh(unit_cube.GenerateMesh(maxh=args.maxh))
        for _ in range(out["nref"]):
            mesh.Refine()
        fes, a, f = build(cfg, mesh)
        a.Assemble()
        f.Assemble()
        out.update(ne=mesh.ne, ndof=fes.ndof * (3 if cfg == "c2" else 1), rows=a.mat.height, nnz=a.mat.nze, mat_type=type(a.mat).__name__,
                   assemble_s=time.perf_counter() - t0)
        log(cfg, "assembled", out["ndof"], out["nnz"], out["assemble_s"])
        fd = fes.FreeDofs()
        if cfg == "c5":
            free = None
        elif cfg == "c4":
            ones = a.mat.CreateColVector(); res = a.mat.CreateColVector()
            ones.FV().NumPy()[:] = 1.0
            res.data = Projector(fd, True) * ones
            free = res.FV().NumPy() > 0.5
        else:
            free = np.array(fd)
        jac = a.mat.CreateSmoother(fd)
        gfu = GridFunction(fes)
        if args.cpu_iters > 0:
            mk = (lambda k: GMRESSolver(a.mat, jac, printrates=False, precision=1e-30, maxsteps=k)) if cfg == "c5" else \
                 (lambda k: CGSolver(a.mat, jac, precision=1e-30, maxsteps=k, printrates=False))
            inv = mk(2)
            gfu.vec.data = inv * f.vec
            inv = mk(args.cpu_iters + (1 if cfg == "c5" else 0))
            t1 = time.perf_counter()
            gfu.vec.data = inv * f.vec
            dt = time.perf_counter() - t1
            out.update(cpu_reference_it_per_s=args.cpu_iters / dt, cpu_reference_iters=args.cpu_iters)
            log(cfg, "cpu", out["cpu_reference_it_per_s"])
        val, col, rowptr = a.mat.CSR()
        rowptr = np.asarray(rowptr); col = np.asarray(col); val = np.asarray(val)
        fh = np.array(f.vec.FV().NumPy())
    es = 3 if cfg == "c2" else 1
    if args.dry:
        log(cfg, out, val.shape, val.dtype, col.dtype, rowptr.dtype, None if free is None else (free.dtype, free.shape, int(free.sum())))
        continue
    ctx = ctx or la.default_context()
    t1 = time.perf_counter()
    A = la.SparseMatrix(rowptr, col, val, entrysize=es)
    dev = A.CreateDeviceMatrix()
    ctx.sync()
    out["create_device_matrix_s"] = time.perf_counter() - t1
    on, share = dev.ReorderInfo()
    b_alg = dev.MultBytes()
    b_st, c16 = dev.StreamBytes()
    ent, ovf, cap = dev.Layout()
    peak = 6545.3
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    n = dev.height
    rng = np.random.default_rng(1)
    xh = rng.random(n * es) + (1j * rng.random(n) if cfg == "c5" else 0)
    xv = la.BaseVector(xh, entrysize=es)
    yv = dev.CreateColVector()
    for _ in range(3):
        dev.Mult(xv, yv)
    ctx.sync()
    ctx.set_option("timing", 1)
    ctx.kernel_time_reset()
    reps = 10
    for _ in range(reps):
        dev.Mult(xv, yv)
    ms, _ = ctx.kernel_time("spmv")
    ctx.kernel_time_reset()
    ctx.set_option("timing", 0)
    t_k = ms / reps * 1e-3
    out.update(reordered=on, sell_padding=ent / A.nze - 1.0, overflow_rows=ovf, spmv_kernel_ms=t_k * 1e3, spmv_gbs_algorithmic=b_alg / t_k / 1e9,
               spmv_frac_of_peak=b_alg / t_k / 1e9 / peak, spmv_gbs_stored=b_st / t_k / 1e9)
    log(cfg, "spmv", out["spmv_kernel_ms"], out["spmv_frac_of_peak"])
    del xv, yv
    jd = dev.CreateSmoother(la.BitArray(free) if free is not None else None)
    fv = la.BaseVector(fh, entrysize=es)
    uv = fv.CreateVector()
    S = 16 if cfg == "c5" else 8
    if cfg == "c5":
        K = args.gmres_steps
        inv = la.GMRESSolver(dev, jd, precision=1e-30, maxsteps=K)
        ab = {}
        for orth in (0, 1):           # the reference's serial modified Gram-Schmidt loop, then the one-reduction form (default)
            ctx.set_option("gmres_orth", orth)
            inv.Mult(fv, uv)
            ctx.sync()
            t1 = time.perf_counter()
            inv.Mult(fv, uv)
            ctx.sync()
            dt = time.perf_counter() - t1
            ab[orth] = dict(steps=inv.GetSteps(), steps_per_s=inv.GetSteps() / dt, last_residual=float(inv.history[-1]),
                            u_norm=float(np.linalg.norm(uv.NumPy())))
            log(cfg, "gmres_orth", orth, ab[orth])
        out["gmres_orth_ab"] = ab
        steps = inv.GetSteps()
        # SURVEY 8d byte model of step j: B_spmv + 3 N S + (2 (j+1) + 4) N S, summed over the steps
        bytes_model = sum(b_alg + 3 * n * S + (2 * (j + 1) + 4) * n * S for j in range(steps))
        bytes_mgs = sum(b_alg + (4 * (j + 1) + 8) * n * S for j in range(steps))
        out.update(solver="GMRES (Jacobi, no restart)", steps=steps, steps_per_s=steps / dt, frac_of_peak_survey_model=bytes_model / dt / 1e9 / peak,
                   frac_of_peak_mgs_bytes=bytes_mgs / dt / 1e9 / peak)
    else:
        inv = la.CGSolver(dev, jd, precision=0.0, maxsteps=args.iters)
        inv.Mult(fv, uv)
        ctx.sync()
        t1 = time.perf_counter()
        inv.Mult(fv, uv)
        ctx.sync()
        dt = time.perf_counter() - t1
        its = inv.GetSteps() - 1
        b_cg = b_alg + (11 + (6 if cfg == "c2" else 0)) * n * es * S
        out.update(solver="CG (Jacobi)", cg_it_per_s=its / dt, cg_frac_of_peak=b_cg * its / dt / 1e9 / peak)
    if out.get("cpu_reference_it_per_s"):
        out["speedup_vs_cpu_reference_same_system"] = (out.get("cg_it_per_s") or out.get("steps_per_s")) / out["cpu_reference_it_per_s"]
    log(cfg, out.get("cg_it_per_s") or out.get("steps_per_s"))
    lines.append(out)
    print(json.dumps(out), flush=True)
    del dev, A, jd, fv, uv, inv, a, f, fes, mesh, jac, gfu, val, col, rowptr
    import gc
    gc.collect()
if args.out:
    open(args.out, "w").write("\n".join(json.dumps(l) for l in lines) + "\n")
