#!/usr/bin/env python
"""BASELINE configs[2] on a REAL netgen mesh, one GPU: unit_cube maxh=0.05 + 3x Refine, H1 order 3 -> ~1.1e8 dofs, assembled
by the reference (oracle/_ref/ngs), handed to the library in NGSolve's own dof numbering straight from NGSolve's memory
(no file in between), multiplied / solved after the automatic Cuthill-McKee reordering.  Run under the reference environment:

    source oracle/_ref/ngs/env.sh && python tools/netgen_big.py --nref 3 --out gpurun_out/r2_netgen_110M.json

`import ngsolve` has to precede numpy (SURVEY.md 8c pitfall 4)."""
import argparse
import hashlib
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
ap.add_argument("--maxh", type=float, default=0.05)
ap.add_argument("--nref", type=int, default=3)
ap.add_argument("--order", type=int, default=3)
ap.add_argument("--iters", type=int, default=100)
ap.add_argument("--cpu-iters", type=int, default=5)
ap.add_argument("--cpu-warmup", type=int, default=2)
ap.add_argument("--cpu-only", action="store_true", help="the reference's CPU CG only (bench.py --impl reference)")
ap.add_argument("--full", action="store_true")
ap.add_argument("--cps-sweep", type=int, nargs="*", default=[], help="also time the product and 50 CG iterations with these grids (CTAs per SM)")
ap.add_argument("--out", default=None)
args = ap.parse_args()
ngsolve.ngsglobals.msg_level = 0
T = os.cpu_count()
SetNumThreads(T)
out = {"ngsolve": ngsolve.__version__, "maxh": args.maxh, "nref": args.nref, "order": args.order, "threads": T}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


t0 = time.perf_counter()
with TaskManager():
    mesh = Mesh(unit_cube.GenerateMesh(maxh=args.maxh))
    for _ in range(args.nref):
        mesh.Refine()
    out.update(ne=mesh.ne, nv=mesh.nv, mesh_s=time.perf_counter() - t0)
    log("mesh", mesh.ne, out["mesh_s"])
    fes = H1(mesh, order=args.order, dirichlet=".*")
    u, v = fes.TnT()
    a = BilinearForm(grad(u) * grad(v) * dx).Assemble()
    f = LinearForm(1 * v * dx).Assemble()
    out.update(ndof=fes.ndof, nnz=a.mat.nze, assemble_s=time.perf_counter() - t0)
    log("assembled", fes.ndof, a.mat.nze, out["assemble_s"])
    val, col, rowptr = a.mat.CSR()
    rowptr = np.asarray(rowptr); col = np.asarray(col); val = np.asarray(val)
    ones = a.mat.CreateColVector(); res = a.mat.CreateColVector()
    ones.FV().NumPy()[:] = 1.0
    res.data = Projector(fes.FreeDofs(), True) * ones
    bits = np.packbits(res.FV().NumPy() > 0.5, bitorder="little")
    fh = np.array(f.vec.FV().NumPy())
    if args.cpu_iters > 0:
        jac = a.mat.CreateSmoother(fes.FreeDofs())
        gfu = GridFunction(fes)
        inv = CGSolver(a.mat, jac, precision=1e-30, maxsteps=max(1, args.cpu_warmup), printrates=False)
        gfu.vec.data = inv * f.vec
        inv = CGSolver(a.mat, jac, precision=1e-30, maxsteps=args.cpu_iters, printrates=False)
        t1 = time.perf_counter()
        gfu.vec.data = inv * f.vec
        dt = time.perf_counter() - t1
        out.update(cpu_reference_it_per_s=(inv.GetSteps() - 1) / dt, cpu_reference_iters=inv.GetSteps() - 1, cpu_reference_s=dt)
        log("cpu cg", out["cpu_reference_it_per_s"])
        del jac, gfu, inv
out["sha256_rowptr"] = hashlib.sha256(rowptr.tobytes()).hexdigest()[:16]
if args.cpu_only:
    s = json.dumps(out)
    if args.out:
        open(args.out, "w").write(s + "\n")
    print(s)
    sys.exit(0)

from ngsolve_b200 import la     # noqa: E402
ctx = la.default_context()
n = len(rowptr) - 1
t1 = time.perf_counter()
A = la.SparseMatrix(rowptr, col, val)
dev = A.CreateDeviceMatrix()
ctx.sync()
out["create_device_matrix_s"] = time.perf_counter() - t1
log("device matrix", out["create_device_matrix_s"])
on, share = dev.ReorderInfo()
csr_b, sell_b, resident = dev.Memory()
b_alg = dev.MultBytes()
b_st, c16 = dev.StreamBytes()
ent, ovf, cap = dev.Layout()
out.update(reordered=on, natural_c16_share=share, csr_bytes_resident=csr_b, sell_bytes=sell_b, csr_arrays_resident=resident,
           sell_padding=ent / A.nze - 1.0, overflow_rows=ovf, c16_share_of_entries=c16 / max(1, ent))
peak = 6545.3
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
x = la.BaseVector(np.random.default_rng(1).random(n))
y = dev.CreateColVector()
for _ in range(3):
    dev.Mult(x, y)
ctx.sync()
ctx.set_option("timing", 1)
ctx.kernel_time_reset()
reps = 10
for _ in range(reps):
    dev.Mult(x, y)
ms, nl = ctx.kernel_time("spmv")
vms, vn = ctx.kernel_time("all")
ctx.kernel_time_reset()
ctx.set_option("timing", 0)
t_k, t_all = ms / reps * 1e-3, vms / reps * 1e-3
out.update(spmv_kernel_ms=t_k * 1e3, spmv_call_ms_incl_gather=t_all * 1e3, spmv_gbs_algorithmic=b_alg / t_k / 1e9,
           spmv_frac_of_peak=b_alg / t_k / 1e9 / peak, spmv_gbs_stored=b_st / t_k / 1e9, algorithmic_bytes=b_alg, stored_bytes=b_st)
log("spmv", out["spmv_kernel_ms"], out["spmv_frac_of_peak"])
del x, y
jac = dev.CreateSmoother(la.BitArray(bits))
fv = la.BaseVector(fh)
uv = fv.CreateVector()
inv = la.CGSolver(dev, jac, precision=0.0, maxsteps=args.iters)
inv.Mult(fv, uv)
ctx.sync()
t1 = time.perf_counter()
inv.Mult(fv, uv)
ctx.sync()
dt = time.perf_counter() - t1
its = inv.GetSteps() - 1
b_cg = b_alg + 11 * n * 8
out.update(cg_it_per_s=its / dt, cg_gbs=b_cg * its / dt / 1e9, cg_frac_of_peak=b_cg * its / dt / 1e9 / peak)
log("cg", out["cg_it_per_s"])
sweep = []
for cps in args.cps_sweep:
    ctx.set_option("spmv_ctas_per_sm", cps)
    xs = la.BaseVector(np.random.default_rng(1).random(n)); ys = dev.CreateColVector()
    for _ in range(2):
        dev.Mult(xs, ys)
    ctx.sync()
    ctx.set_option("timing", 1)
    ctx.kernel_time_reset()
    for _ in range(5):
        dev.Mult(xs, ys)
    ms5, _ = ctx.kernel_time("spmv")
    ctx.kernel_time_reset()
    ctx.set_option("timing", 0)
    del xs, ys
    inv = la.CGSolver(dev, jac, precision=0.0, maxsteps=50)
    inv.Mult(fv, uv)
    ctx.sync()
    t1 = time.perf_counter()
    inv.Mult(fv, uv)
    ctx.sync()
    dt = time.perf_counter() - t1
    sweep.append(dict(ctas_per_sm=cps, spmv_kernel_ms=ms5 / 5, cg_it_per_s=(inv.GetSteps() - 1) / dt))
    log("cps", cps, sweep[-1])
ctx.set_option("spmv_ctas_per_sm", 0)
if sweep:
    out["grid_sweep"] = sweep
if args.full:
    inv = la.CGSolver(dev, jac, precision=1e-8, maxsteps=20000)
    t1 = time.perf_counter()
    inv.Mult(fv, uv)
    ctx.sync()
    out.update(full_solve_steps=inv.GetSteps(), full_solve_s=time.perf_counter() - t1)
s = json.dumps(out)
if args.out:
    open(args.out, "w").write(s + "\n")
print(s)
