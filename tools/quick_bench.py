"""Throw-away first measurement: banded synthetic CSR, SpMV both algorithms + CG iterations."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ngsolve_b200.la as la

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
k = 47
rng = np.random.default_rng(0)
offs = np.unique(np.concatenate([[0], rng.integers(-3000, 3000, size=3 * k)]))[:k]
offs = np.unique(np.concatenate([[0], offs]))
k = len(offs)
i = np.arange(n, dtype=np.int64)
cols = np.sort((i[:, None] + offs[None, :]) % n, axis=1).astype(np.int32)
rowptr = (np.arange(n + 1, dtype=np.uint64) * k)
val = -rng.random(n * k)
diagpos = (cols == i[:, None].astype(np.int32)).reshape(-1)
val[diagpos] = k + 1.0
A = la.SparseMatrix(rowptr, cols.reshape(-1), val)
t0 = time.time(); dev = A.CreateDeviceMatrix(); print("upload s", time.time() - t0)
ctx = dev.ctx
x = la.BaseVector(rng.random(n)); y = dev.CreateColVector()
byts = dev.MultBytes()
import torch
for algo in (2, 1):
    ctx.set_option("spmv_algo", algo)
    for cps in ((1, 2, 3) if algo == 2 else (0,)):
        ctx.set_option("spmv_ctas_per_sm", cps)
        for _ in range(5): dev.Mult(x, y)
        ctx.sync()
        st = torch.cuda.ExternalStream(ctx.stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record(st)
            for _ in range(50): dev.Mult(x, y)
            e1.record(st)
        ctx.sync()
        ms = e0.elapsed_time(e1) / 50
        print("algo", algo, "ctas/sm", cps, "ms", ms, "GB/s", byts / ms / 1e6)
ctx.set_option("spmv_algo", 0); ctx.set_option("spmv_ctas_per_sm", 0)
jac = dev.CreateSmoother()
f = la.BaseVector(rng.random(n))
inv = la.CGSolver(dev, jac, precision=1e-30, maxsteps=200)
u = f.CreateVector()
inv.Mult(f, u)
ctx.sync(); t0 = time.time(); inv.Mult(f, u); ctx.sync(); dt = time.time() - t0
it = inv.GetSteps() - 1
print("CG iterations", it, "s", dt, "it/s", it / dt, "GB/s (B_cg)", (byts + 11 * n * 8) * it / dt / 1e9)
