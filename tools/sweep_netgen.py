#!/usr/bin/env python
"""Option sweep of the SELL product on a cached netgen system (tools/netgen_scale.py --cache): sigma window, inner-loop
variant, grid size.  One JSON line per setting."""
import argparse
import itertools
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from ngsolve_b200 import la  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cache", default="/dev/shm/ng2")
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--sigma", type=int, nargs="+", default=[-1])
ap.add_argument("--variant", type=int, nargs="+", default=[0])
ap.add_argument("--cps", type=int, nargs="+", default=[0])
ap.add_argument("--schedule", type=int, nargs="+", default=[1])
ap.add_argument("--out", default=None)
args = ap.parse_args()
ctx = la.default_context()
A = la.SparseMatrix(np.load(os.path.join(args.cache, "rowptr.npy")), np.load(os.path.join(args.cache, "col.npy")), np.load(os.path.join(args.cache, "val.npy")))
n = A.height
x = la.BaseVector(np.random.default_rng(1).random(n))
ctx.set_option("reorder", 1)
lines = []
for sigma, sched in itertools.product(args.sigma, args.schedule):
    ctx.set_option("sell_sigma", sigma)
    ctx.set_option("sell_schedule", sched)
    t0 = time.perf_counter()
    dev = A.CreateDeviceMatrix()
    ctx.sync()
    create_s = time.perf_counter() - t0
    ent, ovf, cap = dev.Layout()
    b_alg = dev.MultBytes()
    b_st, c16 = dev.StreamBytes()
    y = dev.CreateColVector()
    for var, cps in itertools.product(args.variant, args.cps):
        ctx.set_option("sell_variant", var)
        ctx.set_option("spmv_ctas_per_sm", cps)
        for _ in range(3):
            dev.Mult(x, y)
        ctx.sync()
        ctx.set_option("timing", 1)
        ctx.kernel_time_reset()
        for _ in range(args.reps):
            dev.Mult(x, y)
        ms, nl = ctx.kernel_time("spmv")
        ctx.kernel_time_reset()
        ctx.set_option("timing", 0)
        t = ms / args.reps * 1e-3
        r = dict(sigma=sigma, schedule=sched, variant=var, cps=cps, padding=ent / A.nze - 1.0, overflow_rows=ovf, cap=cap, create_s=create_s,
                 spmv_ms=t * 1e3, gbs_algorithmic=b_alg / t / 1e9, frac_algorithmic=b_alg / t / 1e9 / 6545.3, gbs_stored=b_st / t / 1e9)
        lines.append(r)
        print(json.dumps(r), flush=True)
    del dev, y
if args.out:
    open(args.out, "w").write("\n".join(json.dumps(r) for r in lines) + "\n")
