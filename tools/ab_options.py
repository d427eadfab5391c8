"""A/B of the opt-in options on one B200, one process (one torch import), results as JSON lines:

  python tools/ab_options.py [--m 100] [--mc 50] [--mb 28] [--out gpurun_out/ab.jsonl] [--only fold,c16c,c16b]

  fold : Jacobi-PCG on the bench family (m^3 cubes, H1 order 3) with cg_fold_u = 0 / 1, it/s over K iterations
  c16c : complex SpMV (order 4, mc^3 cubes) with sell_c16_all = 0 / 1 (the option is read when the matrix is created)
  c16b : 3x3-block SpMV (order 4, mb^3 cubes) with sell_c16_all = 0 / 1
Every line is flushed as soon as it is measured, so a run cut short by a time limit keeps what it has."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ngsolve_b200.la as la
from ngsolve_b200 import workloads as W


def emit(out, **kw):
    line = json.dumps(kw)
    print(line, flush=True)
    if out:
        with open(out, "a") as fh:
            fh.write(line + "\n")


def time_region(ctx, fn):
    st = torch.cuda.ExternalStream(ctx.stream)
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    fn()
    e1.record(st)
    ctx.sync()
    return e0.elapsed_time(e1)


def ab_fold(ctx, m, K, out):
    box = W.FemBox(m, order=3)
    A, f = box.device_system(ctx)
    jac = A.CreateSmoother(box.freedofs())
    u = f.CreateVector()
    res = {0: [], 1: []}
    for rep in range(3):
        for fold in (0, 1):
            ctx.set_option("cg_fold_u", fold)
            inv = la.CGSolver(A, jac, precision=0.0, maxsteps=K)
            inv.Mult(f, u)                   # untimed: the cached CUDA graph is re-captured when the option flips
            ms = time_region(ctx, lambda: inv.Mult(f, u))
            res[fold].append((inv.GetSteps() - 1) / (ms * 1e-3))
    ctx.set_option("cg_fold_u", 0)
    emit(out, test="cg_fold_u", dofs=A.height, iterations=K, it_per_s_off=res[0], it_per_s_on=res[1],
         gain=max(res[1]) / max(res[0]) - 1.0)


def ab_c16(ctx, m, kind, name, out):
    r = {}
    for on in (0, 1):
        ctx.set_option("sell_c16_all", on)
        box = W.FemBox(m, order=4, kind=kind, mass=(1.0 + 0.5j) if kind == W.COMPLEX else 0.5, lame=(1.0, 0.7))
        A, f = box.device_system(ctx)
        x = f.CreateVector()
        x.SetRandom(1)
        y = A.CreateColVector()
        for _ in range(3):
            A.Mult(x, y)
        ms = min(time_region(ctx, lambda: [A.Mult(x, y) for _ in range(10)]) / 10 for _ in range(3))
        sb, n16 = A.StreamBytes()
        r[on] = dict(ms=ms, gbs_algorithmic=A.MultBytes() / ms / 1e6, gbs_stored=sb / ms / 1e6, c16_share=n16 / max(1, A.Layout()[0]))
        rows, nnz = A.height, A.nze
        del A, f, x, y, box
    ctx.set_option("sell_c16_all", 0)
    emit(out, test="sell_c16_all " + name, rows=rows, nnz=nnz, off=r[0], on=r[1], gain=r[0]["ms"] / r[1]["ms"] - 1.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=100)
    ap.add_argument("--mc", type=int, default=50)
    ap.add_argument("--mb", type=int, default=28)
    ap.add_argument("--steps", type=int, default=48)
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default="fold,c16c,c16b")
    a = ap.parse_args()
    ctx = la.default_context()
    which = a.only.split(",")
    if "fold" in which:
        ab_fold(ctx, a.m, a.steps, a.out)
    if "c16c" in which:
        ab_c16(ctx, a.mc, W.COMPLEX, "complex", a.out)
    if "c16b" in which:
        ab_c16(ctx, a.mb, W.BLOCK3, "3x3", a.out)


if __name__ == "__main__":
    main()
