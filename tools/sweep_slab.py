"""One GPU, the per-GPU problem of the N-GPU bench run: the z-slab a rank owns when the 160^3 grid is split N ways
(same generator, same local matrix, no neighbours), swept over the SELL grid size.  The 8-GPU run multiplies 13.9 M rows
per GPU in 1.30 ms (0.85 of the copy rate on stored bytes, against 0.94 for the 111 M rows of one GPU): this tool shows on
a single-GPU box how much of that is the kernel at the smaller size and which grid it wants.

  python tools/sweep_slab.py [--size 160] [--parts 8 4 2] [--cps 16 32 64 96] [--out gpurun_out/slab.jsonl]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ngsolve_b200.la as la
from ngsolve_b200 import workloads as W


def timed(ctx, fn, reps):
    st = torch.cuda.ExternalStream(ctx.stream)
    ctx.sync()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record(st)
    for k in range(reps):
        fn()
        ev[k + 1].record(st)
    ctx.sync()
    ms = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(reps))
    return ms[len(ms) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=160)
    ap.add_argument("--parts", type=int, nargs="*", default=[8, 4, 2])
    ap.add_argument("--cps", type=int, nargs="*", default=[16, 32, 64, 96])
    ap.add_argument("--steps", type=int, default=48)
    ap.add_argument("--out", default=None)
    ap.add_argument("--opt", action="append", default=[])
    a = ap.parse_args()
    ctx = la.default_context()
    for o in a.opt:
        k, v = o.split("=")
        ctx.set_option(k, int(v))
    G = (a.size,) * 3
    for parts in a.parts:
        slabs = W.slab_partition(G, parts)
        n, off = slabs[len(slabs) // 2]                 # a middle rank: two interfaces
        box = W.FemBox(n, order=3, offset=off, global_n=G)
        A, f = box.device_system(ctx)
        jac = A.CreateSmoother(box.freedofs())
        x = f.CreateVector()
        x.SetRandom(1)
        y = A.CreateColVector()
        u = f.CreateVector()
        sb, _ = A.StreamBytes()
        for cps in a.cps:
            ctx.set_option("spmv_ctas_per_sm", cps)
            for _ in range(5):
                A.Mult(x, y)
            ms = timed(ctx, lambda: A.Mult(x, y), 50)
            inv = la.CGSolver(A, jac, precision=0.0, maxsteps=a.steps)
            inv.Mult(f, u)
            ms_cg = timed(ctx, lambda: inv.Mult(f, u), 3) / a.steps
            line = dict(options=a.opt, parts=parts, rows=A.height, nnz=A.nze, ctas_per_sm=cps, spmv_ms=ms, gbs_algorithmic=A.MultBytes() / ms / 1e6,
                        gbs_stored=sb / ms / 1e6, cg_ms_per_iteration=ms_cg, it_per_s_times_parts=parts * 1e3 / ms_cg)
            print(json.dumps(line), flush=True)
            if a.out:
                with open(a.out, "a") as fh:
                    fh.write(json.dumps(line) + "\n")
        ctx.set_option("spmv_ctas_per_sm", 0)
        del A, f, jac, x, y, u, box


if __name__ == "__main__":
    main()
