"""Sweep of the launch-time knobs of the compressed (16-bit column) SELL kernel on the bench system.
  python tools/sweep_c16.py [size=160] > profiles/<name>.txt"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ngsolve_b200.la as la
from ngsolve_b200 import workloads as W

m = int(sys.argv[1]) if len(sys.argv) > 1 else 160
ctx = la.default_context()
box = W.FemBox(m, order=3)
A, f = box.device_system(ctx)
x = f.CreateVector(); x.SetRandom(1)
y = A.CreateColVector()
st = torch.cuda.ExternalStream(ctx.stream)
alg = A.MultBytes(); stored, c16 = A.StreamBytes()
print("dofs %d  nnz %d  algorithmic %.2f GB  stored %.2f GB  c16 share %.3f" % (A.height, A.nze, alg / 1e9, stored / 1e9, c16 / A.Layout()[0]))


def t(reps=20):
    for _ in range(3):
        A.Mult(x, y)
    ctx.sync()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record(st)
    for k in range(reps):
        A.Mult(x, y)
        ev[k + 1].record(st)
    ctx.sync()
    ms = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(reps))
    return ms[len(ms) // 2]


def run(tag, **opts):
    for k, v in opts.items():
        ctx.set_option(k, v)
    ms = t()
    print("%-46s %8.3f ms  alg %7.1f GB/s  stored %7.1f GB/s" % (tag, ms, alg / ms / 1e6, (stored if opts.get("sell_c16", 1) else alg) / ms / 1e6), flush=True)


mode = sys.argv[2] if len(sys.argv) > 2 else "async"
run("c16 off", sell_c16=0)
run("c16 on, default kernel (registers)", sell_c16=1, spmv_ctas_per_sm=0)
if mode == "prefetch":
    for cps in (8, 6, 5, 4, 12):
        for pf, nx in ((0, 0), (1, 0), (2, 0), (3, 0), (0, 8), (2, 8), (1, 8), (2, 16), (3, 16), (4, 8)):
            run("c16 ctas/sm=%d pf_steps=%d pf_next=%d" % (cps, pf, nx), sell_c16=1, spmv_ctas_per_sm=cps, sell_pf_steps=pf, sell_pf_next=nx)
elif mode == "grid":
    for rep in range(2):
        for cps in (8, 16, 20, 24, 28, 32, 36, 40, 48, 56, 64, 80, 96, 110):
            run("default kernel, grid = %d CTAs/SM" % cps, sell_c16=1, sell_variant=0, spmv_ctas_per_sm=cps)
else:
    for rep in range(2):
        for var in (0, 4, 5, 6):
            for cps in (0, 5, 6, 8, 10, 12, 15):
                run("variant %d (3 = L2 eviction policies) ctas/sm=%d" % (var, cps), sell_c16=1, sell_variant=var, spmv_ctas_per_sm=cps)
    ctx.set_option("sell_variant", 0)
