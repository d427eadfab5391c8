"""SpMV parameter sweep on the generated FE system (tool for tuning, not a test)."""
import sys, os, json, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ngsolve_b200.la as la
from ngsolve_b200 import workloads as W

m = int(sys.argv[1]) if len(sys.argv) > 1 else 60
order = int(sys.argv[2]) if len(sys.argv) > 2 else 3
kind = int(sys.argv[3]) if len(sys.argv) > 3 else 0
mode = sys.argv[4] if len(sys.argv) > 4 else "full"
ctx = la.default_context()
box = W.FemBox(m, order=order, kind=kind, lame=(1.0, 0.5), mass=(0.1 - 0.2j) if kind == 1 else 0.0)
st = torch.cuda.ExternalStream(ctx.stream)
results = []


def run(tag, reps=20, **opts):
    for k, v in opts.items():
        ctx.set_option(k, v)
    A, f = box.device_system(ctx)
    x = f
    y = A.CreateColVector()
    for _ in range(3):
        A.Mult(x, y)
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        A.Mult(x, y)
    e1.record(st)
    ctx.sync()
    ms = e0.elapsed_time(e1) / reps
    gbs = A.MultBytes() / ms / 1e6
    results.append(dict(tag=tag, ms=ms, gbs=gbs, **opts))
    print("%-40s %8.3f ms %8.1f GB/s  %s" % (tag, ms, gbs, opts), flush=True)
    for k in opts:
        ctx.set_option(k, 1 if k == "sell_schedule" else (-1 if k == "sell_sigma" else 0))
    del A, y


print("system: m=%d order=%d kind=%d ndof=%d" % (m, order, kind, box.ndof))
if mode == "ncu":
    run("sell", reps=2, spmv_algo=3)
    run("stream default", reps=2, spmv_algo=2)
    sys.exit(0)
if mode == "sigma":
    for sigma in (-1, 0):
        run("sell sigma=%d" % sigma, reps=10, spmv_algo=3, sell_sigma=sigma)
    ctx.set_option("sell_sigma", -1)
    if kind == 3:
        for sched in (2, 0):
            run("sell block3 sched=%d" % sched, reps=10, spmv_algo=3, sell_schedule=sched)
    sys.exit(0)
if mode == "schedncu":
    for sched in (1, 0):
        run("sell sched=%d" % sched, reps=1, spmv_algo=3, sell_schedule=sched)
    sys.exit(0)
if mode == "sched":
    for sched in (1, 0):
        for var in ((0, 3) if kind == 0 else (0,)):
            for cps in (0, 8, 12):
                run("sell sched=%d var=%d ctas/sm=%d" % (sched, var, cps), reps=10, spmv_algo=3, sell_schedule=sched, sell_variant=var, spmv_ctas_per_sm=cps)
    sys.exit(0)
for var in ((0, 1, 2, 3, 4, 5) if kind == 0 else (0,)):
    for cps in (0, 4, 5, 6, 8, 10, 12, 16):
        run("sell var=%d ctas/sm=%d" % (var, cps), spmv_algo=3, sell_variant=var, spmv_ctas_per_sm=cps)
run("subwarp W=8", spmv_algo=1, spmv_subwarp=8)
run("stream default", spmv_algo=2)
if mode == "quick":
    sys.exit(0)
base_tile = {0: 2048, 1: 1024, 3: 512}[kind]
for tile, ncw, stages, sw in itertools.product((base_tile, base_tile // 2, base_tile // 4), (8, 16), (3, 4, 6), (4, 8, 16)):
    vb = {0: 8, 1: 16, 3: 72}[kind]
    smem = 128 + stages * (tile * (vb + 4) + 1200)
    if smem > 220 * 1024:
        continue
    try:
        run("stream tile=%d ncw=%d st=%d W=%d" % (tile, ncw, stages, sw), spmv_algo=2, spmv_tile=tile, spmv_ncw=ncw, spmv_stages=stages, spmv_subwarp=sw)
    except Exception as e:
        print("FAILED", tile, ncw, stages, sw, e)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(results, open("gpurun_out/sweep_m%d_o%d_k%d.json" % (m, order, kind), "w"), indent=1)
best = sorted(results, key=lambda r: -r["gbs"])[:8]
print("BEST:")
for r in best:
    print(r)
