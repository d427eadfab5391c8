"""Measure the five BASELINE.json configs on one B200 (parity-test cases of the bench, not bench lines):
SpMV GB/s (algorithmic bytes of SURVEY.md 8d / CUDA-event time, median of the launches) and solver steps/s.

  python tools/bench_configs.py [c1 c2 c3 c4 c5] [--out profiles/r1_configs.jsonl] [--scale 1.0]

Systems: C1/C2/C3/C5 from the library's FE generator at the config's size (structured Kuhn mesh, NGSolve numbering);
C4 (HCurl order 2 on a netgen mesh: irregular row lengths) = the reference-assembled fixture
tests/golden/maxwell_hcurlp2.npz (1836 rows) repeated block-diagonally to ~20 M rows: the row-length distribution
and the per-row column spread are the reference's, the coupling between copies is absent (so x locality is
better than on one big mesh; the SELL padding / sigma-sorting behaviour is what this case exercises).
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import ngsolve_b200.la as la
from ngsolve_b200 import workloads as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def time_mult(ctx, A, x, y, reps=50):
    # SURVEY 8d metric 1: median of >= 50 CUDA-event-timed Mult calls after 5 warm-ups, same x/y buffers
    st = torch.cuda.ExternalStream(ctx.stream)
    for _ in range(5):
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


def time_solve(ctx, inv, f, u):
    inv.Mult(f, u)                       # warm-up (graph build, workspace)
    ctx.sync()
    st = torch.cuda.ExternalStream(ctx.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    inv.Mult(f, u)
    e1.record(st)
    ctx.sync()
    return e0.elapsed_time(e1) * 1e-3


def tiled_fixture(name, copies):
    g = np.load(os.path.join(ROOT, "tests", "golden", name))
    rp, col, val = g["rowptr"].astype(np.int64), g["col"].astype(np.int32), g["val"]
    n, nnz = len(rp) - 1, int(rp[-1])
    RP = (rp[None, :-1] + (np.arange(copies, dtype=np.int64) * nnz)[:, None]).reshape(-1)
    RP = np.concatenate([RP, [copies * nnz]]).astype(np.uint64)
    COL = (col[None, :] + (np.arange(copies, dtype=np.int32) * np.int32(n))[:, None]).reshape(-1)
    VAL = np.tile(val, copies)
    bits = np.unpackbits(g["freebits"], bitorder="little")[:n].astype(bool)
    return RP, COL, VAL, np.tile(g["f"], copies), np.tile(bits, copies)


def run(cfg, scale, spmv_only=False):
    ctx = la.default_context()
    t0 = time.perf_counter()
    name = cfg["name"]
    if "fixture" in cfg:
        copies = max(1, int(cfg["copies"] * scale))
        rp, col, val, f_np, free = tiled_fixture(cfg["fixture"], copies)
        A = la.SparseMatrix(rp, col, val).CreateDeviceMatrix()
        f = la.BaseVector(f_np, ctx=ctx)
        jac = A.CreateSmoother(la.BitArray(free))
        del rp, col, val
        es, cplx = 1, False
    else:
        m = max(2, int(round(cfg["m"] * scale ** (1.0 / 3.0))))
        box = W.FemBox(m, order=cfg["order"], kind=cfg["kind"], lame=cfg.get("lame", (0.0, 1.0)), mass=cfg.get("mass", 0.0))
        A, f = box.device_system(ctx)
        jac = A.CreateSmoother(box.freedofs())
        es, cplx = box.entrysize, cfg["kind"] == W.COMPLEX
    ctx.sync()
    setup = time.perf_counter() - t0
    x = f.CreateVector()
    x.SetRandom(1)
    y = A.CreateColVector()
    ms = time_mult(ctx, A, x, y)
    b = A.MultBytes()
    gbs = b / ms / 1e6
    ent, ovf, cap = A.Layout()
    sb, c16 = A.StreamBytes()
    out = dict(config=name, rows=A.height, scalars=A.height * es * (2 if cplx else 1), nnz=A.nze, entry="complex" if cplx else ("3x3" if es == 3 else "real"),
               spmv_ms=ms, spmv_bytes=b, spmv_gbs=gbs, spmv_frac_of_measured_peak=gbs / peak(), spmv_pct_of_8TBs=gbs / 80.0,
               sell_padding=ent / max(1, A.nze) - 1.0, sell_overflow_rows=ovf, setup_s=setup,
               stored_bytes=sb, spmv_gbs_on_stored_bytes=sb / ms / 1e6, c16_share_of_entries=c16 / max(1, ent))
    if spmv_only:
        return out
    u = f.CreateVector()
    K = cfg["steps"]
    if cfg["solver"] == "cg":
        inv = la.CGSolver(A, jac, precision=0.0, maxsteps=K)
        s = time_solve(ctx, inv, f, u)
        its = inv.GetSteps() - 1
        S = 16 if cplx else 8
        b_cg = b + 11 * A.height * es * S + (6 * A.height * es * S if es == 3 else 0)
        out.update(solver="CG+Jacobi", iterations=its, it_per_s=its / s, cg_gbs=b_cg * its / s / 1e9, cg_frac_of_measured_peak=b_cg * its / s / 1e9 / peak())
        inv = la.CGSolver(A, jac, precision=1e-8, maxsteps=50000)
        s = time_solve(ctx, inv, f, u)
        out.update(full_solve_steps=inv.GetSteps(), full_solve_s=s)
    else:
        inv = la.GMRESSolver(A, jac, precision=0.0, maxsteps=K)
        s = time_solve(ctx, inv, f, u)
        its = inv.GetSteps()
        S = 16 if cplx else 8
        # step j moves B_spmv + 3 N S + (2 (j+1) + 4) N S
        bytes_tot = sum(b + (2 * (j + 1) + 7) * A.height * es * S for j in range(its))
        # modified Gram-Schmidt as the reference does it (cg.cpp:927-932) cannot do better than 4 passes per projection
        # (read w, v_i, v_i+1; write w): the traffic this implementation really needs
        mgs_bytes = sum(b + (4 * (j + 1) + 3 + 3 + 2) * A.height * es * S for j in range(its))
        out.update(solver="GMRES+Jacobi (no restart)", iterations=its, it_per_s=its / s, gmres_gbs=bytes_tot / s / 1e9,
                   gmres_frac_of_measured_peak=bytes_tot / s / 1e9 / peak(), gmres_mgs_gbs=mgs_bytes / s / 1e9,
                   gmres_mgs_frac_of_measured_peak=mgs_bytes / s / 1e9 / peak(), solve_s=s)
    ctx.set_option("timing", 1)
    ctx.kernel_time_reset()
    t0 = time.perf_counter()
    inv.Mult(f, u)
    ctx.sync()
    wall = time.perf_counter() - t0
    out["kernel_ms_by_class"] = {k: ctx.kernel_time(k) for k in ("spmv", "cgupdate", "vec", "other", "all")}
    out["timed_solve_wall_ms"] = wall * 1e3
    ctx.kernel_time_reset()
    ctx.set_option("timing", 0)
    return out


CONFIGS = {
    "c1": dict(name="C1 Poisson H1 p3 ~1M dofs, Jacobi-CG", m=33, order=3, kind=W.REAL, solver="cg", steps=200),
    "c2": dict(name="C2 elasticity H1 p4 dim 3, ~10M scalar dofs, 3x3 block CSR, Jacobi-CG", m=37, order=4, kind=W.BLOCK3, lame=(58.3, 87.5),
               solver="cg", steps=100),
    "c3": dict(name="C3 Poisson H1 p3 ~100M dofs, Jacobi-CG (the bench workload)", m=160, order=3, kind=W.REAL, solver="cg", steps=50),
    "c4": dict(name="C4 Maxwell HCurl p2 (reference-assembled fixture tiled to ~20M rows), Jacobi-CG", fixture="maxwell_hcurlp2.npz", copies=10900,
               solver="cg", steps=100),
    "c5": dict(name="C5 Helmholtz H1 p4 complex ~30M dofs, Jacobi-GMRES", m=77, order=4, kind=W.COMPLEX, mass=-100.0 - 10.0j, solver="gmres", steps=40),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["c1", "c2", "c4", "c5"])
    ap.add_argument("--out", default=None)
    ap.add_argument("--scale", type=float, default=1.0, help="scale the number of rows (smoke runs)")
    ap.add_argument("--opt", action="append", default=[], help="context option name=value (e.g. spmv_ctas_per_sm=96), repeatable")
    ap.add_argument("--spmv-only", action="store_true")
    a = ap.parse_args()
    lines = []
    for o in a.opt:
        name, val = o.split("=")
        la.default_context().set_option(name, int(val))
    for k in a.which:
        r = run(CONFIGS[k], a.scale, a.spmv_only)
        r["options"] = a.opt
        print(json.dumps(r), flush=True)
        lines.append(r)
        import gc
        gc.collect()
    if a.out:
        with open(a.out, "w") as fh:
            for r in lines:
                fh.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
