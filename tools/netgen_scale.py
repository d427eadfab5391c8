#!/usr/bin/env python
"""SpMV / CG of the library on a netgen-numbered system: the reference build that travels with the
repo (oracle/_ref/ngs) meshes and assembles, the library multiplies it as numbered (reorder = 0) and after its
Cuthill-McKee reordering (reorder = 1), and the reference's CPU CGSolver is timed on the same system.

    python tools/netgen_scale.py --nref 2 --out gpurun_out/r2_netgen_14M.json          # maxh 0.05 + 2x Refine: 13.9 M dofs
    ncu ... python tools/netgen_scale.py --cache /dev/shm/ngsys --modes 1 --spmv-only  # profile the product

One JSON object: per mode SpMV ms / GB/s on algorithmic and stored bytes, share of 16-bit-offset entries, CG it/s, steps."""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def reference_env():
    pfx = os.path.join(ROOT, "oracle", "_ref", "ngs")
    site = os.path.join(pfx, "lib", "python3.12", "site-packages")
    if not os.path.isdir(os.path.join(site, "ngsolve")):
        return None
    import importlib.util
    spec = importlib.util.find_spec("cv2")
    libs = os.path.join(os.path.dirname(os.path.dirname(spec.origin)), "opencv_python_headless.libs") if spec and spec.origin else ""
    env = dict(os.environ)
    env["PYTHONPATH"] = site + os.pathsep + env.get("PYTHONPATH", "")
    env["LD_LIBRARY_PATH"] = os.pathsep.join([os.path.join(pfx, "lib"), os.path.join(site, "netgen"), libs, env.get("LD_LIBRARY_PATH", "")])
    return env


def generate(args):
    if os.path.exists(os.path.join(args.cache, "meta.json")):
        return json.load(open(os.path.join(args.cache, "meta.json")))
    env = reference_env()
    if env is None:
        raise SystemExit("netgen_scale: the reference build oracle/_ref/ngs is not here")
    cmd = [sys.executable, os.path.join(ROOT, "tools", "netgen_system.py"), "--maxh", str(args.maxh), "--nref", str(args.nref), "--order", str(args.order),
           "--cpu-iters", str(args.cpu_iters), "--out", args.cache] + (["--cpu-full"] if args.cpu_full else []) + (["--gpu-reference"] if args.gpu_reference else []) + (["--threads", str(args.threads)] if args.threads else [])
    r = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if r.returncode != 0:
        raise SystemExit("netgen_system failed:\n" + r.stdout[-2000:] + r.stderr[-4000:])
    return json.load(open(os.path.join(args.cache, "meta.json")))


def measure(args, meta):
    from ngsolve_b200 import la
    ctx = la.default_context()
    for o in args.opt:
        k, v = o.split("=")
        ctx.set_option(k, int(v))
    rowptr = np.load(os.path.join(args.cache, "rowptr.npy"), mmap_mode="r")
    col = np.load(os.path.join(args.cache, "col.npy"), mmap_mode="r")
    val = np.load(os.path.join(args.cache, "val.npy"), mmap_mode="r")
    fh = np.load(os.path.join(args.cache, "f.npy"))
    bits = np.load(os.path.join(args.cache, "freebits.npy"))
    A = la.SparseMatrix(np.asarray(rowptr), np.asarray(col), np.asarray(val))
    n = A.height
    out = {}
    peak = 6545.3
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    for mode in args.modes:
        ctx.set_option("reorder", mode)
        t0 = time.perf_counter()
        dev = A.CreateDeviceMatrix()
        ctx.sync()
        up = time.perf_counter() - t0
        on, share = dev.ReorderInfo()
        b_alg = dev.MultBytes()
        b_st, c16 = dev.StreamBytes()
        ent, ovf, cap = dev.Layout()
        x = la.BaseVector(np.random.default_rng(1).random(n))
        y = dev.CreateColVector()
        for _ in range(3):
            dev.Mult(x, y)
        ctx.sync()
        ctx.set_option("timing", 1)
        ctx.kernel_time_reset()
        for _ in range(args.reps):
            dev.Mult(x, y)
        ms, nl = ctx.kernel_time("spmv")
        vms, vn = ctx.kernel_time("all")
        ctx.kernel_time_reset()
        ctx.set_option("timing", 0)
        t_k = ms / args.reps * 1e-3                 # the SELL kernel alone
        t_all = vms / args.reps * 1e-3              # + the gather of x into the permuted numbering
        r = {"reordered": on, "natural_c16_share": share, "create_s": up, "sell_padding": ent / A.nze - 1.0, "overflow_rows": ovf,
             "c16_share_of_entries": c16 / max(1, ent), "algorithmic_bytes": b_alg, "stored_bytes": b_st,
             "spmv_kernel_ms": t_k * 1e3, "spmv_call_ms": t_all * 1e3,
             "spmv_gbs_algorithmic": b_alg / t_k / 1e9, "spmv_frac_algorithmic": b_alg / t_k / 1e9 / peak,
             "spmv_gbs_stored": b_st / t_k / 1e9, "spmv_frac_stored": b_st / t_k / 1e9 / peak,
             "spmv_call_gbs_algorithmic": b_alg / t_all / 1e9}
        if not args.spmv_only:
            jac = dev.CreateSmoother(la.BitArray(bits))
            f = la.BaseVector(fh)
            u = f.CreateVector()
            inv = la.CGSolver(dev, jac, precision=0.0, maxsteps=args.iters)
            inv.Mult(f, u)
            ctx.sync()
            t0 = time.perf_counter()
            inv.Mult(f, u)
            ctx.sync()
            dt = time.perf_counter() - t0
            its = inv.GetSteps() - 1
            b_cg = b_alg + 11 * n * 8
            r.update(cg_it_per_s=its / dt, cg_gbs=b_cg * its / dt / 1e9, cg_frac=b_cg * its / dt / 1e9 / peak)
            if args.full:
                inv = la.CGSolver(dev, jac, precision=1e-8, maxsteps=20000)
                t0 = time.perf_counter()
                inv.Mult(f, u)
                ctx.sync()
                r.update(full_steps=inv.GetSteps(), full_s=time.perf_counter() - t0)
                ur = os.path.join(args.cache, "u_ref.npy")
                if os.path.exists(ur):
                    uref = np.load(ur)
                    r["full_rel_diff_vs_reference"] = float(np.max(np.abs(u.NumPy() - uref)) / np.max(np.abs(uref)))
            del jac, f, u, inv
        out["reorder=%d" % mode] = r
        del dev, x, y
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--maxh", type=float, default=0.05)
    ap.add_argument("--nref", type=int, default=2)
    ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--cpu-iters", type=int, default=20)
    ap.add_argument("--cpu-full", action="store_true")
    ap.add_argument("--gpu-reference", action="store_true")
    ap.add_argument("--cache", default="/dev/shm/ngsys")
    ap.add_argument("--modes", type=int, nargs="+", default=[0, 1])
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--spmv-only", action="store_true")
    ap.add_argument("--opt", action="append", default=[])
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    t0 = time.perf_counter()
    meta = generate(args)
    res = {"system": meta, "generate_wall_s": time.perf_counter() - t0}
    res.update(measure(args, meta))
    s = json.dumps(res)
    if args.out:
        open(args.out, "w").write(s + "\n")
    print(s)


if __name__ == "__main__":
    main()
