#!/usr/bin/env python
"""bench.py -- CG iterations/s (and SpMV HBM GB/s) of the assembled-system solve path on B200.

Workload (BASELINE.json configs[2], the config the metric and target are quoted on; it fits one
GPU): Poisson, H1 order 3, unit cube, ~100 M dofs, Jacobi-preconditioned CG.  The system is
assembled on the device by the library's synthetic FE generator (structured Kuhn tetrahedral
mesh, NGSolve dof numbering, see include/ngsb200_workloads.h) because netgen/NGSolve do not exist
on the GPU box.  One step = one CG iteration (SpMV + fused vector updates / dots / Jacobi).
At N > 1 the SAME global grid is split into N slabs of elements, one per rank (strong scaling,
the reference's ParallelDofs / ParallelMatrix split; interface values move by NCCL over NVLink).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size M] [--impl reference]

--impl reference times the reference itself -- NGSolve's C++ CGSolver + JacobiPrecond under its TaskManager on all host
cores (oracle/ref_cpu_cg.py, the build of oracle/build_reference.sh travels to the GPU box in oracle/_ref/ngs) -- on a
bounded sample of the same workload family, scaled to the full size by the nnz ratio; only when that build is absent
the oracle port of the same path (oracle/ngs_oracle.c, OpenMP) stands in (cpu_baseline.kind says which).

config.netgen_check (N = 1, when the reference build is on the box): the same library on a netgen-meshed system in
NGSolve's own dof numbering -- unit_cube maxh=0.05 + 2x Refine, H1 order 3, 13.6 M dofs (the C3 fallback) --
with the automatic Cuthill-McKee reordering, SpMV and CG against the same roofline, the reference's CPU CG beside it.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "cg_iterations_per_s"
UNIT = "iterations/s"
SAMPLE_M_T1 = 20         # single-thread CPU sample: 0.23 M dofs, 10.7 M non-zeros (128 MB of matrix)
SAMPLE_M = 50            # CPU sample: (3*50+1)^3 = 3.4 M dofs, 163 M non-zeros (2 GB of matrix: far out of the host caches)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=160, help="cubes per axis of the global grid; dofs = (3*size+1)^3 "
                    "(160 -> 111 M dofs; divisible by 8 so that the z-slabs of the multi-GPU runs are equal)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--opt", action="append", default=[], help="library option name=value set on the context before the matrix "
                    "is created (e.g. dist_overlap=1, spmv_ctas_per_sm=64); repeatable; recorded in config.options")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-full-solve", action="store_true")
    ap.add_argument("--no-netgen-check", action="store_true")
    ap.add_argument("--netgen-nref", type=int, default=3, help="Refine() steps of the netgen system (3 -> 108 M dofs = configs[2], 2 -> 13.6 M)")
    ap.add_argument("--cpu-iters", type=int, default=150)
    ap.add_argument("--cpu-sample", type=int, default=SAMPLE_M, help="cubes per axis of the CPU arm's sample system")
    ap.add_argument("--cpu-sample-t1", type=int, default=SAMPLE_M_T1, help="the same for the single-thread run")
    return ap.parse_args()


def workload_name(m):
    return "Poisson unit cube H1 order 3, %d^3 cubes x 6 tets, %.1fM dofs, Jacobi-PCG" % (m, (3 * m + 1) ** 3 / 1e6)


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path
# ---------------------------------------------------------------------------------------------
def _reference_env():
    """environment of the reference build (oracle/build_reference.sh -> oracle/_ref/ngs), or None when it is absent"""
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


def cpu_cg_reference(iters, warmup, full_nnz):
    """the reference itself (NGSolve's C++ CGSolver under its TaskManager) on the sample system, in a subprocess
    (`import ngsolve` has to precede numpy/torch).  None when the build is not on this box or does not run."""
    env = _reference_env()
    if env is None:
        return None
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_cpu_cg.py"), "--m", str(SAMPLE_M), "--iters", str(iters),
                            "--warmup", str(warmup)], env=env, capture_output=True, text=True, timeout=900)
        d = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:                    # noqa: BLE001  (report and fall back to the port)
        sys.stderr.write("reference CPU arm unavailable: %r\n" % (e,))
        return None
    raw, nnz, n = d["it_per_s"], d["nnz"], d["ndof"]
    b_iter = nnz * 12 + n * (4 + 8 + 8) + 21 * n * 8
    # SURVEY 8d asks for T = 1 beside T = all cores: a second process (one thread count per process, 8a1) on a smaller sample
    one = None
    try:
        r1 = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_cpu_cg.py"), "--m", str(SAMPLE_M_T1), "--iters", "20",
                             "--warmup", "2", "--threads", "1"], env=env, capture_output=True, text=True, timeout=300)
        d1 = json.loads(r1.stdout.strip().splitlines()[-1])
        one = dict(value=d1["it_per_s"] * d1["nnz"] / full_nnz, raw_iterations_per_s=d1["it_per_s"], cores=1, sample_ndof=d1["ndof"],
                   sample_nnz=d1["nnz"], spmv_gbs=d1["spmv_gbs"],
                   sample="same solver with SetNumThreads(1), 20 iterations on the %.2fM-dof system (%d^3 cubes), scaled by nnz"
                          % (d1["ndof"] / 1e6, SAMPLE_M_T1))
    except Exception as e:                    # noqa: BLE001
        sys.stderr.write("single-thread reference run unavailable: %r\n" % (e,))
    return dict(value=raw * nnz / full_nnz, unit=UNIT, cores=d["threads"], kind="reference", single_thread=one,
                sample="NGSolve %s C++ CGSolver + JacobiPrecond under TaskManager(%d threads), %d iterations on the %.2fM-dof system of the same "
                       "family (%d^3 cubes, injected by CreateFromCOO): %.1f it/s = %.1f GB/s of the reference's per-iteration traffic, SpMV alone "
                       "%.2f ms = %.1f GB/s; scaled by nnz ratio %.4g to the full size"
                       % (d["ngsolve_version"], d["threads"], d["iterations"], n / 1e6, SAMPLE_M, raw, raw * b_iter / 1e9, d["spmv_ms"], d["spmv_gbs"],
                          nnz / full_nnz),
                raw_iterations_per_s=raw, sample_ndof=n, sample_nnz=nnz, reference_setup_s=d["setup_s"])


def cpu_cg(iters, full_nnz, full_ndof, warmup=3):
    """CPU baseline on the sample system: the reference itself when its build travelled with the repo, else the
    oracle port (reference recurrences, 16-chunk dots, OpenMP rows).  dict(value scaled to the full size, raw numbers)."""
    ref = cpu_cg_reference(iters, warmup, full_nnz)
    if ref is not None:
        return ref
    from oracle import pyoracle as orc
    from ngsolve_b200 import workloads as W
    box = W.FemBox(SAMPLE_M, order=3)
    rp, col, val, rhs = box.host_csr()          # host loop of the generator (no GPU involved)
    free = box.freedofs()
    A = orc.Csr(rp, col, val, 0)
    J = orc.Jacobi(A, free.bytes)
    cores = orc.max_threads()
    orc.cg_solve(A, J, rhs, prec=0.0, maxsteps=warmup)                     # warm-up
    t0 = time.perf_counter()
    _, steps, _ = orc.cg_solve(A, J, rhs, prec=0.0, maxsteps=iters)
    dt = time.perf_counter() - t0
    its = steps - 1
    raw = its / dt
    nnz = int(rp[-1])
    b_iter = nnz * 12 + box.ndof * (4 + 8 + 8) + 21 * box.ndof * 8      # reference's unfused sequence
    return dict(value=raw * nnz / full_nnz, unit=UNIT, cores=cores, kind="port",
                sample="oracle Jacobi-PCG, %d iterations on the %.2fM-dof system of the same family (%d^3 cubes): %.1f it/s "
                       "= %.1f GB/s of the reference's per-iteration traffic; scaled by nnz ratio %.4g to the full size"
                       % (its, box.ndof / 1e6, SAMPLE_M, raw, raw * b_iter / 1e9, nnz / full_nnz),
                raw_iterations_per_s=raw, sample_ndof=box.ndof, sample_nnz=nnz)


def full_sizes(m):
    """(ndof, nnz) of the full single-box system without assembling it: row pointers of the sample
    family are cheap, but for the full size use the generator's count on a small box and scale?  No:
    ask the generator itself (row-pointer pass only is O(ndof) on the host, too slow at 100 M), so
    use the closed forms: ndof = (3m+1)^3; nnz from the entity stencils = measured by the GPU arm and
    cached next to this file, else estimated from the sample's nnz/dof."""
    ndof = (3 * m + 1) ** 3
    cache = os.path.join(ROOT, "profiles", "workload_sizes.json")
    if os.path.exists(cache):
        try:
            d = json.load(open(cache))
            if str(m) in d:
                return ndof, int(d[str(m)])
        except Exception:
            pass
    return ndof, None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from ngsolve_b200 import workloads as W
    m = args.size
    ndof, nnz = full_sizes(m)
    if nnz is None:
        # nnz per dof of this mesh family is size independent up to surface terms: take it from the sample
        box = W.FemBox(SAMPLE_M, order=3)
        rp = np.empty(box.ndof + 1, dtype=np.uint64)
        W.check(W._lib().ngsb_femgen_host(box.handle, rp.ctypes.data, None, None, None))
        nnz = int(int(rp[-1]) / box.ndof * ndof)
    iters = max(1, args.steps)
    t0 = time.perf_counter()
    base = None
    if not args.no_netgen_check:
        # the reference at FULL SIZE: NGSolve assembles the netgen configs[2] system (108 M dofs) on this box and its own
        # CGSolver runs `steps` iterations after `warmup` -- measured, nothing extrapolated
        d = netgen_big(args.netgen_nref, ["--cpu-only", "--cpu-iters", str(iters), "--cpu-warmup", str(max(1, args.warmup))])
        if d and d.get("cpu_reference_it_per_s"):
            base = dict(value=d["cpu_reference_it_per_s"], unit=UNIT, cores=d["threads"], kind="reference",
                        sample="NGSolve %s C++ CGSolver + JacobiPrecond under TaskManager(%d threads): %d iterations (after %d warm-up) on the netgen "
                               "unit_cube maxh=%g + %dx Refine H1 order-%d system, %.1f M dofs, %.2f G non-zeros, assembled on this box in %.0f s -- the "
                               "full-size configs[2] system, measured, not extrapolated"
                               % (d["ngsolve"], d["threads"], d["cpu_reference_iters"], max(1, args.warmup), d["maxh"], d["nref"], d["order"], d["ndof"] / 1e6,
                                  d["nnz"] / 1e9, d["assemble_s"]),
                        sample_ndof=d["ndof"], sample_nnz=d["nnz"], reference_setup_s=d["assemble_s"])
            ndof, nnz = d["ndof"], d["nnz"]
    if base is None:
        base = cpu_cg(iters, nnz, ndof, warmup=max(1, args.warmup))
    dt = time.perf_counter() - t0
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 / base["value"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64",
        "data": "netgen mesh (unit_cube, maxh 0.05, refined), assembled by the reference on this box" if "netgen" in base["sample"] else "synthetic",
        "config": {"workload": workload_name(m), "global_dofs": ndof, "nnz": nnz, "precond": "Jacobi (freedofs-masked)",
                   "note": "CPU arm = the reference itself (NGSolve CGSolver under TaskManager, build of oracle/build_reference.sh travels with the "
                           "repo) on the full-size netgen configs[2] system -- the same operator and size class as the GPU arm's timed workload "
                           "(whose config.netgen_check times the library on this very system); without the build: the oracle port on a bounded "
                           "sample scaled by nnz.  kind = " + base["kind"] + "; " + base["sample"][:200]},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": dt,
    }
    emit(line)


def netgen_big(nref, extra, timeout=1500):
    """tools/netgen_big.py in a subprocess under the reference environment (`import ngsolve` precedes numpy; its own CUDA
    context): netgen unit_cube maxh=0.05 + nref x Refine, H1 order 3, assembled by the reference; None without the build."""
    env = _reference_env()
    if env is None:
        return None
    if nref >= 3:
        # the 108 M-dof system needs about 80 GB of host memory while NGSolve holds it: fall back to two refinements (13.6 M dofs)
        try:
            avail = [int(l.split()[1]) for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0] / 2 ** 20
            if avail < 110:
                nref = 2
        except Exception:
            pass
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "netgen_big.py"), "--nref", str(nref)] + extra, env=env,
                           capture_output=True, text=True, timeout=timeout)
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:                    # noqa: BLE001 -- reported, never fatal for the bench line
        return {"error": repr(e)[:300]}


def netgen_check(nref, cpu_iters):
    """BASELINE configs[2] on a REAL netgen mesh (nref = 3: 23.8 M tets, 108.0 M dofs, 5.22 G non-zeros) in NGSolve's own dof
    numbering: the library with its automatic Cuthill-McKee reordering, and the reference's CPU CGSolver on the very same
    system (measured at full size, nothing extrapolated)."""
    d = netgen_big(nref, ["--full", "--cpu-iters", str(cpu_iters)])
    if d is None or "error" in d:
        return d
    keep = ("ne", "nv", "ndof", "nnz", "sha256_rowptr", "mesh_s", "assemble_s", "create_device_matrix_s", "reordered", "natural_c16_share",
            "c16_share_of_entries", "sell_padding", "csr_arrays_resident", "sell_bytes", "spmv_kernel_ms", "spmv_call_ms_incl_gather",
            "spmv_gbs_algorithmic", "spmv_frac_of_peak", "spmv_gbs_stored", "cg_it_per_s", "cg_frac_of_peak", "full_solve_steps", "full_solve_s",
            "cpu_reference_it_per_s", "cpu_reference_iters", "threads")
    out = {"mesh": "netgen unit_cube maxh=%g + %dx Refine, H1 order %d, NGSolve dof numbering (assembled by NGSolve %s on this box)"
                   % (d["maxh"], d["nref"], d["order"], d["ngsolve"])}
    out.update({k: d[k] for k in keep if k in d})
    if d.get("cpu_reference_it_per_s"):
        out["speedup_vs_cpu_reference_same_system"] = d["cg_it_per_s"] / d["cpu_reference_it_per_s"]
    return out


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                pass

    def summary(self, t0, t1):
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 and len(r) >= 7] or [r for (_, r) in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "power_w_max": max(float(r[2]) for r in rows),
                "samples": len(rows), "reasons": reasons}


def run_b200(args):
    import torch
    import torch.distributed as dist
    import ctypes as C
    from ngsolve_b200 import la, workloads as W, _capi
    from ngsolve_b200 import parallel as par

    # NCCL prints its version banner to stdout when NCCL_DEBUG=VERSION is in the environment; stdout carries
    # exactly one JSON line
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)" % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    ng_check = None
    if world == 1 and not args.no_netgen_check:
        ng_check = netgen_check(args.netgen_nref, 0 if args.no_cpu_baseline else 4)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = la.Context(local_rank)
    for o in args.opt:
        name, val = o.split("=")
        ctx.set_option(name, int(val))
    m, K, Wm = args.size, args.steps, args.warmup
    G = (m, m, m)
    t_setup = time.perf_counter()
    if world == 1:
        box = W.FemBox(G, order=3)
        A, f = box.device_system(ctx)
        jac = A.CreateSmoother(box.freedofs())
        solver = None
    else:
        n, off = W.slab_partition(G, world)[rank]
        boxes = [W.FemBox(nn, order=3, offset=oo, global_n=G) for nn, oo in W.slab_partition(G, world)]
        box = boxes[rank]
        A, f = box.device_system(ctx)
        comm = par.Communicator(ctx, world, rank, dist)
        pd = par.ParallelDofs(*W.exchange_tables(boxes, rank), ndof=box.ndof, nranks=world, rank=rank)
        pmat = par.ParallelMatrix(A, pd, comm)
        jac = pmat.CreateSmoother(box.freedofs())
        solver = pmat
    ctx.sync()
    setup_s = time.perf_counter() - t_setup
    nnz_local, ndof_local = A.nze, A.height
    counts = torch.tensor([nnz_local, ndof_local], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(counts)
    nnz_sum = int(counts[0].item())
    global_ndof = box.global_ndof

    stream = torch.cuda.ExternalStream(ctx.stream)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    u = f.CreateVector()

    def solve(maxsteps, prec=0.0):
        if world == 1:
            inv = la.CGSolver(A, jac, precision=prec, maxsteps=maxsteps)
            inv.Mult(f, u)
            return inv
        return pmat.cg_solve(jac, f, u, precision=prec, maxsteps=maxsteps)

    def timed(fn):
        """fn() between CUDA events on the library's stream, barrier + sync on both sides; ms, max over ranks"""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        out = fn()
        e1.record(stream)
        barrier()
        t1 = time.perf_counter()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out, (t0, t1)

    # ---- warm-up (W iterations; also builds the CUDA graph of a batch) then the timed K iterations
    solve(max(Wm, 3))
    solve(K)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    l0 = ctx.launches
    ms, inv, span = timed(lambda: solve(K))
    launches = ctx.launches - l0
    its = inv.GetSteps() - 1
    assert its == K, "CG stopped after %d of %d iterations" % (its, K)
    value = its / (ms * 1e-3)

    # ---- roofline of the dominant kernel: same K iterations with an event pair around every launch
    ctx.set_option("timing", 1)
    ctx.kernel_time_reset()
    solve(K)
    spmv_ms, spmv_n = ctx.kernel_time("spmv")
    upd_ms, upd_n = ctx.kernel_time("cgupdate")
    all_ms, all_n = ctx.kernel_time("all")
    ctx.kernel_time_reset()
    ctx.set_option("timing", 0)
    b_spmv = A.MultBytes()                                  # nnz*12 + 4h + 8h + 8h
    sell_entries, sell_ovf, sell_cap = A.Layout()
    stream_bytes, c16_entries = A.StreamBytes()          # what the kernel streams as stored (16-bit column offsets where possible)
    # exactly K launches do work; the rest of the last batch returns at once on the device-side `done` flag
    # (a few microseconds each), so the class total divided by K is the per-launch duration
    t_spmv = spmv_ms / K * 1e-3
    achieved = b_spmv / t_spmv / 1e9
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"

    # DRAM traffic of the dominant kernel per launch from the committed `ncu --set full` capture of this very command
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if world == 1 and str(m) in tj and c16_entries > 0:
            traffic, traffic_src = tj[str(m)]["traffic_bytes"], tj[str(m)]["capture"]
    except Exception:
        pass

    # ---- e2e: the C-ABI call a script's `gfu.vec.data = inv * f.vec` makes, host buffers, copies timed
    e2e = None
    if world == 1:
        f_host = torch.empty(ndof_local, dtype=torch.float64, pin_memory=True)
        u_host = torch.empty(ndof_local, dtype=torch.float64, pin_memory=True)
        f_host.copy_(torch.from_numpy(f.NumPy()))
        steps_c, nh_c = C.c_int(), C.c_int()

        def host_call():
            _capi.check(_capi.lib().ngsb_cg_solve_host(A.handle, jac.handle, f_host.data_ptr(), u_host.data_ptr(), 0.0, K, 0,
                                                       C.byref(steps_c), None, 0, C.byref(nh_c)))
        host_call()
        t0 = time.perf_counter()
        host_call()
        dt = time.perf_counter() - t0            # wall clock: the call returns after the D2H copy completed
        assert steps_c.value - 1 == K
        e2e = {"value": K / dt, "unit": UNIT, "h2d_bytes_per_step": ndof_local * 8 / K, "d2h_bytes_per_step": ndof_local * 8 / K,
               "note": "ngsb_cg_solve_host: f (pinned host) -> device, %d iterations, u -> pinned host; bytes are per call / %d" % (K, K)}
    else:
        f_host = torch.from_numpy(f.NumPy()).pin_memory()
        u_host = torch.empty_like(f_host).pin_memory()
        fv = la.BaseVector(ndof_local, ctx=ctx)

        def host_call():
            _capi.check(_capi.lib().ngsb_vec_h2d(fv.handle, f_host.data_ptr(), 0, ndof_local))
            r = pmat.cg_solve(jac, fv, u, precision=0.0, maxsteps=K)
            _capi.check(_capi.lib().ngsb_vec_d2h(u.handle, u_host.data_ptr(), 0, ndof_local))
            return r
        host_call()
        barrier()
        t0 = time.perf_counter()
        host_call()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": K / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": ndof_local * 8 / K,
               "d2h_bytes_per_step": ndof_local * 8 / K, "note": "per rank: local f host->device, distributed CG, local u device->host"}

    # ---- a full solve to the reference's default tolerance (not the timed number; shows convergence)
    full = None
    if not args.no_full_solve:
        msf, invf, _ = timed(lambda: solve(20000, 1e-8))
        hist = invf.history
        full = {"precision": 1e-8, "steps": invf.GetSteps(), "seconds": msf * 1e-3,
                "wdn_reduction": float(hist[-1] / hist[0]) if len(hist) > 1 and hist[0] > 0 else None}
    if rank == 0:
        sampler.stop()
    clocks = sampler.summary(*span) if rank == 0 else None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:      # the CPU baseline is reported at N = 1 only
        if ng_check and ng_check.get("cpu_reference_it_per_s"):
            # measured at full size on the netgen system of the same operator (108.0 M dofs against the 111.3 M of the timed
            # generator system): nothing extrapolated
            cpu = dict(value=ng_check["cpu_reference_it_per_s"], unit=UNIT, cores=ng_check.get("threads"), kind="reference",
                       sample="NGSolve C++ CGSolver + JacobiPrecond under TaskManager(%s threads), %d iterations on the FULL-SIZE netgen system of "
                              "config.netgen_check (%.1f M dofs, %.2f G non-zeros; Poisson H1 order 3 like the timed workload): measured, not "
                              "extrapolated" % (ng_check.get("threads"), ng_check.get("cpu_reference_iters", 0), ng_check["ndof"] / 1e6, ng_check["nnz"] / 1e9),
                       sample_ndof=ng_check["ndof"], sample_nnz=ng_check["nnz"])
        else:
            cpu = cpu_cg(args.cpu_iters, nnz_sum, global_ndof)
    if rank == 0:
        if world == 1:
            try:
                cache = os.path.join(ROOT, "profiles", "workload_sizes.json")
                d = json.load(open(cache)) if os.path.exists(cache) else {}
                d[str(m)] = nnz_sum
                json.dump(d, open(cache, "w"))
            except Exception:
                pass
        b_cg = b_spmv + 11 * ndof_local * 8
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(Wm, 3), "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(m), "global_dofs": global_ndof, "nnz_total": nnz_sum, "rows_per_gpu": ndof_local,
                       "precond": "Jacobi (freedofs-masked)",
                       "partition": "1 box" if world == 1 else "%d z-slabs of elements, %s" % (world, "interface exchange + scalar all-reduces by P2P stores "
                                    "into peer memory inside the solver kernels (no NCCL call per iteration)" if pmat.peer_memory else "NCCL send/recv halo + NCCL all-reduce")
                                    + ("; interface slices first, pushed while the interior slices are multiplied (%d + %d slices)" % pmat.overlap[1:] if pmat.overlap[0] else ""),
                       "options": list(args.opt),
                       "l2": "inputs (%.1f GB per GPU) far larger than the 126 MB L2" % (b_spmv / 1e9),
                       "setup_s": setup_s, "full_solve": full, "netgen_check": ng_check,
                       "cg_gbs_per_gpu": b_cg * value / 1e9, "cg_bytes_per_iteration_per_gpu": b_cg,
                       "spmv_pct_of_8TBs": achieved / 8000.0 * 100.0,
                       "sell_padding": sell_entries / max(1, nnz_local) - 1.0, "sell_overflow_rows": sell_ovf},
            "roofline": {"bound": "hbm", "kernel": "sell_spmv_kernel (SELL-32 SpMV + fused <s,As> + CG alpha step)", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": b_spmv, "stored_bytes_per_launch": stream_bytes,
                         "achieved_on_stored_bytes": stream_bytes / t_spmv / 1e9,
                         "c16_share_of_entries": c16_entries / max(1, sell_entries),
                         "note": "achieved/frac use the ALGORITHMIC bytes of SURVEY 8d (12 B per entry); the kernel stores 16-bit column "
                                 "offsets for most slices and streams stored_bytes_per_launch, which is why frac can exceed the copy peak", "avg_launch_ms": t_spmv * 1e3, "launches_timed": K, "launches_enqueued": spmv_n,
                         "kernel_share_of_step": spmv_ms / all_ms if all_ms else None,
                         "cg_update_kernels_ms_per_iteration": upd_ms / K},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """the one JSON line, on the process's original stdout"""
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    # stdout carries exactly one JSON line: native libraries (the NCCL version banner, ...) write to fd 1 directly, so
    # fd 1 is pointed at stderr for the whole run and the line goes to a saved duplicate of the original stdout
    global _REAL_STDOUT, SAMPLE_M, SAMPLE_M_T1
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse()
    SAMPLE_M, SAMPLE_M_T1 = args.cpu_sample, args.cpu_sample_t1
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
