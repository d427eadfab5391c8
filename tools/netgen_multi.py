#!/usr/bin/env python
"""CG on a netgen-numbered system across N GPUs (launch with torch.distributed.run, one rank per GPU).

The reference partitions the MESH (METIS) and every rank assembles its sub-domain; neither MPI nor METIS exist here, so the
globally assembled matrix of tools/netgen_system.py is split the way SURVEY.md 8(e) describes: dofs in the library's
Cuthill-McKee order (ngsb_csr_rcm, computed on each rank's GPU -- it is deterministic), N contiguous blocks of owned rows;
rank r holds its owned dofs plus the ghost dofs its rows couple to, its local matrix has the owned rows complete and the
ghost rows empty, so that A_loc x (x CUMULATED) is a DISTRIBUTED vector in exactly the reference's sense and the
ParallelDofs tables (dist_procs -> exchangedofs, lowest rank = master) follow from who holds what.  Everything after that
is the same distributed Jacobi-PCG as bench.py --gpus N (peer-memory exchange inside the solver kernels).

  python -m torch.distributed.run --nproc-per-node 8 ... tools/netgen_multi.py --cache /dev/shm/ng2 [--nref 2]
One JSON line from rank 0: it/s over --iters iterations, steps and time of the full solve, per-rank sizes."""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cache", default="/dev/shm/ng2")
    ap.add_argument("--nref", type=int, default=2)
    ap.add_argument("--maxh", type=float, default=0.05)
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--cpu-full", action="store_true", help="let the reference solve the system on the CPU first (u_ref for the comparison)")
    ap.add_argument("--opt", action="append", default=[])
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from ngsolve_b200 import la, parallel as par
    world, rank, lrank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lrank)
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    if rank == 0 and not os.path.exists(os.path.join(args.cache, "meta.json")):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "netgen_scale.py"), "--nref", str(args.nref), "--maxh", str(args.maxh),
                               "--cache", args.cache, "--modes", "0", "--spmv-only", "--reps", "1", "--cpu-iters", "10"] + (["--cpu-full"] if args.cpu_full else []),
                              stdout=subprocess.DEVNULL)
    dist.barrier()
    meta = json.load(open(os.path.join(args.cache, "meta.json")))
    rowptr = np.load(os.path.join(args.cache, "rowptr.npy"), mmap_mode="r")
    col = np.load(os.path.join(args.cache, "col.npy"), mmap_mode="r")
    val = np.load(os.path.join(args.cache, "val.npy"), mmap_mode="r")
    fglob = np.load(os.path.join(args.cache, "f.npy"))
    bits = np.load(os.path.join(args.cache, "freebits.npy"))
    n = len(rowptr) - 1
    ctx = la.Context(lrank)
    for o in args.opt:
        k, v = o.split("=")
        ctx.set_option(k, int(v))
    t0 = time.perf_counter()
    # ---- the ordering (deterministic: every rank gets the same permutation from its own GPU)
    ctx.set_option("reorder", 0)
    ctx.set_option("csr_keep", 1)
    G = la.DevSparseMatrix(la.SparseMatrix(np.asarray(rowptr), np.asarray(col), np.asarray(val), ctx=ctx), ctx=ctx)
    perm = G.RCM().astype(np.int64)                 # new -> old
    del G
    t_rcm = time.perf_counter() - t0
    # ---- my block of owned rows, ghosts, local matrix; dist_procs from everybody's ghost list
    cuts = [(n * r) // world for r in range(world + 1)]
    lo, hi = cuts[rank], cuts[rank + 1]
    lrp, lcol, gval, loc2glob, ghosts = par.row_block_local_system(rowptr, col, val, perm, cuts, rank)
    nloc, nloc_nnz = len(loc2glob), len(lcol)
    all_ghosts = [None] * world
    dist.all_gather_object(all_ghosts, ghosts)
    dp_first, dp = par.row_block_dist_procs(loc2glob, cuts, rank, all_ghosts)
    pd = par.ParallelDofs.from_dist_procs(dp_first, dp, world, rank)
    # ---- local objects
    free_glob = np.unpackbits(bits, bitorder="little")[:n].astype(bool)
    free_loc = free_glob[perm[loc2glob]]
    f_loc = np.where((loc2glob >= lo) & (loc2glob < hi), fglob[perm[loc2glob]], 0.0)         # DISTRIBUTED: every dof's load once
    A = la.DevSparseMatrix(la.SparseMatrix(lrp, lcol, gval, nloc, nloc, ctx=ctx), ctx=ctx)
    comm = par.Communicator(ctx, world, rank, dist)
    pmat = par.ParallelMatrix(A, pd, comm)
    jac = pmat.CreateSmoother(la.BitArray(free_loc))
    f = la.BaseVector(f_loc, ctx=ctx)
    u = f.CreateVector()
    ctx.sync()
    setup_s = time.perf_counter() - t0

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.ExternalStream(ctx.stream)
    pmat.cg_solve(jac, f, u, precision=0.0, maxsteps=3)
    pmat.cg_solve(jac, f, u, precision=0.0, maxsteps=args.iters)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    res = pmat.cg_solve(jac, f, u, precision=0.0, maxsteps=args.iters)
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    its = res.GetSteps() - 1
    barrier()
    t1 = time.perf_counter()
    full = pmat.cg_solve(jac, f, u, precision=1e-8, maxsteps=20000)
    barrier()
    full_s = time.perf_counter() - t1
    # the solution against the reference's CPU solve, when tools/netgen_system.py --cpu-full stored it
    rel = None
    ur = os.path.join(args.cache, "u_ref.npy")
    if os.path.exists(ur):
        uref = np.load(ur)[perm[loc2glob]]
        rel = float(np.max(np.abs(u.NumPy() - uref)) / max(1e-300, np.max(np.abs(uref))))
    sizes = [None] * world
    dist.all_gather_object(sizes, dict(rank=rank, owned=hi - lo, ghosts=len(ghosts), nnz=nloc_nnz, neighbours=len(pd.GetDistantProcs()), rel=rel))
    if rank == 0:
        peak = 6545.3
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass
        b_cg = max(s["nnz"] for s in sizes) * 12 + max(s["owned"] + s["ghosts"] for s in sizes) * (20 + 88)
        line = {"system": {k: meta[k] for k in ("maxh", "nref", "order", "ne", "ndof", "nnz", "sha256_rowptr", "sha256_col")},
                "n_gpus": world, "partition": "contiguous blocks of the Cuthill-McKee order, owned rows complete + ghost dofs (SURVEY 8e)",
                "cg_it_per_s": its / (float(ms.item()) * 1e-3), "ms_per_iteration": float(ms.item()) / its,
                "cg_frac_of_peak_slowest_rank": b_cg / (float(ms.item()) / its * 1e-3) / 1e9 / peak,
                "full_solve_steps": full.GetSteps(), "full_solve_s": full_s, "rcm_s": t_rcm, "setup_s": setup_s, "ranks": sizes,
                "options": args.opt, "peer_memory": pmat.peer_memory}
        s = json.dumps(line)
        if args.out:
            open(args.out, "w").write(s + "\n")
        print(s)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
