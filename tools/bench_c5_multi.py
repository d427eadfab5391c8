"""BASELINE config 5 (complex Helmholtz H1 order 4, ~30 M dofs, Jacobi-GMRES) on N GPUs of one box (strong scaling).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_c5_multi.py [--m 77] [--steps 40]
N = 1 works without torchrun.  Prints one JSON line on rank 0."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=77)
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--orth", type=int, nargs="+", default=[1], help="gmres_orth modes to time (0 serial MGS, 1 batched)")
a = ap.parse_args()
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
from ngsolve_b200 import la, workloads as W, parallel as par
ctx = la.Context(lr)
G = (a.m, a.m, a.m)
kw = dict(order=4, kind=W.COMPLEX, mass=-100.0 - 10.0j)
if world == 1:
    box = W.FemBox(G, **kw)
    A, f = box.device_system(ctx)
    jac = A.CreateSmoother(box.freedofs())
else:
    boxes = [W.FemBox(n, offset=o, global_n=G, **kw) for n, o in W.slab_partition(G, world)]
    box = boxes[rank]
    A, f = box.device_system(ctx)
    comm = par.Communicator(ctx, world, rank, dist)
    pd = par.ParallelDofs(*W.exchange_tables(boxes, rank), ndof=box.ndof, nranks=world, rank=rank)
    pmat = par.ParallelMatrix(A, pd, comm)
    jac = pmat.CreateSmoother(box.freedofs())
u = f.CreateVector()
stream = torch.cuda.ExternalStream(ctx.stream)


def solve():
    if world == 1:
        inv = la.GMRESSolver(A, jac, precision=0.0, maxsteps=a.steps)
        inv.Mult(f, u)
        return inv.GetSteps()
    return pmat.gmres_solve(jac, f, u, precision=0.0, maxsteps=a.steps).GetSteps()


def barrier():
    ctx.sync(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for orth in a.orth:
    ctx.set_option("gmres_orth", orth)
    solve()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    steps = solve()
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"config": "C5 Helmholtz H1 p4 complex, Jacobi-GMRES (no restart), %d^3 cubes" % a.m, "n_gpus": world, "global_dofs": box.global_ndof,
                          "rows_per_gpu": A.height, "gmres_steps": steps, "seconds": float(ms.item()) * 1e-3, "steps_per_s": steps / (float(ms.item()) * 1e-3),
                          "gmres_orth": orth, "reductions_per_step": "2" if orth else "j+2",
                          "data_path": "peer memory" if world > 1 and pmat.peer_memory else ("single GPU" if world == 1 else "nccl")}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
