"""GPU, 2 ranks over NCCL (skipped on a single-GPU box): ParallelMatrix / Cumulate / distributed
Jacobi-PCG through the C ABI against the single-GPU solve of the same global system."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = (12, 10, 16)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, order, kind, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import ngsolve_b200.la as la
    from ngsolve_b200 import workloads as W, parallel as par
    ctx = la.Context(rank)
    kw = dict(order=order, kind=kind, lame=(1.0, 0.6))
    boxes = [W.FemBox(n, offset=o, global_n=G, **kw) for n, o in W.slab_partition(G, world)]
    box = boxes[rank]
    A, f = box.device_system(ctx)
    comm = par.Communicator(ctx, world, rank, dist)
    pd = par.ParallelDofs(*W.exchange_tables(boxes, rank), ndof=box.ndof, nranks=world, rank=rank)
    pmat = par.ParallelMatrix(A, pd, comm)
    assert np.array_equal(pmat.MasterDofs(), pd.MasterDofs())
    jac = pmat.CreateSmoother(box.freedofs())
    u = f.CreateVector()
    res = pmat.cg_solve(jac, f, u, precision=1e-9, maxsteps=3000)
    # Cumulate and the two flavours of the parallel inner product
    ones = la.BaseVector(np.ones(box.ndof * box.entrysize), entrysize=box.entrysize, ctx=ctx)
    cnt = la.BaseVector(np.ones(box.ndof * box.entrysize), entrysize=box.entrysize, ctx=ctx)
    pmat.Cumulate(cnt)                                           # = number of sharers per dof
    n_cum = pmat.InnerProduct(ones, ones, both_cumulated=True)   # masked: counts every global dof once
    out[rank] = (res.GetSteps(), box.dof_info()[0].astype(np.int64), u.NumPy().reshape(-1), res.history, cnt.NumPy().reshape(-1), n_cum,
                 box.global_ndof * box.entrysize)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("order,kind", [(3, 0), (2, 3)])
def test_two_gpu_cg_equals_one_gpu_cg(order, kind):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import ngsolve_b200.la as la
    from ngsolve_b200 import workloads as W
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), order, kind, out), nprocs=2, join=True)
    glob = W.FemBox(G, order=order, kind=kind, lame=(1.0, 0.6))
    A, f = glob.device_system()
    inv = la.CGSolver(A, A.CreateSmoother(glob.freedofs()), precision=1e-9, maxsteps=3000)
    u = (inv * f).Evaluate().NumPy().reshape(-1)
    es = glob.entrysize
    for r in range(2):
        steps, gi, ur, hist, cnt, n_cum, n_glob = out[r]
        assert abs(steps - inv.GetSteps()) <= 2, (steps, inv.GetSteps())
        k = min(len(hist), len(inv.history), 25)
        assert np.allclose(hist[:k], inv.history[:k], rtol=1e-8)
        full = np.repeat(gi * es, es) + np.tile(np.arange(es), len(gi))
        assert np.max(np.abs(ur - u[full])) <= 1e-7 * np.max(np.abs(u))
        assert set(np.unique(cnt)) <= {1.0, 2.0} and (cnt == 2.0).any()
        assert n_cum == n_glob
