"""GPU, 3 and 4 ranks sharing cuda:0 (peer-memory data path over CUDA IPC, gloo bootstrap): the distributed index maps and the
neighbour exchange of the library against the REFERENCE's ParallelDofs run on in-process ranks
(tests/golden/pardofs_reference.npz <- oracle/ref_pardofs/harness.cpp + linalg/paralleldofs.cpp).

  * ngsb_parmat_masterdofs                 == ParallelDofs::ismasterdof              (bit-exact, paralleldofs.cpp:61-66)
  * ngsb_parmat_cumulate(data)             == AllReduceDofData(data, SUM)            (bit-exact: copies summed in ascending rank
                                              order on every sharer = the master's order, paralleldofs.hpp:213-334)
  * ngsb_parmat_jacobi_create (diag = data) == 1 / AllReduceDofData(diag)            (bit-exact, linalg/jacobi.cpp:60-61)
  * master-masked inner product            == sum over the global dofs counted once  (parallelvvector.cpp:305-314)
"""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ngsolve_b200.la as la
    from ngsolve_b200 import parallel as par
    g = load_golden("pardofs_reference")
    pre = "%s_r%d_" % (name, rank)
    data = g[pre + "data"]
    n = len(data)
    ctx = la.Context(0)
    ctx.set_option("reorder", 0)
    # local matrix: diag(data) -- the exchange tables, not the matrix, are under test
    A = la.DevSparseMatrix(la.SparseMatrix(np.arange(n + 1, dtype=np.uint64), np.arange(n, dtype=np.int32), data.copy()), ctx=ctx)
    comm = par.Communicator(ctx, world, rank, dist, bootstrap="allgather", p2p=1)
    pd = par.ParallelDofs.from_dist_procs(g[pre + "dp_first"], g[pre + "dp"], world, rank)
    pmat = par.ParallelMatrix(A, pd, comm)
    master = pmat.MasterDofs()
    v = la.BaseVector(data.copy(), ctx=ctx)
    pmat.Cumulate(v)
    jac = pmat.CreateSmoother(None)
    inv = np.empty(n)
    from ngsolve_b200 import _capi
    _capi.check(_capi.lib().ngsb_jacobi_download(jac.handle, inv.ctypes.data))
    ones = la.BaseVector(np.ones(n), ctx=ctx)
    count = pmat.InnerProduct(ones, ones, both_cumulated=True)
    out[rank] = (master.astype(np.uint8), v.NumPy().copy(), inv, count)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["grid4", "rand3"])
def test_exchange_against_reference_paralleldofs(name):
    import torch.multiprocessing as mp
    g = load_golden("pardofs_reference")
    world = int(g[name + "_nranks"])
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), name, out), nprocs=world, join=True)
    for r in range(world):
        pre = "%s_r%d_" % (name, r)
        master, cum, inv, count = out[r]
        assert np.array_equal(master, g[pre + "master"])
        assert np.array_equal(cum, g[pre + "allred"]), np.max(np.abs(cum - g[pre + "allred"]))       # bit-exact on every sharer
        assert np.array_equal(inv, 1.0 / g[pre + "allred"])
        assert count == float(g[name + "_nglobal"])
