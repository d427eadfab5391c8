"""CPU, world_size 2 over gloo: the host-side logic of the N > 1 path (slab partition, ParallelDofs
exchange tables, master dofs) driven through the reference's distributed CG choreography
(Cumulate = neighbour exchange + add, masked dots + all-reduce, summed Jacobi
diagonal), with the oracle doing the local arithmetic.  The 2-rank run must reproduce the 1-rank
oracle CG on the global system."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

G = (4, 3, 6)
PREC, MAXSTEPS = 1e-10, 400


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cumulate(v, pd, rank):
    """ParallelBaseVector::Cumulate: send my interface values, add the neighbours' (ascending rank)"""
    reqs, recv = [], {}
    for p in pd.GetDistantProcs():
        idx = pd.GetExchangeDofs(p)
        recv[p] = torch.zeros(len(idx), dtype=torch.float64)
        reqs.append(dist.isend(torch.from_numpy(v[idx].copy()), p))
        reqs.append(dist.irecv(recv[p], p))
    for r in reqs:
        r.wait()
    for p in sorted(recv):
        v[pd.GetExchangeDofs(p)] += recv[p].numpy()


def _allsum(x):
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t)
    return float(t[0])


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as orc
    from ngsolve_b200 import workloads as W
    from ngsolve_b200.parallel import ParallelDofs
    boxes = [W.FemBox(n, order=2, offset=o, global_n=G) for n, o in W.slab_partition(G, world)]
    box = boxes[rank]
    rp, col, val, f = box.host_csr()
    gi, surf, free = box.dof_info()
    pd = ParallelDofs(*W.exchange_tables(boxes, rank), ndof=box.ndof, nranks=world, rank=rank)
    master = pd.MasterDofs()
    A = orc.Csr(rp, col, val, 0)
    # Jacobi: diagonal summed over sharers, then inverted on the free dofs (linalg/jacobi.cpp:49-67)
    diag = np.array([val[int(rp[i]) + int(np.searchsorted(col[int(rp[i]):int(rp[i + 1])], i))] for i in range(box.ndof)])
    _cumulate(diag, pd, rank)
    inv = np.where(free.astype(bool), 1.0 / diag, 0.0)
    # CGSolver::Mult on parallel vectors
    u = np.zeros(box.ndof)
    d = f.copy()
    _cumulate(d, pd, rank)                       # Jacobi cumulates its input
    w = inv * d
    s = w.copy()
    wdn = _allsum(float(np.dot(w[master], d[master])))
    hist = [abs(wdn)]
    err = PREC ** 2 * abs(wdn)
    n = 0
    while True:
        cont = n < MAXSTEPS and abs(wdn) > err
        n += 1
        if not cont:
            break
        as_ = A.mult(s)                          # DISTRIBUTED
        wd = wdn
        kss = _allsum(float(np.dot(s, as_)))     # (CUMULATED, DISTRIBUTED): local dot + all-reduce
        al = wd / kss
        u += al * s
        _cumulate(as_, pd, rank)                 # d -= al*as cumulates as
        d -= al * as_
        w = inv * d
        wdn = _allsum(float(np.dot(d[master], w[master])))
        be = wdn / wd
        s = be * s + w
        hist.append(abs(wdn))
    out[rank] = (n, gi.astype(np.int64), u, np.array(hist))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_cg_equals_one_rank_cg():
    from oracle import pyoracle as orc
    from ngsolve_b200 import workloads as W
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    glob = W.FemBox(G, order=2)
    rp, col, val, f = glob.host_csr()
    A = orc.Csr(rp, col, val, 0)
    J = orc.Jacobi(A, glob.freedofs().bytes)
    u, steps, hist = orc.cg_solve(A, J, f, prec=PREC, maxsteps=MAXSTEPS)
    for r in range(2):
        n, gi, ur, hr = out[r]
        assert abs(n - steps) <= 1
        k = min(len(hr), len(hist)) - 2
        assert np.allclose(hr[:k], hist[:k], rtol=1e-8)
        assert np.max(np.abs(ur - u[gi])) <= 1e-9 * np.max(np.abs(u))   # CUMULATED solution = global solution restricted


def test_masterdofs_partition_global_dofs():
    from ngsolve_b200 import workloads as W
    from ngsolve_b200.parallel import ParallelDofs
    boxes = [W.FemBox(n, order=3, offset=o, global_n=(3, 3, 7)) for n, o in W.slab_partition((3, 3, 7), 4)]
    seen = np.zeros(boxes[0].global_ndof, dtype=int)
    for r, b in enumerate(boxes):
        pd = ParallelDofs(*W.exchange_tables(boxes, r), ndof=b.ndof, nranks=4, rank=r)
        seen[b.dof_info()[0][pd.MasterDofs()].astype(np.int64)] += 1
    assert np.all(seen == 1)          # every global dof has exactly one master (global_ndof = sum of master counts)


# ---- pinned against the REFERENCE's ParallelDofs (tests/golden/pardofs_reference.npz, produced by running
# linalg/paralleldofs.cpp on in-process ranks: tests/golden/make_golden_pardofs.py) --------------------------------------
@pytest.mark.parametrize("name", ["grid4", "rand3"])
def test_exchange_tables_equal_the_reference_paralleldofs(name):
    from conftest import load_golden
    from oracle import pyoracle as orc
    from ngsolve_b200.parallel import ParallelDofs
    g = load_golden("pardofs_reference")
    nr = int(g[name + "_nranks"])
    masters = 0
    for r in range(nr):
        pre = "%s_r%d_" % (name, r)
        first, dp = g[pre + "dp_first"], g[pre + "dp"]
        # the oracle's restatement of the constructor ...
        dist_procs = [dp[first[i]:first[i + 1]] for i in range(len(first) - 1)]
        exch, ismaster = orc.pardofs_build(nr, r, dist_procs)
        gf = g[pre + "ex_first"]
        for p in range(nr):                                                   # bit-exact halo maps
            assert np.array_equal(exch[p], g[pre + "ex_dofs"][gf[p]:gf[p + 1]])
        assert np.array_equal(ismaster.astype(np.uint8), g[pre + "master"])
        # ... and the host side of the product
        pd = ParallelDofs.from_dist_procs(first, dp, nr, r)
        assert np.array_equal(pd.ex_first.astype(np.int64), g[pre + "ex_first"]) and np.array_equal(pd.ex_dofs, g[pre + "ex_dofs"])
        assert np.array_equal(pd.MasterDofs().astype(np.uint8), g[pre + "master"])
        masters += int(g[pre + "master"].sum())
        assert int(g[pre + "global_ndof"]) == int(g[name + "_nglobal"])
    assert masters == int(g[name + "_nglobal"])        # global_ndof = all-reduced master count (paralleldofs.cpp:104-107)


def test_row_block_split_reproduces_the_global_product():
    """host logic of tools/netgen_multi.py: a globally assembled matrix split into row blocks of a permuted numbering.
    Sum over the ranks of the DISTRIBUTED local products = the global product; exchange tables pair up; masters partition."""
    from conftest import load_golden
    from oracle import pyoracle as orc
    from ngsolve_b200 import parallel as par
    g = load_golden("reorder_netgen_h1p3")
    rowptr, col, val, perm = g["rowptr"], g["col"], g["val"], g["perm"].astype(np.int64)
    n = len(rowptr) - 1
    world = 4
    cuts = [(n * r) // world for r in range(world + 1)]
    loc = [par.row_block_local_system(rowptr, col, val, perm, cuts, r) for r in range(world)]
    all_ghosts = [l[4] for l in loc]
    x = g["x"]
    y = np.zeros(n)
    masters = np.zeros(n, dtype=int)
    pds = []
    for r in range(world):
        lrp, lcol, lval, l2g, ghosts = loc[r]
        dp_first, dp = par.row_block_dist_procs(l2g, cuts, r, all_ghosts)
        pd = par.ParallelDofs.from_dist_procs(dp_first, dp, world, r)
        pds.append(pd)
        yl = orc.Csr(lrp, lcol, lval, 0).mult(x[perm[l2g]])          # x CUMULATED: the global values at the local dofs
        np.add.at(y, perm[l2g], yl)                                   # DISTRIBUTED results summed over the ranks
        masters[perm[l2g][pd.MasterDofs()]] += 1
    assert np.max(np.abs(y - g["y"])) <= 1e-12 * np.max(np.abs(g["y"]))
    assert np.all(masters == 1)
    for r in range(world):                     # both sides of every exchange list the same global dofs in the same order
        for q in range(world):
            a = loc[r][3][pds[r].GetExchangeDofs(q)]
            b = loc[q][3][pds[q].GetExchangeDofs(r)]
            assert np.array_equal(a, b)
