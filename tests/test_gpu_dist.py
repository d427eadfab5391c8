"""GPU, 2 ranks: ParallelMatrix / Cumulate / distributed Jacobi-PCG and GMRES through the C ABI against
the single-GPU solve of the same global system.

Two ways to get two ranks:
  * "same": both processes use cuda:0 and bootstrap over gloo with the all-gather callback -- there is no
    NCCL (it refuses two ranks on one device), so this exercises exactly the peer-memory data path
    (CUDA IPC mailboxes, P2P stores, sequence flags); runs on the single-GPU box.
  * "two": one GPU per rank, NCCL bootstrap; peer memory (default) and the NCCL data path are both run
    (skipped on a single-GPU box)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G_DEFAULT = (12, 10, 16)
REAL, COMPLEX, BLOCK3 = 0, 1, 3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _box_kw(cfg):
    return dict(order=cfg["order"], kind=cfg["kind"], lame=(1.0, 0.6), mass=cfg.get("mass", 0.0))


def _worker(rank, world, port, cfg, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    same = cfg["devices"] == "same"
    dev = 0 if same else rank
    torch.cuda.set_device(dev)
    if same:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
    import ngsolve_b200.la as la
    from ngsolve_b200 import workloads as W, parallel as par
    ctx = la.Context(dev)
    if cfg.get("overlap"):
        ctx.set_option("dist_overlap", 1)
    if cfg.get("fold"):
        ctx.set_option("cg_fold_u", 1)
    if "fused_push" in cfg:
        ctx.set_option("dist_fused_push", cfg["fused_push"])
    if "gmres_orth" in cfg:
        ctx.set_option("gmres_orth", cfg["gmres_orth"])
    G = tuple(cfg.get("G", G_DEFAULT))
    boxes = [W.FemBox(n, offset=o, global_n=G, **_box_kw(cfg)) for n, o in W.slab_partition(G, world)]
    box = boxes[rank]
    A, f = box.device_system(ctx)
    comm = par.Communicator(ctx, world, rank, dist, bootstrap="allgather" if same else "nccl", p2p=cfg["p2p"])
    pd = par.ParallelDofs(*W.exchange_tables(boxes, rank), ndof=box.ndof, nranks=world, rank=rank)
    pmat = par.ParallelMatrix(A, pd, comm)
    assert pmat.peer_memory == (cfg["p2p"] != 0), "data path: peer_memory=%s, wanted p2p=%d" % (pmat.peer_memory, cfg["p2p"])
    assert np.array_equal(pmat.MasterDofs(), pd.MasterDofs())
    if cfg.get("overlap"):
        on, nb, ni = pmat.overlap
        # tiny slabs may have no interior slice at all: that rank then keeps the single launch (ranks may mix both forms)
        assert on or cfg["overlap"] == "maybe"
        assert not on or (nb > 0 and ni > 0 and nb + ni == (box.ndof + 31) // 32), (on, nb, ni, box.ndof)
    else:
        assert not pmat.overlap[0]
    jac = pmat.CreateSmoother(box.freedofs())
    u = f.CreateVector()
    if cfg["solver"] == "cg":
        res = pmat.cg_solve(jac, f, u, precision=1e-9, maxsteps=3000, conjugate=cfg.get("conjugate", False))
        res2 = pmat.cg_solve(jac, f, u, precision=1e-9, maxsteps=3000, conjugate=cfg.get("conjugate", False))   # cached graph, same answer
        assert res2.GetSteps() == res.GetSteps() and np.array_equal(res2.history, res.history)
    else:
        res = pmat.gmres_solve(jac, f, u, precision=1e-8, maxsteps=cfg["maxsteps"])
    # Cumulate and the two flavours of the parallel inner product
    cplx = cfg["kind"] == COMPLEX
    one = np.ones(box.ndof * box.entrysize) if not cplx else np.full(box.ndof, 1.0 + 2.0j)
    ones = la.BaseVector(one, entrysize=box.entrysize, ctx=ctx)
    cnt = la.BaseVector(one, entrysize=box.entrysize, ctx=ctx)
    pmat.Cumulate(cnt)                                           # = number of sharers per dof (times the entry)
    n_cum = pmat.InnerProduct(ones, ones, both_cumulated=True)   # masked: counts every global dof once
    n_mix = pmat.InnerProduct(cnt, ones, both_cumulated=False)   # (cumulated, "distributed"): sum over all local copies
    out[rank] = (res.GetSteps(), box.dof_info()[0].astype(np.int64), u.NumPy().reshape(-1), res.history, cnt.NumPy().reshape(-1),
                 n_cum, n_mix, box.global_ndof * box.entrysize)
    dist.barrier()
    dist.destroy_process_group()


def _run(cfg, world=2):
    import torch.multiprocessing as mp
    import ngsolve_b200.la as la
    from ngsolve_b200 import workloads as W
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), cfg, out), nprocs=world, join=True)
    glob = W.FemBox(tuple(cfg.get("G", G_DEFAULT)), **_box_kw(cfg))
    A, f = glob.device_system()
    jac = A.CreateSmoother(glob.freedofs())
    if cfg["solver"] == "cg":
        inv = la.CGSolver(A, jac, precision=1e-9, maxsteps=3000, conjugate=cfg.get("conjugate", False))
    else:
        inv = la.GMRESSolver(A, jac, precision=1e-8, maxsteps=cfg["maxsteps"])
    u = (inv * f).Evaluate().NumPy().reshape(-1)
    es = glob.entrysize
    cplx = cfg["kind"] == COMPLEX
    cplx = cfg["kind"] == COMPLEX
    # every interface dof of the slab partition has exactly two copies; each rank sees its own interfaces
    total_shared = sum(int(((out[r][4] / (1.0 + 2.0j) if cplx else out[r][4]).real == 2.0).sum()) for r in range(world)) // 2
    for r in range(world):
        steps, gi, ur, hist, cnt, n_cum, n_mix, n_glob = out[r]
        assert abs(steps - inv.GetSteps()) <= 2, (steps, inv.GetSteps())
        k = min(len(hist), len(inv.history), 25)
        assert np.allclose(hist[:k], inv.history[:k], rtol=1e-8)
        full = np.repeat(gi * es, es) + np.tile(np.arange(es), len(gi))
        assert np.max(np.abs(ur - u[full])) <= 1e-7 * np.max(np.abs(u))     # CUMULATED solution = global solution restricted
        sharers = cnt / (1.0 + 2.0j) if cplx else cnt
        assert set(np.unique(sharers.real)) <= {1.0, 2.0} and (sharers.real == 2.0).any() and np.all(sharers.imag == 0)
        if cplx:      # <1+2i, conj(1+2i)> = 5 per dof
            assert n_cum == 5.0 * n_glob
        else:
            assert n_cum == n_glob
        # sum over ranks of local <cnt, 1> counts a shared dof 2 (copies) x 2 (value) times: n_glob + 3 * (#interface dofs of the whole partition)
        expect = (n_glob + 3 * total_shared) * (5.0 if cplx else 1.0)
        assert abs(n_mix - expect) <= 1e-9 * abs(expect)


CASES = [
    dict(order=3, kind=REAL, solver="cg"),
    dict(order=2, kind=BLOCK3, solver="cg"),
    dict(order=2, kind=COMPLEX, mass=1.0 + 0.5j, solver="cg"),
    dict(order=2, kind=COMPLEX, mass=2.0, solver="cg", conjugate=True),
    dict(order=2, kind=COMPLEX, mass=-30.0 - 8.0j, solver="gmres", maxsteps=300),
    dict(order=2, kind=REAL, solver="gmres", maxsteps=200),
]
IDS = ["cg-real-p3", "cg-block3-p2", "cg-complex-bilinear", "cg-complex-conjugate", "gmres-helmholtz-complex", "gmres-real"]


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_two_ranks_on_one_gpu_peer_memory(case):
    """peer-memory data path without NCCL: both ranks on cuda:0 (the two contexts time-slice the GPU, every wait costs a
    context switch -> keep the GMRES systems small: step j has j+2 reductions)"""
    extra = dict(G=(6, 5, 8)) if case["solver"] == "gmres" else {}
    _run(dict(case, devices="same", p2p=1, **extra))


@pytest.mark.parametrize("case", CASES[4:], ids=IDS[4:])
def test_two_ranks_gmres_serial_orthogonalisation(case):
    """option gmres_orth = 0: the reference's loop, j+2 single reductions per step (the default batches them into one exchange)"""
    _run(dict(case, devices="same", p2p=1, G=(6, 5, 8), gmres_orth=0))


def test_three_ranks_gmres_batched_reduction():
    """the vector all-reduce of the batched orthogonalisation with a middle rank that has two neighbours"""
    _run(dict(CASES[4], devices="same", p2p=1, G=(5, 4, 9)), world=3)


@pytest.mark.parametrize("case", CASES[:3], ids=IDS[:3])
def test_two_ranks_interface_first_overlap(case):
    """option dist_overlap: interface slices, push, interior slices, unpack -- same solve as the one-GPU solver"""
    _run(dict(case, devices="same", p2p=1, overlap=True))


def test_four_ranks_interface_first_overlap():
    """... together with cg_fold_u (u += al s in the direction kernel)"""
    _run(dict(order=2, kind=REAL, solver="cg", devices="same", p2p=1, G=(4, 4, 12), overlap="maybe", fold=True), world=4)


def test_two_ranks_cg_fold_u():
    _run(dict(CASES[1], devices="same", p2p=1, fold=True))


@pytest.mark.parametrize("case", CASES[:3], ids=IDS[:3])
def test_two_ranks_separate_push_kernel(case):
    """option dist_fused_push = 0: product, then halo_push_kernel, then unpack, then a finalize launch per reduction (the round-1
    sequence); the default since round 2 pushes the interface rows from inside the product kernel and all-reduces <d,w> in the
    update kernel (4 launches per iteration instead of 6) -- both must reproduce the one-GPU solve"""
    _run(dict(case, devices="same", p2p=1, fused_push=0))


def test_four_ranks_separate_push_kernel():
    _run(dict(order=2, kind=REAL, solver="cg", devices="same", p2p=1, G=(4, 4, 8), fused_push=0), world=4)


def test_four_ranks_on_one_gpu_peer_memory():
    """4 slabs: the middle ranks have two neighbours each, the scalar all-reduce has 4 contributions (the layout of the
    4- and 8-GPU bench runs), still without NCCL"""
    _run(dict(order=2, kind=REAL, solver="cg", devices="same", p2p=1, G=(4, 4, 8)), world=4)


@pytest.mark.parametrize("world", [4, 8])
def test_many_gpu_solve_equals_one_gpu_solve(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    _run(dict(order=3, kind=REAL, solver="cg", devices="two", p2p=1, G=(8, 8, 16)), world=world)


@pytest.mark.parametrize("p2p", [1, 0], ids=["peer-memory", "nccl"])
@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_two_gpu_solve_equals_one_gpu_solve(case, p2p):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run(dict(case, devices="two", p2p=p2p))
