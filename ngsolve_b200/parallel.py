"""One GPU per process: ParallelDofs / ParallelMatrix / Cumulate and the distributed Krylov solvers
(linalg/paralleldofs.cpp, parallel/parallel_matrices.cpp, parallel/parallelvvector.cpp), bound to
the C ABI.  torch.distributed is only the bootstrap plumbing: it hands the ncclUniqueId to the ranks,
or (bootstrap="allgather") carries the CUDA IPC handles of the peer-memory data path itself."""
import ctypes as C
import weakref

import numpy as np

from . import _capi, la
from ._capi import check


def _make_allgather(dist, nranks):
    """ngsb_allgather_fn on top of torch.distributed (any backend)."""
    import torch

    def fn(_user, send, recv, nbytes):
        try:
            src = torch.frombuffer((C.c_ubyte * nbytes).from_address(send), dtype=torch.uint8).clone()
            cuda = dist.get_backend() == "nccl"
            if cuda:
                src = src.cuda()
            outs = [torch.empty_like(src) for _ in range(nranks)]
            dist.all_gather(outs, src)
            flat = torch.cat(outs).cpu().numpy().tobytes()
            C.memmove(recv, flat, nbytes * nranks)
            return 0
        except Exception:      # never raise through the C frame
            import traceback
            traceback.print_exc()
            return 1
    return _capi.ALLGATHER_FN(fn)


class Communicator:
    """NgMPI_Comm stand-in.  bootstrap="nccl": an NCCL communicator on the context's stream (also the
    fall-back data path); bootstrap="allgather": no NCCL at all, peer memory only.
    p2p: -1 auto, 0 NCCL data path, 1 peer memory required."""

    def __init__(self, ctx, nranks, rank, dist=None, bootstrap="nccl", p2p=-1):
        self.ctx, self.nranks, self.rank, self.dist = ctx, nranks, rank, dist
        self._cb = _make_allgather(dist, nranks) if (nranks > 1 and bootstrap == "allgather") else None
        uid = None
        if nranks > 1 and bootstrap == "nccl":
            import torch
            uid = (C.c_ubyte * 128)()
            if rank == 0:
                check(_capi.lib().ngsb_comm_unique_id(uid))
            t = torch.tensor(list(uid), dtype=torch.uint8)
            if dist.get_backend() == "nccl":
                t = t.cuda()
            dist.broadcast(t, 0)
            uid = (C.c_ubyte * 128)(*t.cpu().tolist())
        h = C.c_void_p()
        check(_capi.lib().ngsb_comm_create_ex(ctx.handle, nranks, rank, uid, C.cast(self._cb, C.c_void_p) if self._cb else None,
                                              None, p2p, C.byref(h)))
        self.handle = h
        self._fin = weakref.finalize(self, _capi.lib().ngsb_comm_destroy, h)

    @property
    def peer_memory(self):
        pm = C.c_int()
        check(_capi.lib().ngsb_comm_info(self.handle, None, None, C.byref(pm), None))
        return bool(pm.value)


class ParallelDofs:
    """exchangedofs as a table over ranks (ascending local dofs per neighbour), ismasterdof derived
    by the lowest-rank rule (linalg/paralleldofs.cpp:46-66)."""

    def __init__(self, ex_first, ex_dofs, ndof, nranks, rank):
        self.ex_first = np.ascontiguousarray(ex_first, dtype=np.uint64)
        self.ex_dofs = np.ascontiguousarray(ex_dofs, dtype=np.int32)
        self.ndof, self.nranks, self.rank = ndof, nranks, rank
        assert len(self.ex_first) == nranks + 1

    @classmethod
    def from_dist_procs(cls, dp_first, dp, nranks, rank):
        """the reference constructor's input: per local dof the other ranks that hold it (ParallelDofs(comm, Table<int>
        dist_procs), linalg/paralleldofs.cpp:20-59): exchangedofs[p] = the local dofs listing p, ascending"""
        dp_first = np.asarray(dp_first, dtype=np.int64)
        dp = np.asarray(dp, dtype=np.int64)
        ndof = len(dp_first) - 1
        dof_of = np.repeat(np.arange(ndof, dtype=np.int64), np.diff(dp_first))
        order = np.lexsort((dof_of, dp))                      # by rank, then by local dof
        cnt = np.bincount(dp, minlength=nranks) if len(dp) else np.zeros(nranks, dtype=np.int64)
        ex_first = np.concatenate([[0], np.cumsum(cnt)]).astype(np.uint64)
        return cls(ex_first, dof_of[order].astype(np.int32), ndof, nranks, rank)

    def GetExchangeDofs(self, proc):
        return self.ex_dofs[int(self.ex_first[proc]):int(self.ex_first[proc + 1])]

    def GetDistantProcs(self):
        return [p for p in range(self.nranks) if self.ex_first[p + 1] > self.ex_first[p]]

    def MasterDofs(self):
        m = np.ones(self.ndof, dtype=bool)
        for p in range(self.rank):
            m[self.GetExchangeDofs(p)] = False
        return m


def row_block_local_system(rowptr, col, val, perm, cuts, rank):
    """Rank `rank`'s share of a globally assembled matrix split by rows: dofs renumbered by `perm`
    (new -> old), rank r owns the new indices [cuts[r], cuts[r+1]).  Its local dofs are the owned ones plus the ghost dofs
    its rows couple to, numbered ascending in the new global index; its local matrix holds the owned rows complete and
    the ghost rows empty, so A_loc * x (x CUMULATED) is DISTRIBUTED in the reference's sense.  Index work only (host, numpy).
    Returns (lrowptr uint64, lcol int32, lval, loc2glob (new global index per local dof), ghosts)."""
    n = len(rowptr) - 1
    perm = np.asarray(perm, dtype=np.int64)
    iperm = np.empty(n, dtype=np.int64)
    iperm[perm] = np.arange(n)
    lo, hi = int(cuts[rank]), int(cuts[rank + 1])
    rp = np.asarray(rowptr).astype(np.int64)
    old_rows = perm[lo:hi]
    lens = rp[old_rows + 1] - rp[old_rows]
    nnz = int(lens.sum())
    idx = np.repeat(rp[old_rows] - np.concatenate([[0], np.cumsum(lens)[:-1]]), lens) + np.arange(nnz)
    gcol = iperm[np.asarray(col)[idx]]
    gval = np.asarray(val)[idx]
    grow = np.repeat(np.arange(lo, hi), lens)
    ghosts = np.unique(gcol[(gcol < lo) | (gcol >= hi)])
    loc2glob = np.concatenate([ghosts[ghosts < lo], np.arange(lo, hi), ghosts[ghosts >= hi]])
    lcol = np.searchsorted(loc2glob, gcol)
    lrow = np.searchsorted(loc2glob, grow)
    order = np.lexsort((lcol, lrow))
    cnt = np.bincount(lrow, minlength=len(loc2glob))
    lrp = np.concatenate([[0], np.cumsum(cnt)]).astype(np.uint64)
    return lrp, lcol[order].astype(np.int32), gval[order], loc2glob, ghosts


def row_block_dist_procs(loc2glob, cuts, rank, all_ghosts):
    """dist_procs of the row-block split: for every local dof the OTHER ranks that hold it -- the ranks that ghost it and,
    for a ghost, its owner.  all_ghosts[q] = sorted ghost list (new global indices) of rank q.  Returns (dp_first, dp)."""
    world = len(cuts) - 1
    nloc = len(loc2glob)
    owner = np.searchsorted(np.asarray(cuts[1:]), loc2glob, side="right")
    pi, pp = [np.flatnonzero(owner != rank)], [owner[owner != rank]]
    for q in range(world):
        if q == rank:
            continue
        _, mine, _ = np.intersect1d(loc2glob, all_ghosts[q], assume_unique=True, return_indices=True)
        pi.append(mine)
        pp.append(np.full(len(mine), q))
    pi, pp = np.concatenate(pi).astype(np.int64), np.concatenate(pp).astype(np.int64)
    key = np.unique(pi * world + pp)               # by local dof, then rank; a ghost's owner may also appear as "ghosting" rank: once
    pi, pp = key // world, key % world
    dp_first = np.concatenate([[0], np.cumsum(np.bincount(pi, minlength=nloc))]).astype(np.int64)
    return dp_first, pp


class ParallelMatrix(la.BaseMatrix):
    """ParallelMatrix(local matrix, pardofs, C2D): cumulated in, distributed out."""

    def __init__(self, local, pardofs, comm):
        self.local, self.pardofs, self.comm, self.ctx = local, pardofs, comm, local.ctx
        self.height = self.width = local.height
        self.is_complex, self.entrysize = local.is_complex, local.entrysize
        h = C.c_void_p()
        check(_capi.lib().ngsb_parmat_create_ex(comm.handle, local.handle, la._np_ptr(pardofs.ex_first),
                                                la._np_ptr(pardofs.ex_dofs) if len(pardofs.ex_dofs) else None,
                                                C.cast(comm._cb, C.c_void_p) if comm._cb else None, None, C.byref(h)))
        self.handle = h
        self._fin = weakref.finalize(self, _capi.lib().ngsb_parmat_destroy, h)

    @property
    def peer_memory(self):
        pm = C.c_int()
        check(_capi.lib().ngsb_parmat_info(self.handle, C.byref(pm), None, None, None))
        return bool(pm.value)

    @property
    def overlap(self):
        """(active, interface slices, interior slices) of the interface-first split (Context option dist_overlap=1)"""
        on, nb, ni = C.c_int(), C.c_size_t(), C.c_size_t()
        check(_capi.lib().ngsb_parmat_overlap_info(self.handle, C.byref(on), C.byref(nb), C.byref(ni)))
        return bool(on.value), nb.value, ni.value

    def MasterDofs(self):
        out = np.empty(self.height, dtype=np.uint8)
        check(_capi.lib().ngsb_parmat_masterdofs(self.handle, la._np_ptr(out)))
        return out.astype(bool)

    def Mult(self, x, y):
        x._dev_read()
        y._dev_write()
        check(_capi.lib().ngsb_parmat_mult(self.handle, x.handle, y.handle))

    def Cumulate(self, v):
        v._dev_write()
        check(_capi.lib().ngsb_parmat_cumulate(self.handle, v.handle))

    def InnerProduct(self, x, y, both_cumulated, conjugate=True):
        x._dev_read()
        y._dev_read()
        out = (C.c_double * 2)()
        check(_capi.lib().ngsb_parmat_dot(self.handle, x.handle, y.handle, 1 if both_cumulated else 0, 1 if conjugate else 0, out))
        return complex(out[0], out[1]) if self.is_complex else out[0]

    def CreateSmoother(self, freedofs=None):
        bits = la._freebits(freedofs)
        h = C.c_void_p()
        check(_capi.lib().ngsb_parmat_jacobi_create(self.handle, la._np_ptr(bits) if bits is not None else None, C.byref(h)))
        return _ParJacobi(self, h)

    def cg_solve(self, jac, f, u, precision=1e-8, maxsteps=200, conjugate=False):
        """CGSolver::Mult on parallel vectors: f DISTRIBUTED in, u CUMULATED out"""
        f._dev_read()
        u._dev_write()
        cap = min(maxsteps, 1 << 20) + 2
        hist = np.zeros(cap)
        steps, nh = C.c_int(), C.c_int()
        ip = _capi.IP_REAL if not self.is_complex else (_capi.IP_COMPLEX_CONJ if conjugate else _capi.IP_COMPLEX)
        check(_capi.lib().ngsb_parmat_cg_solve(self.handle, jac.handle if jac is not None else None, f.handle, u.handle, precision,
                                               maxsteps, ip, C.byref(steps), la._np_ptr(hist), cap, C.byref(nh)))
        return _Result(steps.value, hist[:min(nh.value, cap)].copy())

    def gmres_solve(self, jac, f, x, precision=1e-8, maxsteps=200):
        """GMRESSolver::Mult on parallel vectors: f DISTRIBUTED in, x CUMULATED out"""
        f._dev_read()
        x._dev_write()
        cap = maxsteps + 2
        hist = np.zeros(cap)
        steps, nh = C.c_int(), C.c_int()
        check(_capi.lib().ngsb_parmat_gmres_solve(self.handle, jac.handle if jac is not None else None, f.handle, x.handle, precision,
                                                  maxsteps, C.byref(steps), la._np_ptr(hist), cap, C.byref(nh)))
        return _Result(steps.value, hist[:min(nh.value, cap)].copy())


class _ParJacobi(la.DevJacobiMatrix):
    def __init__(self, pmat, handle):
        self.ctx, self.handle = pmat.ctx, handle
        self.height = self.width = pmat.height
        self.is_complex, self.entrysize = pmat.is_complex, pmat.entrysize
        self._fin = weakref.finalize(self, _capi.lib().ngsb_jacobi_destroy, handle)


class _Result:
    def __init__(self, steps, history):
        self.steps, self.history = steps, history

    def GetSteps(self):
        return self.steps
