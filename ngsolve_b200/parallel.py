"""One GPU per process: ParallelDofs / ParallelMatrix / Cumulate and the distributed Jacobi-PCG
(linalg/paralleldofs.cpp, parallel/parallel_matrices.cpp, parallel/parallelvvector.cpp), bound to
the C ABI.  torch.distributed is only the plumbing that hands the ncclUniqueId to the ranks."""
import ctypes as C
import weakref

import numpy as np

from . import _capi, la
from ._capi import check


class Communicator:
    """NgMPI_Comm stand-in: an NCCL communicator on the context's stream."""

    def __init__(self, ctx, nranks, rank, dist=None):
        self.ctx, self.nranks, self.rank = ctx, nranks, rank
        uid = (C.c_ubyte * 128)()
        if nranks > 1:
            import torch
            if rank == 0:
                check(_capi.lib().ngsb_comm_unique_id(uid))
            t = torch.tensor(list(uid), dtype=torch.uint8)
            if dist.get_backend() == "nccl":
                t = t.cuda()
            dist.broadcast(t, 0)
            uid = (C.c_ubyte * 128)(*t.cpu().tolist())
        h = C.c_void_p()
        check(_capi.lib().ngsb_comm_create(ctx.handle, nranks, rank, uid, C.byref(h)))
        self.handle = h
        self._fin = weakref.finalize(self, _capi.lib().ngsb_comm_destroy, h)


class ParallelDofs:
    """exchangedofs as a table over ranks (ascending local dofs per neighbour), ismasterdof derived
    by the lowest-rank rule (linalg/paralleldofs.cpp:46-66)."""

    def __init__(self, ex_first, ex_dofs, ndof, nranks, rank):
        self.ex_first = np.ascontiguousarray(ex_first, dtype=np.uint64)
        self.ex_dofs = np.ascontiguousarray(ex_dofs, dtype=np.int32)
        self.ndof, self.nranks, self.rank = ndof, nranks, rank
        assert len(self.ex_first) == nranks + 1

    def GetExchangeDofs(self, proc):
        return self.ex_dofs[int(self.ex_first[proc]):int(self.ex_first[proc + 1])]

    def GetDistantProcs(self):
        return [p for p in range(self.nranks) if self.ex_first[p + 1] > self.ex_first[p]]

    def MasterDofs(self):
        m = np.ones(self.ndof, dtype=bool)
        for p in range(self.rank):
            m[self.GetExchangeDofs(p)] = False
        return m


class ParallelMatrix(la.BaseMatrix):
    """ParallelMatrix(local matrix, pardofs, C2D): cumulated in, distributed out."""

    def __init__(self, local, pardofs, comm):
        self.local, self.pardofs, self.comm, self.ctx = local, pardofs, comm, local.ctx
        self.height = self.width = local.height
        self.is_complex, self.entrysize = local.is_complex, local.entrysize
        h = C.c_void_p()
        check(_capi.lib().ngsb_parmat_create(comm.handle, local.handle, la._np_ptr(pardofs.ex_first),
                                             la._np_ptr(pardofs.ex_dofs) if len(pardofs.ex_dofs) else None, C.byref(h)))
        self.handle = h
        self._fin = weakref.finalize(self, _capi.lib().ngsb_parmat_destroy, h)

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

    def InnerProduct(self, x, y, both_cumulated):
        x._dev_read()
        y._dev_read()
        out = C.c_double()
        check(_capi.lib().ngsb_parmat_dot(self.handle, x.handle, y.handle, 1 if both_cumulated else 0, C.byref(out)))
        return out.value

    def CreateSmoother(self, freedofs=None):
        bits = la._freebits(freedofs)
        h = C.c_void_p()
        check(_capi.lib().ngsb_parmat_jacobi_create(self.handle, la._np_ptr(bits) if bits is not None else None, C.byref(h)))
        return _ParJacobi(self, h)

    def cg_solve(self, jac, f, u, precision=1e-8, maxsteps=200):
        """CGSolver::Mult on parallel vectors: f DISTRIBUTED in, u CUMULATED out"""
        f._dev_read()
        u._dev_write()
        cap = min(maxsteps, 1 << 20) + 2
        hist = np.zeros(cap)
        steps, nh = C.c_int(), C.c_int()
        check(_capi.lib().ngsb_parmat_cg_solve(self.handle, jac.handle if jac is not None else None, f.handle, u.handle, precision,
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
