"""Synthetic finite-element systems (include/ngsb200_workloads.h): H1 order-p elements on a
Kuhn-triangulated box, assembled by the library on the device (or by its host loop for tests).
Benchmark / test input only -- the reference gets these matrices from BilinearForm.Assemble on
netgen meshes, which are not available on the GPU box."""
import ctypes as C

import numpy as np

from . import _capi, la
from ._capi import check, REAL, COMPLEX, BLOCK3  # noqa: F401


class _Desc(C.Structure):
    _fields_ = [("order", C.c_int), ("kind", C.c_int), ("n", C.c_int * 3), ("offset", C.c_int * 3), ("global_", C.c_int * 3),
                ("h", C.c_double), ("mass_re", C.c_double), ("mass_im", C.c_double), ("lame_lambda", C.c_double),
                ("lame_mu", C.c_double)]


def _lib():
    L = _capi.lib()
    if not getattr(L, "_femgen_ready", False):
        vp = C.c_void_p
        L.ngsb_femgen_create.argtypes = [C.POINTER(_Desc), C.POINTER(vp)]
        L.ngsb_femgen_destroy.argtypes = [vp]
        L.ngsb_femgen_sizes.argtypes = [vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.ngsb_femgen_dof_info.argtypes = [vp, vp, vp, vp]
        L.ngsb_femgen_host.argtypes = [vp, vp, vp, vp, vp]
        L.ngsb_femgen_device.argtypes = [vp, vp, C.POINTER(vp), C.POINTER(vp)]
        for f in ("create", "destroy", "sizes", "dof_info", "host", "device"):
            getattr(L, "ngsb_femgen_" + f).restype = C.c_int
        L._femgen_ready = True
    return L


class FemBox:
    """One box of cubes (optionally a sub-domain of a global grid)."""

    def __init__(self, n, order=3, kind=REAL, offset=(0, 0, 0), global_n=None, h=None, mass=0.0, lame=(0.0, 1.0)):
        n = tuple(int(v) for v in (n if np.ndim(n) else (n, n, n)))
        global_n = tuple(global_n) if global_n is not None else n
        d = _Desc()
        d.order, d.kind = order, kind
        d.n[:], d.offset[:], d.global_[:] = n, tuple(offset), global_n
        d.h = h if h is not None else 1.0 / max(global_n)
        d.mass_re, d.mass_im = complex(mass).real, complex(mass).imag
        d.lame_lambda, d.lame_mu = lame
        self.desc, self.kind, self.order = d, kind, order
        h_ = C.c_void_p()
        check(_lib().ngsb_femgen_create(C.byref(d), C.byref(h_)))
        self.handle = h_
        a, b = C.c_size_t(), C.c_size_t()
        check(_lib().ngsb_femgen_sizes(h_, C.byref(a), C.byref(b)))
        self.ndof, self.global_ndof = a.value, b.value
        self.entrysize = 3 if kind == BLOCK3 else 1

    def __del__(self):
        try:
            _lib().ngsb_femgen_destroy(self.handle)
        except Exception:
            pass

    def dof_info(self):
        """(global dof index, on-box-surface flag, free-dof flag) per local dof"""
        gi = np.empty(self.ndof, dtype=np.uint64)
        surf = np.empty(self.ndof, dtype=np.uint8)
        free = np.empty(self.ndof, dtype=np.uint8)
        check(_lib().ngsb_femgen_dof_info(self.handle, la._np_ptr(gi), la._np_ptr(surf), la._np_ptr(free)))
        return gi, surf, free

    def freedofs(self):
        return la.BitArray(self.dof_info()[2].astype(bool))

    def host_csr(self):
        """(rowptr, col, val, rhs) assembled by the library's host loop (tests / small systems)"""
        rowptr = np.empty(self.ndof + 1, dtype=np.uint64)
        check(_lib().ngsb_femgen_host(self.handle, la._np_ptr(rowptr), None, None, None))
        nnz = int(rowptr[-1])
        nv = {REAL: 1, COMPLEX: 1, BLOCK3: 9}[self.kind]
        dt = np.complex128 if self.kind == COMPLEX else np.float64
        col = np.empty(nnz, dtype=np.int32)
        val = np.empty(nnz * nv, dtype=dt)
        rhs = np.empty(self.ndof * self.entrysize, dtype=dt)
        check(_lib().ngsb_femgen_host(self.handle, la._np_ptr(rowptr), la._np_ptr(col), la._np_ptr(val), la._np_ptr(rhs)))
        return rowptr, col, val, rhs

    def device_system(self, ctx=None):
        """(DevSparseMatrix, rhs BaseVector) assembled on the device"""
        ctx = ctx or la.default_context()
        A, f = C.c_void_p(), C.c_void_p()
        check(_lib().ngsb_femgen_device(ctx.handle, self.handle, C.byref(A), C.byref(f)))
        return la.DevSparseMatrix(None, ctx=ctx, _handle=A), la.BaseVector(None, ctx=ctx, _handle=f)


def slab_partition(global_n, nranks):
    """Split the global grid into `nranks` slabs along z (sub-domains of whole cubes, like a mesh
    partition).  Returns [(n, offset)] per rank."""
    gx, gy, gz = global_n
    cuts = [(gz * r) // nranks for r in range(nranks + 1)]
    return [((gx, gy, cuts[r + 1] - cuts[r]), (0, 0, cuts[r])) for r in range(nranks)]


def exchange_tables(boxes, rank):
    """ParallelDofs::exchangedofs of `rank` (linalg/paralleldofs.cpp:46-59) for a list of FemBox
    sub-domains of one global grid: for every other rank the local dofs shared with it, ascending.
    Returns (ex_first[nranks+1] uint64, ex_dofs int32); pairing by position is checked."""
    me = boxes[rank]
    gi, surf, _ = me.dof_info()
    cand = np.flatnonzero(surf)
    keys = gi[cand]
    order = np.argsort(keys, kind="stable")
    skeys, scand = keys[order], cand[order]
    first = [0]
    dofs = []
    for q, other in enumerate(boxes):
        if q == rank:
            first.append(first[-1])
            continue
        ogi, osurf, _ = other.dof_info()
        ocand = np.flatnonzero(osurf)
        okeys = ogi[ocand]
        pos = np.searchsorted(skeys, okeys)
        pos[pos >= len(skeys)] = 0
        hit = skeys[pos] == okeys if len(skeys) else np.zeros(len(okeys), dtype=bool)
        mine = np.sort(scand[pos[hit]])
        theirs = np.sort(ocand[hit])
        # both sides order their list by local dof number; the two orders must pair the same dofs
        assert np.array_equal(gi[mine], ogi[theirs]), "exchange lists of ranks %d/%d do not pair up" % (rank, q)
        dofs.append(mine.astype(np.int32))
        first.append(first[-1] + len(mine))
    return np.array(first, dtype=np.uint64), (np.concatenate(dofs) if dofs else np.zeros(0, dtype=np.int32)).astype(np.int32)
