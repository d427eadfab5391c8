"""ctypes front end of the CPU oracle (oracle/ngs_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py.  The product package (ngsolve_b200/) never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KIND_REAL, KIND_COMPLEX, KIND_BLOCK3 = 0, 1, 3


def build(force=False):
    so = os.path.join(_HERE, "libngs_oracle.so")
    src = os.path.join(_HERE, "ngs_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libngs_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_vec_inner_d.restype = C.c_double
        _LIB.orc_vec_l2norm.restype = C.c_double
        _LIB.orc_get_max_threads.restype = C.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _scal(kind):
    return {0: 1, 1: 2, 3: 3}[kind]


def _dtype(kind):
    return np.complex128 if kind == KIND_COMPLEX else np.float64


def _prep(kind, a):
    return np.ascontiguousarray(a, dtype=_dtype(kind))


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def max_threads():
    return lib().orc_get_max_threads()


class Csr:
    """CSR matrix as SparseMatrix::CSR() hands it out (uint64 rowptr, int32 col, values)."""

    def __init__(self, rowptr, col, val, kind):
        self.kind = kind
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.uint64)
        self.col = np.ascontiguousarray(col, dtype=np.int32)
        self.val = _prep(kind, val)
        self.n = len(self.rowptr) - 1

    def multadd(self, s, x, y):
        """y += s*A*x in place (y modified); s real or complex."""
        x = _prep(self.kind, x)
        assert y.flags.c_contiguous and y.dtype == _dtype(self.kind)
        L = lib()
        if self.kind == KIND_REAL:
            L.orc_csr_multadd_d(C.c_size_t(self.n), _p(self.rowptr), _p(self.col), _p(self.val), C.c_double(s), _p(x), _p(y))
        elif self.kind == KIND_BLOCK3:
            L.orc_csr_multadd_b3(C.c_size_t(self.n), _p(self.rowptr), _p(self.col), _p(self.val), C.c_double(s), _p(x), _p(y))
        elif isinstance(s, complex):
            L.orc_csr_multadd_zs(C.c_size_t(self.n), _p(self.rowptr), _p(self.col), _p(self.val), C.c_double(s.real), C.c_double(s.imag), _p(x), _p(y))
        else:
            L.orc_csr_multadd_z(C.c_size_t(self.n), _p(self.rowptr), _p(self.col), _p(self.val), C.c_double(s), _p(x), _p(y))
        return y

    def mult(self, x):
        x = _prep(self.kind, x)
        y = np.empty(self.n * (3 if self.kind == KIND_BLOCK3 else 1), dtype=_dtype(self.kind))
        lib().orc_csr_mult(C.c_int(self.kind), C.c_size_t(self.n), _p(self.rowptr), _p(self.col), _p(self.val), _p(x), _p(y))
        return y

    def reorder(self, perm):
        perm = np.ascontiguousarray(perm, dtype=np.uint64)
        nrp = np.empty_like(self.rowptr)
        ncol = np.empty_like(self.col)
        nval = np.empty_like(self.val)
        lib().orc_csr_reorder(C.c_int(self.kind), C.c_size_t(self.n), _p(self.rowptr), _p(self.col), _p(self.val), _p(perm),
                              _p(nrp), _p(ncol), _p(nval))
        return Csr(nrp, ncol, nval, self.kind)

    def rcm(self, max_components=64):
        """the serial Cuthill-McKee specification the device ordering (csrc/reorder.cu) is checked against"""
        perm = np.empty(self.n, dtype=np.uint64)
        lib().orc_rcm(C.c_size_t(self.n), _p(self.rowptr), _p(self.col), C.c_int(max_components), _p(perm))
        return perm


def inner(x, y, conjugate=False):
    if np.iscomplexobj(x):
        x = _prep(KIND_COMPLEX, x)
        y = _prep(KIND_COMPLEX, y)
        out = np.zeros(2)
        lib().orc_vec_inner_z(C.c_size_t(len(x)), _p(x), _p(y), C.c_int(1 if conjugate else 0), _p(out))
        return complex(out[0], out[1])
    x = _prep(KIND_REAL, x)
    y = _prep(KIND_REAL, y)
    return lib().orc_vec_inner_d(C.c_size_t(len(x)), _p(x), _p(y))


def l2norm(x):
    if np.iscomplexobj(x):
        x = _prep(KIND_COMPLEX, x)
        return lib().orc_vec_l2norm(C.c_size_t(len(x)), _p(x), C.c_int(1))
    x = _prep(KIND_REAL, x)
    return lib().orc_vec_l2norm(C.c_size_t(len(x)), _p(x), C.c_int(0))


def axpy(y, s, x):
    """y += s*x in place."""
    if np.iscomplexobj(y):
        y += complex(s) * x      # serial expression template in the reference
    else:
        lib().orc_vec_add_d(C.c_size_t(len(y)), _p(y), C.c_double(s), _p(_prep(KIND_REAL, x)))
    return y


class Jacobi:
    def __init__(self, A, freebits=None):
        self.kind = A.kind
        self.n = A.n
        self.freebits = None if freebits is None else np.ascontiguousarray(freebits, dtype=np.uint8)
        ms = {0: 1, 1: 1, 3: 9}[A.kind]
        self.invdiag = np.zeros(A.n * ms, dtype=_dtype(A.kind))
        rc = lib().orc_jacobi_setup(C.c_int(A.kind), C.c_size_t(A.n), _p(A.rowptr), _p(A.col), _p(A.val), _p(self.freebits), _p(self.invdiag))
        if rc == -1:
            raise RuntimeError("Inverse matrix: Matrix singular")

    def multadd(self, s, x, y):
        x = _prep(self.kind, x)
        lib().orc_jacobi_multadd(C.c_int(self.kind), C.c_size_t(self.n), _p(self.invdiag), _p(self.freebits), C.c_double(s), _p(x), _p(y))
        return y

    def mult(self, x):
        y = np.zeros(self.n * (3 if self.kind == KIND_BLOCK3 else 1), dtype=_dtype(self.kind))
        return self.multadd(1.0, x, y)


class BlockJacobi:
    """BlockJacobiPrecond<double>: blocks = list of dof lists"""

    def __init__(self, A, blocks):
        assert A.kind == KIND_REAL
        self.n = A.n
        self.first = np.zeros(len(blocks) + 1, dtype=np.uint64)
        self.first[1:] = np.cumsum([len(b) for b in blocks], dtype=np.uint64)
        self.dofs = np.array([d for b in blocks for d in b] or [0], dtype=np.int32)
        self.nblocks = len(blocks)
        tot = int(sum(len(b) ** 2 for b in blocks))
        self.inv = np.zeros(max(1, tot))
        L = lib()
        L.orc_blockjacobi_setup.restype = C.c_int
        rc = L.orc_blockjacobi_setup(C.c_size_t(A.n), _p(A.rowptr), _p(A.col), _p(A.val), C.c_size_t(self.nblocks), _p(self.first),
                                     _p(self.dofs), _p(self.inv))
        if rc != 0:
            raise RuntimeError("Inverse matrix: Matrix singular (block %d)" % (-1 - rc))

    def inverses(self):
        out, off = [], 0
        for b in range(self.nblocks):
            bs = int(self.first[b + 1] - self.first[b])
            out.append(self.inv[off:off + bs * bs].reshape(bs, bs).copy())
            off += bs * bs
        return out

    def multadd(self, s, x, y, transpose=False):
        x = _prep(KIND_REAL, x)
        assert y.flags.c_contiguous and y.dtype == np.float64
        lib().orc_blockjacobi_multadd(C.c_size_t(self.nblocks), _p(self.first), _p(self.dofs), _p(self.inv), C.c_double(s), _p(x), _p(y),
                                      C.c_int(1 if transpose else 0))
        return y

    def mult(self, x, transpose=False):
        return self.multadd(1.0, x, np.zeros(self.n), transpose)


def multtransadd(A, s, x, y):
    """SparseMatrix::MultTransAdd (serial scatter), y modified in place"""
    x = _prep(A.kind, x)
    assert y.flags.c_contiguous and y.dtype == _dtype(A.kind)
    lib().orc_csr_multtransadd(C.c_int(A.kind), C.c_size_t(A.n), _p(A.rowptr), _p(A.col), _p(A.val), C.c_double(s), _p(x), _p(y))
    return y


def sym_multadd(rowptr, col, val, s, x, y):
    """SparseMatrixSymmetric<double>::MultAdd on lower-triangular storage"""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.uint64)
    col = np.ascontiguousarray(col, dtype=np.int32)
    val = _prep(KIND_REAL, val)
    x = _prep(KIND_REAL, x)
    lib().orc_csrsym_multadd_d(C.c_size_t(len(rowptr) - 1), _p(rowptr), _p(col), _p(val), C.c_double(s), _p(x), _p(y))
    return y


def cg_solve(A, jac, f, prec=1e-8, maxsteps=200, ip_mode=None, initialize=True, u0=None):
    """CGSolver<IPTYPE>::Mult.  Returns (u, steps, history of Abs(wdn))."""
    kind = A.kind
    if ip_mode is None:
        ip_mode = 1 if kind == KIND_COMPLEX else 0
    f = _prep(kind, f)
    u = np.zeros_like(f) if u0 is None else _prep(kind, u0).copy()
    hist = np.zeros(maxsteps + 2)
    nh = C.c_int(0)
    L = lib()
    L.orc_cg_solve.restype = C.c_int
    steps = L.orc_cg_solve(C.c_int(kind), C.c_int(ip_mode), C.c_size_t(A.n), _p(A.rowptr), _p(A.col), _p(A.val),
                           _p(jac.invdiag) if jac is not None else None, _p(jac.freebits) if jac is not None else None,
                           _p(f), _p(u), C.c_double(prec), C.c_int(maxsteps), C.c_int(1 if initialize else 0), _p(hist), C.byref(nh))
    return u, steps, hist[:min(nh.value, maxsteps + 1)].copy()


def gmres_solve(A, jac, f, prec=1e-8, maxsteps=200, initialize=True, x0=None):
    kind = A.kind
    f = _prep(kind, f)
    x = np.zeros_like(f) if x0 is None else _prep(kind, x0).copy()
    hist = np.zeros(maxsteps + 2)
    nh = C.c_int(0)
    L = lib()
    L.orc_gmres_solve.restype = C.c_int
    steps = L.orc_gmres_solve(C.c_int(kind), C.c_size_t(A.n), _p(A.rowptr), _p(A.col), _p(A.val),
                              _p(jac.invdiag) if jac is not None else None, _p(jac.freebits) if jac is not None else None,
                              _p(f), _p(x), C.c_double(prec), C.c_int(maxsteps), C.c_int(1 if initialize else 0), _p(hist), C.byref(nh))
    return x, steps, hist[:min(nh.value, maxsteps + 1)].copy()


def pardofs_build(ntasks, rank, dist_procs):
    """ParallelDofs ctor tables.  dist_procs: list (per local dof) of lists of ranks."""
    ndof = len(dist_procs)
    first = np.zeros(ndof + 1, dtype=np.uint64)
    for i, d in enumerate(dist_procs):
        first[i + 1] = first[i] + len(d)
    data = np.array([p for d in dist_procs for p in d], dtype=np.int32)
    if len(data) == 0:
        data = np.zeros(1, dtype=np.int32)
    ex_first = np.zeros(ntasks + 1, dtype=np.uint64)
    ex_data = np.zeros(max(1, int(first[-1])), dtype=np.int32)
    ismaster = np.zeros(max(1, ndof), dtype=np.uint8)
    lib().orc_pardofs_build(C.c_int(ntasks), C.c_int(rank), C.c_size_t(ndof), _p(first), _p(data), _p(ex_first), _p(ex_data), _p(ismaster))
    exch = [ex_data[int(ex_first[p]):int(ex_first[p + 1])].copy() for p in range(ntasks)]
    return exch, ismaster[:ndof].copy()
