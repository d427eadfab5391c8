"""Host-side mirror of the ngla / ngscuda Python interface for the assembled-system solve path.

Same names, argument meaning and error behaviour as the reference objects a script touches
on this path, bound to libngsb200 through its C ABI (include/ngsb200.h):

  reference (linalg/python_linalg.cpp, ngscuda/python_ngscuda.cpp)        here
  ----------------------------------------------------------------------  ------------------
  BaseVector / UnifiedVector  (.data, +=, InnerProduct, Norm, Range, FV)  BaseVector, UnifiedVector
  SparseMatrixd / ...St7complex / ...Mat<3,3>  (.CSR, CreateFromCOO)      SparseMatrix (host holder)
  BaseMatrix.CreateDeviceMatrix / ngscuda.CreateDevMatrix                 -> DevSparseMatrix
  mat.CreateSmoother(freedofs) -> JacobiPrecond -> DevDiagonalMatrix      JacobiPrecond -> DevJacobiMatrix
  CGSolver(mat, pre, precision, maxsteps) / GetSteps / DevCGSolver        CGSolver, DevCGSolver
  GMRESSolver(mat, pre, precision, maxsteps)                              GMRESSolver

There is no host execution path: a host SparseMatrix only carries the CSR arrays until
CreateDeviceMatrix() uploads them; all arithmetic runs in the CUDA library.
"""
import ctypes as C
import weakref

import numpy as np

from . import _capi
from ._capi import NgsbError, check, scal2, REAL, COMPLEX, BLOCK3

__all__ = ["NgsbError", "Context", "default_context", "BaseVector", "UnifiedVector", "UnifiedScalar", "BaseMatrix",
           "SparseMatrix", "DevSparseMatrix", "JacobiPrecond", "DevJacobiMatrix", "DiagonalMatrix", "Projector", "CGSolver",
           "DevCGSolver", "GMRESSolver", "CreateDevMatrix", "InnerProduct", "Norm", "BitArray"]


def _np_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------------------------------------
# context (InitCUDA + InitCuLinalg, ngscuda/python_ngscuda.cpp:20-25)
# ------------------------------------------------------------------------------------------------
class Context:
    def __init__(self, device=-1):
        h = C.c_void_p()
        check(_capi.lib().ngsb_ctx_create(device, C.byref(h)))
        self.handle = h
        self._fin = weakref.finalize(self, _capi.lib().ngsb_ctx_destroy, h)

    def sync(self):
        check(_capi.lib().ngsb_ctx_sync(self.handle))

    def set_option(self, name, value):
        check(_capi.lib().ngsb_ctx_set_option(self.handle, name.encode(), int(value)))

    @property
    def launches(self):
        n = C.c_uint64()
        check(_capi.lib().ngsb_ctx_launch_count(self.handle, C.byref(n)))
        return n.value

    @property
    def stream(self):
        return _capi.lib().ngsb_ctx_stream(self.handle)

    def device_info(self):
        d, s = C.c_int(), C.c_int()
        check(_capi.lib().ngsb_ctx_device(self.handle, C.byref(d), C.byref(s)))
        return d.value, s.value

    def kernel_time(self, klass="all"):
        ms, n = C.c_double(), C.c_uint64()
        check(_capi.lib().ngsb_ctx_kernel_time(self.handle, klass.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def kernel_time_reset(self):
        check(_capi.lib().ngsb_ctx_kernel_time_reset(self.handle))


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(-1)
    return _default_ctx


def _kind(is_complex, entrysize):
    if is_complex:
        if entrysize != 1:
            raise NgsbError("complex vectors with entry size %d are not supported" % entrysize)
        return COMPLEX
    if entrysize == 1:
        return REAL
    if entrysize == 3:
        return BLOCK3
    raise NgsbError("entry size %d is not supported (1 or 3)" % entrysize)


class BitArray:
    """ngcore BitArray as the path uses it: packed bytes, bit i -> byte i/8, bit i%8."""

    def __init__(self, bits):
        b = np.asarray(bits)
        if b.dtype == np.bool_:
            self.n = len(b)
            self.bytes = np.packbits(b, bitorder="little")
        else:
            self.bytes = np.ascontiguousarray(b, dtype=np.uint8)
            self.n = 8 * len(self.bytes)

    def __getitem__(self, i):
        return bool((self.bytes[i >> 3] >> (i & 7)) & 1)


def _freebits(freedofs):
    if freedofs is None:
        return None
    if isinstance(freedofs, BitArray):
        return freedofs.bytes
    return BitArray(freedofs).bytes


# ------------------------------------------------------------------------------------------------
# vector expressions (DynamicVectorExpression, linalg/basevector.hpp:1141-1373)
# ------------------------------------------------------------------------------------------------
class _Expr:
    """sum of terms  s * v   or   s * (M * v)."""

    __array_ufunc__ = None          # numpy scalars defer to __rmul__ instead of broadcasting

    def __init__(self, terms):
        self.terms = terms          # list of (scalar, matrix or None, BaseVector)

    def __add__(self, o):
        return _Expr(self.terms + _as_expr(o).terms)

    def __sub__(self, o):
        return _Expr(self.terms + (-_as_expr(o)).terms)

    def __neg__(self):
        return _Expr([(-s, m, v) for (s, m, v) in self.terms])

    def __rmul__(self, s):
        return _Expr([(s * t, m, v) for (t, m, v) in self.terms])

    __mul__ = __rmul__

    def _apply(self, scal, target, assign):
        # AssignTo / AddTo: first term may overwrite, the rest accumulate; a term that reads the
        # target while it is overwritten goes through a temporary (the reference evaluates
        # `x.data = A*x` into a temporary as well, basevector.hpp:1350)
        terms = [(scal * s, m, v) for (s, m, v) in self.terms]
        if assign and any(v is target or v._overlaps(target) for (_, _, v) in terms[1:]) or \
           (assign and terms[0][1] is not None and (terms[0][2] is target or terms[0][2]._overlaps(target))):
            tmp = target.CreateVector()
            self._apply(scal, tmp, True)
            target.Set(1.0, tmp)
            return
        first = assign
        for (s, m, v) in terms:
            if m is None:
                if first:
                    target.Set(s, v)
                else:
                    target.Add(s, v)
            else:
                if first:
                    if s == 1:
                        m.Mult(v, target)
                    else:
                        target.SetScalar(0.0)
                        m.MultAdd(s, v, target)
                else:
                    m.MultAdd(s, v, target)
            first = False

    def Evaluate(self):
        s, m, v = self.terms[0]
        out = m.CreateColVector() if m is not None else v.CreateVector()
        self._apply(1.0, out, True)
        return out

    def InnerProduct(self, other, conjugate=True):
        return self.Evaluate().InnerProduct(_as_vec(other), conjugate=conjugate)

    def Norm(self):
        return self.Evaluate().Norm()


def _as_expr(o):
    if isinstance(o, _Expr):
        return o
    if isinstance(o, BaseVector):
        return _Expr([(1.0, None, o)])
    raise TypeError("cannot use %r in a vector expression" % (o,))


def _as_vec(o):
    return o if isinstance(o, BaseVector) else _as_expr(o).Evaluate()


class _DataProxy:
    """`v.data` on the right of += / -=."""

    def __init__(self, vec):
        self.vec = vec

    def __iadd__(self, e):
        _as_expr(e)._apply(1.0, self.vec, False)
        return self

    def __isub__(self, e):
        _as_expr(e)._apply(-1.0, self.vec, False)
        return self


# ------------------------------------------------------------------------------------------------
# vectors
# ------------------------------------------------------------------------------------------------
class BaseVector:
    """Device vector with a host mirror and dirty flags: UnifiedVector,
    ngscuda/unifiedvector.hpp:8-98 (host_uptodate / dev_uptodate, UpdateHost / UpdateDevice)."""

    __array_ufunc__ = None

    def __init__(self, arg, complex=False, entrysize=1, ctx=None, _handle=None, _parent=None):
        self.ctx = ctx or default_context()
        self._host = None
        self._host_uptodate = False
        self._dev_uptodate = True
        self._parent = _parent
        self._views = []            # weak references to live Range() views (aliases of this vector's device storage)
        if _parent is not None:
            _parent._views.append(weakref.ref(self))
        if _handle is not None:
            self.handle = _handle
        else:
            init = None
            if isinstance(arg, BaseVector):
                n, complex, entrysize = arg.size, arg.is_complex, arg.entrysize
                init = arg
            elif isinstance(arg, np.ndarray):
                complex = np.iscomplexobj(arg)
                if arg.ndim == 2:
                    entrysize = arg.shape[1]
                n = arg.shape[0] if arg.ndim == 2 else len(arg) // entrysize
                init = arg
            else:
                n = int(arg)
            h = C.c_void_p()
            check(_capi.lib().ngsb_vec_create(self.ctx.handle, n, _kind(complex, entrysize), C.byref(h)))
            self.handle = h
            if isinstance(init, BaseVector):
                self.Set(1.0, init)
            elif init is not None:
                self._upload(init)
        n, k, ns = C.c_size_t(), C.c_int(), C.c_size_t()
        check(_capi.lib().ngsb_vec_info(self.handle, C.byref(n), C.byref(k), C.byref(ns)))
        self.size, self.kind, self.nscal = n.value, k.value, ns.value
        self.is_complex = self.kind == COMPLEX
        self.entrysize = 3 if self.kind == BLOCK3 else 1
        self._fin = weakref.finalize(self, _capi.lib().ngsb_vec_destroy, self.handle)

    # ---- host mirror -----------------------------------------------------------------------
    def _dtype(self):
        return np.complex128 if self.is_complex else np.float64

    def _upload(self, arr):
        n, ns = C.c_size_t(), C.c_size_t()
        check(_capi.lib().ngsb_vec_info(self.handle, C.byref(n), None, C.byref(ns)))
        a = np.ascontiguousarray(arr, dtype=self._dtype() if hasattr(self, "kind") else
                                 (np.complex128 if np.iscomplexobj(arr) else np.float64)).reshape(-1)
        if a.size * (2 if np.iscomplexobj(a) else 1) != ns.value:
            raise NgsbError("size of vector = %d scalars != size of array = %d" % (ns.value, a.size))
        check(_capi.lib().ngsb_vec_h2d(self.handle, _np_ptr(a), 0, n.value))

    # Range() views alias the parent's DEVICE storage but keep a host mirror of their own.  Parent and views stay coherent in
    # both directions (the reference's UnifiedVectorWrapper updates the device copy and invalidates the host copy of the wrapped
    # vector, ngscuda/unifiedvector.cpp:376-377): before any alias touches the device copy, pending host writes of all aliases
    # are flushed in program order (ancestors, self, views); a device write through one alias invalidates the others' mirrors.
    def _ancestors(self):
        chain, p = [], self._parent
        while p is not None:
            chain.append(p)
            p = p._parent
        return chain[::-1]                      # root first

    def _descendants(self):
        out = []
        self._views = [r for r in self._views if r() is not None]
        for r in self._views:
            v = r()
            if v is not None:
                out.append(v)
                out.extend(v._descendants())
        return out

    def __del__(self):
        # a dying view must not lose host writes that never reached the shared device storage
        try:
            if self._parent is not None and not self._dev_uptodate and self._host is not None:
                self._flush_all(True)
        except Exception:
            pass

    def _flush_own(self):
        if not self._dev_uptodate:
            self._upload(self._host)
            self._dev_uptodate = True

    def _flush_all(self, own=True):
        for a in self._ancestors():
            a._flush_own()
        if own:
            self._flush_own()
        for d in self._descendants():
            d._flush_own()

    def _invalidate_others_host(self):
        for a in self._ancestors() + self._descendants():
            a._host_uptodate = False

    def UpdateDevice(self):
        self._flush_all(True)

    def UpdateHost(self):
        if self._host is None:
            self._host = np.zeros(self.size * self.entrysize, dtype=self._dtype())
            self._host_uptodate = False
        self._flush_all(False)
        if not self._host_uptodate:
            self._flush_own()
            check(_capi.lib().ngsb_vec_d2h(self.handle, _np_ptr(self._host), 0, self.size))
            self._host_uptodate = True

    def _dev_write(self):
        """about to be written on the device"""
        self._flush_all(True)
        self._host_uptodate = False
        self._invalidate_others_host()

    def _dev_read(self):
        self.UpdateDevice()

    def _overlaps(self, other):
        a = _capi.lib().ngsb_vec_devptr(self.handle) or 0
        b = _capi.lib().ngsb_vec_devptr(other.handle) or 0
        return a < b + 8 * other.nscal and b < a + 8 * self.nscal and self.nscal > 0 and other.nscal > 0

    class _FV:
        def __init__(self, vec):
            self.vec = vec

        def NumPy(self):
            v = self.vec
            v.UpdateHost()
            v._invalidate_others_host()  # ... and the aliases' mirrors go stale with it
            v._dev_uptodate = False      # non-const FVDouble(): the host copy may be written
            return v._host.reshape(-1, 3) if v.entrysize == 3 else v._host

        def __len__(self):
            return self.vec.size

    def FV(self):
        return BaseVector._FV(self)

    def NumPy(self):
        """read-only host copy (does not invalidate the device data)"""
        self.UpdateHost()
        out = self._host.copy()
        return out.reshape(-1, 3) if self.entrysize == 3 else out

    # ---- BaseVector interface -------------------------------------------------------------------
    def __len__(self):
        return self.size

    def CreateVector(self):
        return BaseVector(self.size, self.is_complex, self.entrysize, ctx=self.ctx)

    def CreateDeviceVector(self, unified=True, copy=True):
        v = self.CreateVector()
        if copy:
            v.Set(1.0, self)
        return v

    def Range(self, begin, end):
        self.UpdateDevice()
        h = C.c_void_p()
        check(_capi.lib().ngsb_vec_range(self.handle, begin, end, C.byref(h)))
        return BaseVector(None, ctx=self.ctx, _handle=h, _parent=self)

    def SetScalar(self, s):
        self._dev_uptodate = True            # whatever the host mirror held is overwritten as a whole
        self._dev_write()
        check(_capi.lib().ngsb_vec_set_scalar(self.handle, scal2(s)))
        return self

    def Scale(self, s):
        self._dev_write()
        if isinstance(s, UnifiedScalar):
            check(_capi.lib().ngsb_vec_scale_dev(self.handle, s.handle))
        else:
            check(_capi.lib().ngsb_vec_scale(self.handle, scal2(s)))
        return self

    def Set(self, s, x):
        x._dev_read()
        self._dev_write()
        check(_capi.lib().ngsb_vec_set(self.handle, scal2(s), x.handle))
        return self

    def Add(self, s, x):
        x._dev_read()
        self._dev_write()
        if isinstance(s, UnifiedScalar):          # BaseVector::Add(BaseScalar&, v), linalg/basevector.cpp:292-298
            check(_capi.lib().ngsb_vec_axpy_dev(self.handle, s.handle, x.handle))
        else:
            check(_capi.lib().ngsb_vec_axpy(self.handle, scal2(s), x.handle))
        return self

    def CreateScalar(self):
        return UnifiedScalar(ctx=self.ctx)

    def InnerProduct(self, other, scal=None, conjugate=True):
        if isinstance(scal, bool):                # InnerProduct(other, conjugate) positional form
            scal, conjugate = None, scal
        other = _as_vec(other)
        self._dev_read()
        other._dev_read()
        if scal is not None:                      # InnerProduct(v2, BaseScalar&): result stays on the device
            check(_capi.lib().ngsb_vec_dot_dev(self.handle, other.handle, 1 if (conjugate and self.is_complex) else 0, scal.handle))
            return scal
        out = (C.c_double * 2)()
        check(_capi.lib().ngsb_vec_dot(self.handle, other.handle, 1 if (conjugate and self.is_complex) else 0, out))
        return complex(out[0], out[1]) if self.is_complex else out[0]

    def Norm(self):
        self._dev_read()
        out = C.c_double()
        check(_capi.lib().ngsb_vec_nrm2(self.handle, C.byref(out)))
        return out.value

    def SetRandom(self, seed=0):
        rng = np.random.default_rng(seed)
        a = rng.random(self.size * self.entrysize)
        if self.is_complex:
            a = a + 1j * rng.random(self.size * self.entrysize)
        self.FV().NumPy().reshape(-1)[:] = a

    # ---- python operators ------------------------------------------------------------------------
    @property
    def data(self):
        return _DataProxy(self)

    @data.setter
    def data(self, e):
        if isinstance(e, _DataProxy):
            return                      # result of `v.data += ...`
        _as_expr(e)._apply(1.0, self, True)

    def __setitem__(self, key, value):
        if isinstance(key, slice) and key == slice(None, None, None) and np.isscalar(value):
            self.SetScalar(value)
        else:
            self.FV().NumPy()[key] = value

    def __getitem__(self, key):
        self.UpdateHost()
        return (self._host.reshape(-1, 3) if self.entrysize == 3 else self._host)[key]

    def __iadd__(self, e):
        _as_expr(e)._apply(1.0, self, False)
        return self

    def __isub__(self, e):
        _as_expr(e)._apply(-1.0, self, False)
        return self

    def __imul__(self, s):
        return self.Scale(s)

    def __add__(self, o):
        return _as_expr(self) + o

    def __sub__(self, o):
        return _as_expr(self) - o

    def __neg__(self):
        return -_as_expr(self)

    def __rmul__(self, s):
        return s * _as_expr(self)

    def Evaluate(self):
        return self


class UnifiedScalar:
    """Device-resident scalar: ngscuda UnifiedScalar / BaseScalar (ngscuda/unifiedvector.hpp:107-136,
    linalg/basescalar.hpp:17-29).  Lets a solver keep alpha = rz/pq etc. on the device:
    `x.InnerProduct(y, scal)`, `y.Add(scal, x)`, `x.Scale(scal)`, `a.Div(b, c)`, `a.Neg(b)`."""

    def __init__(self, value=0.0, ctx=None):
        self.ctx = ctx or default_context()
        h = C.c_void_p()
        check(_capi.lib().ngsb_scalar_create(self.ctx.handle, C.byref(h)))
        self.handle = h
        self._fin = weakref.finalize(self, _capi.lib().ngsb_scalar_destroy, h)
        self.Set(value)

    def Set(self, v):
        check(_capi.lib().ngsb_scalar_set(self.handle, scal2(v)))

    def Get(self):
        out = (C.c_double * 2)()
        check(_capi.lib().ngsb_scalar_get(self.handle, out))
        return complex(out[0], out[1]) if out[1] != 0.0 else out[0]

    GetD = Get

    def Div(self, a, b):          # self = a / b
        check(_capi.lib().ngsb_scalar_div(self.handle, a.handle, b.handle))
        return self

    def Neg(self, a):             # self = -a
        check(_capi.lib().ngsb_scalar_neg(self.handle, a.handle))
        return self

    def Copy(self, a):
        check(_capi.lib().ngsb_scalar_copy(self.handle, a.handle))
        return self


class UnifiedVector(BaseVector):
    """ngscuda.UnifiedVector(n | BaseVector | numpy array), ngscuda/python_ngscuda.cpp:29-50"""


def InnerProduct(x, y, conjugate=True):
    return _as_vec(x).InnerProduct(y, conjugate=conjugate)


def Norm(x):
    return _as_vec(x).Norm()


# ------------------------------------------------------------------------------------------------
# matrices
# ------------------------------------------------------------------------------------------------
class BaseMatrix:
    """Operator interface (linalg/basematrix.hpp:116-133): Mult, MultAdd, CreateRow/ColVector.
    Python subclasses may override Mult/MultAdd/Height/Width like tests/pytest/test_basematrix.py."""

    is_complex = False
    entrysize = 1
    ctx = None
    __array_ufunc__ = None

    def Height(self):
        return self.height

    def Width(self):
        return self.width

    def IsComplex(self):
        return self.is_complex

    def CreateRowVector(self):
        return BaseVector(self.Width(), self.is_complex, self.entrysize, ctx=self.ctx)

    def CreateColVector(self):
        return BaseVector(self.Height(), self.is_complex, self.entrysize, ctx=self.ctx)

    def CreateVector(self, colvector=False):
        return self.CreateColVector() if colvector else self.CreateRowVector()

    def Mult(self, x, y):
        # BaseMatrix::Mult: y.SetZero(); MultAdd(1, x, y)   linalg/basematrix.cpp:120-127
        y.SetScalar(0.0)
        self.MultAdd(1.0, x, y)

    def MultAdd(self, s, x, y):
        # default for subclasses that only provide Mult (basematrix.cpp: temporary + Add)
        tmp = y.CreateVector()
        self.Mult(x, tmp)
        y.Add(s, tmp)

    def CreateDeviceMatrix(self):
        # no creator registered -> the matrix itself (linalg/basematrix.cpp:388-398)
        return self

    def __mul__(self, x):
        if isinstance(x, MultiVector):
            return _MultiVectorExpr(self, x)
        if isinstance(x, BaseVector):
            return _Expr([(1.0, self, x)])
        if isinstance(x, _Expr):
            if len(x.terms) == 1 and x.terms[0][1] is None:
                s, _, v = x.terms[0]
                return _Expr([(s, self, v)])
            return _Expr([(1.0, self, x.Evaluate())])
        return NotImplemented

    def __rmul__(self, s):
        return _ScaledMatrix(s, self)       # VScaleMatrix, linalg/basematrix.hpp

    # composite operators (SumMatrix / ProductMatrix / Transpose, linalg/basematrix.hpp:560-860); like the reference they
    # hold their operands and recurse in CreateDeviceMatrix()
    def __add__(self, o):
        return SumMatrix(self, o, 1.0, 1.0) if isinstance(o, BaseMatrix) else NotImplemented

    def __sub__(self, o):
        return SumMatrix(self, o, 1.0, -1.0) if isinstance(o, BaseMatrix) else NotImplemented

    def __neg__(self):
        return _ScaledMatrix(-1.0, self)

    def __matmul__(self, o):
        return ProductMatrix(self, o) if isinstance(o, BaseMatrix) else NotImplemented

    @property
    def T(self):
        return TransposeMatrix(self)

    def MultTrans(self, s, x, y):
        # Python binding: y = 0; MultTransAdd(1.0, x, y) -- `s` is ignored (linalg/python_linalg.cpp:1074)
        y.SetScalar(0.0)
        self.MultTransAdd(1.0, x, y)

    def MultTransAdd(self, s, x, y):
        raise NgsbError("%s: MultTransAdd is not implemented" % type(self).__name__)

    def MultScale(self, s, x, y):
        self.Mult(x, y)
        if s != 1.0:
            y.Scale(s)

    def _check(self, x, y, who):
        if x.size != self.Width():
            raise NgsbError("%s: width of matrix = %d != size of x = %d" % (who, self.Width(), x.size))
        if y.size != self.Height():
            raise NgsbError("%s: height of matrix = %d != size of y = %d" % (who, self.Height(), y.size))


class _ScaledMatrix(BaseMatrix):
    def __init__(self, s, m):
        self.s, self.m = s, m
        self.is_complex, self.entrysize, self.ctx = m.is_complex, m.entrysize, m.ctx

    def Height(self):
        return self.m.Height()

    def Width(self):
        return self.m.Width()

    def Mult(self, x, y):
        y.SetScalar(0.0)
        self.m.MultAdd(self.s, x, y)

    def MultAdd(self, s, x, y):
        self.m.MultAdd(s * self.s, x, y)

    def MultTransAdd(self, s, x, y):
        self.m.MultTransAdd(s * self.s, x, y)

    def CreateDeviceMatrix(self):
        return _ScaledMatrix(self.s, self.m.CreateDeviceMatrix())


class SumMatrix(BaseMatrix):
    """a*A + b*B (SumMatrix, linalg/basematrix.hpp:560-600)"""

    def __init__(self, A, B, a=1.0, b=1.0):
        if A.Height() != B.Height() or A.Width() != B.Width():
            raise NgsbError("SumMatrix: sizes don't match: %d x %d and %d x %d" % (A.Height(), A.Width(), B.Height(), B.Width()))
        self.A, self.B, self.a, self.b = A, B, a, b
        self.is_complex, self.entrysize, self.ctx = A.is_complex or B.is_complex, A.entrysize, A.ctx or B.ctx

    def Height(self):
        return self.A.Height()

    def Width(self):
        return self.A.Width()

    def Mult(self, x, y):
        # SumMatrix::Mult: a == 1 ? A.Mult : (y = 0; A.MultAdd(a)); then B.MultAdd(b)
        if self.a == 1:
            self.A.Mult(x, y)
        else:
            y.SetScalar(0.0)
            self.A.MultAdd(self.a, x, y)
        self.B.MultAdd(self.b, x, y)

    def MultAdd(self, s, x, y):
        self.A.MultAdd(s * self.a, x, y)
        self.B.MultAdd(s * self.b, x, y)

    def MultTransAdd(self, s, x, y):
        self.A.MultTransAdd(s * self.a, x, y)
        self.B.MultTransAdd(s * self.b, x, y)

    def CreateDeviceMatrix(self):
        return SumMatrix(self.A.CreateDeviceMatrix(), self.B.CreateDeviceMatrix(), self.a, self.b)


class ProductMatrix(BaseMatrix):
    """A @ B (ProductMatrix, linalg/basematrix.hpp:720-760): y = A (B x) through a temporary of B's column type"""

    def __init__(self, A, B):
        if A.Width() != B.Height():
            raise NgsbError("ProductMatrix: width of A = %d != height of B = %d" % (A.Width(), B.Height()))
        self.A, self.B = A, B
        self.is_complex, self.entrysize, self.ctx = A.is_complex or B.is_complex, A.entrysize, A.ctx or B.ctx
        self._tmp = None

    def Height(self):
        return self.A.Height()

    def Width(self):
        return self.B.Width()

    def CreateRowVector(self):
        return self.B.CreateRowVector()

    def CreateColVector(self):
        return self.A.CreateColVector()

    def _t(self):
        if self._tmp is None:
            self._tmp = self.B.CreateColVector()
        return self._tmp

    def Mult(self, x, y):
        self.B.Mult(x, self._t())
        self.A.Mult(self._t(), y)

    def MultAdd(self, s, x, y):
        self.B.Mult(x, self._t())
        self.A.MultAdd(s, self._t(), y)

    def MultTransAdd(self, s, x, y):
        t = self.A.CreateRowVector()
        self.A.MultTrans(1.0, x, t)
        self.B.MultTransAdd(s, t, y)

    def CreateDeviceMatrix(self):
        return ProductMatrix(self.A.CreateDeviceMatrix(), self.B.CreateDeviceMatrix())


class TransposeMatrix(BaseMatrix):
    """A.T (Transpose, linalg/basematrix.hpp:800-860): Mult = MultTrans of the operand"""

    def __init__(self, A):
        self.A = A
        self.is_complex, self.entrysize, self.ctx = A.is_complex, A.entrysize, A.ctx

    def Height(self):
        return self.A.Width()

    def Width(self):
        return self.A.Height()

    def CreateRowVector(self):
        return self.A.CreateColVector()

    def CreateColVector(self):
        return self.A.CreateRowVector()

    def Mult(self, x, y):
        self.A.MultTrans(1.0, x, y)

    def MultAdd(self, s, x, y):
        self.A.MultTransAdd(s, x, y)

    def MultTransAdd(self, s, x, y):
        self.A.MultAdd(s, x, y)

    def CreateDeviceMatrix(self):
        return TransposeMatrix(self.A.CreateDeviceMatrix())


class IdentityMatrix(BaseMatrix):
    """IdentityMatrix(n) (linalg/special_matrix.hpp): y = x"""

    def __init__(self, size, complex=False, ctx=None):
        self.height = self.width = int(size)
        self.is_complex, self.ctx = complex, ctx or default_context()

    def Mult(self, x, y):
        y.Set(1.0, x)

    def MultAdd(self, s, x, y):
        y.Add(s, x)

    MultTransAdd = MultAdd


class SparseMatrix(BaseMatrix):
    """Host-side CSR holder: what SparseMatrix<TM>::CSR() hands out (linalg/python_linalg.cpp:
    121-138): values, int32 columns (ascending per row), uint64 row pointers.  No arithmetic
    happens on the host; CreateDeviceMatrix() uploads."""

    def __init__(self, rowptr, col, val, height=None, width=None, entrysize=1, ctx=None):
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.uint64)
        self.col = np.ascontiguousarray(col, dtype=np.int32)
        self.is_complex = np.iscomplexobj(val)
        self.entrysize = entrysize
        self.val = np.ascontiguousarray(val, dtype=np.complex128 if self.is_complex else np.float64).reshape(-1)
        self.height = len(self.rowptr) - 1 if height is None else int(height)
        self.width = self.height if width is None else int(width)
        self.nze = len(self.col)
        self.ctx = ctx or default_context()
        if len(self.val) != self.nze * entrysize * entrysize:
            raise NgsbError("SparseMatrix: %d values for %d entries of size %dx%d" % (len(self.val), self.nze, entrysize, entrysize))

    @staticmethod
    def CreateFromCOO(indi, indj, values, h, w):
        """SparseMatrixd.CreateFromCOO (linalg/python_linalg.cpp:144-152): duplicates are summed."""
        i = np.asarray(indi, dtype=np.int64)
        j = np.asarray(indj, dtype=np.int64)
        v = np.asarray(values)
        order = np.lexsort((j, i))
        i, j, v = i[order], j[order], v[order]
        key = i * np.int64(w) + j
        first = np.ones(len(key), dtype=bool)
        first[1:] = key[1:] != key[:-1]
        idx = np.flatnonzero(first)
        vals = np.add.reduceat(v, idx) if len(v) else v
        rows = i[idx]
        rowptr = np.zeros(h + 1, dtype=np.uint64)
        np.add.at(rowptr, rows + 1, 1)
        rowptr = np.cumsum(rowptr).astype(np.uint64)
        return SparseMatrix(rowptr, j[idx].astype(np.int32), vals, h, w)

    def CSR(self):
        return self.val, self.col, self.rowptr

    def CreateDeviceMatrix(self):
        return DevSparseMatrix(self)

    def CreateSmoother(self, freedofs=None):
        return JacobiPrecond(self, freedofs)

    def CreateBlockSmoother(self, blocks):
        """mat.CreateBlockSmoother(blocks) (linalg/python_linalg.cpp): blocks = list of dof lists (may overlap)"""
        return BlockJacobiPrecond(self, blocks)

    def Mult(self, x, y):
        raise NgsbError("SparseMatrix on the host: this package has no CPU path, use CreateDeviceMatrix()")

    MultAdd = Mult


def _mat_kind(is_complex, entrysize):
    return _kind(is_complex, entrysize)


class DevSparseMatrix(BaseMatrix):
    """ngscuda.DevSparseMatrix (ngscuda/cuda_linalg.hpp:50-72) for double, Complex and 3x3 blocks."""

    def __init__(self, mat, ctx=None, _handle=None):
        self.ctx = ctx or (mat.ctx if mat is not None else default_context())
        if _handle is None:
            h = C.c_void_p()
            check(_capi.lib().ngsb_csr_create(self.ctx.handle, mat.height, mat.width, mat.nze, _np_ptr(mat.rowptr),
                                              _np_ptr(mat.col), _np_ptr(mat.val), _mat_kind(mat.is_complex, mat.entrysize),
                                              C.byref(h)))
            _handle = h
        self.handle = _handle
        hh, ww, nn, kk = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_int()
        check(_capi.lib().ngsb_csr_info(self.handle, C.byref(hh), C.byref(ww), C.byref(nn), C.byref(kk)))
        self.height, self.width, self.nze, self.kind = hh.value, ww.value, nn.value, kk.value
        self.is_complex = self.kind == COMPLEX
        self.entrysize = 3 if self.kind == BLOCK3 else 1
        self._fin = weakref.finalize(self, _capi.lib().ngsb_csr_destroy, self.handle)

    def Mult(self, x, y):
        x._dev_read()
        y._dev_write()
        check(_capi.lib().ngsb_csr_mult(self.handle, x.handle, y.handle))

    def MultAdd(self, s, x, y):
        x._dev_read()
        y._dev_write()
        check(_capi.lib().ngsb_csr_multadd(self.handle, scal2(s), x.handle, y.handle))

    def CreateSmoother(self, freedofs=None):
        return DevJacobiMatrix(self, freedofs)

    def CreateBlockSmoother(self, blocks):
        return DevBlockJacobiMatrix(self, blocks)

    def MultTransAdd(self, s, x, y):
        """SparseMatrix::MultTransAdd, linalg/sparsematrix_impl.hpp:344-352 (A^T is built on the device at first use)"""
        x._dev_read()
        y._dev_write()
        check(_capi.lib().ngsb_csr_multtransadd(self.handle, scal2(s), x.handle, y.handle))

    def CreateTranspose(self, sorted=True):
        """mat.CreateTranspose() (linalg/python_linalg.cpp:172)"""
        h = C.c_void_p()
        check(_capi.lib().ngsb_csr_transpose(self.handle, C.byref(h)))
        return DevSparseMatrix(None, ctx=self.ctx, _handle=h)

    def MultAddMulti(self, alpha, xs, ys):
        """SparseMatrix<double>::MultAdd(alpha, MultiVector x, MultiVector y), linalg/sparsematrix.cpp:2274-2351"""
        k = len(xs)
        if len(ys) != k or len(alpha) != k:
            raise NgsbError("MultAdd(MultiVector): %d x vectors, %d y vectors, %d scalars" % (k, len(ys), len(alpha)))
        for v in xs:
            v._dev_read()
        for v in ys:
            v._dev_write()
        al = np.ascontiguousarray(alpha, dtype=np.float64)
        xa = (C.c_void_p * max(1, k))(*[v.handle for v in xs])
        ya = (C.c_void_p * max(1, k))(*[v.handle for v in ys])
        check(_capi.lib().ngsb_csr_multadd_multi(self.handle, k, _np_ptr(al), xa, ya))

    def Reorder(self, perm):
        """SparseMatrix::Reorder, linalg/sparsematrix_impl.hpp:762-783"""
        p = np.ascontiguousarray(perm, dtype=np.uint64)
        h = C.c_void_p()
        check(_capi.lib().ngsb_csr_reorder(self.handle, _np_ptr(p), C.byref(h)))
        return DevSparseMatrix(None, ctx=self.ctx, _handle=h)

    def Archive(self):
        """the bytes SparseMatrix<TM>::DoArchive writes into a BinaryOutArchive (linalg/sparsematrix_impl.hpp:443-452)"""
        n = C.c_size_t()
        check(_capi.lib().ngsb_csr_archive_size(self.handle, C.byref(n)))
        buf = np.empty(n.value, dtype=np.uint8)
        check(_capi.lib().ngsb_csr_archive_write(self.handle, _np_ptr(buf), n.value))
        return buf

    @staticmethod
    def FromArchive(data, kind=REAL, ctx=None):
        """device matrix from such bytes (kind: REAL, COMPLEX or BLOCK3 -- the archive does not name its entry type)"""
        ctx = ctx or default_context()
        buf = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8))
        h = C.c_void_p()
        check(_capi.lib().ngsb_csr_create_from_archive(ctx.handle, _np_ptr(buf), len(buf), kind, C.byref(h)))
        return DevSparseMatrix(None, ctx=ctx, _handle=h)

    def Memory(self):
        """(bytes of CSR arrays + tables resident, bytes of the SELL copy, CSR column/value arrays resident?)"""
        c, s, r = C.c_uint64(), C.c_uint64(), C.c_int()
        check(_capi.lib().ngsb_csr_memory(self.handle, C.byref(c), C.byref(s), C.byref(r)))
        return c.value, s.value, bool(r.value)

    def RCM(self):
        """the Cuthill-McKee permutation option "reorder" uses (ngsb_csr_rcm), in the form Reorder() takes"""
        p = np.empty(self.height, dtype=np.uint64)
        check(_capi.lib().ngsb_csr_rcm(self.handle, _np_ptr(p)))
        return p

    def ReorderInfo(self, want_perm=False):
        """(products run on an internally reordered copy?, natural 16-bit-offset share or -1[, permutation])"""
        r, s = C.c_int(), C.c_double()
        p = np.empty(self.height, dtype=np.uint64) if want_perm else None
        check(_capi.lib().ngsb_csr_reorder_info(self.handle, C.byref(r), _np_ptr(p) if want_perm else None, C.byref(s)))
        return (bool(r.value), s.value, p) if want_perm else (bool(r.value), s.value)

    def CSR(self):
        ms = 9 if self.kind == BLOCK3 else 1
        rowptr = np.empty(self.height + 1, dtype=np.uint64)
        col = np.empty(self.nze, dtype=np.int32)
        val = np.empty(self.nze * ms, dtype=np.complex128 if self.is_complex else np.float64)
        check(_capi.lib().ngsb_csr_download(self.handle, _np_ptr(rowptr), _np_ptr(col), _np_ptr(val)))
        return val, col, rowptr

    def Layout(self):
        """(padded SELL entries, rows with an overflow part, slice cap)"""
        e, o, c = C.c_uint64(), C.c_uint32(), C.c_uint32()
        check(_capi.lib().ngsb_csr_layout(self.handle, C.byref(e), C.byref(o), C.byref(c)))
        return e.value, o.value, c.value

    def StreamBytes(self):
        """(bytes the SELL kernel streams per Mult as stored, padded entries held with 16-bit column offsets)"""
        b, c = C.c_double(), C.c_uint64()
        check(_capi.lib().ngsb_csr_stream_bytes(self.handle, C.byref(b), C.byref(c)))
        return b.value, c.value

    def MultBytes(self):
        b = C.c_double()
        check(_capi.lib().ngsb_csr_mult_bytes(self.handle, C.byref(b)))
        return b.value


class SparseMatrixSymmetric(SparseMatrix):
    """SparseMatrixSymmetric<TM> as its CSR() hands it out: the lower triangle (columns <= row), what
    BilinearForm(symmetric_storage=True) assembles (linalg/sparsematrix.hpp:760-835).  CreateDeviceMatrix() expands it
    to the full matrix on the device; Mult/MultAdd then equal SparseMatrixSymmetric::MultAdd."""

    def CreateDeviceMatrix(self):
        h = C.c_void_p()
        check(_capi.lib().ngsb_csr_create_symmetric(self.ctx.handle, self.height, self.nze, _np_ptr(self.rowptr), _np_ptr(self.col),
                                                    _np_ptr(self.val), _mat_kind(self.is_complex, self.entrysize), C.byref(h)))
        return DevSparseMatrix(None, ctx=self.ctx, _handle=h)


class _MultiVectorExpr:
    def __init__(self, mat, mv):
        self.mat, self.mv = mat, mv


class MultiVector:
    """MultiVector(vec, k) (linalg/multivector.hpp, linalg/python_linalg.cpp:716-880): k vectors of vec's kind.
    `my[:] = mat * mx` runs the multi-right-hand-side product of the device matrix."""

    def __init__(self, arg, k=None):
        if isinstance(arg, BaseVector):
            self.vecs = [arg.CreateVector() for _ in range(int(k))]
        else:
            self.vecs = list(arg)

    def __len__(self):
        return len(self.vecs)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return MultiVector(self.vecs[i])
        return self.vecs[i]

    def __setitem__(self, key, value):
        tgt = self.vecs[key] if isinstance(key, slice) else [self.vecs[key]]
        if isinstance(value, _MultiVectorExpr):
            if len(value.mv) != len(tgt):
                raise NgsbError("MultiVector: %d vectors assigned to %d" % (len(value.mv), len(tgt)))
            for v in tgt:
                v.SetScalar(0.0)
            m = value.mat
            if isinstance(m, DevSparseMatrix):
                m.MultAddMulti(np.ones(len(tgt)), value.mv.vecs, tgt)
            else:
                for x, y in zip(value.mv.vecs, tgt):
                    m.MultAdd(1.0, x, y)
        elif isinstance(value, MultiVector):
            for x, y in zip(value.vecs, tgt):
                y.Set(1.0, x)
        else:
            for y in tgt:
                y.SetScalar(value)

    def __rmul__(self, m):
        return _MultiVectorExpr(m, self) if isinstance(m, BaseMatrix) else NotImplemented

    def InnerProduct(self, other, conjugate=True):
        """matrix of inner products (MultiVector::InnerProduct): out[i, j] = <other[j], self[i]> as `other.InnerProduct`"""
        out = np.zeros((len(self), len(other)), dtype=np.complex128 if self.vecs and self.vecs[0].is_complex else np.float64)
        for i, a in enumerate(self.vecs):
            for j, b in enumerate(other.vecs):
                out[i, j] = a.InnerProduct(b, conjugate=conjugate)
        return out


def CreateDevMatrix(mat):
    """ngscuda.CreateDevMatrix: throws if no device version exists (ngscuda/cuda_linalg.cpp:176-182)"""
    dev = mat.CreateDeviceMatrix()
    if dev is mat and not isinstance(mat, (DevSparseMatrix, DevJacobiMatrix, DevBlockJacobiMatrix)):
        raise NgsbError("CreateDevMatrix: no device matrix for %s" % type(mat).__name__)
    return dev


class JacobiPrecond(BaseMatrix):
    """mat.CreateSmoother(freedofs) on a host matrix; the inverse diagonal is built on the device
    by CreateDeviceMatrix() (linalg/jacobi.cpp:39-68 restated in the CUDA library)."""

    def __init__(self, mat, freedofs=None):
        self.mat, self.freedofs = mat, freedofs
        self.height = self.width = mat.height
        self.is_complex, self.entrysize, self.ctx = mat.is_complex, mat.entrysize, mat.ctx
        self._dev = None

    def CreateDeviceMatrix(self, devmat=None):
        if self._dev is None:
            self._dev = DevJacobiMatrix(devmat or self.mat.CreateDeviceMatrix(), self.freedofs)
        return self._dev

    def Mult(self, x, y):
        raise NgsbError("JacobiPrecond on the host: this package has no CPU path, use CreateDeviceMatrix()")

    MultAdd = Mult


class DevJacobiMatrix(BaseMatrix):
    """DevDiagonalMatrix built from JacobiPrecond (ngscuda/cuda_linalg.cpp:103-115, 321-366)."""

    def __init__(self, devmat=None, freedofs=None, invdiag=None, ctx=None, entrysize=1):
        h = C.c_void_p()
        bits = _freebits(freedofs)
        if devmat is not None:
            self.ctx = devmat.ctx
            check(_capi.lib().ngsb_jacobi_create_from_csr(devmat.handle, _np_ptr(bits) if bits is not None else None, C.byref(h)))
            self.height = self.width = devmat.height
            self.is_complex, self.entrysize = devmat.is_complex, devmat.entrysize
        else:
            self.ctx = ctx or default_context()
            inv = np.ascontiguousarray(invdiag)
            self.is_complex, self.entrysize = np.iscomplexobj(inv), entrysize
            n = inv.size // (entrysize * entrysize)
            if bits is not None and len(bits) < (n + 7) // 8:
                raise NgsbError("Projector/DiagonalMatrix: BitArray of %d bits for %d dofs" % (8 * len(bits), n))
            inv = inv.astype(np.complex128 if self.is_complex else np.float64).reshape(-1)
            check(_capi.lib().ngsb_jacobi_create(self.ctx.handle, n, _np_ptr(inv), _kind(self.is_complex, entrysize),
                                                 _np_ptr(bits) if bits is not None else None, C.byref(h)))
            self.height = self.width = n
        self.handle = h
        self._fin = weakref.finalize(self, _capi.lib().ngsb_jacobi_destroy, h)

    def Mult(self, x, y):
        self._check(x, y, "JacobiPrecond::Mult")
        x._dev_read()
        y._dev_write()
        check(_capi.lib().ngsb_jacobi_mult(self.handle, x.handle, y.handle))

    def MultAdd(self, s, x, y):
        x._dev_read()
        y._dev_write()
        check(_capi.lib().ngsb_jacobi_multadd(self.handle, scal2(s), x.handle, y.handle))

    def MultTransAdd(self, s, x, y):
        if self.entrysize != 1:
            raise NgsbError("DiagonalMatrix::MultTransAdd: only scalar entries")
        self.MultAdd(s, x, y)

    def InvDiag(self):
        ms = 9 if self.entrysize == 3 else 1
        out = np.empty(self.height * ms, dtype=np.complex128 if self.is_complex else np.float64)
        check(_capi.lib().ngsb_jacobi_download(self.handle, _np_ptr(out)))
        return out


def _block_table(blocks):
    """list of dof lists (or (first, dofs) arrays) -> the reference's Table<int> as (first uint64[nb+1], dofs int32)"""
    if isinstance(blocks, tuple) and len(blocks) == 2 and isinstance(blocks[0], np.ndarray):
        return np.ascontiguousarray(blocks[0], dtype=np.uint64), np.ascontiguousarray(blocks[1], dtype=np.int32)
    first = np.zeros(len(blocks) + 1, dtype=np.uint64)
    first[1:] = np.cumsum([len(b) for b in blocks], dtype=np.uint64)
    dofs = np.fromiter((d for b in blocks for d in b), dtype=np.int32, count=int(first[-1]))
    return first, dofs


class BlockJacobiPrecond(BaseMatrix):
    """mat.CreateBlockSmoother(blocks) on a host matrix (BlockJacobiPrecond<double>, linalg/blockjacobi.cpp:380-500);
    the block inverses are built on the device by CreateDeviceMatrix()."""

    def __init__(self, mat, blocks):
        self.mat, self.blocks = mat, blocks
        self.height = self.width = mat.height
        self.is_complex, self.entrysize, self.ctx = mat.is_complex, mat.entrysize, mat.ctx
        self._dev = None

    def CreateDeviceMatrix(self, devmat=None):
        if self._dev is None:
            self._dev = DevBlockJacobiMatrix(devmat or self.mat.CreateDeviceMatrix(), self.blocks)
        return self._dev

    def Mult(self, x, y):
        raise NgsbError("BlockJacobiPrecond on the host: this package has no CPU path, use CreateDeviceMatrix()")

    MultAdd = Mult


class DevBlockJacobiMatrix(BaseMatrix):
    """DevBlockJacobiMatrix (ngscuda/dev_blockjacobi.cpp:21-140): y(block) += s * inv(A(block,block)) * x(block)."""

    def __init__(self, devmat, blocks, inverses=None, n=None, ctx=None):
        first, dofs = _block_table(blocks)
        h = C.c_void_p()
        if inverses is None:
            self.ctx = devmat.ctx
            check(_capi.lib().ngsb_blockjacobi_create(devmat.handle, len(first) - 1, _np_ptr(first), _np_ptr(dofs) if len(dofs) else None, C.byref(h)))
            self.height = self.width = devmat.height
        else:
            # inverse blocks computed elsewhere (list of row-major matrices), e.g. the reference's GetInverses()
            self.ctx = ctx or (devmat.ctx if devmat is not None else default_context())
            flat = np.ascontiguousarray(np.concatenate([np.asarray(m, dtype=np.float64).reshape(-1) for m in inverses] or [np.zeros(1)]))
            self.height = self.width = int(n if n is not None else devmat.height)
            check(_capi.lib().ngsb_blockjacobi_create_from_inverses(self.ctx.handle, self.height, len(first) - 1, _np_ptr(first),
                                                                    _np_ptr(dofs) if len(dofs) else None, _np_ptr(flat), C.byref(h)))
        self.handle = h
        self.is_complex, self.entrysize = False, 1
        self.first, self.dofs = first, dofs
        self._fin = weakref.finalize(self, _capi.lib().ngsb_blockjacobi_destroy, h)

    def Mult(self, x, y):
        self._check(x, y, "BlockJacobiPrecond::Mult")
        x._dev_read()
        y._dev_write()
        check(_capi.lib().ngsb_blockjacobi_mult(self.handle, x.handle, y.handle, 0))

    def MultAdd(self, s, x, y):
        self._check(x, y, "BlockJacobiPrecond::MultAdd")
        x._dev_read()
        y._dev_write()
        check(_capi.lib().ngsb_blockjacobi_multadd(self.handle, float(s), x.handle, y.handle, 0))

    def MultTrans(self, s, x, y):
        # the Python binding takes a `value` and ignores it: y = 0; MultTransAdd(1.0, x, y)  (linalg/python_linalg.cpp:1074)
        self._check(x, y, "BlockJacobiPrecond::MultTrans")
        x._dev_read()
        y._dev_write()
        check(_capi.lib().ngsb_blockjacobi_mult(self.handle, x.handle, y.handle, 1))

    def MultTransAdd(self, s, x, y):
        self._check(x, y, "BlockJacobiPrecond::MultTransAdd")
        x._dev_read()
        y._dev_write()
        check(_capi.lib().ngsb_blockjacobi_multadd(self.handle, float(s), x.handle, y.handle, 1))

    def GetInverses(self):
        """the inverse blocks (row-major numpy matrices), BlockJacobiPrecond::GetInverses"""
        tot = C.c_size_t()
        check(_capi.lib().ngsb_blockjacobi_info(self.handle, None, None, None, None, C.byref(tot)))
        flat = np.empty(max(1, tot.value))
        check(_capi.lib().ngsb_blockjacobi_download(self.handle, _np_ptr(flat)))
        out, off = [], 0
        for b in range(len(self.first) - 1):
            bs = int(self.first[b + 1] - self.first[b])
            out.append(flat[off:off + bs * bs].reshape(bs, bs).copy())
            off += bs * bs
        return out


def DiagonalMatrix(diag, ctx=None):
    """DiagonalMatrix<double|Complex> on the device (linalg/diagonalmatrix.hpp:12-80; DevDiagonalMatrix,
    ngscuda/cuda_linalg.cpp:117-123, 321-366): y = diag .* x"""
    d = diag.NumPy() if isinstance(diag, BaseVector) else np.asarray(diag)
    return DevJacobiMatrix(invdiag=d, ctx=ctx)


def Projector(mask, range=True, ctx=None):
    """Projector(mask, range): keeps (range=True) or clears (range=False) the dofs of the BitArray
    (linalg/special_matrix.hpp; DevProjector, ngscuda/cuda_linalg.cpp:887-967).  It is what
    python/krylovspace.py:79 puts in place of a missing preconditioner: Projector(freedofs, True)."""
    b = mask if isinstance(mask, BitArray) else BitArray(mask)
    n = b.n
    bits = b.bytes if range else np.bitwise_not(b.bytes)
    return DevJacobiMatrix(invdiag=np.ones(n), freedofs=BitArray(bits), ctx=ctx)


# ------------------------------------------------------------------------------------------------
# Krylov solvers (C++ classes of linalg/cg.cpp as bound in linalg/python_linalg.cpp:1773-1830)
# ------------------------------------------------------------------------------------------------
class _KrylovSolver(BaseMatrix):
    def __init__(self, mat, pre=None, printrates=False, precision=1e-8, maxsteps=200, conjugate=False):
        self.mat = mat.CreateDeviceMatrix() if isinstance(mat, SparseMatrix) else mat
        if isinstance(pre, (JacobiPrecond, BlockJacobiPrecond)):
            pre = pre.CreateDeviceMatrix(self.mat if isinstance(self.mat, DevSparseMatrix) else None)
        self.pre = pre
        self.precision, self.maxsteps, self.conjugate, self.printrates = precision, maxsteps, conjugate, printrates
        self.steps = 0
        self.history = np.zeros(0)
        self.height, self.width = self.mat.Width(), self.mat.Height()
        self.is_complex, self.entrysize, self.ctx = self.mat.is_complex, self.mat.entrysize, self.mat.ctx

    def GetSteps(self):
        return self.steps

    def SetPrecision(self, p):
        self.precision = p

    def SetMaxSteps(self, m):
        self.maxsteps = m

    def _fused(self):
        return isinstance(self.mat, DevSparseMatrix) and (self.pre is None or isinstance(self.pre, DevJacobiMatrix))

    def MultAdd(self, s, x, y):
        tmp = y.CreateVector()
        self.Mult(x, tmp)
        y.Add(s, tmp)


class CGSolver(_KrylovSolver):
    """CGSolver<IPTYPE>::Mult, linalg/cg.cpp:503-633.  With a device sparse matrix and a device
    Jacobi (or no) preconditioner the whole loop runs in the library (fused kernels, device-side
    stopping rule); any other operator pair is driven op by op through Mult/InnerProduct/Add with
    the same recurrences."""

    def Mult(self, f, u, initialize=True):
        if self._fused():
            f._dev_read()
            u._dev_write()
            ip = 0 if not self.is_complex else (2 if self.conjugate else 1)
            cap = self.maxsteps + 2
            hist = np.zeros(cap)
            steps, nh = C.c_int(), C.c_int()
            check(_capi.lib().ngsb_cg_solve(self.mat.handle, self.pre.handle if self.pre is not None else None, f.handle, u.handle,
                                            self.precision, self.maxsteps, ip, 1 if initialize else 0, C.byref(steps),
                                            _np_ptr(hist), cap, C.byref(nh)))
            self.steps = steps.value
            self.history = hist[:min(nh.value, cap)].copy()
            return
        self._generic(f, u, initialize)

    def _generic(self, f, u, initialize):
        A, Cm, conj = self.mat, self.pre, self.conjugate
        d, w, s, as_ = f.CreateVector(), f.CreateVector(), f.CreateVector(), f.CreateVector()
        if initialize:
            u.SetScalar(0.0)
            d.Set(1.0, f)
        else:
            A.Mult(u, as_)
            d.Set(1.0, f)
            d.Add(-1.0, as_)
        if Cm is not None:
            Cm.Mult(d, w)
        else:
            w.Set(1.0, d)
        s.Set(1.0, w)
        wdn = w.InnerProduct(d, conjugate=conj)
        hist = [abs(wdn)]
        if wdn == 0:
            wdn = 1
        err = self.precision ** 2 * abs(wdn)
        n = 0
        while True:
            cont = n < self.maxsteps and abs(wdn) > err
            n += 1
            if not cont:
                break
            A.Mult(s, as_)
            wd = wdn
            kss = s.InnerProduct(as_, conjugate=conj)
            if kss == 0:
                break
            al = wd / kss
            u.Add(al, s)
            d.Add(-al, as_)
            if Cm is not None:
                Cm.Mult(d, w)
            else:
                w.Set(1.0, d)
            wdn = d.InnerProduct(w, conjugate=conj)
            be = wdn / wd
            s.Scale(be)
            s.Add(1.0, w)
            hist.append(abs(wdn))
        self.steps = n
        self.history = np.array(hist)


class DevCGSolver(CGSolver):
    """ngscuda.DevCGSolver(mat, pre, maxsteps, precision), ngscuda/python_ngscuda.cpp:243-266"""

    def __init__(self, mat, pre, maxsteps=200, precision=1e-8):
        super().__init__(mat, pre, precision=precision, maxsteps=maxsteps)


class GMRESSolver(_KrylovSolver):
    """GMRESSolver<IPTYPE>::Mult, linalg/cg.cpp:854-1022 (device-resident in the library)."""

    def Mult(self, f, x, initialize=True):
        if not self._fused():
            raise NgsbError("GMRESSolver: operands must be a device sparse matrix and a device Jacobi preconditioner")
        f._dev_read()
        x._dev_write()
        cap = self.maxsteps + 2
        hist = np.zeros(cap)
        steps, nh = C.c_int(), C.c_int()
        check(_capi.lib().ngsb_gmres_solve(self.mat.handle, self.pre.handle if self.pre is not None else None, f.handle, x.handle,
                                           self.precision, self.maxsteps, 1 if initialize else 0, C.byref(steps), _np_ptr(hist),
                                           cap, C.byref(nh)))
        self.steps = steps.value
        self.history = hist[:min(nh.value, cap)].copy()


def cg_solve_host(mat, pre, f_host, precision=1e-8, maxsteps=200, conjugate=False):
    """`gfu.vec.data = inv * f.vec` with host buffers in one call (ngsb_cg_solve_host): copies f
    to the device, solves, copies u back.  Returns (u, steps, history)."""
    is_c = mat.is_complex
    f = np.ascontiguousarray(f_host, dtype=np.complex128 if is_c else np.float64).reshape(-1)
    u = np.empty_like(f)
    cap = maxsteps + 2
    hist = np.zeros(cap)
    steps, nh = C.c_int(), C.c_int()
    ip = 0 if not is_c else (2 if conjugate else 1)
    check(_capi.lib().ngsb_cg_solve_host(mat.handle, pre.handle if pre is not None else None, _np_ptr(f), _np_ptr(u), precision,
                                         maxsteps, ip, C.byref(steps), _np_ptr(hist), cap, C.byref(nh)))
    return u, steps.value, hist[:min(nh.value, cap)].copy()
