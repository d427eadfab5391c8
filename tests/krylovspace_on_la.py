"""TEST HELPER (not part of the product package): pure-Python Krylov solvers in the style of ngsolve.krylovspace
(python/krylovspace.py), so that the op-by-op device path of the ctypes mirror ngsolve_b200.la can be driven on a box
without NGSolve.  The reference's OWN file runs on the real adapter in tests/test_gpu_dropin.py.

they only use the BaseMatrix / BaseVector interface (`w.data = A * s`, `InnerProduct`,
`+=`), so they exercise the op-by-op device path -- the way an unchanged NGSolve script
reaches the library through the generic virtual calls.

CGSolver follows python/krylovspace.py:263-290, GMResSolver python/krylovspace.py:988-1095
(no restart), both with the residual bookkeeping of LinearSolver.CheckResidual (:129-156):
`iterations` counts residual checks, `residuals` is the list of checked values.
"""
from math import sqrt

import numpy as np

from ngsolve_b200.la import BaseMatrix, BaseVector, Norm, Projector  # noqa: F401


class LinearSolver(BaseMatrix):
    name = "LinearSolver"

    def __init__(self, mat, pre=None, freedofs=None, tol=None, maxiter=100, atol=None, callback=None, printrates=False):
        if atol is None and tol is None:
            tol = 1e-12
        assert (freedofs is None) != (pre is None)      # either pre or freedofs must be given (python/krylovspace.py:78)
        self.mat = mat.CreateDeviceMatrix()
        if pre is None:
            pre = Projector(freedofs, True, ctx=self.mat.ctx)
        self.pre = pre.CreateDeviceMatrix() if hasattr(pre, "CreateDeviceMatrix") else pre
        self.tol, self.atol, self.maxiter, self.callback, self.printrates = tol, atol, maxiter, callback, printrates
        self.residuals, self.iterations = [], 0
        self.height, self.width = self.mat.Width(), self.mat.Height()
        self.is_complex, self.entrysize, self.ctx = self.mat.is_complex, self.mat.entrysize, self.mat.ctx

    def Solve(self, rhs, sol=None, initialize=True):
        self.iterations, self.residuals = 0, []
        if sol is None:
            sol = rhs.CreateVector()
            initialize = True
        if initialize:
            sol[:] = 0
        self.sol = sol
        self._SolveImpl(rhs=rhs, sol=sol)
        return sol

    def Mult(self, x, y):
        self.Solve(rhs=x, sol=y, initialize=True)

    def CheckResidual(self, residual):
        self.iterations += 1
        self.residuals.append(residual)
        if len(self.residuals) == 1:
            if self.tol is None:
                self._final_residual = self.atol
            else:
                self._final_residual = residual * self.tol
                if self.atol is not None:
                    self._final_residual = max(self._final_residual, self.atol)
        elif self.callback is not None:
            self.callback(self.iterations, residual)
        if self.printrates:
            print("%s iteration %d, residual = %g" % (self.name, self.iterations, residual))
        return self.iterations >= self.maxiter or residual <= self._final_residual


class CGSolver(LinearSolver):
    name = "CG"

    def __init__(self, *args, conjugate=False, **kwargs):
        super().__init__(*args, **kwargs)
        self.conjugate = conjugate

    def _SolveImpl(self, rhs, sol):
        A, pre, conj = self.mat, self.pre, self.conjugate
        d, w, s = sol.CreateVector(), sol.CreateVector(), sol.CreateVector()
        d.data = rhs - A * sol
        w.data = pre * d
        s.data = w
        wdn = w.InnerProduct(d, conjugate=conj)
        if self.CheckResidual(sqrt(abs(wdn))):
            return
        while True:
            w.data = A * s
            wd = wdn
            as_s = s.InnerProduct(w, conjugate=conj)
            if as_s == 0 or wd == 0:
                break
            alpha = wd / as_s
            sol.data += alpha * s
            d.data += (-alpha) * w
            w.data = pre * d
            wdn = w.InnerProduct(d, conjugate=conj)
            if self.CheckResidual(sqrt(abs(wdn))):
                return
            beta = wdn / wd
            s *= beta
            s.data += w


class GMResSolver(LinearSolver):
    name = "GMRes"

    def _SolveImpl(self, rhs, sol):
        A, pre, m = self.mat, self.pre, self.maxiter
        cplx = rhs.is_complex
        dt = np.complex128 if cplx else np.float64
        ip = lambda x, y: y.InnerProduct(x, conjugate=True)      # noqa: E731
        sn, cs = np.zeros(m, dtype=dt), np.zeros(m, dtype=dt)
        tmp, r = rhs.CreateVector(), rhs.CreateVector()
        tmp.data = rhs - A * sol
        r.data = pre * tmp
        Q, H = [rhs.CreateVector()], []
        r_norm = Norm(r)
        if self.CheckResidual(abs(r_norm)):
            return
        Q[0].data = (1.0 / r_norm) * r
        beta = np.zeros(m + 1, dtype=dt)
        beta[0] = r_norm

        def givens(v1, v2):
            if v2 == 0:
                return 1, 0
            if v1 == 0:
                return 0, v2 / abs(v2)
            t = sqrt((np.conj(v1) * v1 + np.conj(v2) * v2).real)
            return abs(v1) / t, v1 / abs(v1) * np.conj(v2) / t

        k = 0
        for k in range(m):
            q = rhs.CreateVector()
            tmp.data = A * Q[k]
            q.data = pre * tmp
            h = np.zeros(m + 1, dtype=dt)
            for i in range(k + 1):
                h[i] = ip(Q[i], q)
                q -= h[i] * Q[i]
            h[k + 1] = Norm(q)
            H.append(h)
            if abs(h[k + 1]) < 1e-12:
                break
            q *= 1.0 / h[k + 1].real
            Q.append(q)
            for i in range(k):
                t = cs[i] * h[i] + sn[i] * h[i + 1]
                h[i + 1] = -np.conj(sn[i]) * h[i] + np.conj(cs[i]) * h[i + 1]
                h[i] = t
            cs[k], sn[k] = givens(h[k], h[k + 1])
            h[k] = cs[k] * h[k] + sn[k] * h[k + 1]
            h[k + 1] = 0
            beta[k + 1] = -np.conj(sn[k]) * beta[k]
            beta[k] = cs[k] * beta[k]
            if self.CheckResidual(abs(beta[k + 1])):
                break
        Hm = np.zeros((k + 1, k + 1), dtype=dt)
        for i in range(k + 1):
            Hm[:, i] = H[i][:k + 1]
        y = np.linalg.solve(Hm, beta[:k + 1])       # (k+1)x(k+1) Hessenberg system: host scalars, like the reference
        for i in range(k + 1):
            sol.data += (complex(y[i]) if cplx else float(y[i])) * Q[i]
