#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the REFERENCE itself.

Run in the build container only (needs the reference built from /root/reference, see
DESIGN.md "Oracle / reference build"):

    D=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs
    PYTHONPATH=/tmp/ngs/lib/python3.12/site-packages LD_LIBRARY_PATH=/tmp/ngs/lib:$D \
        python tests/golden/make_golden.py

Every array stored here is an output of NGSolve's own CPU code path
(SparseMatrix::MultAdd, BaseVector::InnerProduct, JacobiPrecond, CGSolver, GMRESSolver,
krylovspace.CGSolver): the oracle (oracle/ngs_oracle.c) and the CUDA path are both checked
against them.  `import ngsolve` must come before numpy (SURVEY.md 8c pitfall 4).
"""
import os
import sys

import ngsolve
from ngsolve import *          # noqa: F401,F403
from ngsolve import krylovspace
from netgen.csg import unit_cube
from netgen.geom2d import unit_square

import numpy as np

OUT = os.path.dirname(os.path.abspath(__file__))
REFTESTS = "/root/reference/tests/pytest"
ngsolve.ngsglobals.msg_level = 0


def csr_of(mat):
    vals, cols, rp = mat.CSR()
    return (np.array(rp, dtype=np.uint64), np.array(cols, dtype=np.int32), np.array(vals))


def freebits(fes, n):
    fd = fes.FreeDofs()
    bits = np.zeros((n + 7) // 8, dtype=np.uint8)
    for i in range(n):
        if fd[i]:
            bits[i >> 3] |= np.uint8(1 << (i & 7))
    return bits, fd


def vec_np(v, cplx=False):
    return np.array(v.FV().NumPy(), dtype=np.complex128 if cplx else np.float64).copy()


def set_vec(v, arr):
    v.FV().NumPy()[:] = arr


def fixture(name, fes, a, f, cplx, block, solver="cg", prec=1e-8, maxsteps=2000, gmres_maxsteps=60):
    mat = a.mat
    n = mat.height
    es = block
    rp, cols, vals = csr_of(mat)
    bits, fd = freebits(fes, n)
    rng = np.random.default_rng(12345)
    N = n * es
    dt = np.complex128 if cplx else np.float64

    def rnd():
        r = rng.random(N)
        if cplx:
            r = r + 1j * rng.random(N)
        return r.astype(dt)

    x = mat.CreateRowVector()
    y = mat.CreateColVector()
    xv, y0 = rnd(), rnd()
    set_vec(x, xv)
    out = dict(rowptr=rp, col=cols, val=vals, freebits=bits, n=np.int64(n), block=np.int64(es), is_complex=np.int64(cplx), x=xv, y0=y0)
    # Mult and MultAdd
    mat.Mult(x, y)
    out["y_mult"] = vec_np(y, cplx)
    set_vec(y, y0)
    mat.MultAdd(0.7, x, y)
    out["y_multadd"] = vec_np(y, cplx)
    if cplx:
        set_vec(y, y0)
        mat.MultAdd(0.3 - 0.9j, x, y)
        out["y_multadd_cs"] = vec_np(y, cplx)
    # reductions / updates on BaseVectors of this kind
    x2 = mat.CreateRowVector()
    set_vec(x2, y0)
    if cplx:
        out["dot_xy"] = np.array([x.InnerProduct(x2, conjugate=False)])
        out["dot_xy_conj"] = np.array([x.InnerProduct(x2, conjugate=True)])
    else:
        out["dot_xy"] = np.array([x.InnerProduct(x2)])
    out["norm_x"] = np.array([x.Norm()])
    x2.data += 0.5 * x
    out["axpy_05"] = vec_np(x2, cplx)
    x2.data = 3.0 * x
    out["set_3"] = vec_np(x2, cplx)
    # Jacobi (masked by freedofs)
    jac = mat.CreateSmoother(fd)
    jac.Mult(x, y)
    out["jac_mult"] = vec_np(y, cplx)
    set_vec(y, y0)
    jac.MultAdd(0.25, x, y)
    out["jac_multadd"] = vec_np(y, cplx)
    # right-hand side and solves
    fv = vec_np(f.vec, cplx)
    out["f"] = fv
    u = f.vec.CreateVector()
    if solver in ("cg", "both"):
        inv = CGSolver(mat, jac, precision=prec, maxsteps=maxsteps, printrates=False)
        u.data = inv * f.vec
        out["cg_steps"] = np.int64(inv.GetSteps())
        out["cg_u"] = vec_np(u, cplx)
        out["cg_prec"] = np.float64(prec)
        out["cg_maxsteps"] = np.int64(maxsteps)
        # python CG gives the residual history sqrt(|<d,w>|); same iteration sequence
        pinv = krylovspace.CGSolver(mat, jac, tol=prec, maxiter=maxsteps, printrates=False)
        up = pinv.Solve(rhs=f.vec)
        out["pycg_iterations"] = np.int64(pinv.iterations)
        out["pycg_residuals"] = np.array(list(pinv.residuals), dtype=np.float64)
        out["pycg_u"] = vec_np(up, cplx)
        if cplx:
            invc = CGSolver(mat, jac, precision=prec, maxsteps=maxsteps, printrates=False, conjugate=True)
            u.data = invc * f.vec
            out["cgconj_steps"] = np.int64(invc.GetSteps())
            out["cgconj_u"] = vec_np(u, cplx)
        # a capped run pins the maxsteps exit
        invm = CGSolver(mat, jac, precision=1e-30, maxsteps=7, printrates=False)
        u.data = invm * f.vec
        out["cg7_steps"] = np.int64(invm.GetSteps())
        out["cg7_u"] = vec_np(u, cplx)
        # no preconditioner is not constructible for the C++ CGSolver from python (pre is
        # mandatory), so that branch stays oracle-only.
    if solver in ("gmres", "both"):
        ginv = GMRESSolver(mat, jac, printrates=False, precision=prec, maxsteps=gmres_maxsteps)
        u.data = ginv * f.vec
        out["gmres_steps"] = np.int64(ginv.GetSteps())
        out["gmres_u"] = vec_np(u, cplx)
        out["gmres_prec"] = np.float64(prec)
        out["gmres_maxsteps"] = np.int64(gmres_maxsteps)
    out["ngsolve_version"] = np.array([ngsolve.__version__])
    out["mat_type"] = np.array([type(mat).__name__])
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: n={n} es={es} nnz={len(cols)} type={type(mat).__name__}"
          f" cg_steps={out.get('cg_steps')} gmres_steps={out.get('gmres_steps')} -> {os.path.getsize(path)/1e6:.2f} MB")


def main():
    # C1-like: Poisson, H1 order 3, unit cube
    mesh = Mesh(unit_cube.GenerateMesh(maxh=0.28))
    fes = H1(mesh, order=3, dirichlet=".*")
    u, v = fes.TnT()
    a = BilinearForm(fes)
    a += grad(u) * grad(v) * dx
    a.Assemble()
    f = LinearForm(fes)
    f += 1 * v * dx
    f.Assemble()
    fixture("poisson_h1p3", fes, a, f, False, 1, solver="both", gmres_maxsteps=80)

    # C2-like: elasticity, H1 order 4, dim 3 -> SparseMatrix<Mat<3,3,double>>
    mesh = Mesh(unit_cube.GenerateMesh(maxh=0.7))
    fes = H1(mesh, order=4, dim=3, dirichlet="back")
    u, v = fes.TnT()
    E, nu = 210.0, 0.2
    mu = E / 2 / (1 + nu)
    lam = E * nu / ((1 + nu) * (1 - 2 * nu))

    def eps(w):
        return 0.5 * (grad(w) + grad(w).trans)

    a = BilinearForm(fes)
    a += (2 * mu * InnerProduct(eps(u), eps(v)) + lam * Trace(grad(u)) * Trace(grad(v))) * dx
    a.Assemble()
    f = LinearForm(fes)
    f += CF((0, 0, -1)) * v * dx
    f.Assemble()
    fixture("elasticity_h1p4_dim3", fes, a, f, False, 3, solver="cg", maxsteps=5000)

    # C4-like: Maxwell curl-curl + mass, HCurl order 2
    mesh = Mesh(unit_cube.GenerateMesh(maxh=0.35))
    fes = HCurl(mesh, order=2, dirichlet=".*")
    u, v = fes.TnT()
    a = BilinearForm(fes)
    a += (curl(u) * curl(v) + u * v) * dx
    a.Assemble()
    f = LinearForm(fes)
    f += CF((1, 0.5, -0.25)) * v * dx
    f.Assemble()
    fixture("maxwell_hcurlp2", fes, a, f, False, 1, solver="cg", maxsteps=5000)

    # C5-like: complex Helmholtz, H1 order 4, impedance boundary
    mesh = Mesh(unit_cube.GenerateMesh(maxh=0.5))
    fes = H1(mesh, order=4, complex=True)
    u, v = fes.TnT()
    omega = 10.0
    a = BilinearForm(fes)
    a += (grad(u) * grad(v) - omega * omega * u * v) * dx
    a += -1j * omega * u * v * ds
    a.Assemble()
    f = LinearForm(fes)
    f += exp(-20 * ((x - 0.5) ** 2 + (y - 0.5) ** 2 + (z - 0.5) ** 2)) * v * dx
    f.Assemble()
    fixture("helmholtz_h1p4_complex", fes, a, f, True, 1, solver="both", maxsteps=300, gmres_maxsteps=120)

    # complex-symmetric but definite system (shifted Laplace): CG<Complex> converges
    a = BilinearForm(fes)
    a += (grad(u) * grad(v) + (1 + 2j) * u * v) * dx
    a.Assemble()
    fixture("shifted_laplace_complex", fes, a, f, True, 1, solver="both", maxsteps=2000, gmres_maxsteps=100)

    # tests/pytest/test_solvers.py:59-79 problem (unit square, H1 p4), Jacobi instead of BDDC
    mesh = Mesh(unit_square.GenerateMesh(maxh=0.2))
    fes = H1(mesh, order=4, dirichlet=".*")
    u, v = fes.TnT()
    f = LinearForm(32 * (y * (1 - y) + x * (1 - x)) * v * dx).Assemble()
    a = BilinearForm(grad(u) * grad(v) * dx).Assemble()
    fixture("square_h1p4_testsolvers", fes, a, f, False, 1, solver="both", prec=1e-13, maxsteps=3000, gmres_maxsteps=200)
    # exact discrete solution from the reference's direct solver, and its L2 error functional
    gfu = GridFunction(fes)
    gfu.vec.data = a.mat.Inverse(fes.FreeDofs(), inverse="sparsecholesky") * f.vec
    exact = 16 * x * (1 - x) * y * (1 - y)
    err = sqrt(Integrate((gfu - exact) * (gfu - exact), mesh))
    m = BilinearForm(u * v * dx).Assemble()      # mass matrix: error norm without ngsolve at test time
    rp, cols, vals = csr_of(m.mat)
    gex = GridFunction(fes)
    gex.Set(exact)
    np.savez_compressed(os.path.join(OUT, "square_h1p4_exact.npz"), u_direct=vec_np(gfu.vec), l2err_direct=np.float64(err),
                        mass_rowptr=rp, mass_col=cols, mass_val=vals, u_interp=vec_np(gex.vec))
    print("square_h1p4_exact: direct-solve L2 error", err)

    # tests/pytest/test_matrix.py:57-100 golden entries (stored meshes of the reference)
    os.chdir(REFTESTS)
    mesh = Mesh("cube.vol.gz")
    fes = H1(mesh, dim=3)
    u, v = fes.TrialFunction(), fes.TestFunction()
    a = BilinearForm(fes)
    a += SymbolicBFI(InnerProduct(u, v))
    a.Assemble()
    rp, cols, vals = csr_of(a.mat)
    blk = np.array(a.mat[1, 1])
    fesc = H1(mesh, complex=True)
    u, v = fesc.TrialFunction(), fesc.TestFunction()
    ac = BilinearForm(fesc)
    ac += SymbolicBFI(InnerProduct(u, v))
    ac.Assemble()
    rpc, colsc, valsc = csr_of(ac.mat)
    np.savez_compressed(os.path.join(OUT, "test_matrix_cube_h1dim3.npz"), rowptr=rp, col=cols, val=vals, n=np.int64(a.mat.height),
                        block_1_1=blk, golden_x=np.float64(0.0002480226944690391),
                        c_rowptr=rpc, c_col=colsc, c_val=valsc, c_entry_1_1=np.array([ac.mat[1, 1]]))
    print("test_matrix_cube_h1dim3: a.mat[1,1] =", blk.tolist())


if __name__ == "__main__":
    main()
