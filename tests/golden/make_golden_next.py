#!/usr/bin/env python
"""Golden fixtures for the SURVEY.md 8(f) rows, produced by the REFERENCE itself (oracle/build_reference.sh build):

    source oracle/_ref/ngs/env.sh && python tests/golden/make_golden_next.py

  next_blockjacobi.npz   BlockJacobiPrecond<double> (mat.CreateBlockSmoother): Mult / MultAdd / MultTrans outputs and
                         CGSolver(mat, blockjacobi) steps + solution, vertex-patch blocks of an H1 order-3 cube problem
  next_transpose.npz     SparseMatrix::MultTransAdd for double / Complex / Mat<3,3> matrices (non-symmetric values),
                         CreateTranspose() CSR
  next_symmetric.npz     SparseMatrixSymmetric<double> (BilinearForm(symmetric=True)): lower-triangle CSR + Mult/MultAdd
  next_multivector.npz   SparseMatrix<double>::MultAdd(alpha, MultiVector x, MultiVector y) with 6 vectors (one 4-group + 2)
  next_operators.npz     composite operators of linalg/basematrix.hpp: (A + 2 B) x, (A @ C) x, A.T x, (3 A) x, Projector
`import ngsolve` must come before numpy (SURVEY.md 8c pitfall 4).
"""
import os

import ngsolve
from ngsolve import *          # noqa: F401,F403
from netgen.csg import unit_cube

import numpy as np

OUT = os.path.dirname(os.path.abspath(__file__))
ngsolve.ngsglobals.msg_level = 0


def csr_of(mat):
    vals, cols, rp = mat.CSR()
    return (np.array(rp, dtype=np.uint64), np.array(cols, dtype=np.int32), np.array(vals))


def vec_np(v, cplx=False):
    return np.array(v.FV().NumPy(), dtype=np.complex128 if cplx else np.float64).copy()


def set_vec(v, arr):
    v.FV().NumPy()[:] = arr


def freebits(fes):
    fd = fes.FreeDofs()
    n = fes.ndof
    bits = np.zeros((n + 7) // 8, dtype=np.uint8)
    for i in range(n):
        if fd[i]:
            bits[i >> 3] |= np.uint8(1 << (i & 7))
    return bits


rng = np.random.default_rng(2024)
mesh = Mesh(unit_cube.GenerateMesh(maxh=0.35))

# ---------------------------------------------------------------- block Jacobi
fes = H1(mesh, order=3, dirichlet=".*")
u, v = fes.TnT()
a = BilinearForm(grad(u) * grad(v) * dx).Assemble()
f = LinearForm(1 * v * dx).Assemble()
n = fes.ndof
blocks = []
fd = fes.FreeDofs()
for vert in mesh.vertices:
    dofs = set(d for d in fes.GetDofNrs(vert) if d >= 0 and fd[d])
    for ed in mesh[vert].edges:
        dofs |= set(d for d in fes.GetDofNrs(ed) if d >= 0 and fd[d])
    for fa in mesh[vert].faces:
        dofs |= set(d for d in fes.GetDofNrs(fa) if d >= 0 and fd[d])
    if dofs:
        blocks.append(sorted(dofs))
for el in mesh.Elements(VOL):
    dofs = [d for d in fes.GetDofNrs(NodeId(CELL, el.nr)) if d >= 0 and fd[d]]
    if dofs:
        blocks.append(dofs)
bj = a.mat.CreateBlockSmoother(blocks)
x = a.mat.CreateColVector()
y = a.mat.CreateColVector()
xn = rng.random(n)
y0 = rng.random(n)
set_vec(x, xn)
bj.Mult(x, y)
bj_mult = vec_np(y)
set_vec(y, y0)
bj.MultAdd(0.5, x, y)
bj_multadd = vec_np(y)
bj.MultTrans(1.0, x, y)
bj_multtrans = vec_np(y)
inv = CGSolver(a.mat, bj, precision=1e-8, maxsteps=2000)
sol = a.mat.CreateColVector()
sol.data = inv * f.vec
rp, col, val = csr_of(a.mat)
bfirst = np.zeros(len(blocks) + 1, dtype=np.uint64)
bfirst[1:] = np.cumsum([len(b) for b in blocks])
np.savez_compressed(os.path.join(OUT, "next_blockjacobi.npz"), rowptr=rp, col=col, val=val, freebits=freebits(fes),
                    bfirst=bfirst, bdofs=np.array([d for b in blocks for d in b], dtype=np.int32), x=xn, y0=y0,
                    bj_mult=bj_mult, bj_multadd_05=bj_multadd, bj_multtrans=bj_multtrans, f=vec_np(f.vec),
                    cg_steps=inv.GetSteps(), cg_u=vec_np(sol), cg_prec=1e-8, cg_maxsteps=2000,
                    ngsolve_version=ngsolve.__version__)
print("blockjacobi: n", n, "blocks", len(blocks), "max bs", max(len(b) for b in blocks), "cg steps", inv.GetSteps())

# ---------------------------------------------------------------- MultTransAdd / CreateTranspose (non-symmetric values)
out = {}
for tag, space, cplx in (("d", H1(mesh, order=2), False), ("z", H1(mesh, order=2, complex=True), True),
                         ("b3", H1(mesh, order=1, dim=3), False)):
    uu, vv = space.TnT()
    if tag == "b3":
        b = BilinearForm(InnerProduct(grad(uu), grad(vv)) * dx + (uu[0] * vv[1] + 2 * uu[2] * vv[0] - grad(uu)[0, 0] * vv[2]) * dx).Assemble()
    elif cplx:
        b = BilinearForm(grad(uu) * grad(vv) * dx + (1 + 2j) * grad(uu)[0] * vv * dx).Assemble()
    else:
        b = BilinearForm(grad(uu) * grad(vv) * dx + grad(uu)[0] * vv * dx + 3 * uu * grad(vv)[1] * dx).Assemble()
    m = b.mat
    rp, col, val = csr_of(m)
    es = 3 if tag == "b3" else 1
    xs = rng.random(m.height * es) + (1j * rng.random(m.height * es) if cplx else 0)
    ys = rng.random(m.height * es) + (1j * rng.random(m.height * es) if cplx else 0)
    xv, yv = m.CreateColVector(), m.CreateRowVector()
    set_vec(xv, xs)
    set_vec(yv, ys)
    m.MultTransAdd(0.75, xv, yv)
    out.update({tag + "_rowptr": rp, tag + "_col": col, tag + "_val": val, tag + "_x": xs, tag + "_y0": ys, tag + "_multtransadd_075": vec_np(yv, cplx)})
    if tag == "d":
        t = m.CreateTranspose()
        trp, tcol, tval = csr_of(t)
        out.update(d_t_rowptr=trp, d_t_col=tcol, d_t_val=tval)
    print("transpose", tag, "h", m.height, "nze", len(col), type(m).__name__)
np.savez_compressed(os.path.join(OUT, "next_transpose.npz"), ngsolve_version=ngsolve.__version__, **out)

# ---------------------------------------------------------------- symmetric storage
fes2 = H1(mesh, order=2, dirichlet=".*")
uu, vv = fes2.TnT()
s = BilinearForm(fes2, symmetric=True, symmetric_storage=True)
s += grad(uu) * grad(vv) * dx + uu * vv * dx
s.Assemble()
rp, col, val = csr_of(s.mat)
xs, ys = rng.random(fes2.ndof), rng.random(fes2.ndof)
xv, yv = s.mat.CreateColVector(), s.mat.CreateColVector()
set_vec(xv, xs)
s.mat.Mult(xv, yv)
sym_mult = vec_np(yv)
set_vec(yv, ys)
s.mat.MultAdd(-1.5, xv, yv)
np.savez_compressed(os.path.join(OUT, "next_symmetric.npz"), rowptr=rp, col=col, val=val, x=xs, y0=ys, y_mult=sym_mult,
                    y_multadd_m15=vec_np(yv), mat_type=type(s.mat).__name__, ngsolve_version=ngsolve.__version__)
print("symmetric:", type(s.mat).__name__, "n", fes2.ndof, "stored nze", len(col))

# ---------------------------------------------------------------- MultiVector
m = a.mat
K = 6
mx = MultiVector(m.CreateColVector(), K)
my = MultiVector(m.CreateColVector(), K)
X = rng.random((K, n))
Y0 = rng.random((K, n))
for k in range(K):
    set_vec(mx[k], X[k])
    set_vec(my[k], Y0[k])
my[:] = m * mx
Ymult = np.array([vec_np(my[k]) for k in range(K)])
np.savez_compressed(os.path.join(OUT, "next_multivector.npz"), X=X, Y_mult=Ymult, ngsolve_version=ngsolve.__version__)
print("multivector: K", K)

# ---------------------------------------------------------------- composite operators
fesA = H1(mesh, order=2)
uu, vv = fesA.TnT()
A = BilinearForm(grad(uu) * grad(vv) * dx + grad(uu)[0] * vv * dx).Assemble().mat
B = BilinearForm(uu * vv * dx).Assemble().mat
rpA, colA, valA = csr_of(A)
rpB, colB, valB = csr_of(B)
xs = rng.random(fesA.ndof)
xv, yv = A.CreateColVector(), A.CreateColVector()
set_vec(xv, xs)
res = {}
yv.data = (A + 2 * B) * xv
res["sum"] = vec_np(yv)
yv.data = (A @ B) * xv
res["prod"] = vec_np(yv)
yv.data = A.T * xv
res["trans"] = vec_np(yv)
yv.data = (3 * A) * xv
res["scaled"] = vec_np(yv)
yv.data = (A - B) * xv + 0.5 * xv
res["expr"] = vec_np(yv)
mask = BitArray(fesA.ndof)
mask.Clear()
for i in range(0, fesA.ndof, 3):
    mask.Set(i)
yv.data = Projector(mask, True) * xv
res["proj_range"] = vec_np(yv)
yv.data = Projector(mask, False) * xv
res["proj_kernel"] = vec_np(yv)
yv.data = (IdentityMatrix(fesA.ndof) - Projector(mask, True) @ A) * xv
res["id_minus_pa"] = vec_np(yv)
np.savez_compressed(os.path.join(OUT, "next_operators.npz"), a_rowptr=rpA, a_col=colA, a_val=valA, b_rowptr=rpB, b_col=colB, b_val=valB,
                    x=xs, mask=np.array([bool(mask[i]) for i in range(fesA.ndof)]), ngsolve_version=ngsolve.__version__, **res)
print("operators: n", fesA.ndof)
