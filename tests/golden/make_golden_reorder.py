#!/usr/bin/env python
"""Golden fixture for the dof reordering, produced by the REFERENCE itself:

    source oracle/_ref/ngs/env.sh && python tests/golden/make_golden_reorder.py

  reorder_netgen_h1p3.npz   netgen unit_cube (maxh=0.4, one Refine()) H1 order-3 stiffness matrix in NGSolve's own dof
                            numbering; `perm` = the Cuthill-McKee permutation of oracle/ngs_oracle.c (orc_rcm, the serial
                            specification of csrc/reorder.cu); r_rowptr/r_col/r_val = SparseMatrix<double>::Reorder(perm)
                            of the reference (linalg/sparsematrix_impl.hpp:762-783, reached through tests/golden/
                            ref_helpers.cpp because it has no Python binding); y = Mult of the reference on x;
                            CGSolver(mat, jacobi) steps and solution.
`import ngsolve` must come before numpy (SURVEY.md 8c pitfall 4).
"""
import os
import subprocess
import sys
import sysconfig
import tempfile

import ngsolve
from ngsolve import *          # noqa: F401,F403
from netgen.csg import unit_cube

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pyoracle as orc     # noqa: E402

ngsolve.ngsglobals.msg_level = 0
pfx = os.path.join(ROOT, "oracle", "_ref", "ngs")
tmp = tempfile.mkdtemp()
so = os.path.join(tmp, "ref_helpers" + sysconfig.get_config_var("EXT_SUFFIX"))
subprocess.check_call([os.path.join(pfx, "bin", "ngscxx"), "-shared", os.path.join(HERE, "ref_helpers.cpp"), "-L" + os.path.join(pfx, "lib"),
                       "-lngla", "-lngstd", "-lngbla", "-L" + os.path.join(pfx, "lib", "python3.12", "site-packages", "netgen"), "-lngcore", "-o", so])
sys.path.insert(0, tmp)
import ref_helpers             # noqa: E402

mesh = Mesh(unit_cube.GenerateMesh(maxh=0.4))
mesh.Refine()
fes = H1(mesh, order=3, dirichlet=".*")
u, v = fes.TnT()
a = BilinearForm(grad(u) * grad(v) * dx).Assemble()
f = LinearForm(1 * v * dx).Assemble()
val, col, rowptr = a.mat.CSR()
rowptr = np.array(rowptr, dtype=np.uint64); col = np.array(col, dtype=np.int32); val = np.array(val)
n = fes.ndof
perm = orc.Csr(rowptr, col, val, 0).rcm()
rmat = ref_helpers.reorder(a.mat, [int(p) for p in perm])
rval, rcol, rrowptr = rmat.CSR()
fd = fes.FreeDofs()
bits = np.zeros((n + 7) // 8, dtype=np.uint8)
for i in range(n):
    if fd[i]:
        bits[i >> 3] |= np.uint8(1 << (i & 7))
rng = np.random.default_rng(7)
x = rng.random(n)
xv = a.mat.CreateRowVector(); yv = a.mat.CreateColVector()
xv.FV().NumPy()[:] = x
yv.data = a.mat * xv
jac = a.mat.CreateSmoother(fd)
inv = CGSolver(a.mat, jac, precision=1e-8, maxsteps=5000, printrates=False)
gfu = GridFunction(fes)
gfu.vec.data = inv * f.vec
np.savez_compressed(os.path.join(HERE, "reorder_netgen_h1p3.npz"), rowptr=rowptr, col=col, val=val, perm=perm,
                    r_rowptr=np.array(rrowptr, dtype=np.uint64), r_col=np.array(rcol, dtype=np.int32), r_val=np.array(rval),
                    freebits=bits, x=x, y=np.array(yv.FV().NumPy()), f=np.array(f.vec.FV().NumPy()), u=np.array(gfu.vec.FV().NumPy()),
                    cg_steps=inv.GetSteps(), ne=mesh.ne, nv=mesh.nv, ngsolve=ngsolve.__version__)
d = np.abs(np.repeat(np.arange(n), np.diff(np.array(rrowptr, dtype=np.int64))) - np.array(rcol))
d0 = np.abs(np.repeat(np.arange(n), np.diff(rowptr.astype(np.int64))) - col)
print("n", n, "nnz", len(col), "steps", inv.GetSteps(), "mean |i-j| natural", d0.mean(), "reordered", d.mean())
