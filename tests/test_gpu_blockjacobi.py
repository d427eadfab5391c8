"""GPU: block-Jacobi preconditioner through the C ABI against the oracle and the reference's fixtures.
BlockJacobiPrecond<double> ctor + MultAdd/MultTransAdd, linalg/blockjacobi.cpp:380-500, 594-681;
DevBlockJacobiMatrix, ngscuda/dev_blockjacobi.cpp:21-140."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _fixture(name="poisson_h1p3.npz"):
    return np.load(os.path.join(GOLD, name))


def _blocks(n, rng):
    # overlapping blocks of very different sizes (1 .. 150 dofs), one empty block, dofs not sorted
    sizes = [1, 2, 3, 5, 17, 31, 32, 33, 64, 100, 150, 0, 7]
    return [list(map(int, rng.choice(n, size=k, replace=False))) for k in sizes]


def _rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def test_inverses_and_products_match_the_oracle():
    import ngsolve_b200.la as la
    from oracle import pyoracle as orc
    g = _fixture()
    rng = np.random.default_rng(7)
    n = len(g["rowptr"]) - 1
    free = np.unpackbits(g["freebits"], bitorder="little")[:n].astype(bool)
    blocks = [[d for d in b if free[d]] for b in _blocks(n, rng)]
    A = la.SparseMatrix(g["rowptr"], g["col"], g["val"])
    dev = A.CreateDeviceMatrix()
    bj = dev.CreateBlockSmoother(blocks)
    obj = orc.BlockJacobi(orc.Csr(g["rowptr"], g["col"], g["val"], 0), blocks)
    for mine, ref in zip(bj.GetInverses(), obj.inverses()):
        assert mine.shape == ref.shape
        if ref.size:
            assert _rel(mine, ref) <= 1e-11           # same Gauss-Jordan, same pivots; only FMA contraction differs
    x = rng.random(n)
    xv = la.BaseVector(x)
    y = bj.CreateColVector()
    bj.Mult(xv, y)
    assert _rel(y.NumPy(), obj.mult(x)) <= 1e-12
    bj.MultTrans(1.0, xv, y)
    assert _rel(y.NumPy(), obj.mult(x, transpose=True)) <= 1e-12
    y0 = rng.random(n)
    y = la.BaseVector(y0)
    bj.MultAdd(-0.75, xv, y)
    assert _rel(y.NumPy(), obj.multadd(-0.75, x, y0.copy())) <= 1e-12
    y = la.BaseVector(y0)
    bj.MultTransAdd(2.5, xv, y)
    assert _rel(y.NumPy(), obj.multadd(2.5, x, y0.copy(), transpose=True)) <= 1e-12
    # the same preconditioner from inverses computed elsewhere (the adapter's path: BlockJacobiPrecond::GetInverses)
    bj2 = la.DevBlockJacobiMatrix(None, blocks, inverses=obj.inverses(), n=n)
    bj2.Mult(xv, y)
    assert _rel(y.NumPy(), obj.mult(x)) <= 1e-13
    # deterministic: no atomics on the path
    y1, y2 = bj.CreateColVector(), bj.CreateColVector()
    bj.Mult(xv, y1)
    bj.Mult(xv, y2)
    assert np.array_equal(y1.NumPy(), y2.NumPy())


def test_host_precond_object_and_cg():
    """mat.CreateBlockSmoother(blocks) -> CGSolver(mat, pre) (op-by-op CG with a non-Jacobi device preconditioner): same
    solution as the Jacobi-preconditioned solve (step counts against the reference: test_reference_fixture)"""
    import ngsolve_b200.la as la
    g = _fixture()
    n = len(g["rowptr"]) - 1
    free = np.unpackbits(g["freebits"], bitorder="little")[:n].astype(bool)
    # blocks = each free dof together with the free dofs of its matrix row (vertex-patch-like, heavily overlapping)
    rp, col = g["rowptr"].astype(np.int64), g["col"]
    blocks = [[int(c) for c in col[rp[i]:rp[i + 1]] if free[c]] for i in range(0, n, 9) if free[i]]
    covered = np.zeros(n, dtype=bool)
    for b in blocks:
        covered[b] = True
    blocks += [[int(i)] for i in np.flatnonzero(free & ~covered)]
    A = la.SparseMatrix(g["rowptr"], g["col"], g["val"])
    pre = A.CreateBlockSmoother(blocks)
    f = la.BaseVector(g["f"])
    inv = la.CGSolver(A, pre, precision=1e-8, maxsteps=2000)
    u = (inv * f).Evaluate().NumPy()
    jac = la.CGSolver(A, A.CreateSmoother(la.BitArray(g["freebits"])), precision=1e-8, maxsteps=2000)
    uj = (jac * f).Evaluate().NumPy()
    assert 1 < inv.GetSteps() < 2000 and 1 < jac.GetSteps() < 2000
    assert _rel(u, uj) <= 1e-6
    assert isinstance(la.CreateDevMatrix(pre), la.DevBlockJacobiMatrix)


def test_reference_fixture():
    """golden vectors produced by the reference's BlockJacobiPrecond (tests/golden/make_golden_next.py)"""
    path = os.path.join(GOLD, "next_blockjacobi.npz")
    if not os.path.exists(path):
        pytest.skip("fixture not generated")
    import ngsolve_b200.la as la
    g = np.load(path)
    A = la.SparseMatrix(g["rowptr"], g["col"], g["val"]).CreateDeviceMatrix()
    bj = A.CreateBlockSmoother((g["bfirst"], g["bdofs"]))
    x = la.BaseVector(g["x"])
    y = bj.CreateColVector()
    bj.Mult(x, y)
    assert _rel(y.NumPy(), g["bj_mult"]) <= 1e-12
    y = la.BaseVector(g["y0"])
    bj.MultAdd(0.5, x, y)
    assert _rel(y.NumPy(), g["bj_multadd_05"]) <= 1e-12
    bj.MultTrans(1.0, x, y)
    assert _rel(y.NumPy(), g["bj_multtrans"]) <= 1e-12
    inv = la.CGSolver(A, bj, precision=float(g["cg_prec"]), maxsteps=int(g["cg_maxsteps"]))
    u = (inv * la.BaseVector(g["f"])).Evaluate().NumPy()
    assert abs(inv.GetSteps() - int(g["cg_steps"])) <= 2
    assert _rel(u, g["cg_u"]) <= 1e-6


def test_errors_are_loud():
    import ngsolve_b200.la as la
    g = _fixture()
    n = len(g["rowptr"]) - 1
    A = la.SparseMatrix(g["rowptr"], g["col"], g["val"]).CreateDeviceMatrix()
    with pytest.raises(la.NgsbError, match="out of range"):
        A.CreateBlockSmoother([[0, n]])
    # a block of two dofs whose 2x2 matrix is singular: build a tiny matrix by hand
    S = la.SparseMatrix(np.array([0, 2, 4], dtype=np.uint64), np.array([0, 1, 0, 1], dtype=np.int32), np.array([1.0, 2.0, 2.0, 4.0]))
    with pytest.raises(la.NgsbError, match="singular"):
        S.CreateDeviceMatrix().CreateBlockSmoother([[0, 1]])
    bj = A.CreateBlockSmoother([[0, 1, 2]])
    x = la.BaseVector(np.ones(n + 1))
    with pytest.raises(la.NgsbError):
        bj.Mult(x, bj.CreateColVector())
    gc = _fixture("helmholtz_h1p4_complex.npz")
    Ac = la.SparseMatrix(gc["rowptr"], gc["col"], gc["val"]).CreateDeviceMatrix()
    with pytest.raises(la.NgsbError, match="double"):
        Ac.CreateBlockSmoother([[0, 1]])
