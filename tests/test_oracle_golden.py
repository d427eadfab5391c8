"""CPU: the oracle (oracle/ngs_oracle.c) against golden vectors produced by the reference.

This is what pins the oracle: every fixture value was computed by NGSolve's own code
(tests/golden/make_golden.py).  Tolerances are stated per check; integer results
(iteration counts) must match exactly.
"""
import numpy as np
import pytest

from conftest import SYSTEMS, kind_of, load_golden, relerr
from oracle import pyoracle as orc


def system(name):
    g = load_golden(name)
    kind = kind_of(g)
    A = orc.Csr(g["rowptr"], g["col"], g["val"], kind)
    return g, kind, A


@pytest.mark.parametrize("name", SYSTEMS)
def test_mult_and_multadd(name):
    g, kind, A = system(name)
    assert relerr(A.mult(g["x"]), g["y_mult"]) <= 1e-14
    y = g["y0"].copy()
    A.multadd(0.7, g["x"], y)
    assert relerr(y, g["y_multadd"]) <= 1e-14
    if kind == 1:
        y = g["y0"].copy()
        A.multadd(0.3 - 0.9j, g["x"], y)
        assert relerr(y, g["y_multadd_cs"]) <= 1e-14


@pytest.mark.parametrize("name", SYSTEMS)
def test_vector_ops(name):
    g, kind, A = system(name)
    x, y0 = g["x"], g["y0"]
    if kind == 1:
        assert abs(orc.inner(x, y0, False) - g["dot_xy"][0]) <= 1e-13 * abs(g["dot_xy"][0])
        assert abs(orc.inner(x, y0, True) - g["dot_xy_conj"][0]) <= 1e-13 * abs(g["dot_xy_conj"][0])
    else:
        # 16-chunk summation restated exactly -> bitwise equal unless the reference build contracted FMAs
        assert abs(orc.inner(x, y0) - g["dot_xy"][0]) <= 4e-16 * abs(g["dot_xy"][0]) * 16
    assert abs(orc.l2norm(x) - g["norm_x"][0]) <= 1e-14 * g["norm_x"][0]
    y = y0.copy()
    orc.axpy(y, 0.5, x)
    assert relerr(y, g["axpy_05"]) <= 1e-15


@pytest.mark.parametrize("name", SYSTEMS)
def test_jacobi(name):
    g, kind, A = system(name)
    J = orc.Jacobi(A, g["freebits"])
    assert relerr(J.mult(g["x"]), g["jac_mult"]) <= 1e-13
    y = g["y0"].copy()
    J.multadd(0.25, g["x"], y)
    assert relerr(y, g["jac_multadd"]) <= 1e-13


@pytest.mark.parametrize("name", SYSTEMS)
def test_cg_matches_reference(name):
    g, kind, A = system(name)
    J = orc.Jacobi(A, g["freebits"])
    u, steps, hist = orc.cg_solve(A, J, g["f"], prec=float(g["cg_prec"]), maxsteps=int(g["cg_maxsteps"]))
    # the bar of BASELINE.json: within +-2 iterations of the reference CGSolver
    assert abs(steps - int(g["cg_steps"])) <= 2, (steps, int(g["cg_steps"]))
    assert relerr(u, g["cg_u"]) <= 1e-6
    # python CGSolver: iterations and residual history sqrt(|<d,w>|)
    res = g["pycg_residuals"]
    m = min(len(res), len(hist)) - 2
    assert m >= 3
    # CG amplifies rounding differences along the run: the early history must agree tightly,
    # the whole history within a factor of 10 (same convergence curve)
    k = min(m, 20)
    assert np.allclose(np.sqrt(hist[:k]), res[:k], rtol=1e-6, atol=0)
    if name != "helmholtz_h1p4_complex":      # indefinite: CG wanders, only the early part is comparable
        ratio = np.sqrt(hist[:m]) / res[:m]
        assert ratio.max() < 10 and ratio.min() > 0.1
    # exit by maxsteps: 7 iterations -> GetSteps() == 8
    u7, steps7, _ = orc.cg_solve(A, J, g["f"], prec=1e-30, maxsteps=7)
    assert steps7 == int(g["cg7_steps"]) == 8
    assert relerr(u7, g["cg7_u"]) <= 1e-11
    if kind == 1:
        uc, stepsc, _ = orc.cg_solve(A, J, g["f"], prec=float(g["cg_prec"]), maxsteps=int(g["cg_maxsteps"]), ip_mode=2)
        assert abs(stepsc - int(g["cgconj_steps"])) <= 2
        if int(g["cgconj_steps"]) < int(g["cg_maxsteps"]):
            assert relerr(uc, g["cgconj_u"]) <= 1e-5


@pytest.mark.parametrize("name", ["poisson_h1p3", "helmholtz_h1p4_complex", "shifted_laplace_complex", "square_h1p4_testsolvers"])
def test_gmres_matches_reference(name):
    g, kind, A = system(name)
    J = orc.Jacobi(A, g["freebits"])
    x, steps, hist = orc.gmres_solve(A, J, g["f"], prec=float(g["gmres_prec"]), maxsteps=int(g["gmres_maxsteps"]))
    assert abs(steps - int(g["gmres_steps"])) <= 2, (steps, int(g["gmres_steps"]))
    assert relerr(x, g["gmres_u"]) <= 1e-6


def test_reference_test_solvers_problem():
    """tests/pytest/test_solvers.py:59-79 (Jacobi instead of BDDC): p4 is exact, error < 1e-12."""
    g, kind, A = system("square_h1p4_testsolvers")
    ex = load_golden("square_h1p4_exact")
    J = orc.Jacobi(A, g["freebits"])
    u, steps, _ = orc.cg_solve(A, J, g["f"], prec=1e-13, maxsteps=3000)
    M = orc.Csr(ex["mass_rowptr"], ex["mass_col"], ex["mass_val"], 0)
    e = u - ex["u_interp"]
    assert np.sqrt(abs(orc.inner(e, M.mult(e)))) < 1e-12
    assert relerr(u, ex["u_direct"]) < 1e-11


def test_reference_test_matrix_golden_entry():
    """tests/pytest/test_matrix.py:85-93: H1(dim=3) mass matrix, a.mat[1,1] == x * I3."""
    g = load_golden("test_matrix_cube_h1dim3")
    rp, col, val = g["rowptr"], g["col"], g["val"].reshape(-1, 3, 3)
    j = int(rp[1]) + int(np.searchsorted(col[int(rp[1]):int(rp[2])], 1))
    assert col[j] == 1
    x = float(g["golden_x"])
    assert np.linalg.norm(val[j] - x * np.eye(3)) < 1e-8
    # and through the oracle's block SpMV: A * e_(1,c) has x in component c of row 1
    A = orc.Csr(rp, col, g["val"], 3)
    for c in range(3):
        e = np.zeros(3 * A.n)
        e[3 + c] = 1.0
        y = A.mult(e)
        assert abs(y[3 + c] - x) < 1e-8 and abs(y[3 + (c + 1) % 3]) < 1e-12


def test_reorder_is_permutation_similarity():
    g, kind, A = system("poisson_h1p3")
    rng = np.random.default_rng(3)
    perm = rng.permutation(A.n).astype(np.uint64)
    B = A.reorder(perm)
    # rows sorted ascending, same nnz
    for i in range(0, A.n, 97):
        c = B.col[int(B.rowptr[i]):int(B.rowptr[i + 1])]
        assert np.all(np.diff(c) > 0)
    x = g["x"]
    # (P A P^T)(P x) = P (A x)   with (P x)[i] = x[perm[i]]
    assert relerr(B.mult(x[perm.astype(np.int64)]), A.mult(x)[perm.astype(np.int64)]) <= 1e-14


def test_pardofs_tables():
    # 3 ranks; local dofs of rank 1 shared as listed (linalg/paralleldofs.cpp:46-66)
    dist = [[], [0], [2], [0, 2], [], [2]]
    exch, master = orc.pardofs_build(3, 1, dist)
    assert [e.tolist() for e in exch] == [[1, 3], [], [2, 3, 5]]
    assert master.tolist() == [1, 0, 1, 0, 1, 1]
