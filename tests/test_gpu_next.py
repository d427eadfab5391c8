"""GPU: the SURVEY.md 8(f) rows through the C ABI against fixtures produced by the reference
(tests/golden/make_golden_next.py) and against the oracle: MultTransAdd / CreateTranspose, symmetric storage,
MultiVector products (4 right-hand sides per sweep), composite operators."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    return np.max(np.abs(a - b)) / np.max(np.abs(b))


@pytest.mark.parametrize("tag,es", [("d", 1), ("z", 1), ("b3", 3)])
def test_multtransadd_matches_reference(tag, es):
    import ngsolve_b200.la as la
    g = np.load(os.path.join(GOLD, "next_transpose.npz"))
    A = la.SparseMatrix(g[tag + "_rowptr"], g[tag + "_col"], g[tag + "_val"], entrysize=es).CreateDeviceMatrix()
    x = la.BaseVector(g[tag + "_x"], entrysize=es)
    y = la.BaseVector(g[tag + "_y0"], entrysize=es)
    A.MultTransAdd(0.75, x, y)
    assert _rel(y.NumPy().reshape(-1), g[tag + "_multtransadd_075"]) <= 1e-12
    A.MultTransAdd(0.75, x, y)            # cached transpose, accumulates again
    again = 2 * g[tag + "_multtransadd_075"] - g[tag + "_y0"]
    assert _rel(y.NumPy().reshape(-1), again) <= 1e-12
    # Python binding quirk kept: MultTrans(value, x, y) ignores value
    A.MultTrans(123.0, x, y)
    assert _rel(y.NumPy().reshape(-1), (g[tag + "_multtransadd_075"] - g[tag + "_y0"]) / 0.75) <= 1e-11


def test_create_transpose_pattern_is_bit_exact():
    import ngsolve_b200.la as la
    g = np.load(os.path.join(GOLD, "next_transpose.npz"))
    A = la.SparseMatrix(g["d_rowptr"], g["d_col"], g["d_val"]).CreateDeviceMatrix()
    val, col, rp = A.CreateTranspose().CSR()
    assert np.array_equal(rp, g["d_t_rowptr"]) and np.array_equal(col, g["d_t_col"]) and np.array_equal(val, g["d_t_val"])
    # rectangular + empty rows/columns
    rng = np.random.default_rng(3)
    import scipy.sparse as sp
    M = sp.random(37, 91, density=0.08, random_state=5, format="csr")
    M.sort_indices()
    R = la.SparseMatrix(M.indptr.astype(np.uint64), M.indices.astype(np.int32), M.data, height=37, width=91).CreateDeviceMatrix()
    T = R.CreateTranspose()
    tv, tc, tr = T.CSR()
    MT = M.T.tocsr()
    MT.sort_indices()
    assert T.height == 91 and T.width == 37
    assert np.array_equal(tr, MT.indptr.astype(np.uint64)) and np.array_equal(tc, MT.indices) and np.array_equal(tv, MT.data)
    x = rng.random(37)
    y = la.BaseVector(np.zeros(91))
    R.MultTransAdd(2.0, la.BaseVector(x), y)
    assert _rel(y.NumPy(), 2.0 * (M.T @ x)) <= 1e-13
    with pytest.raises(la.NgsbError):
        R.MultTransAdd(1.0, la.BaseVector(np.ones(91)), y)


def test_symmetric_storage_matches_reference():
    import ngsolve_b200.la as la
    from oracle import pyoracle as orc
    g = np.load(os.path.join(GOLD, "next_symmetric.npz"))
    S = la.SparseMatrixSymmetric(g["rowptr"], g["col"], g["val"]).CreateDeviceMatrix()
    x = la.BaseVector(g["x"])
    y = S.CreateColVector()
    S.Mult(x, y)
    assert _rel(y.NumPy(), g["y_mult"]) <= 1e-12
    y = la.BaseVector(g["y0"])
    S.MultAdd(-1.5, x, y)
    assert _rel(y.NumPy(), g["y_multadd_m15"]) <= 1e-12
    assert _rel(y.NumPy(), orc.sym_multadd(g["rowptr"], g["col"], g["val"], -1.5, g["x"], g["y0"].copy())) <= 1e-12
    # the expanded matrix is the full symmetric matrix: pattern symmetric, twice the strict lower part + diagonal
    val, col, rp = S.CSR()
    n = len(rp) - 1
    import scipy.sparse as sp
    F = sp.csr_matrix((val, col, rp.astype(np.int64)), shape=(n, n))
    assert abs(F - F.T).max() == 0.0
    ndiag = int(sum(1 for i in range(n) if g["rowptr"][i + 1] > g["rowptr"][i] and g["col"][int(g["rowptr"][i + 1]) - 1] == i))
    assert S.nze == 2 * len(g["col"]) - ndiag
    with pytest.raises(la.NgsbError, match="lower triangle"):
        la.SparseMatrixSymmetric(np.array([0, 2, 3], dtype=np.uint64), np.array([0, 1, 1], dtype=np.int32), np.ones(3)).CreateDeviceMatrix()


def test_multivector_product_matches_reference_and_single_products():
    import ngsolve_b200.la as la
    g = np.load(os.path.join(GOLD, "next_multivector.npz"))
    b = np.load(os.path.join(GOLD, "next_blockjacobi.npz"))
    A = la.SparseMatrix(b["rowptr"], b["col"], b["val"]).CreateDeviceMatrix()
    K, n = g["X"].shape
    mx = la.MultiVector(A.CreateColVector(), K)
    my = la.MultiVector(A.CreateColVector(), K)
    for k in range(K):
        mx[k].FV().NumPy()[:] = g["X"][k]
    my[:] = A * mx
    for k in range(K):
        assert _rel(my[k].NumPy(), g["Y_mult"][k]) <= 1e-12
        y1 = A.CreateColVector()
        A.Mult(mx[k], y1)
        assert np.array_equal(my[k].NumPy(), y1.NumPy())      # same summation order as the single-vector kernel
    # alpha and accumulation
    al = np.arange(1, K + 1) * 0.5
    ys = [la.BaseVector(np.full(n, 2.0)) for _ in range(K)]
    A.MultAddMulti(al, mx.vecs, ys)
    for k in range(K):
        assert _rel(ys[k].NumPy(), 2.0 + al[k] * g["Y_mult"][k]) <= 1e-12
    with pytest.raises(la.NgsbError):
        A.MultAddMulti(al, mx.vecs, mx.vecs)                  # aliasing


def test_composite_operators_match_reference():
    import ngsolve_b200.la as la
    g = np.load(os.path.join(GOLD, "next_operators.npz"))
    A = la.SparseMatrix(g["a_rowptr"], g["a_col"], g["a_val"]).CreateDeviceMatrix()
    B = la.SparseMatrix(g["b_rowptr"], g["b_col"], g["b_val"]).CreateDeviceMatrix()
    x = la.BaseVector(g["x"])
    y = A.CreateColVector()
    y.data = (A + 2 * B) * x
    assert _rel(y.NumPy(), g["sum"]) <= 1e-12
    y.data = (A @ B) * x
    assert _rel(y.NumPy(), g["prod"]) <= 1e-12
    y.data = A.T * x
    assert _rel(y.NumPy(), g["trans"]) <= 1e-12
    y.data = (3 * A) * x
    assert _rel(y.NumPy(), g["scaled"]) <= 1e-12
    y.data = (A - B) * x + 0.5 * x
    assert _rel(y.NumPy(), g["expr"]) <= 1e-12
    mask = la.BitArray(g["mask"])
    y.data = la.Projector(mask, True) * x
    assert np.array_equal(y.NumPy(), g["proj_range"])
    y.data = la.Projector(mask, False) * x
    assert np.array_equal(y.NumPy(), g["proj_kernel"])
    y.data = (la.IdentityMatrix(len(g["x"])) - la.Projector(mask, True) @ A) * x
    assert _rel(y.NumPy(), g["id_minus_pa"]) <= 1e-12
    # host operands recurse in CreateDeviceMatrix like the reference's composite operators
    hA = la.SparseMatrix(g["a_rowptr"], g["a_col"], g["a_val"])
    hB = la.SparseMatrix(g["b_rowptr"], g["b_col"], g["b_val"])
    dev = (hA + 2 * hB).CreateDeviceMatrix()
    y.data = dev * x
    assert _rel(y.NumPy(), g["sum"]) <= 1e-12
    with pytest.raises(la.NgsbError):
        A @ la.IdentityMatrix(3)
