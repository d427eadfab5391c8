"""Second opinion on the CPU oracle (the checker of the GPU tests): its products, transposed products, symmetric-storage
product, dots and Jacobi against scipy/numpy on random ragged matrices -- empty rows, one very long row, rectangular
shapes, all three entry kinds.  The oracle sums a row in storage order with one accumulator like the reference
(linalg/sparsematrix.hpp:625-632); scipy may order differently, hence 1e-13 relative instead of bit equality."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import pyoracle as orc

REAL, COMPLEX, BLOCK3 = orc.KIND_REAL, orc.KIND_COMPLEX, orc.KIND_BLOCK3


def ragged_csr(rng, h, w, kind, long_row=True):
    lens = rng.integers(0, 9, size=h)
    lens[rng.integers(0, h, size=max(1, h // 7))] = 0            # empty rows
    if long_row and w >= 64:
        lens[h // 2] = min(w, 300)                               # longer than a SELL slice is wide
    rowptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    col = np.concatenate([np.sort(rng.choice(w, size=int(l), replace=False)) for l in lens] + [np.zeros(0, dtype=np.int64)]).astype(np.int32)
    nnz = int(rowptr[-1])
    if kind == COMPLEX:
        val = rng.standard_normal(nnz) + 1j * rng.standard_normal(nnz)
    elif kind == BLOCK3:
        val = rng.standard_normal(nnz * 9)
    else:
        val = rng.standard_normal(nnz)
    return rowptr, col, val


def to_scipy(rowptr, col, val, h, w, kind):
    if kind == BLOCK3:
        return sp.bsr_matrix((val.reshape(-1, 3, 3), col, rowptr.astype(np.int64)), shape=(3 * h, 3 * w)).tocsr()
    return sp.csr_matrix((val, col, rowptr.astype(np.int64)), shape=(h, w))


def vec(rng, n, kind):
    if kind == COMPLEX:
        return rng.standard_normal(n) + 1j * rng.standard_normal(n)
    return rng.standard_normal(n * (3 if kind == BLOCK3 else 1))


@pytest.mark.parametrize("kind", [REAL, COMPLEX, BLOCK3], ids=["real", "complex", "block3"])
@pytest.mark.parametrize("shape", [(1, 1), (37, 37), (500, 500), (64, 200), (200, 64)])
def test_products_against_scipy(kind, shape):
    h, w = shape
    rng = np.random.default_rng(1000 * h + w + kind)
    rowptr, col, val = ragged_csr(rng, h, w, kind)
    S = to_scipy(rowptr, col, val, h, w, kind)
    A = orc.Csr(rowptr, col, val, kind)
    x, y0 = vec(rng, w, kind), vec(rng, h, kind)
    ref = S @ x
    scale = max(1.0, np.max(np.abs(ref)))
    if h == w:
        assert np.max(np.abs(A.mult(x) - ref)) <= 1e-13 * scale           # Mult sizes its result by n (square use in the tests)
    y = y0.copy()
    A.multadd(-0.75, x, y)
    assert np.max(np.abs(y - (y0 - 0.75 * ref))) <= 1e-13 * scale
    if kind == COMPLEX:
        y = y0.copy()
        A.multadd(0.5 - 2.0j, x, y)
        assert np.max(np.abs(y - (y0 + (0.5 - 2.0j) * ref))) <= 1e-13 * scale * 3
    # MultTransAdd: y += s A^T x (plain transpose, no conjugation: linalg/sparsematrix_impl.hpp:344-375)
    xt, yt0 = vec(rng, h, kind), vec(rng, w, kind)
    yt = yt0.copy()
    orc.multtransadd(A, 1.25, xt, yt)
    reft = yt0 + 1.25 * (S.T @ xt)
    assert np.max(np.abs(yt - reft)) <= 1e-13 * max(1.0, np.max(np.abs(reft)))


def test_symmetric_storage_against_scipy():
    rng = np.random.default_rng(3)
    n = 300
    M = sp.random(n, n, density=0.03, random_state=5, format="csr")
    M = (M + M.T + sp.diags(rng.random(n) + 1.0)).tocsr()
    L = sp.tril(M).tocsr()
    L.sort_indices()
    x, y0 = rng.standard_normal(n), rng.standard_normal(n)
    y = y0.copy()
    orc.sym_multadd(L.indptr.astype(np.uint64), L.indices.astype(np.int32), L.data, 2.0, x, y)
    ref = y0 + 2.0 * (M @ x)
    assert np.max(np.abs(y - ref)) <= 1e-13 * np.max(np.abs(ref))


def test_dots_norms_and_jacobi_against_numpy():
    rng = np.random.default_rng(4)
    for n in (1, 15, 16, 17, 1000, 4099):                       # around the 16-chunk split of S_BaseVector::InnerProduct
        x, y = rng.standard_normal(n), rng.standard_normal(n)
        assert abs(orc.inner(x, y) - float(x @ y)) <= 1e-13 * max(1.0, float(np.abs(x) @ np.abs(y)))
        assert abs(orc.l2norm(x) - np.linalg.norm(x)) <= 1e-13 * np.linalg.norm(x)
        zx, zy = x + 1j * rng.standard_normal(n), y - 1j * rng.standard_normal(n)
        b = max(1.0, float(np.abs(zx) @ np.abs(zy)))
        assert abs(orc.inner(zx, zy) - np.sum(zx * zy)) <= 1e-13 * b                       # bilinear
        assert abs(orc.inner(zx, zy, conjugate=True) - np.sum(zx * np.conj(zy))) <= 1e-13 * b   # conjugation on the argument
        assert abs(orc.l2norm(zx) - np.linalg.norm(zx)) <= 1e-13 * np.linalg.norm(zx)
    # JacobiPrecond: inverse diagonal on the free dofs, zero elsewhere (linalg/jacobi.cpp:39-68)
    n = 203
    M = (sp.random(n, n, density=0.05, random_state=9) + sp.diags(rng.random(n) + 2.0)).tocsr()
    M.sort_indices()
    free = rng.random(n) < 0.8
    bits = np.packbits(free, bitorder="little")
    A = orc.Csr(M.indptr.astype(np.uint64), M.indices.astype(np.int32), M.data, REAL)
    J = orc.Jacobi(A, bits)
    x = rng.standard_normal(n)
    ref = np.where(free, x / M.diagonal(), 0.0)
    assert np.max(np.abs(J.mult(x) - ref)) <= 1e-15 * np.max(np.abs(ref))
