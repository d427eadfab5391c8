"""GPU: option csr_keep = 0 -- the uploaded column / value arrays are released once the SELL copy exists and come back,
bit-identical, from the SELL copy when something asks for them (csrc/csrview.cu); the Jacobi constructor reads the diagonal
out of the SELL copy.  Checked against the reference fixtures for all entry kinds, with and without the internal reordering,
and on rows longer than the slice cap (overflow part)."""
import numpy as np
import pytest

from conftest import kind_of, load_golden, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def la():
    import ngsolve_b200.la as la
    la.default_context()
    return la


@pytest.fixture()
def released(la):
    ctx = la.default_context()
    ctx.set_option("csr_keep", 0)
    yield ctx
    ctx.set_option("csr_keep", -1)
    ctx.set_option("reorder", -1)
    ctx.set_option("sell_cap", 0)


@pytest.mark.parametrize("reorder", [0, 1])
@pytest.mark.parametrize("name", ["poisson_h1p3", "elasticity_h1p4_dim3", "maxwell_hcurlp2", "helmholtz_h1p4_complex"])
def test_csr_comes_back_bit_identical(la, released, name, reorder):
    g = load_golden(name)
    k = kind_of(g)
    es = 3 if k == 3 else 1
    released.set_option("reorder", reorder)
    A = la.SparseMatrix(g["rowptr"], g["col"], g["val"], entrysize=es)
    dev = A.CreateDeviceMatrix()
    csr_b, sell_b, resident = dev.Memory()
    assert not resident and csr_b < 16 * (dev.height + 1) + 8 * dev.height + 64       # row pointers (+ permutation tables) only
    assert dev.ReorderInfo()[0] == bool(reorder)
    # products and the Jacobi constructor never need the arrays back
    x = la.BaseVector(np.asarray(g["x"]), entrysize=es)
    y = dev.CreateColVector()
    dev.Mult(x, y)
    assert relerr(y.NumPy().reshape(-1), g["y_mult"]) <= 1e-12
    jac = dev.CreateSmoother(la.BitArray(g["freebits"]))
    assert not dev.Memory()[2]
    released.set_option("csr_keep", 1)
    keep = A.CreateDeviceMatrix()
    jac_keep = keep.CreateSmoother(la.BitArray(g["freebits"]))
    released.set_option("csr_keep", 0)
    assert keep.Memory()[2]
    yj, yk = dev.CreateColVector(), dev.CreateColVector()
    jac.Mult(x, yj)
    jac_keep.Mult(x, yk)
    assert np.array_equal(yj.NumPy(), yk.NumPy())                        # same diagonal, bit for bit
    assert relerr(yj.NumPy().reshape(-1), g["jac_mult"]) <= 1e-13
    # CSR(): rebuilt from the SELL copy
    val, col, rowptr = dev.CSR()
    assert dev.Memory()[2]
    assert np.array_equal(rowptr, g["rowptr"]) and np.array_equal(col, g["col"])
    assert np.array_equal(val.reshape(-1), np.asarray(g["val"]).reshape(-1))


def test_rows_longer_than_the_slice_cap(la, released):
    """overflow part of long rows (slice cap forced to 8) in the rebuild and in the diagonal"""
    rng = np.random.default_rng(17)
    n = 700
    rows = [np.unique(np.concatenate([[i], rng.choice(n, rng.integers(1, 40))])) for i in range(n)]
    rows[5] = np.arange(n)                                    # one dense row
    rowptr = np.zeros(n + 1, dtype=np.uint64)
    rowptr[1:] = np.cumsum([len(r) for r in rows])
    col = np.concatenate(rows).astype(np.int32)
    val = rng.random(len(col)) + 0.1
    for reorder in (0, 1):
        released.set_option("reorder", reorder)
        released.set_option("sell_cap", 8)
        dev = la.SparseMatrix(rowptr, col, val).CreateDeviceMatrix()
        assert not dev.Memory()[2] and dev.Layout()[1] > 0          # released, and rows with an overflow part exist
        jac = dev.CreateSmoother(None)
        x = la.BaseVector(np.ones(n))
        y = dev.CreateColVector()
        jac.Mult(x, y)
        diag = np.array([val[int(rowptr[i]) + int(np.searchsorted(rows[i], i))] for i in range(n)])
        assert np.array_equal(y.NumPy(), 1.0 / diag)
        v2, c2, r2 = dev.CSR()
        assert np.array_equal(r2, rowptr) and np.array_equal(c2, col) and np.array_equal(v2, val)


def test_transpose_and_reorder_after_release(la, released):
    g = load_golden("poisson_h1p3")
    dev = la.SparseMatrix(g["rowptr"], g["col"], g["val"]).CreateDeviceMatrix()
    assert not dev.Memory()[2]
    perm = np.random.default_rng(3).permutation(dev.height).astype(np.uint64)
    from oracle import pyoracle as orc
    ref = orc.Csr(g["rowptr"], g["col"], g["val"], 0).reorder(perm)
    val, col, rowptr = dev.Reorder(perm).CSR()
    assert np.array_equal(rowptr, ref.rowptr) and np.array_equal(col, ref.col) and np.array_equal(val, ref.val)


@pytest.mark.parametrize("name,kind", [("real", 0), ("complex", 1), ("block3", 3)])
def test_archive_wire_format_equals_the_reference(la, name, kind):
    """SparseMatrix<TM>::DoArchive (linalg/sparsematrix_impl.hpp:443-452): the library writes the bytes the reference's
    BinaryOutArchive holds for the same matrix, and reads what the reference wrote (tests/golden/make_golden_archive.py)"""
    g = load_golden("archive_wire")
    rowptr, col, val, ref = g[name + "_rowptr"], g[name + "_col"], g[name + "_val"], g[name + "_bytes"]
    dev = la.SparseMatrix(rowptr, col, val, entrysize=3 if kind == 3 else 1).CreateDeviceMatrix()
    assert np.array_equal(dev.Archive(), ref)
    back = la.DevSparseMatrix.FromArchive(ref.tobytes(), kind=kind)
    v2, c2, r2 = back.CSR()
    assert np.array_equal(r2, rowptr) and np.array_equal(c2, col) and np.array_equal(v2.reshape(-1), np.asarray(val).reshape(-1))
    with pytest.raises(la.NgsbError):
        la.DevSparseMatrix.FromArchive(ref.tobytes()[:-8], kind=kind)
