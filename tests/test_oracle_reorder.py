"""CPU: the oracle's Reorder restatement against the REFERENCE's SparseMatrix::Reorder (golden produced through
tests/golden/ref_helpers.cpp), and the serial Cuthill-McKee specification (orc_rcm) -- pinned by the committed permutation,
checked for the properties the device library relies on, and compared with scipy's reverse_cuthill_mckee."""
import numpy as np
import scipy.sparse as sp
from scipy.sparse.csgraph import reverse_cuthill_mckee

from conftest import load_golden
from oracle import pyoracle as orc


def _bandwidth(rowptr, col):
    n = len(rowptr) - 1
    return np.abs(np.repeat(np.arange(n), np.diff(rowptr.astype(np.int64))) - col)


def test_oracle_reorder_equals_reference_reorder():
    g = load_golden("reorder_netgen_h1p3")
    B = orc.Csr(g["rowptr"], g["col"], g["val"], 0).reorder(g["perm"])
    assert np.array_equal(B.rowptr, g["r_rowptr"])        # bit-exact: pattern ...
    assert np.array_equal(B.col, g["r_col"])
    assert np.array_equal(B.val, g["r_val"])              # ... and values


def test_rcm_pinned_and_is_a_permutation():
    g = load_golden("reorder_netgen_h1p3")
    A = orc.Csr(g["rowptr"], g["col"], g["val"], 0)
    perm = A.rcm()
    assert np.array_equal(perm, g["perm"])
    assert np.array_equal(np.sort(perm), np.arange(A.n, dtype=np.uint64))


def test_rcm_reduces_the_bandwidth_like_scipy():
    g = load_golden("reorder_netgen_h1p3")
    n = len(g["rowptr"]) - 1
    nat = _bandwidth(g["rowptr"], g["col"]).mean()
    ours = _bandwidth(g["r_rowptr"], g["r_col"]).mean()
    M = sp.csr_matrix((np.ones(len(g["col"]), dtype=np.int8), g["col"], g["rowptr"].astype(np.int64)), shape=(n, n))
    p = reverse_cuthill_mckee(M, symmetric_mode=True)
    S = M[p][:, p]
    S.sort_indices()
    theirs = _bandwidth(S.indptr, S.indices).mean()
    assert ours < 0.3 * nat                  # netgen numbering: mean |i-j| ~ n/3
    assert ours < 1.15 * theirs


def _multi_component(rng, blocks, isolated):
    """block-diagonal pattern of random symmetric blocks + isolated dofs + one dof with an empty row"""
    rows = []
    off = 0
    for b in blocks:
        M = sp.random(b, b, density=min(1.0, 6.0 / b), random_state=rng, format="csr")
        M = ((M + M.T) != 0).astype(np.int8) + sp.eye(b, dtype=np.int8, format="csr")
        M = sp.csr_matrix(M)
        M.sort_indices()
        for i in range(b):
            rows.append(M.indices[M.indptr[i]:M.indptr[i + 1]] + off)
        off += b
    for _ in range(isolated):
        rows.append(np.array([off]))
        off += 1
    rows.append(np.zeros(0, dtype=np.int64))       # empty row
    off += 1
    rowptr = np.zeros(off + 1, dtype=np.uint64)
    rowptr[1:] = np.cumsum([len(r) for r in rows])
    col = np.concatenate(rows).astype(np.int32)
    return rowptr, col


def test_rcm_components_and_cutoff():
    rng = np.random.default_rng(11)
    rowptr, col = _multi_component(rng, [40, 17, 300, 5], isolated=3)
    n = len(rowptr) - 1
    A = orc.Csr(rowptr, col, np.ones(len(col)), 0)
    full = A.rcm(max_components=64)
    assert np.array_equal(np.sort(full), np.arange(n, dtype=np.uint64))
    # components are laid out one after the other (reversed as a whole): the first block's dofs come last
    order = full[::-1].astype(np.int64)
    assert set(order[:40]) == set(range(40))
    assert set(order[40:57]) == set(range(40, 57))
    # cutoff after two components: the remaining dofs in ascending order
    cut = A.rcm(max_components=2)[::-1].astype(np.int64)
    assert np.array_equal(cut[:57], order[:57])
    assert np.array_equal(cut[57:], np.arange(57, n))


def test_archive_wire_layout_of_the_reference():
    """SparseMatrix<TM>::DoArchive into a BinaryOutArchive (linalg/sparsematrix_impl.hpp:443-452), restated: size_t size, width,
    nze; Array<size_t> firsti; Array<int> colnr; Array<TM> data -- each array as size_t count + raw elements"""
    import struct
    g = load_golden("archive_wire")
    for name in ("real", "complex", "block3"):
        rp, col, val = g[name + "_rowptr"], g[name + "_col"], g[name + "_val"]
        n, nnz = len(rp) - 1, len(col)
        exp = (struct.pack("<QQQ", n, n, nnz) + struct.pack("<Q", n + 1) + rp.tobytes() + struct.pack("<Q", nnz) + col.tobytes()
               + struct.pack("<Q", nnz) + np.ascontiguousarray(val).view(np.float64).tobytes())
        assert exp == g[name + "_bytes"].tobytes()
