"""CPU: the oracle's restatement of the SURVEY.md 8(f) rows against fixtures produced by the reference itself
(tests/golden/make_golden_next.py): BlockJacobiPrecond, MultTransAdd / CreateTranspose, SparseMatrixSymmetric,
MultiVector products."""
import os

import numpy as np

from oracle import pyoracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    return np.max(np.abs(a - b)) / np.max(np.abs(b))


def test_blockjacobi_against_reference():
    g = np.load(os.path.join(GOLD, "next_blockjacobi.npz"))
    A = orc.Csr(g["rowptr"], g["col"], g["val"], 0)
    bf = g["bfirst"].astype(np.int64)
    blocks = [list(g["bdofs"][bf[b]:bf[b + 1]]) for b in range(len(bf) - 1)]
    assert max(len(b) for b in blocks) >= 100        # the reference inverts these with LAPACK: rounding-level difference only
    bj = orc.BlockJacobi(A, blocks)
    assert _rel(bj.mult(g["x"]), g["bj_mult"]) <= 1e-12
    assert _rel(bj.multadd(0.5, g["x"], g["y0"].copy()), g["bj_multadd_05"]) <= 1e-12
    assert _rel(bj.mult(g["x"], transpose=True), g["bj_multtrans"]) <= 1e-12


def test_multtransadd_against_reference():
    g = np.load(os.path.join(GOLD, "next_transpose.npz"))
    for tag, kind in (("d", 0), ("z", 1), ("b3", 3)):
        A = orc.Csr(g[tag + "_rowptr"], g[tag + "_col"], g[tag + "_val"], kind)
        y = orc.multtransadd(A, 0.75, g[tag + "_x"], g[tag + "_y0"].copy())
        assert _rel(y, g[tag + "_multtransadd_075"]) <= 1e-14
    # CreateTranspose: A^T x through the transposed CSR equals MultTransAdd
    T = orc.Csr(g["d_t_rowptr"], g["d_t_col"], g["d_t_val"], 0)
    y = T.multadd(0.75, g["d_x"], g["d_y0"].copy())
    assert _rel(y, g["d_multtransadd_075"]) <= 1e-13


def test_symmetric_storage_against_reference():
    g = np.load(os.path.join(GOLD, "next_symmetric.npz"))
    assert "Symmetric" in str(g["mat_type"])
    y = orc.sym_multadd(g["rowptr"], g["col"], g["val"], 1.0, g["x"], np.zeros_like(g["x"]))
    assert _rel(y, g["y_mult"]) <= 1e-14
    y = orc.sym_multadd(g["rowptr"], g["col"], g["val"], -1.5, g["x"], g["y0"].copy())
    assert _rel(y, g["y_multadd_m15"]) <= 1e-14


def test_multivector_against_reference():
    g = np.load(os.path.join(GOLD, "next_multivector.npz"))
    b = np.load(os.path.join(GOLD, "next_blockjacobi.npz"))
    A = orc.Csr(b["rowptr"], b["col"], b["val"], 0)
    for k in range(g["X"].shape[0]):
        assert _rel(A.mult(g["X"][k]), g["Y_mult"][k]) <= 1e-14
