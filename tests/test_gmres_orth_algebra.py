"""CPU: the algebra behind option gmres_orth = 1 (csrc/krylov.cu, gmres_dots_kernel / gmres_orth_finish_kernel / gmres_project_kernel).

The reference orthogonalises w = C A v_j by modified Gram-Schmidt (linalg/cg.cpp:927-932):  h_i = <v_i, w>;  w -= h_i v_i,  i = 0..j,
with the bilinear product of GMRESSolver<IPTYPE> (no conjugation).  Written out, h_i = <v_i, w> - sum_{k<i} <v_i, v_k> h_k, i.e.
(I + L) h = V^T w with L the strict lower part of V^T V: one batched reduction (V^T w and the new row of L), a unit triangular solve
and one projection pass.  This file restates that in numpy and checks, on the reference-generated fixtures, that the whole GMRES
built on it stops at the reference's step count with the reference's solution -- the claim the GPU tests then check on the device."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import load_golden


def gmres(A, cinv, f, prec, maxsteps, batched):
    """GMRESSolver::Mult (cg.cpp:854-1022); batched: the orthogonalisation as (I + L) h = V^T w, else the reference's loop.
    Also returns the largest |<v_k, w>| / |w| seen after a projection (loss of orthogonality)."""
    cplx = np.iscomplexobj(A.data) or np.iscomplexobj(f)
    dt = np.complex128 if cplx else np.float64
    sqrt = (lambda z: np.sqrt(z + 0j)) if cplx else np.sqrt
    n, ms = len(f), maxsteps
    x = np.zeros(n, dtype=dt)
    r = cinv * f.astype(dt)
    norm = np.sqrt(np.sum(np.abs(r) ** 2))
    V = [r / sqrt(np.sum(r * r))]
    H = np.zeros((ms + 1, ms), dtype=dt)
    L = np.zeros((ms + 1, ms + 1), dtype=dt)
    gam, ci, si = (np.zeros(ms + 2, dtype=dt) for _ in range(3))
    gam[0] = norm
    err = prec * abs(norm)
    j = -1
    lost = 0.0
    while True:
        go = (j < ms - 2) and (norm > err)
        j += 1
        if not go:
            break
        w = cinv * (A @ V[j])
        Vm = np.array(V[:j + 1])
        if batched:
            rhs = Vm @ w                                  # pass 1: all <v_k, w> ...
            if j > 0:
                L[j, :j] = Vm[:j] @ V[j]                  # ... and the new row of L, from the same read of V
            h = np.zeros(j + 1, dtype=dt)
            for k in range(j + 1):                        # forward substitution, ascending i like the finish kernel
                m = rhs[k]
                for i in range(k):
                    m -= L[k, i] * h[i]
                h[k] = m
            H[:j + 1, j] = h
            for k in range(j + 1):                        # pass 2
                w = w - h[k] * V[k]
        else:
            for i in range(j + 1):
                H[i, j] = np.sum(V[i] * w)
                w = w - H[i, j] * V[i]
        lost = max(lost, float(np.max(np.abs(Vm @ w))) / float(np.sqrt(np.sum(np.abs(w) ** 2))))
        H[j + 1, j] = sqrt(np.sum(w * w))
        V.append(w / H[j + 1, j])
        for i in range(j):
            hi, hip = H[i, j], H[i + 1, j]
            H[i, j] = ci[i + 1] * hi + si[i + 1] * hip
            H[i + 1, j] = si[i + 1] * hi - ci[i + 1] * hip
        beta = sqrt(H[j, j] ** 2 + H[j + 1, j] ** 2)
        si[j + 1], ci[j + 1] = H[j + 1, j] / beta, H[j, j] / beta
        H[j, j] = beta
        gam[j + 1] = si[j + 1] * gam[j]
        gam[j] = ci[j + 1] * gam[j]
        norm = abs(gam[j])
    jf = j - 1
    y = np.zeros(ms + 2, dtype=dt)
    for i in range(jf, -1, -1):
        y[i] = (gam[i] - np.sum(H[i, i + 1:jf + 1] * y[i + 1:jf + 1])) / H[i, i]
    for i in range(jf + 1):
        x += y[i] * V[i]
    return x, jf, lost


@pytest.mark.parametrize("name", ["poisson_h1p3", "helmholtz_h1p4_complex", "shifted_laplace_complex", "square_h1p4_testsolvers"])
def test_batched_orthogonalisation_reproduces_the_reference_gmres(name):
    g = load_golden(name)
    n = len(g["rowptr"]) - 1
    A = sp.csr_matrix((g["val"], g["col"], g["rowptr"].astype(np.int64)), shape=(n, n))
    free = np.unpackbits(g["freebits"], bitorder="little")[:n].astype(bool)
    d = A.diagonal()
    cinv = np.where(free, 1.0 / np.where(d == 0, 1, d), 0)
    xb, sb, lost_b = gmres(A, cinv, g["f"], float(g["gmres_prec"]), int(g["gmres_maxsteps"]), True)
    xm, sm, lost_m = gmres(A, cinv, g["f"], float(g["gmres_prec"]), int(g["gmres_maxsteps"]), False)
    assert sb == sm == int(g["gmres_steps"])
    assert np.linalg.norm(xb - g["gmres_u"]) <= 1e-9 * np.linalg.norm(g["gmres_u"])
    assert np.linalg.norm(xb - xm) <= 1e-9 * np.linalg.norm(xm)
    assert lost_b <= 10.0 * lost_m + 1e-13        # w leaves the projection as orthogonal to the v_k as it does after MGS
