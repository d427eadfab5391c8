"""CPU: the synthetic FE system generator (include/ngsb200_workloads.h, host loop of the library)
against an independent brute-force assembly (numpy, numerical quadrature, dictionary numbering)
and against properties of the continuous problem.  The device assembly is compared bit for bit
with the host loop in tests/test_gpu_workloads.py."""
import itertools

import numpy as np
import pytest

from ngsolve_b200 import workloads as W

PERMS = list(itertools.permutations(range(3)))


def _duffy_rule(n=6):
    """conical product rule on the unit tetrahedron (weights sum to 1/6), exact to degree 2n-3"""
    x, w = np.polynomial.legendre.leggauss(n)
    x, w = 0.5 * (x + 1), 0.5 * w
    pts, wts = [], []
    for a, wa in zip(x, w):
        for b, wb in zip(x, w):
            for c, wc in zip(x, w):
                l1 = a
                l2 = b * (1 - a)
                l3 = c * (1 - a) * (1 - b)
                pts.append((1 - l1 - l2 - l3, l1, l2, l3))
                wts.append(wa * wb * wc * (1 - a) ** 2 * (1 - b))
    return np.array(pts), np.array(wts)


def _basis(order, lam):
    """values and d/dlambda of the hierarchical basis of one tet at barycentric point lam;
    returns [(entity key as tuple of local vertices, sub, value, dlam[4])]"""
    out = []
    for v in range(4):
        d = np.zeros(4)
        d[v] = 1
        out.append(((v,), 0, lam[v], d))
    for a, b in itertools.combinations(range(4), 2):
        for k in range(order - 1):
            t = lam[b] - lam[a]
            val = lam[a] * lam[b] * t ** k
            d = np.zeros(4)
            dk = k * t ** (k - 1) if k > 0 else 0.0
            d[a] = lam[b] * t ** k - lam[a] * lam[b] * dk
            d[b] = lam[a] * t ** k + lam[a] * lam[b] * dk
            out.append(((a, b), k, val, d))
    for a, b, c in itertools.combinations(range(4), 3):
        sub = 0
        for j in range(order - 2):
            for i in range(order - 2 - j):
                s, t = lam[b] - lam[a], lam[c] - lam[a]
                P = lam[a] * lam[b] * lam[c]
                Q, R = s ** i, t ** j
                dQ = i * s ** (i - 1) if i > 0 else 0.0
                dR = j * t ** (j - 1) if j > 0 else 0.0
                d = np.zeros(4)
                d[a] = lam[b] * lam[c] * Q * R - P * dQ * R - P * Q * dR
                d[b] = lam[a] * lam[c] * Q * R + P * dQ * R
                d[c] = lam[a] * lam[b] * Q * R + P * Q * dR
                out.append(((a, b, c), sub, P * Q * R, d))
                sub += 1
    if order >= 4:
        P = lam[0] * lam[1] * lam[2] * lam[3]
        d = np.array([P / lam[q] if lam[q] != 0 else 0.0 for q in range(4)])
        out.append(((0, 1, 2, 3), 0, P, d))
    return out


def brute_force(n, order, h, mass=0.0, elasticity=None):
    """dense matrix + load vector, arbitrary (dictionary) numbering"""
    pts, wts = _duffy_rule()
    numbering = {}
    rows, cols, vals = [], [], []
    load = {}
    nx, ny, nz = n
    vid = lambda p: p[0] + (nx + 1) * (p[1] + (ny + 1) * p[2])      # noqa: E731
    for cz, cy, cx in itertools.product(range(nz), range(ny), range(nx)):
        for pi in PERMS:
            verts = [np.array([cx, cy, cz])]
            for ax in pi:
                e = np.zeros(3, dtype=int)
                e[ax] = 1
                verts.append(verts[-1] + e)
            gv = [vid(v) for v in verts]
            assert gv == sorted(gv)
            X = np.array(verts, dtype=float) * h
            T = np.vstack([np.ones(4), X.T])
            glam = np.linalg.inv(T)[:, 1:]                      # grad lambda_k (rows)
            vol = abs(np.linalg.det(T)) / 6.0
            ndl = len(_basis(order, pts[0]))
            dim = 3 if elasticity else 1
            Ke = np.zeros((ndl * dim, ndl * dim))
            be = np.zeros(ndl)
            keys = None
            for lam, w in zip(pts, wts):
                B = _basis(order, lam)
                keys = [(tuple(gv[q] for q in ent), sub) for ent, sub, _, _ in B]
                phi = np.array([b[2] for b in B])
                grad = np.array([b[3] @ glam for b in B])       # ndl x 3
                ww = w * 6.0 * vol
                be += ww * phi
                if elasticity:
                    lamc, mu = elasticity
                    for a in range(3):
                        for b in range(3):
                            blk = lamc * np.outer(grad[:, a], grad[:, b]) + mu * np.outer(grad[:, b], grad[:, a])
                            if a == b:
                                blk = blk + mu * grad @ grad.T
                            Ke[a::3, b::3] += ww * blk
                else:
                    Ke += ww * (grad @ grad.T + mass * np.outer(phi, phi))
            idx = [numbering.setdefault(k, len(numbering)) for k in keys]
            for i, gi in enumerate(idx):
                load[gi] = load.get(gi, 0.0) + be[i]
            full = np.repeat(np.array(idx) * dim, dim) + np.tile(np.arange(dim), len(idx))
            rows.append(np.repeat(full, len(full)))
            cols.append(np.tile(full, len(full)))
            vals.append(Ke.reshape(-1))
    N = len(numbering) * (3 if elasticity else 1)
    A = np.zeros((N, N))
    np.add.at(A, (np.concatenate(rows), np.concatenate(cols)), np.concatenate(vals))
    graph = np.zeros((N, N), dtype=bool)                        # dofs sharing an element (MatrixGraph)
    graph[np.concatenate(rows), np.concatenate(cols)] = True
    return A, np.array([load[i] for i in range(len(numbering))]), graph


def dense_of(box):
    rp, col, val, rhs = box.host_csr()
    n = box.ndof
    if box.kind == W.BLOCK3:
        A = np.zeros((3 * n, 3 * n))
        v = val.reshape(-1, 3, 3)
        for i in range(n):
            for j in range(int(rp[i]), int(rp[i + 1])):
                A[3 * i:3 * i + 3, 3 * col[j]:3 * col[j] + 3] = v[j]
    else:
        A = np.zeros((n, n), dtype=val.dtype)
        for i in range(n):
            A[i, col[int(rp[i]):int(rp[i + 1])]] = val[int(rp[i]):int(rp[i + 1])]
    return A, rp, col, rhs


@pytest.mark.parametrize("order,n", [(1, (3, 2, 2)), (2, (2, 2, 2)), (3, (2, 2, 1)), (3, (2, 2, 2)), (4, (2, 1, 1))])
def test_matches_brute_force_up_to_numbering(order, n):
    h = 0.37
    box = W.FemBox(n, order=order, h=h, mass=0.5)
    A, rp, col, rhs = dense_of(box)
    B, load, graph = brute_force(n, order, h, mass=0.5)
    assert A.shape == B.shape
    assert np.allclose(A, A.T, rtol=0, atol=1e-13 * np.abs(A).max())
    for i in range(box.ndof):                                   # ascending, unique columns per row
        assert np.all(np.diff(col[int(rp[i]):int(rp[i + 1])]) > 0)
    assert np.count_nonzero(B) <= int(rp[-1])                   # stored pattern covers the coupling graph
    ea, eb = np.linalg.eigvalsh(A), np.linalg.eigvalsh(B)
    assert np.allclose(ea, eb, rtol=1e-10, atol=1e-12 * abs(eb).max())
    assert np.allclose(np.sort(np.diag(A)), np.sort(np.diag(B)), rtol=1e-11)
    assert np.allclose(np.sort(rhs), np.sort(load), rtol=1e-11, atol=1e-15)
    # the stored pattern is exactly the element coupling graph (entries that cancel to 0.0 stay, like NGSolve's MatrixGraph)
    assert sorted(np.diff(rp).tolist()) == sorted(np.count_nonzero(graph, axis=1).tolist())


def test_elasticity_blocks_match_brute_force():
    n, h = (2, 1, 1), 0.5
    box = W.FemBox(n, order=2, kind=W.BLOCK3, h=h, lame=(1.3, 0.7))
    A, rp, col, rhs = dense_of(box)
    B, _, _ = brute_force(n, 2, h, elasticity=(1.3, 0.7))
    assert np.allclose(A, A.T, atol=1e-13)
    assert np.allclose(np.linalg.eigvalsh(A), np.linalg.eigvalsh(B), rtol=1e-10, atol=1e-12)
    # rigid body motions are in the kernel: translation = vertex dofs 1, higher-order dofs 0
    gi, surf, free = box.dof_info()
    nv = 3 * 2 * 2
    t = np.zeros((box.ndof, 3))
    t[:nv, 1] = 1.0
    assert np.abs(A @ t.reshape(-1)).max() < 1e-12


def test_continuous_problem_properties():
    m = 4
    box = W.FemBox(m, order=3)                                  # unit cube, h = 1/4
    A, rp, col, rhs = dense_of(box)
    nvert = (m + 1) ** 3
    c = np.zeros(box.ndof)
    c[:nvert] = 1.0
    assert np.abs(A @ c).max() < 1e-12                          # constants are in the kernel
    # u = x: vertex dofs = coordinate; energy = |grad x|^2 * volume = 1
    xs = np.tile(np.arange(m + 1) / m, (m + 1) ** 2)
    u = np.zeros(box.ndof)
    u[:nvert] = xs
    assert u @ A @ u == pytest.approx(1.0, rel=1e-12)
    gi, surf, free = box.dof_info()
    assert np.abs((A @ u)[free.astype(bool)]).max() < 1e-12     # discrete harmonic in the interior
    assert np.array_equal(gi, np.arange(box.ndof, dtype=np.uint64))          # single box: local == global
    assert np.array_equal(surf.astype(bool), ~free.astype(bool))
    # NGSolve H1 order 3: ndof = nv + 2 ne + nf;  row length 20 is the minimum (one tet)
    assert box.ndof == (3 * m + 1) ** 3 and np.diff(rp).min() == 20
    # Poisson with the masked Jacobi is SPD on the free dofs
    F = free.astype(bool)
    assert np.linalg.eigvalsh(A[np.ix_(F, F)]).min() > 0


@pytest.mark.parametrize("order,kind", [(3, W.REAL), (2, W.BLOCK3), (2, W.COMPLEX)])
def test_subdomains_sum_to_global(order, kind):
    """The reference's MPI split: local matrices of local elements, interface dofs duplicated; summing
    them through the global numbering gives the global matrix, and the exchange tables pair up."""
    G = (3, 2, 4)
    kw = dict(order=order, kind=kind, mass=(0.3 - 0.2j) if kind == W.COMPLEX else 0.0, lame=(1.0, 0.5))
    glob = W.FemBox(G, **kw)
    Ag, _, _, fg = dense_of(glob)
    es = 3 if kind == W.BLOCK3 else 1
    parts = [W.FemBox(n, offset=o, global_n=G, **kw) for n, o in W.slab_partition(G, 3)]
    S = np.zeros_like(Ag)
    fs = np.zeros_like(fg)
    for b in parts:
        assert b.global_ndof == glob.ndof
        A, _, _, f = dense_of(b)
        gi = b.dof_info()[0].astype(np.int64)
        full = (np.repeat(gi * es, es) + np.tile(np.arange(es), len(gi)))
        S[np.ix_(full, full)] += A
        np.add.at(fs, full, f)
    assert np.allclose(S, Ag, rtol=1e-12, atol=1e-13)
    assert np.allclose(fs, fg, rtol=1e-12, atol=1e-15)
    # exchange tables: ascending, symmetric between neighbours, master = lowest rank
    from oracle import pyoracle as orc
    tables = [W.exchange_tables(parts, r) for r in range(3)]
    for r, (first, dofs) in enumerate(tables):
        for q in range(3):
            mine = dofs[int(first[q]):int(first[q + 1])]
            ofirst, odofs = tables[q]
            theirs = odofs[int(ofirst[r]):int(ofirst[r + 1])]
            assert len(mine) == len(theirs) and (abs(r - q) == 1) == (len(mine) > 0)
            assert np.all(np.diff(mine) > 0)
            assert np.array_equal(parts[r].dof_info()[0][mine], parts[q].dof_info()[0][theirs])
        # the same tables through the oracle's ParallelDofs restatement (dist_procs -> exchangedofs)
        dist = [[] for _ in range(parts[r].ndof)]
        for q in range(3):
            for dof in dofs[int(first[q]):int(first[q + 1])]:
                dist[dof].append(q)
        exch, master = orc.pardofs_build(3, r, dist)
        for q in range(3):
            assert np.array_equal(exch[q], dofs[int(first[q]):int(first[q + 1])])
