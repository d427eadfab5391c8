"""GPU: the device assembly of the synthetic FE systems against the library's host loop (bit for
bit), and size-independent properties of the solve path on systems far beyond what the oracle can
check entry by entry (linearity, symmetry, constants in the kernel, CG against the oracle at a
mid size)."""
import numpy as np
import pytest

from oracle import pyoracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import ngsolve_b200.la as la
    from ngsolve_b200 import workloads as W
    la.default_context()
    return la, W


@pytest.mark.parametrize("order,kind,n", [(3, 0, (7, 5, 6)), (4, 0, (3, 4, 3)), (2, 3, (5, 5, 4)), (4, 3, (3, 3, 3)), (3, 1, (4, 6, 5)), (1, 0, (9, 9, 9))])
def test_device_assembly_equals_host_assembly(mods, order, kind, n):
    la, W = mods
    box = W.FemBox(n, order=order, kind=kind, mass=(0.4 - 0.3j) if kind == 1 else 0.25, lame=(1.2, 0.8))
    rp, col, val, rhs = box.host_csr()
    A, f = box.device_system()
    dval, dcol, drp = A.CSR()
    assert np.array_equal(drp, rp) and np.array_equal(dcol, col)
    assert np.array_equal(dval.view(np.float64), val.view(np.float64))
    assert np.array_equal(f.NumPy().reshape(-1).view(np.float64), rhs.view(np.float64))
    # and the SpMV of the generated system against the oracle
    rng = np.random.default_rng(1)
    es = 3 if kind == 3 else 1
    xs = rng.standard_normal(box.ndof * es) + (1j * rng.standard_normal(box.ndof * es) if kind == 1 else 0)
    y = (A * la.BaseVector(xs, entrysize=es)).Evaluate().NumPy().reshape(-1)
    yref = orc.Csr(rp, col, val, kind).mult(xs)
    assert np.max(np.abs(y - yref)) <= 1e-12 * np.max(np.abs(yref))


def test_cg_matches_oracle_on_generated_poisson(mods):
    """~0.2 M dofs: the oracle (reference recurrences) finishes in seconds; steps within +-2."""
    la, W = mods
    box = W.FemBox(19, order=3)
    rp, col, val, rhs = box.host_csr()
    free = box.freedofs()
    A, f = box.device_system()
    inv = la.CGSolver(A, A.CreateSmoother(free), precision=1e-8, maxsteps=5000)
    u = (inv * f).Evaluate().NumPy()
    oA = orc.Csr(rp, col, val, 0)
    ou, osteps, ohist = orc.cg_solve(oA, orc.Jacobi(oA, free.bytes), rhs, prec=1e-8, maxsteps=5000)
    assert abs(inv.GetSteps() - osteps) <= 2, (inv.GetSteps(), osteps)
    assert np.max(np.abs(u - ou)) <= 1e-6 * np.max(np.abs(ou))
    k = min(len(ohist), len(inv.history), 30)
    assert np.allclose(inv.history[:k], ohist[:k], rtol=1e-8)
    # true residual on the free dofs went down by the requested factor (Dirichlet dofs are never touched)
    F = np.unpackbits(free.bytes, bitorder="little")[:box.ndof].astype(bool)
    r = rhs - oA.mult(u)
    assert np.linalg.norm(r[F]) <= 1e-6 * np.linalg.norm(rhs[F])
    assert not u[~F].any()


@pytest.mark.parametrize("m", [64])
def test_large_system_properties(mods, m):
    """7 M dofs / 0.34 G entries (4 GB): properties that do not need an entrywise reference."""
    la, W = mods
    box = W.FemBox(m, order=3)
    A, f = box.device_system()
    n = box.ndof
    assert n == (3 * m + 1) ** 3 and A.nze > 46 * n
    rng = np.random.default_rng(5)
    x, y = la.BaseVector(rng.standard_normal(n)), la.BaseVector(rng.standard_normal(n))
    Ax, Ay = (A * x).Evaluate(), (A * y).Evaluate()
    # symmetry of the assembled operator
    assert la.InnerProduct(Ax, y) == pytest.approx(la.InnerProduct(x, Ay), rel=1e-10)
    # linearity
    z = x.CreateVector()
    z.data = 2.0 * x - 0.5 * y
    Az = (A * z).Evaluate()
    Az.data -= 2.0 * Ax - 0.5 * Ay
    assert Az.Norm() <= 1e-12 * Ax.Norm()
    # constants are in the kernel of the stiffness matrix: vertex dofs 1, hierarchical dofs 0
    c = np.zeros(n)
    c[:(m + 1) ** 3] = 1.0
    assert (A * la.BaseVector(c)).Evaluate().Norm() <= 1e-10 * Ax.Norm()
    # both SpMV kernels agree
    A.ctx.set_option("spmv_algo", 1)
    Ax1 = (A * x).Evaluate()
    A.ctx.set_option("spmv_algo", 0)
    Ax1.data -= Ax
    assert Ax1.Norm() <= 1e-13 * Ax.Norm()
    # Jacobi-CG reduces <d,w> monotonically enough to converge; exact step count checked at small size
    inv = la.CGSolver(A, A.CreateSmoother(box.freedofs()), precision=1e-6, maxsteps=3000)
    u = (inv * f).Evaluate()
    assert 50 < inv.GetSteps() < 3000 and inv.history[-1] <= 1e-12 * inv.history[0]
    # ngsb_cg_solve_host (host buffers) gives the same answer as the device-vector call
    uh, steps, _ = la.cg_solve_host(A, A.CreateSmoother(box.freedofs()), f.NumPy(), precision=1e-6, maxsteps=3000)
    assert steps == inv.GetSteps() and np.array_equal(uh, u.NumPy())


def test_elasticity_cg_and_helmholtz_gmres_match_oracle_on_generated_systems(mods):
    """configs 2 and 5 in small: 3x3-block elasticity CG and complex GMRes on generated systems, steps within
    +-2 of the oracle's restatement of the reference solvers."""
    la, W = mods
    box = W.FemBox((8, 8, 8), order=3, kind=W.BLOCK3, lame=(1.0, 0.7))
    rp, col, val, rhs = box.host_csr()
    free = box.freedofs()
    A, f = box.device_system()
    inv = la.CGSolver(A, A.CreateSmoother(free), precision=1e-8, maxsteps=5000)
    u = (inv * f).Evaluate().NumPy().reshape(-1)
    oA = orc.Csr(rp, col, val, 3)
    ou, osteps, ohist = orc.cg_solve(oA, orc.Jacobi(oA, free.bytes), rhs, prec=1e-8, maxsteps=5000)
    assert abs(inv.GetSteps() - osteps) <= 2, (inv.GetSteps(), osteps)
    assert np.max(np.abs(u - ou)) <= 1e-6 * np.max(np.abs(ou))
    # complex shifted Laplace (definite enough for GMRes to converge quickly)
    boxc = W.FemBox((7, 7, 7), order=3, kind=W.COMPLEX, mass=5.0 + 3.0j)
    rp, col, val, rhs = boxc.host_csr()
    freec = boxc.freedofs()
    Ac, fc = boxc.device_system()
    g = la.GMRESSolver(Ac, Ac.CreateSmoother(freec), precision=1e-8, maxsteps=150)
    x = (g * fc).Evaluate().NumPy()
    oAc = orc.Csr(rp, col, val, 1)
    ox, osteps, _ = orc.gmres_solve(oAc, orc.Jacobi(oAc, freec.bytes), rhs, prec=1e-8, maxsteps=150)
    assert abs(g.GetSteps() - osteps) <= 2, (g.GetSteps(), osteps)
    assert np.max(np.abs(x - ox)) <= 1e-6 * np.max(np.abs(ox))


def test_c16_column_compression_is_bit_exact_and_partial():
    """16-bit column offsets (SELL slices whose entry steps span < 65536 columns) change the bytes streamed, not the
    arithmetic: products with and without are bit-identical; on a grid whose plane stride exceeds 65535 dofs some slices
    must keep their 32-bit columns"""
    import ngsolve_b200.la as la
    from ngsolve_b200 import workloads as W
    ctx = la.default_context()
    box = W.FemBox((88, 88, 3), order=3)            # plane stride (3*88+1)^2 = 70 225 > 65 535
    rng = np.random.default_rng(5)
    ys = []
    for c16 in (1, 0):
        ctx.set_option("sell_c16", c16)
        A, f = box.device_system(ctx)
        x = la.BaseVector(rng.random(A.height) if not ys else xs, ctx=ctx)
        xs = x.NumPy().copy()
        y = A.CreateColVector()
        A.Mult(x, y)
        sb, n16 = A.StreamBytes()
        ent = A.Layout()[0]
        if c16:
            assert 0 < n16 < ent and n16 > 0.5 * ent
            assert sb < A.MultBytes() * 1.02
        else:
            assert n16 == 0
        ys.append(y.NumPy().copy())
        inv = la.CGSolver(A, A.CreateSmoother(box.freedofs()), precision=1e-8, maxsteps=500)
        u = f.CreateVector()
        inv.Mult(f, u)
        ys.append(u.NumPy().copy())
        ys.append(inv.GetSteps())
    ctx.set_option("sell_c16", 1)
    assert np.array_equal(ys[0], ys[3]) and np.array_equal(ys[1], ys[4]) and ys[2] == ys[5]


@pytest.mark.parametrize("n", [(88, 88, 3), (5, 4, 3), (1, 1, 1)])
def test_kernel_variants_are_bit_exact(n):
    """the tuning variants of the real SELL kernel (3 = L2 eviction policies on the compressed loop, 1/2 = other inner
    loops) compute the same row sums in the same order: products and accumulating products are bit-identical, CG solves agree"""
    import ngsolve_b200.la as la
    from ngsolve_b200 import workloads as W
    ctx = la.default_context()
    box = W.FemBox(n, order=3)
    A, f = box.device_system(ctx)
    jac = A.CreateSmoother(box.freedofs())
    rng = np.random.default_rng(11)
    x = la.BaseVector(rng.random(A.height), ctx=ctx)
    y0 = rng.random(A.height)
    out = {}
    try:
        for var in (0, 1, 2, 3, 4, 5, 6):
            ctx.set_option("sell_variant", var)
            y = A.CreateColVector()
            A.Mult(x, y)
            z = la.BaseVector(y0, ctx=ctx)
            A.MultAdd(-0.5, x, z)
            inv = la.CGSolver(A, jac, precision=1e-8, maxsteps=300)
            u = f.CreateVector()
            inv.Mult(f, u)
            out[var] = (y.NumPy().copy(), z.NumPy().copy(), u.NumPy().copy(), inv.GetSteps())
    finally:
        ctx.set_option("sell_variant", 0)
    ref = out[0]
    for key, val in out.items():
        assert np.array_equal(val[0], ref[0]) and np.array_equal(val[1], ref[1]), key
        # the fused <s, A s> is summed per CTA and the variants use different grids: same steps, solution to rounding
        assert abs(val[3] - ref[3]) <= 1 and np.max(np.abs(val[2] - ref[2])) <= 1e-9 * np.max(np.abs(ref[2])), key


@pytest.mark.parametrize("kind", [0, 1, 3], ids=["real", "complex", "block3"])
@pytest.mark.parametrize("stop", ["precision", "maxsteps", "batches"])
def test_cg_fold_u_is_the_same_solve(kind, stop):
    """option cg_fold_u: `u += al s` runs in the direction kernel (which reads s anyway) instead of the update kernel.
    Same arithmetic per entry -> same steps, the same residual history bit for bit, and the same solution (bit-identical
    for real entries; complex products may contract differently, 1e-14); the update owed by the iteration that ends the
    loop (precision reached or maxsteps) must not be lost, and later batch iterations must not repeat it."""
    import ngsolve_b200.la as la
    from ngsolve_b200 import workloads as W
    ctx = la.default_context()
    box = W.FemBox((9, 8, 7), order=2, kind=kind, mass=(1.0 + 0.5j) if kind == 1 else 0.5, lame=(1.0, 0.7))
    A, f = box.device_system(ctx)
    jac = A.CreateSmoother(box.freedofs())
    kw = dict(precision=1e-9, maxsteps=2000) if stop != "maxsteps" else dict(precision=1e-30, maxsteps=37)
    out = []
    try:
        ctx.set_option("cg_persistent", 0)      # the option belongs to the three-kernel loop (the persistent kernel never folds)
        for fold in (0, 1):
            ctx.set_option("cg_fold_u", fold)
            for batch in ((16,) if stop != "batches" else (1, 5)):      # 1: no CUDA graph, 5: the loop ends inside a batch
                ctx.set_option("cg_batch", batch)
                inv = la.CGSolver(A, jac, **kw)
                u = f.CreateVector()
                inv.Mult(f, u)
                out.append((inv.GetSteps(), np.array(inv.history), u.NumPy().copy()))
    finally:
        ctx.set_option("cg_fold_u", 0)
        ctx.set_option("cg_batch", 16)
        ctx.set_option("cg_persistent", -1)
    ref = out[0]
    assert ref[0] > 10
    if stop == "maxsteps":
        assert ref[0] == 38        # GetSteps() = maxsteps + 1 (linalg/cg.cpp:593: n++ in the failing test)
    for steps, hist, u in out[1:]:
        assert steps == ref[0] and np.array_equal(hist, ref[1])
        if kind == 1:
            assert np.max(np.abs(u - ref[2])) <= 1e-14 * np.max(np.abs(ref[2]))
        else:
            assert np.array_equal(u, ref[2])


@pytest.mark.parametrize("kind", [1, 3], ids=["complex", "block3"])
@pytest.mark.parametrize("n", [(32, 32, 2), (6, 5, 4)], ids=["150k-rows-p4", "small-p3"])
def test_c16_all_complex_and_block_entries(kind, n):
    """option sell_c16_all: 16-bit column offsets for Complex and Mat<3,3> matrices too (18.1 / 74.1 instead of 20 / 76
    bytes per entry).  Same summation order: Mult, MultAdd and the CG solve are bit-identical with the option off."""
    import ngsolve_b200.la as la
    from ngsolve_b200 import workloads as W
    ctx = la.default_context()
    box = W.FemBox(n, order=4 if n[0] > 10 else 3, kind=kind, mass=(1.0 + 0.5j) if kind == 1 else 0.5, lame=(1.0, 0.7))
    rng = np.random.default_rng(7)
    xs = rng.random(box.ndof * box.entrysize) if kind != 1 else rng.random(box.ndof) + 1j * rng.random(box.ndof)
    y0 = rng.random(box.ndof * box.entrysize) if kind != 1 else rng.random(box.ndof) - 1j * rng.random(box.ndof)
    res = []
    try:
        for on in (1, 0):
            ctx.set_option("sell_c16_all", on)
            A, f = box.device_system(ctx)
            x = la.BaseVector(xs, entrysize=box.entrysize, ctx=ctx)
            y = A.CreateColVector()
            A.Mult(x, y)
            z = la.BaseVector(y0, entrysize=box.entrysize, ctx=ctx)
            A.MultAdd(-0.75, x, z)
            sb, n16 = A.StreamBytes()
            ent = A.Layout()[0]
            if on:
                assert n16 > 0.5 * ent and sb < A.MultBytes() * 1.03, (n16, ent, sb, A.MultBytes())
                if n[0] < 10:
                    assert n16 == ent
            else:
                assert n16 == 0
            inv = la.CGSolver(A, A.CreateSmoother(box.freedofs()), precision=1e-8, maxsteps=400)
            u = f.CreateVector()
            inv.Mult(f, u)
            res.append((y.NumPy().copy(), z.NumPy().copy(), u.NumPy().copy(), inv.GetSteps()))
    finally:
        ctx.set_option("sell_c16_all", 0)
    for a, b in zip(res[0][:3], res[1][:3]):
        assert np.array_equal(a, b)
    assert res[0][3] == res[1][3]
