"""`conjgrad` on the GPU (LazyAlgebra's routine re-exported by
src/OptimPackNextGen.jl:18-19, call site test/conjgrad-tests.jl:30): the fused passes of
libopkb200 (opkb_cg_update / _curvature / _diag_step, opkb_smooth2d_apply) against the
oracle's `vupdate!` / `vcombine!` / `vdot` sequence, through the C ABI."""
import numpy as np
import pytest

from test_conjgrad import normal_equations, spd_problem

pytestmark = pytest.mark.gpu

F64_RTOL, F32_RTOL = 1e-10, 1e-4


@pytest.fixture(scope="module")
def opk():
    import opkb200
    return opkb200


def host_operator(A):
    """A dense matrix as a user `A!(dst, src)` on device Vectors (through the host:
    the product has no dense kernels, tiny n)."""
    def apply(dst, src):
        dst.upload((A @ src.download()).astype(A.dtype))
        return dst
    return apply


def dense(A):
    def apply(dst, src):
        dst[:] = A @ src
        return dst
    return apply


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [1, 2, 3, 5, 255, 1024, 4099, 300001])
def test_fused_passes_bit_exact(opk, oracle, dtype, n):
    """x += alpha p, r -= alpha q element-wise bit-exact; the sums to the rounding of a
    double sum (exact products); odd lengths exercise the scalar tail."""
    from opkb200 import _lib
    from opkb200.conjgrad import _cg_curvature, _cg_update
    ctx = opk.default_context()
    rng = np.random.default_rng(n)
    x, r, p, q, a = (rng.normal(size=n).astype(dtype) for _ in range(5))
    alpha, beta = 0.37123456789, 1.6180339887
    xv, rv, pv, qv, av = (ctx.from_numpy(v) for v in (x, r, p, q, a))
    # curvature
    pq, pp = _cg_curvature(ctx, pv, qv)
    assert abs(pq - oracle.vdot(p, q)) <= 1e-13 * np.abs(p.astype(np.float64) * q).sum()
    assert abs(pp - oracle.vdot(p, p)) <= 1e-13 * pp
    # update
    rho, xx = _cg_update(ctx, xv, rv, pv, qv, alpha)
    xo, ro = oracle.vupdate(x.copy(), alpha, p), oracle.vupdate(r.copy(), -alpha, q)
    assert np.array_equal(xv.download(), xo) and np.array_equal(rv.download(), ro)
    assert abs(rho - oracle.vdot(ro, ro)) <= 1e-13 * rho and abs(xx - oracle.vdot(xo, xo)) <= 1e-13 * xx
    assert np.array_equal(pv.download(), p) and np.array_equal(qv.download(), q)
    # diagonal step: p = beta p + r; q = a p   (then the restart form p = r)
    op = opk.DiagonalOperator(av)
    for restart in (False, True):
        pv.upload(p)
        pq, pp = op.direction_apply(pv, rv, qv, beta, restart)
        po = ro.copy() if restart else oracle.vcombine(p.copy(), beta, p, 1.0, ro)
        qo = oracle.vproduct(np.empty_like(po), a, po)
        assert np.array_equal(pv.download(), po) and np.array_equal(qv.download(), qo)
        assert abs(pq - oracle.vdot(po, qo)) <= 1e-13 * np.abs(po.astype(np.float64) * qo).sum()
        assert abs(pp - oracle.vdot(po, po)) <= 1e-13 * pp
    with pytest.raises(opk.OpkbError):
        _cg_update(ctx, xv, xv, pv, qv, alpha)               # aliased vectors
    with pytest.raises(opk.OpkbError):
        _cg_curvature(ctx, pv, ctx.vector(n + 1, dtype))     # another space


@pytest.mark.parametrize("T,prec", [(np.float32, 1e-5), (np.float64, 1e-7)])
def test_reference_assertion_user_operator(opk, oracle, T, prec):
    """test/conjgrad-tests.jl:21-34 with a user operator `A!(dst, src)`."""
    for seed in range(3):
        A, b, x0 = normal_equations(T, seed)
        x = opk.conjgrad(host_operator(A), b, ftol=1e-16, maxiter=len(b), quiet=True)
        assert x.dtype == T and np.abs(x - x0).max() <= prec
        xo, info, _ = oracle.conjgrad(dense(A), b, ftol=1e-16, maxiter=len(b))
        # (an ill-conditioned system run to stagnation: two roundings of the sums away from each
        # other is as far as from the direct solution)
        assert np.abs(x - xo).max() <= 2 * prec


CASES = [dict(), dict(ftol=0.0, gtol=(0.0, 1e-6)), dict(ftol=0.0, xtol=1e-5), dict(ftol=0.0, maxiter=17),
         dict(ftol=1e-14, restart=7)]


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("kw", CASES)
def test_trajectory_user_operator(opk, oracle, T, kw):
    """Same iterate sequence as the oracle for the first 20 iterations (a user operator:
    direction, operator, curvature kernel, fused update)."""
    A, b = spd_problem(60, T)
    x0 = np.random.default_rng(5).normal(size=60).astype(T)
    rtol = F64_RTOL if T == np.float64 else F32_RTOL
    for start in (None, x0):
        xo, info, tr = oracle.conjgrad(dense(A), b, start, trace=True, **kw)
        got = []
        s = opk.ConjGrad(host_operator(A), b, start, quiet=True,
                         observer=lambda s: got.append((s.iter, s.rho, s.x.download())), **kw)
        s.solve()
        assert s.status == info["status"]
        # the stopping iteration may move by one when a test is met to rounding
        assert abs(s.iter - info["iter"]) <= (0 if "maxiter" in kw else 1)
        for (k, rho, xk), w in list(zip(got, tr))[:20]:
            assert k == w["iter"]
            assert np.linalg.norm(xk - w["x"]) <= rtol * max(np.linalg.norm(w["x"]), 1e-300)
        assert np.linalg.norm(s.result() - xo) <= 10 * rtol * np.linalg.norm(xo)


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("n", [1000, 200001])
def test_trajectory_diagonal_operator(opk, oracle, T, n):
    """The one-kernel direction + product + curvature path against the oracle's
    `vproduct!` operator."""
    a, c = opk.Quadratic.synthetic(n, T, kappa=1e3)
    b = (a * c).astype(T)
    rtol = F64_RTOL if T == np.float64 else F32_RTOL
    kw = dict(ftol=0.0, maxiter=25, restart=10)
    xo, info, tr = oracle.conjgrad(("diag", a), b, trace=True, **kw)
    got = []
    ctx = opk.default_context()
    launches = ctx.launch_count
    s = opk.ConjGrad(opk.DiagonalOperator(a), b, quiet=True,
                     observer=lambda s: got.append((s.iter, s.rho, s.x.download())), **kw)
    s.solve()
    assert (s.iter, s.status) == (info["iter"], info["status"]) == (25, -1)
    assert ctx.launch_count - launches <= 2 * 25 + 2 + 2 * 3 + 4     # two kernels per iteration
    for (k, rho, xk), w in list(zip(got, tr))[:21]:
        assert k == w["iter"] and abs(rho - w["rho"]) <= 100 * rtol * w["rho"]
        assert np.linalg.norm(xk - w["x"]) <= rtol * max(np.linalg.norm(w["x"]), 1e-300)
    # and the solution of the diagonal system is c
    x = opk.conjgrad(opk.DiagonalOperator(a), b, ftol=0.0, gtol=(0.0, 1e-12 if T == np.float64 else 1e-6),
                     maxiter=2000)
    assert np.abs(x - c).max() <= (1e-7 if T == np.float64 else 5e-3)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(37, 53), (1, 300), (300, 1), (2, 2), (64, 1000)])
def test_smooth2d_apply_matches_array_expression(opk, oracle, dtype, shape):
    rows, cols = shape
    w, _ = opk.Smooth2D.synthetic(rows, cols, dtype)
    p = np.random.default_rng(3).normal(size=rows * cols).astype(dtype)
    ctx = opk.default_context()
    op = opk.Smooth2DOperator(w, 0.7, cols)
    pv, qv = ctx.from_numpy(p), ctx.vector(rows * cols, dtype)
    pq, pp = op.apply_dot(qv, pv)
    q_ref = opk.Smooth2DOperator.reference(w, 0.7, cols)(np.empty_like(p), p)
    assert np.array_equal(qv.download(), q_ref)                       # element-wise: bit-exact
    assert abs(pq - oracle.vdot(p, q_ref)) <= 1e-13 * np.abs(p.astype(np.float64) * q_ref).sum()
    assert abs(pp - oracle.vdot(p, p)) <= 1e-13 * pp
    with pytest.raises(opk.OpkbError):
        op.apply_dot(pv, pv)                                          # q must not alias p


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(37, 53), (1, 300), (300, 1), (2, 2), (64, 1000), (9, 2048), (17, 4100)])
def test_smooth2d_step_matches_array_expression(opk, oracle, dtype, shape):
    """opkb_cg_smooth2d_step: p <- beta p + r (or r), q = A.p, p.q, p.p in one pass; p and q
    bit-identical to `vcombine!` followed by the array expression of the operator."""
    rows, cols = shape
    n = rows * cols
    w, _ = opk.Smooth2D.synthetic(rows, cols, dtype)
    rng = np.random.default_rng(7)
    p, r = rng.normal(size=n).astype(dtype), rng.normal(size=n).astype(dtype)
    beta = 0.6180339887
    ctx = opk.default_context()
    op = opk.Smooth2DOperator(w, 0.7, cols)
    A = opk.Smooth2DOperator.reference(w, 0.7, cols)
    for restart in (False, True):
        pv, rv, qv = ctx.from_numpy(p), ctx.from_numpy(r), ctx.vector(n, dtype)
        pq, pp = op.direction_apply(pv, rv, qv, beta, restart)
        p_ref = r.copy() if restart else oracle.vcombine(np.empty_like(p), beta, p, 1.0, r)
        q_ref = A(np.empty_like(p), p_ref)
        assert np.array_equal(pv.download(), p_ref) and np.array_equal(qv.download(), q_ref)
        assert np.array_equal(rv.download(), r)
        assert abs(pq - oracle.vdot(p_ref, q_ref)) <= 1e-13 * np.abs(p_ref.astype(np.float64) * q_ref).sum()
        assert abs(pp - oracle.vdot(p_ref, p_ref)) <= 1e-13 * pp
        # a second step on top of the first one (the buffers were swapped)
        pq2, pp2 = op.direction_apply(pv, rv, qv, beta, False)
        p2 = oracle.vcombine(np.empty_like(p), beta, p_ref, 1.0, r)
        assert np.array_equal(pv.download(), p2) and np.array_equal(qv.download(), A(np.empty_like(p), p2))
    with pytest.raises(opk.OpkbError):
        op.direction_apply(pv, rv, pv, beta, False)                   # q must not alias p


@pytest.mark.parametrize("dtype,shape", [(np.float64, (40, 64)), (np.float64, (129, 77)), (np.float32, (40, 64))])
def test_trajectory_smooth2d_operator(opk, oracle, dtype, shape):
    """conjgrad(A, w.*y) with the stencil operator: the oracle's iterates for 20
    iterations; the solution is the unconstrained minimiser of Smooth2D (g = 0)."""
    rows, cols = shape
    n = rows * cols
    mu = 2.0
    w, y = opk.Smooth2D.synthetic(rows, cols, dtype)
    b = (w * y).astype(dtype)
    rtol = F64_RTOL if dtype == np.float64 else F32_RTOL
    xo, info, tr = oracle.conjgrad(opk.Smooth2DOperator.reference(w, mu, cols), b, ftol=0.0, maxiter=25,
                                   trace=True)
    got = []
    s = opk.ConjGrad(opk.Smooth2DOperator(w, mu, cols), b, ftol=0.0, maxiter=25, quiet=True,
                     observer=lambda s: got.append((s.iter, s.rho, s.x.download())))
    s.solve()
    assert (s.iter, s.status) == (info["iter"], info["status"])
    for (k, rho, xk), wt in list(zip(got, tr))[:21]:
        assert k == wt["iter"]
        assert np.linalg.norm(xk - wt["x"]) <= rtol * max(np.linalg.norm(wt["x"]), 1e-300)
    # run to convergence: the gradient of Smooth2D vanishes there
    x = opk.conjgrad(opk.Smooth2DOperator(w, mu, cols), b, ftol=0.0,
                     gtol=(0.0, 1e-10 if dtype == np.float64 else 1e-5), maxiter=5000)
    g = np.empty_like(x)
    opk.Smooth2D.reference(w, y, mu, cols)(x, g)
    assert np.linalg.norm(g) <= (1e-8 if dtype == np.float64 else 1e-3) * np.linalg.norm(b)


def test_inplace_and_device_vectors(opk):
    """`conjgrad!(x, A, b, x0, p, q, r)` on device Vectors, workspace supplied."""
    ctx = opk.default_context()
    n = 5000
    a, c = opk.Quadratic.synthetic(n, np.float64, kappa=10.0)
    op = opk.DiagonalOperator(a)
    bv = ctx.from_numpy(a * c)
    xv = ctx.from_numpy(np.ones(n))
    p, q, r = (ctx.vector(n) for _ in range(3))
    out = opk.conjgrad_(xv, op, bv, xv, p, q, r, ftol=0.0, gtol=(0.0, 1e-12), maxiter=500)
    assert out is xv and np.abs(xv.download() - c).max() <= 1e-9
    x0 = ctx.from_numpy(np.ones(n))
    x2 = opk.conjgrad(op, bv, x0, ftol=0.0, gtol=(0.0, 1e-12), maxiter=500)
    assert isinstance(x2, opk.Vector) and np.array_equal(x0.download(), np.ones(n))   # x0 untouched
    assert np.array_equal(x2.download(), xv.download())                                 # reproducible
    with pytest.raises(ArithmeticError):
        opk.conjgrad(opk.DiagonalOperator(-a), bv)                                      # not positive definite


def test_large_grid_properties(opk):
    """4096 x 4096 stencil system: the residual norm the solver tracks by recurrence equals
    the true one, two runs are bit-identical (two kernels per iteration)."""
    rows = cols = 4096
    n = rows * cols
    ctx = opk.default_context()
    w, y = opk.Smooth2D.synthetic(rows, cols)
    wv = ctx.from_numpy(w)
    bv = ctx.from_numpy(w * y)
    op = opk.Smooth2DOperator(wv, 4.0, cols)
    res = []
    for _ in range(2):
        rhos = []
        s = opk.ConjGrad(op, bv, ftol=0.0, maxiter=40, restart=1000, quiet=True,
                         observer=lambda s: rhos.append(s.rho))
        s.solve()
        res.append((s.rho, s.x.download()))
        assert rhos[-1] < 1e-6 * rhos[0]
    assert res[0][0] == res[1][0] and np.array_equal(res[0][1], res[1][1])
    q = ctx.vector(n)
    op(q, s.x)
    true_r = bv.download() - q.download()
    assert abs(np.dot(true_r, true_r) - res[1][0]) <= 1e-6 * res[1][0]


def test_conjgrad_golden(opk):
    """The CUDA path against the committed fixture (tests/golden/conjgrad_traces.json,
    traces of the oracle): nothing outside the repository is needed."""
    import json
    import os
    import sys
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    sys.path.insert(0, gdir)
    from make_golden import cg_problem
    with open(os.path.join(gdir, "conjgrad_traces.json")) as fh:
        gold = json.load(fh)
    for case in gold["cases"]:
        dtype = np.dtype(case["dtype"])
        rows, cols = case["shape"]
        _, b, w = cg_problem(case["operator"], (rows, cols), case["dtype"])
        A = opk.Smooth2DOperator(w, 2.0, cols) if case["operator"] == "stencil" else opk.DiagonalOperator(w)
        rtol = F64_RTOL if dtype == np.float64 else F32_RTOL
        got = []
        s = opk.ConjGrad(A, b, quiet=True, observer=lambda s: got.append((s.iter, s.rho, s.x.download()[:8])),
                         **case["kw"])
        s.solve()
        assert s.status == case["final"]["status"] and abs(s.iter - case["final"]["iter"]) <= 1
        for (k, rho, xh), wt in list(zip(got, case["trace"]))[:21]:
            assert k == wt["iter"] and abs(rho - wt["rho"]) <= 100 * rtol * wt["rho"]
            ref = np.array(wt["x_head"])
            assert np.linalg.norm(xh - ref) <= rtol * max(np.linalg.norm(ref), 1e-300) + 1e-300


def test_edge_cases(opk, oracle):
    """Empty and one-element systems, b = 0, x0 already the solution, a 1 x 1 grid."""
    ctx = opk.default_context()
    assert not ctx.uses_mailbox                                        # one rank: no exchange at all
    # n = 0: nothing to do, converged at once
    x = opk.conjgrad(opk.DiagonalOperator(np.zeros(0)), np.zeros(0))
    assert x.shape == (0,)
    # n = 1: a scalar equation, solved in one iteration
    s = opk.ConjGrad(opk.DiagonalOperator(np.array([4.0])), np.array([2.0]), ftol=0.0, maxiter=5)
    s.solve()
    assert s.result()[0] == 0.5 and s.iter == 1 and s.status == opk.CG_GTEST
    # b = 0 and x0 = 0: |r| = 0 <= gtest = 0
    s = opk.ConjGrad(opk.DiagonalOperator(np.ones(10)), np.zeros(10))
    s.solve()
    assert s.iter == 0 and s.status == opk.CG_GTEST and not s.result().any()
    # x0 is the solution already: r = b - A.x0 = 0
    a, c = opk.Quadratic.synthetic(1000, np.float64, kappa=10.0)
    s = opk.ConjGrad(opk.DiagonalOperator(a), a * c, c.copy(), ftol=0.0, gtol=(1e-12, 0.0), maxiter=10)
    s.solve()
    assert s.iter <= 1 and np.abs(s.result() - c).max() <= 1e-12
    # a 1 x 1 grid: the stencil operator is w itself
    w = np.array([3.0])
    x = opk.conjgrad(opk.Smooth2DOperator(w, 2.0, 1), np.array([6.0]), ftol=0.0, maxiter=3)
    assert x[0] == 2.0
    xo, info, _ = oracle.conjgrad(opk.Smooth2DOperator.reference(w, 2.0, 1), np.array([6.0]), ftol=0.0, maxiter=3)
    assert xo[0] == 2.0
