"""`conjgrad` on the CPU: the oracle against the reference's own test
(test/conjgrad-tests.jl:16-36) and against SciPy's `cg`, and the product's host loop
(`opkb200.ConjGrad` on the numpy back end of tests/cpu_backend.py) against the oracle,
bit for bit.  The GPU twin is tests/test_gpu_conjgrad.py."""
import numpy as np
import pytest

import opkb200 as opk
from cpu_backend import NumpyCgOps


def normal_equations(T, seed=0):
    """The problem of test/conjgrad-tests.jl:21-28: A = H'.H, b = H'.y, dims = (70, 39)."""
    rng = np.random.default_rng(seed)
    H = rng.normal(size=(70, 39)).astype(T)
    H *= T(1 / np.abs(H).max())
    x = rng.normal(size=39).astype(T)
    y = H @ x + T(0.01) * rng.normal(size=70).astype(T)
    A = (H.T @ H).astype(T)
    b = (H.T @ y).astype(T)
    return A, b, np.linalg.solve(A.astype(np.float64), b.astype(np.float64))


def dense_operator(A):
    def apply(dst, src):
        dst[:] = A @ src
        return dst
    return apply


def spd_problem(n, T, seed=1):
    rng = np.random.default_rng(seed)
    H = rng.normal(size=(2 * n, n))
    A = (H.T @ H / n + 0.1 * np.eye(n)).astype(T)
    return A, rng.normal(size=n).astype(T)


@pytest.mark.parametrize("T,prec", [(np.float32, 1e-5), (np.float64, 1e-7)])
def test_oracle_meets_the_reference_assertion(oracle, T, prec):
    """test/conjgrad-tests.jl:30-34: conjgrad(A!, b; ftol=1e-16, maxiter=length(b)),
    max|x - A\\b| <= prec (Float16 is not a type of this path)."""
    for seed in range(5):
        A, b, x0 = normal_equations(T, seed)
        x, info, _ = oracle.conjgrad(dense_operator(A), b, ftol=1e-16, maxiter=len(b))
        assert np.abs(x - x0).max() <= prec
        # the product's host loop on numpy vectors: the same run, bit for bit
        s = opk.ConjGrad(dense_operator(A), b.copy(), np.zeros_like(b), ops=NumpyCgOps(), ftol=1e-16,
                         maxiter=len(b), quiet=True)
        s.solve()
        assert (s.iter, s.status) == (info["iter"], info["status"])
        assert np.array_equal(s.x, x)


def test_oracle_follows_scipy_cg(oracle):
    """An implementation that is not ours: scipy.sparse.linalg.cg runs the same
    Hestenes-Stiefel recurrences; without restart the iterates agree to rounding."""
    cg = pytest.importorskip("scipy.sparse.linalg").cg
    A, b = spd_problem(100, np.float64)
    ref = []
    cg(A, b, rtol=0.0, atol=0.0, maxiter=30, callback=lambda xk: ref.append(xk.copy()))
    x, info, tr = oracle.conjgrad(dense_operator(A), b, ftol=0.0, maxiter=30, restart=1000, trace=True)
    assert info["iter"] == 30 and info["status"] == -1
    for k in range(1, 31):
        assert np.linalg.norm(tr[k]["x"] - ref[k - 1]) <= 1e-13 * np.linalg.norm(ref[k - 1])


CASES = [
    dict(),                                   # defaults: ftol = 1e-8
    dict(ftol=0.0, gtol=(0.0, 1e-6)),         # gtest
    dict(ftol=0.0, xtol=1e-5),                # xtest
    dict(ftol=0.0, maxiter=17),               # too many iterations
    dict(ftol=1e-14, restart=7),              # restarts on the true residuals
]


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("kw", CASES)
def test_host_loop_equals_oracle(oracle, T, kw):
    A, b = spd_problem(60, T)
    rng = np.random.default_rng(5)
    for x0 in (None, rng.normal(size=60).astype(T)):
        x, info, tr = oracle.conjgrad(dense_operator(A), b, x0, trace=True, **kw)
        got = []
        s = opk.ConjGrad(dense_operator(A), b.copy(), np.zeros_like(b) if x0 is None else x0.copy(),
                         ops=NumpyCgOps(), quiet=True, observer=lambda s: got.append((s.iter, s.rho, s.x.copy())),
                         **kw)
        s.solve()
        assert (s.iter, s.status, s.rho, s.phi) == (info["iter"], info["status"], info["rho"], info["phi"])
        assert np.array_equal(s.x, x)
        assert len(got) == len(tr)
        for (k, rho, xk), w in zip(got, tr):
            assert (k, rho) == (w["iter"], w["rho"]) and np.array_equal(xk, w["x"])


def test_statuses_and_errors(oracle):
    A, b = spd_problem(30, np.float64)
    run = lambda **kw: opk.ConjGrad(dense_operator(A), b.copy(), np.zeros_like(b), ops=NumpyCgOps(), quiet=True, **kw)
    s = run(ftol=0.0, gtol=(0.0, 1e-8)); s.solve(); assert s.status == opk.CG_GTEST
    s = run(); s.solve(); assert s.status == opk.CG_FTEST
    s = run(ftol=0.0, xtol=1e-4); s.solve(); assert s.status == opk.CG_XTEST
    s = run(ftol=0.0, maxiter=3); s.solve(); assert s.status == opk.CG_TOO_MANY_ITERATIONS and s.iter == 3
    # not positive definite: an error when strict (default), a status otherwise
    neg = dense_operator(-A)
    with pytest.raises(ArithmeticError):
        opk.ConjGrad(neg, b.copy(), np.zeros_like(b), ops=NumpyCgOps()).solve()
    s = opk.ConjGrad(neg, b.copy(), np.zeros_like(b), ops=NumpyCgOps(), strict=False, quiet=True)
    s.solve()
    assert s.status == opk.CG_NOT_POSITIVE_DEFINITE
    x, info, _ = oracle.conjgrad(neg, b)
    assert info["status"] == -2
    with pytest.raises(ValueError):
        run(ftol=-1.0)
    with pytest.raises(ValueError):
        run(restart=0)
    # b = 0: converged at once (gtest with rho = 0)
    s = opk.ConjGrad(dense_operator(A), np.zeros(30), np.zeros(30), ops=NumpyCgOps())
    s.solve()
    assert s.status == opk.CG_GTEST and s.iter == 0


def test_smooth2d_operator_is_the_hessian_of_smooth2d():
    """`Smooth2DOperator.reference` (the numpy twin of opkb_smooth2d_apply) against the
    gradient of `Smooth2D.reference`: g(x) = A.x - w.*y, and A is symmetric."""
    rows, cols = 13, 24
    w, y = opk.Smooth2D.synthetic(rows, cols)
    fg = opk.Smooth2D.reference(w, y, 2.0, cols)
    A = opk.Smooth2DOperator.reference(w, 2.0, cols)
    rng = np.random.default_rng(2)
    x, z = rng.normal(size=rows * cols), rng.normal(size=rows * cols)
    g, Ax, Az = np.empty_like(x), np.empty_like(x), np.empty_like(x)
    fg(x, g)
    A(Ax, x)
    A(Az, z)
    assert np.allclose(g, Ax - w * y, rtol=0, atol=1e-12)
    assert abs(np.dot(z, Ax) - np.dot(x, Az)) <= 1e-10 * abs(np.dot(z, Ax))


def test_oracle_reproduces_conjgrad_golden(oracle):
    """tests/golden/conjgrad_traces.json (make_golden.py conjgrad): the committed fixture
    the GPU suite is checked against must still be what the oracle produces."""
    import json
    import os
    import sys
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    sys.path.insert(0, gdir)
    from make_golden import cg_problem
    with open(os.path.join(gdir, "conjgrad_traces.json")) as fh:
        gold = json.load(fh)
    assert len(gold["cases"]) >= 4
    for case in gold["cases"]:
        A, b, _ = cg_problem(case["operator"], tuple(case["shape"]), case["dtype"])
        kw = dict(case["kw"])
        x, info, tr = oracle.conjgrad(A, b, trace=True, **kw)
        assert {k: info[k] for k in ("iter", "status", "rho", "phi")} == case["final"]
        assert [float(v) for v in x[:8]] == case["x_final_head"]
        assert len(tr) == len(case["trace"])
        for t, w in zip(tr, case["trace"]):
            assert (t["iter"], t["rho"], t["phi"]) == (w["iter"], w["rho"], w["phi"])
            assert [float(v) for v in t["x"][:8]] == w["x_head"]
