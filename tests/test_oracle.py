"""Pins the CPU oracle: analytic known answers, the reference's own test
assertions (test/rosenbrock.jl:61-123) and the spec of the vector operations
(test/vectors-tests.jl:15-81, doc/algebra.md:84-116).  CPU only."""
import math

import numpy as np
import pytest

DTYPES = (np.float64, np.float32)


@pytest.mark.parametrize("dtype", DTYPES)
def test_rosenbrock_known_answers(oracle, dtype):
    # f0 = 10*24.2 = 242, |g0| = 736.3923 at x0 = (-1.2, 1, ...)
    x0 = oracle.rosenbrock_init(20, dtype)
    assert x0[0] == dtype(-1.2) and x0[1] == dtype(1.0)
    g = np.empty_like(x0)
    f = oracle.rosenbrock_fg(x0, g)
    rtol = 1e-12 if dtype == np.float64 else 1e-6
    assert f == pytest.approx(242.0, rel=rtol)
    assert oracle.vnorm2(g) == pytest.approx(math.sqrt(542273.6), rel=rtol)
    # projected start (lower = 0): f0 = 1010, |p0| = 632.4872
    x = oracle.project_variables(np.empty_like(x0), x0, 0.0, None)
    f = oracle.rosenbrock_fg(x, g)
    p = oracle.project_direction(np.empty_like(x), x, 0.0, None, -1, g)
    assert f == pytest.approx(1010.0, rel=rtol)
    assert oracle.vnorm2(p) == pytest.approx(632.4872, rel=1e-6)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("kw", [dict(), dict(mem=15), dict(lower=0.0)],
                         ids=["default", "mem15", "lower0"])
def test_reference_assertions_vmlmb(oracle, dtype, kw):
    # test/rosenbrock.jl:76-79, :89-92, :101-104: max|x - 1| < 1e-3, n = 20
    x0 = oracle.rosenbrock_init(20, dtype)
    x, res, _ = oracle.vmlmb("rosenbrock", x0, **kw)
    assert res["rc"] == 0 and res["stage"] == 3
    assert np.abs(x - 1).max() < 1e-3


def test_survey_anchor_counts(oracle):
    # restatement-derived anchors of SURVEY 8(c) (not reference output)
    x0 = oracle.rosenbrock_init(20, np.float64)
    x, res, tr = oracle.vmlmb("rosenbrock", x0, trace=True)
    assert (res["iter"], res["eval"], res["rejects"]) == (35, 49, 0)
    assert tr[1]["f"] == pytest.approx(124.25818092431233, rel=1e-13)
    assert tr[1]["stp"] == pytest.approx(1.0 / 736.3922867602565, rel=1e-13)
    assert tr[2]["stp"] == 1.0
    x, res, _ = oracle.vmlmb("rosenbrock", x0, mem=15)
    assert (res["iter"], res["eval"]) == (34, 48)
    x, res, tr = oracle.vmlmb("rosenbrock", x0, lower=0.0, trace=True)
    assert (res["iter"], res["eval"]) == (25, 31)
    assert tr[4]["stp"] == pytest.approx(0.1)  # More-Toraldo gamma1 clamp
    xb, resb, _ = oracle.vmlmb("rosenbrock", x0, lower=0.0, blmvm=True)
    assert np.array_equal(x, xb)  # all variables stay free
    x, info, _ = oracle.spg("rosenbrock", x0, 10, lower=0.0)
    assert (info["iter"], info["fcnt"], info["pcnt"], info["status"]) == (64, 82, 130, 1)
    assert np.abs(x - 1).max() < 1e-3


def test_bounded_solution(oracle):
    # lower = 0, upper = 0.9 -> x* = (0.9, 0.81), f* = 0.01 n/2
    n = 20
    x0 = oracle.rosenbrock_init(n, np.float64)
    x, res, tr = oracle.vmlmb("rosenbrock", x0, lower=0.0, upper=0.9, trace=True)
    assert res["stage"] == 3
    assert np.allclose(x[0::2], 0.9) and np.allclose(x[1::2], 0.81, atol=1e-6)
    assert res["f"] == pytest.approx(0.01 * n / 2, rel=1e-8)
    assert not tr[-1]["mask"][0::2].any() and tr[-1]["mask"][1::2].all() or res["gnorm"] < 1e-4


@pytest.mark.parametrize("dtype", DTYPES)
def test_vector_ops_spec(oracle, dtype):
    # test/vectors-tests.jl:26-81 (exact equalities)
    rng = np.random.default_rng(1)
    a = rng.standard_normal(60).astype(dtype)
    b = rng.standard_normal(60).astype(dtype)
    T = dtype
    pi = T(math.pi)
    for alpha, want in ((0, 0 * a), (1, a), (-1, -a), (math.pi, pi * a)):
        assert np.array_equal(oracle.vscale(a.copy(), alpha), want)
    for alpha, want in ((0, a), (1, a + b), (-1, a - b), (math.pi, a + pi * b)):
        assert np.array_equal(oracle.vupdate(a.copy(), alpha, b), want)
    assert np.array_equal(oracle.vproduct(np.empty_like(a), a, b), a * b)
    al, be = T(rng.standard_normal()), T(rng.standard_normal())
    cases = [(0, 0, np.zeros_like(a)), (0, 1, b), (0, -1, -b), (0, math.pi, pi * b),
             (1, 0, a), (-1, 0, -a), (math.pi, 0, pi * a), (1, 1, a + b),
             (-1, 1, b - a), (1, -1, a - b), (-1, -1, -(a + b)),
             (float(al), float(be), al * a + be * b)]
    for alpha, beta, want in cases:
        got = oracle.vcombine(np.empty_like(a), alpha, a, beta, b)
        assert np.array_equal(got, want), (alpha, beta)


@pytest.mark.parametrize("dtype", DTYPES)
def test_reductions_modes(oracle, dtype):
    rng = np.random.default_rng(2)
    n = 100003
    x = rng.standard_normal(n).astype(dtype)
    y = rng.standard_normal(n).astype(dtype)
    exact = math.fsum(float(u) * float(v) for u, v in zip(x[:5000], y[:5000]))
    assert oracle.vdot(x[:5000].copy(), y[:5000].copy()) == pytest.approx(exact, rel=1e-13, abs=1e-13)
    ref = float(np.dot(x.astype(np.float64), y.astype(np.float64)))
    assert oracle.vdot(x, y) == pytest.approx(ref, rel=1e-12, abs=1e-10)
    sel = np.flatnonzero(rng.random(n) < 0.3)
    refs = float(np.dot(x[sel].astype(np.float64), y[sel].astype(np.float64)))
    assert oracle.vdot_sel(sel, x, y) == pytest.approx(refs, rel=1e-12, abs=1e-10)
    assert oracle.vnorm2(x) == pytest.approx(math.sqrt(float(np.dot(x.astype(np.float64), x.astype(np.float64)))), rel=1e-13)
    assert oracle.vnorminf(x) == float(np.abs(x).max())
    try:
        tol = 1e-9 if dtype == np.float64 else 1e-3
        for mode in (oracle.SUM_SEQUENTIAL, oracle.SUM_PAIRWISE):
            oracle.set_sum_mode(mode)
            assert oracle.vdot(x, y) == pytest.approx(ref, rel=tol, abs=tol * 100)
    finally:
        oracle.set_sum_mode(oracle.SUM_ACCURATE)
    # thread-count independence of the accurate mode
    v1 = oracle.vdot(x, y)
    oracle.set_num_threads(1)
    v2 = oracle.vdot(x, y)
    oracle.set_num_threads(0)
    assert v1 == v2


@pytest.mark.parametrize("dtype", DTYPES)
def test_bounds_semantics(oracle, dtype):
    T = dtype
    nan, inf = T(np.nan), T(np.inf)
    # fastclamp keeps NaN (src/bounds.jl:41, :53)
    src = np.array([-2, 0.5, 3, nan], dtype=T)
    out = oracle.project_variables(np.empty_like(src), src, 0.0, 1.0)
    assert np.array_equal(out[:3], np.array([0, 0.5, 1], dtype=T)) and np.isnan(out[3])
    lo = np.array([0, 0, 0, 0], dtype=T)
    hi = np.array([1, 1, 2, 1], dtype=T)
    out = oracle.project_variables(np.empty_like(src), src, lo, hi)
    assert np.array_equal(out[:3], np.array([0, 0.5, 2], dtype=T))
    # project_direction with orientation `-` (src/bounds.jl:647-648):
    # keep d if (d > 0 and x > lo) or (d < 0 and x < hi)
    x = np.array([0, 0, 1, 1, 0.5, 0.5], dtype=T)
    d = np.array([1, -1, 1, -1, 1, -1], dtype=T)
    p = oracle.project_direction(np.empty_like(x), x, 0.0, 1.0, -1, d)
    assert np.array_equal(p, np.array([0, -1, 1, 0, 1, -1], dtype=T))
    p = oracle.project_direction(np.empty_like(x), x, 0.0, None, -1, d)
    assert np.array_equal(p, np.array([0, -1, 1, -1, 1, -1], dtype=T))
    p = oracle.project_direction(np.empty_like(x), x, None, 1.0, +1, d)
    assert np.array_equal(p, np.array([1, -1, 0, -1, 1, -1], dtype=T))
    with pytest.raises(ValueError):
        oracle.project_direction(np.empty_like(x), x, 1.0, 0.0, -1, d)
    # step_limits (src/bounds.jl:600-624): search direction is -d
    x = np.array([0.5, 0.5, 0.25, 0.0], dtype=T)
    d = np.array([1.0, -1.0, 0.0, -2.0], dtype=T)
    smin, smax = oracle.step_limits(x, 0.0, 1.0, -1, d)
    assert smin == 0.5 and smax == 0.5
    smin, smax = oracle.step_limits(x, 0.0, None, -1, d)   # unbounded above
    assert smin == 0.5 and smax == math.inf
    smin, smax = oracle.step_limits(x, None, None, -1, d)
    assert smin == math.inf and smax == math.inf
    d0 = np.zeros(4, dtype=T)
    smin, smax = oracle.step_limits(x, 0.0, 1.0, -1, d0)
    assert smin == math.inf and smax == 0.0
    # get_free_variables (src/bounds.jl:701-714), 0-based here
    gp = np.array([0, 1, -0.0, 2, nan], dtype=T)
    assert oracle.get_free_variables(gp).tolist() == [1, 3, 4]
    # get_free_variables(x, lo, hi, orient, d) (src/bounds.jl:716-844) == projdir != 0
    x = np.array([0, 0, 1, 1, 0.5, 0.5], dtype=T)
    d = np.array([1, -1, 1, -1, 0, -1], dtype=T)
    assert oracle.get_free_variables_dir(x, 0.0, 1.0, -1, d).tolist() == [1, 2, 5]
    assert oracle.get_free_variables_dir(x, None, None, -1, d).tolist() == [0, 1, 2, 3, 4, 5]
    assert oracle.get_free_variables_dir(x, 0.0, None, +1, d).tolist() == [0, 2, 3, 5]
    # empty input
    e = np.empty(0, dtype=T)
    assert oracle.vdot(e, e) == 0.0 and oracle.get_free_variables(e).size == 0


def test_linesearch_state_machines(oracle):
    O = oracle
    # start! argument errors (src/linesearches.jl:334-349)
    ls = O.LineSearch(O.LS_MORE_TORALDO, ftol=1e-3)
    assert ls.start(1.0, -1.0, 1.0, 2.0, 1.0) == O.TASK_ERROR and ls.reason() == "STPMAX < STPMIN"
    assert ls.start(1.0, +1.0, 1.0, 0.0, 2.0) == O.TASK_ERROR
    assert ls.start(1.0, -1.0, 3.0, 0.0, 2.0) == O.TASK_ERROR and ls.reason() == "STP > STPMAX"
    # quadratic phi(t) = 1 - t + t^2: More-Toraldo interpolation
    phi = lambda t: 1 - t + t * t
    assert ls.start(phi(0), -1.0, 4.0, 0.0, 10.0) == O.TASK_SEARCH and not ls.usederivatives
    assert ls.iterate(4.0, phi(4.0), 0.0) == O.TASK_SEARCH
    # q = 4, r = f - (f0 - q) = 13 - (1 - 4) = 16, q/(2r) = 0.125 in [0.1, 0.5]
    assert ls.stp == pytest.approx(0.5)
    assert ls.iterate(ls.stp, phi(ls.stp), 0.0) == O.TASK_CONVERGENCE
    # Armijo halves the step
    ls = O.LineSearch(O.LS_ARMIJO, ftol=1e-4)
    ls.start(phi(0), -1.0, 4.0, 0.0, 10.0)
    assert ls.iterate(4.0, phi(4.0), 0.0) == O.TASK_SEARCH and ls.stp == 2.0
    assert ls.iterate(2.0, phi(2.0), 0.0) == O.TASK_SEARCH and ls.stp == 1.0
    assert ls.iterate(1.0, phi(1.0), 0.0) == O.TASK_SEARCH and ls.stp == 0.5
    assert ls.iterate(0.5, phi(0.5), 0.0) == O.TASK_CONVERGENCE
    # More-Thuente finds a strong-Wolfe point of the quadratic (minimum 0.5)
    dphi = lambda t: -1 + 2 * t
    ls = O.LineSearch(O.LS_MORE_THUENTE, ftol=1e-3, gtol=0.1, xtol=0.1)
    assert ls.usederivatives
    task = ls.start(phi(0), dphi(0), 4.0, 0.0, 1e20)
    n = 0
    while task == O.TASK_SEARCH:
        t = ls.stp
        task = ls.iterate(t, phi(t), dphi(t))
        n += 1
        assert n < 20
    assert task == O.TASK_CONVERGENCE
    assert phi(t) <= phi(0) + 1e-3 * t * dphi(0) and abs(dphi(t)) <= 0.1 * abs(dphi(0))
    # iterate! with a step different from getstep is the reference's @assert
    ls = O.LineSearch(O.LS_ARMIJO)
    ls.start(1.0, -1.0, 1.0, 0.0, 2.0)
    assert ls.iterate(0.7, 1.0, 0.0) == O.TASK_ERROR


@pytest.mark.parametrize("dtype", DTYPES)
def test_quadratic_box_solution(oracle, dtype):
    # separable quadratic in a box: x* = clamp(c, lo, hi)
    rng = np.random.default_rng(5)
    n = 1000
    a = (10.0 ** (3 * rng.random(n))).astype(dtype)
    c = (-1 + 3 * rng.random(n)).astype(dtype)
    x0 = np.full(n, 0.5, dtype=dtype)
    lo, hi = np.zeros(n, dtype=dtype), np.ones(n, dtype=dtype)
    x, res, tr = oracle.vmlmb(("quadratic", a, c), x0, lo, hi, mem=5, trace=True,
                              maxiter=500, ftol=(0, 0), xtol=(0, 0), gtol=(0, 1e-10))
    want = np.clip(c, 0, 1)
    tol = 1e-6 if dtype == np.float64 else 2e-3
    assert np.abs(x - want).max() < tol
    xs, info, _ = oracle.spg(("quadratic", a, c), x0, 10, lo, hi, maxit=5000)
    assert np.abs(xs - want).max() < (1e-4 if dtype == np.float64 else 2e-3)


def test_python_callable_objective(oracle):
    def fg(x, g):
        g[:] = 2 * (x - 3)
        return float(((x - 3) ** 2).sum())
    x, res, _ = oracle.vmlmb(fg, np.zeros(7))
    assert np.allclose(x, 3.0, atol=1e-6)
    x, res, _ = oracle.vmlmb(fg, np.zeros(7), upper=2.0)
    assert np.allclose(x, 2.0)


def test_oracle_reproduces_golden(oracle):
    """tests/golden/*.json were generated from this oracle (make_golden.py);
    the GPU suite compares the CUDA path with the same files."""
    import json
    import os
    from helpers import problem
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    with open(os.path.join(gdir, "vmlmb_traces.json")) as fh:
        gold = json.load(fh)
    for case in gold["cases"]:
        pb = problem(case["kind"], case["n"], np.dtype(case["dtype"]))
        x, res, tr = oracle.vmlmb(pb["oracle_obj"], pb["x0"], pb["lower"], pb["upper"],
                                  trace=True, maxiter=case["maxiter"], **case["kw"])
        assert len(tr) == len(case["trace"])
        for t, w in zip(tr, case["trace"]):
            assert (t["iter"], t["eval"], t["f"], t["gnorm"], t["stp"]) == \
                (w["iter"], w["eval"], w["f"], w["gnorm"], w["stp"])
            assert int(t["mask"].sum()) == w["nfree"]
    with open(os.path.join(gdir, "spg_traces.json")) as fh:
        gold = json.load(fh)
    for case in gold["cases"]:
        pb = problem(case["kind"], case["n"], np.dtype(case["dtype"]))
        x, info, tr = oracle.spg(pb["oracle_obj"], pb["x0"], case["m"], pb["lower"], pb["upper"],
                                 trace=True, maxit=case["maxit"])
        assert (info["iter"], info["fcnt"], info["status"]) == \
            (case["final"]["iter"], case["final"]["fcnt"], case["final"]["status"])
        for t, w in zip(tr, case["trace"]):
            assert (t["f"], t["pgtwon"], t["pginfn"]) == (w["f"], w["pgtwon"], w["pginfn"])


def test_float32_trajectories_depend_on_the_reduction_order(oracle):
    """What the 1e-4 Float32 gate can mean.  The reference accumulates its dot
    products in the element type (a generic Julia loop, or BLAS sdot: SURVEY 8c);
    the canonical oracle accumulates exact products in Float64.  On the chaotic
    Rosenbrock trajectory the three summation modes of the oracle itself drift
    apart by more than 1e-3 in x within 20 iterations, i.e. by MORE than the
    Gram-form two-loop of the GPU path deviates from the canonical oracle
    (~2e-3, tests/test_gpu_vmlmb.py).  The Float32 gate is therefore applied
    against ONE fixed reduction order (the canonical one), and the convex
    quadratic, which is not chaotic, stays within 1e-4 in every mode."""
    from helpers import problem

    def run(kind, n, mode, mem):
        oracle.set_sum_mode(mode)
        try:
            pb = problem(kind, n, np.float32)
            return oracle.vmlmb(pb["oracle_obj"], pb["x0"], pb["lower"], pb["upper"], mem=mem, maxiter=25,
                                trace=True)[2]
        finally:
            oracle.set_sum_mode(oracle.SUM_ACCURATE)

    def drift(tr, ref, it):
        a = [t for t in tr if t["iter"] == it][0]
        b = [t for t in ref if t["iter"] == it][0]
        return np.linalg.norm(a["x"] - b["x"]) / np.linalg.norm(b["x"])

    ref = run("rosen_box", 1000, oracle.SUM_ACCURATE, 7)
    seq = run("rosen_box", 1000, oracle.SUM_SEQUENTIAL, 7)
    pw = run("rosen_box", 1000, oracle.SUM_PAIRWISE, 7)
    assert drift(seq, ref, 5) < 1e-4 and drift(pw, ref, 5) < 1e-4          # same algorithm ...
    assert drift(seq, ref, 20) > 1e-3 and drift(pw, ref, 20) > 1e-4        # ... chaotic amplification
    refq = run("quad_box", 100001, oracle.SUM_ACCURATE, 5)
    seqq = run("quad_box", 100001, oracle.SUM_SEQUENTIAL, 5)
    assert drift(seqq, refq, 20) < 1e-4
