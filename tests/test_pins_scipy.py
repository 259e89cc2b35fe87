"""An independent pin of the Moré-Thuente line search (SURVEY 8a L3,
src/linesearches.jl:413-647, `cstep` :701-870): the reference restates MINPACK-2
`dcsrch`/`dcstep`, and SciPy ships a line-by-line port of those two routines
(`scipy.optimize._dcsrch.DCSRCH`).  On randomly drawn 1-D functions the oracle's
state machine, the product's host-side `MoreThuenteLineSearch` and SciPy's DCSRCH
must propose bit-identical step sequences and stop at the same step."""
import math

import numpy as np
import pytest

from opkb200.linesearches import MoreThuenteLineSearch

DCSRCH = pytest.importorskip("scipy.optimize._dcsrch").DCSRCH


def _functions(rng):
    a, b, c = (float(v) for v in rng.normal(size=3))
    k = int(rng.integers(0, 3))
    a, b = abs(a), abs(b)
    if k == 0:     # shifted parabola times a quadratic modulation
        def phi(t):
            return (t - a - 0.3) ** 2 * (1 + 0.3 * b * t + 0.1 * c * c * t * t)

        def dphi(t):
            return (2 * (t - a - 0.3) * (1 + 0.3 * b * t + 0.1 * c * c * t * t)
                    + (t - a - 0.3) ** 2 * (0.3 * b + 0.2 * c * c * t))
    elif k == 1:   # Moré-Thuente's first test function, tilted
        def phi(t):
            return -t / (t * t + a + 0.1) + 0.01 * b * b * t

        def dphi(t):
            return (t * t - a - 0.1) / (t * t + a + 0.1) ** 2 + 0.01 * b * b
    else:          # damped oscillation in a bowl
        def phi(t):
            return -math.exp(-(a + 0.1) * t) * math.cos(3 * b * t) + 0.2 * t * t

        def dphi(t):
            e = math.exp(-(a + 0.1) * t)
            return e * ((a + 0.1) * math.cos(3 * b * t) + 3 * b * math.sin(3 * b * t)) + 0.4 * t
    return phi, dphi


def _scipy_steps(phi, dphi, stp0, ftol, gtol, xtol, stpmin, stpmax):
    d = DCSRCH(phi, dphi, ftol, gtol, xtol, stpmin, stpmax)
    steps, task, stp, f, g = [], b"START", stp0, phi(0.0), dphi(0.0)
    for _ in range(60):
        stp, f, g, task = d._iterate(stp, f, g, task)
        if task[:2] != b"FG":
            break
        steps.append(stp)
        f, g = phi(stp), dphi(stp)
    return steps, task[:4]


def _oracle_steps(O, phi, dphi, stp0, ftol, gtol, xtol, stpmin, stpmax):
    ls = O.LineSearch(O.LS_MORE_THUENTE, ftol=ftol, gtol=gtol, xtol=xtol)
    ls.start(phi(0.0), dphi(0.0), stp0, stpmin, stpmax)
    steps = []
    for _ in range(60):
        if ls.task != O.TASK_SEARCH:
            break
        steps.append(ls.stp)
        ls.iterate(ls.stp, phi(ls.stp), dphi(ls.stp))
    return steps, O.TASK_NAMES[ls.task]


def _host_steps(phi, dphi, stp0, ftol, gtol, xtol, stpmin, stpmax):
    ls = MoreThuenteLineSearch(ftol=ftol, gtol=gtol, xtol=xtol)
    task = ls.start(phi(0.0), dphi(0.0), stp0, stpmin=stpmin, stpmax=stpmax)
    steps = []
    for _ in range(60):
        if task != "SEARCH":
            break
        stp = ls.getstep()
        steps.append(stp)
        task = ls.iterate(stp, phi(stp), dphi(stp))
    return steps, task


def test_more_thuente_equals_minpack2_dcsrch(oracle):
    rng = np.random.default_rng(1)
    done = conv = 0
    for case in range(600):
        phi, dphi = _functions(rng)
        if not dphi(0.0) < 0:
            continue
        stp0 = float(10 ** rng.uniform(-2, 1))
        ftol, gtol, xtol = (1e-3, 0.9, 0.1) if case % 2 else (1e-4, 0.1, 1e-10)   # vmlmb's defaults / strict
        args = (phi, dphi, stp0, ftol, gtol, xtol, 0.0, stp0 * 1e3)
        s_ref, t_ref = _scipy_steps(*args)
        s_orc, t_orc = _oracle_steps(oracle, *args)
        s_hst, t_hst = _host_steps(*args)
        assert s_orc == s_ref, (case, s_orc[:5], s_ref[:5])     # bit-identical steps
        assert s_hst == s_ref, (case, s_hst[:5], s_ref[:5])
        assert (t_ref == b"CONV") == (t_orc == "CONVERGENCE") == (t_hst == "CONVERGENCE")
        done += 1
        conv += t_ref == b"CONV"
    assert done >= 300 and conv >= 250


def test_two_loop_equals_scipy_inverse_hessian_product(oracle):
    """K9 (`apply_lbfgs!`, src/quasinewton.jl:716-745) against SciPy's own two-loop
    recursion (`scipy.optimize.LbfgsInvHessProduct`, H0 = I): same H*v to rounding,
    whatever the position of `mark` in the circular pair storage."""
    from scipy.optimize import LbfgsInvHessProduct

    rng = np.random.default_rng(3)
    n, mem = 1000, 7
    for m in (1, 3, 7):
        S = [rng.normal(size=n) for _ in range(mem)]
        Y = [s * rng.uniform(0.5, 2.0, size=n) + 0.05 * rng.normal(size=n) for s in S]   # y.s > 0
        for mark in (m - 1, mem - 1, 1 if m > 1 else 0):
            # the m most recent pairs end at slot `mark`, oldest first for SciPy
            slots = [(mark - (m - 1) + j) % mem for j in range(m)]
            rho = np.array([float(np.dot(Y[k], S[k])) for k in range(mem)])
            v = rng.normal(size=n)
            want = LbfgsInvHessProduct(np.array([S[k] for k in slots]), np.array([Y[k] for k in slots])).matvec(v)
            got = v.copy()
            modif, _ = oracle.apply_lbfgs(S, Y, rho, 1.0, m, mark, got)
            assert modif
            assert np.linalg.norm(got - want) <= 1e-12 * np.linalg.norm(want)


# ---------------------------------------------------------------------------
# K11, method 0 (`_vmlmb!` without bounds = L-BFGS + More-Thuente,
# src/quasinewton.jl:381-601 with :273 method 0 and the defaults of :276-278)
# against SciPy's L-BFGS-B (the C translation of Zhu/Byrd/Lu/Nocedal's Fortran
# code): without bounds L-BFGS-B reduces to L-BFGS with `dcsrch` (ftol = 1e-3,
# gtol = 0.9, xtol = 0.1: the same three numbers as vmlmb's default line search), a
# first step 1/|g| then 1, and H0 = (s.y / y.y) I -- the same algorithm, coded
# independently (compact representation instead of the two-loop recursion).  The
# iterate sequences must agree to rounding: 1e-10 over the first 20 iterations, the
# north_star's own tolerance.
def _rosen_np(x):
    x1, x2 = x[0::2], x[1::2]
    g2 = 200 * (x2 - x1 * x1)
    g = np.empty_like(x)
    g[1::2] = g2
    g[0::2] = -2 * (x1 * g2 + (1 - x1))
    return float(np.sum((1 - x1) ** 2) + np.sum((10 * (x2 - x1 * x1)) ** 2)), g


def lbfgsb_trajectory(fg, x0, mem, niter):
    """[(f_k, x_k)] for k = 1..niter from scipy.optimize's L-BFGS-B, no bounds."""
    so = pytest.importorskip("scipy.optimize")
    tr = []
    so.minimize(fg, x0, jac=True, method="L-BFGS-B", callback=lambda xk: tr.append((fg(xk)[0], xk.copy())),
                options=dict(maxcor=mem, ftol=0.0, gtol=0.0, maxiter=niter, maxls=20))
    return tr


def lbfgsb_cases():
    import opkb200 as opk

    for n, mem in ((20, 5), (20, 15), (1000, 5), (1000, 7), (100000, 5)):
        x0 = opk.Rosenbrock.x0(n, np.float64, perturb_seed=7 if n > 20 else None)
        yield "rosen%d-m%d" % (n, mem), _rosen_np, "rosenbrock", opk.Rosenbrock(), x0, mem
    for n, mem in ((1003, 10), (5000, 3)):
        a, c = opk.Quadratic.synthetic(n, np.float64, kappa=1e3)
        yield ("quad%d-m%d" % (n, mem), lambda x, a=a, c=c: (float(0.5 * np.sum(a * (x - c) ** 2)), a * (x - c)),
               ("quadratic", a, c), opk.Quadratic(a, c), np.full(n, 0.5), mem)


def assert_follows_lbfgsb(trace, ref, niter=20, rtol=1e-10):
    trace = [t for t in trace if t["iter"] > 0]
    assert len(trace) >= niter and len(ref) >= niter
    for k in range(niter):
        f, x = ref[k]
        assert trace[k]["iter"] == k + 1
        assert abs(trace[k]["f"] - f) <= rtol * abs(f), (k, trace[k]["f"], f)
        assert np.linalg.norm(trace[k]["x"] - x) <= rtol * np.linalg.norm(x), k


NOSTOP = dict(xtol=(0.0, 0.0), ftol=(0.0, 0.0), gtol=(0.0, 0.0))


def test_unconstrained_vmlmb_follows_scipy_lbfgsb(oracle):
    from cpu_backend import NumpyOps

    import opkb200 as opk

    for name, fg, oobj, gobj, x0, mem in lbfgsb_cases():
        ref = lbfgsb_trajectory(fg, x0, mem, 25)
        _, ro, tro = oracle.vmlmb(oobj, x0, None, None, mem=mem, maxiter=25, trace=True, **NOSTOP)
        assert_follows_lbfgsb(tro, ref)
        if len(x0) > 5000:
            continue
        # the product's host stage machine (numpy vector ops: tests/cpu_backend.py)
        tr = []
        s = opk.VMLMB(gobj, x0, mem=mem, maxiter=25, mode=lambda solver, x, o=oobj: NumpyOps(solver, x, o),
                      observer=lambda s: tr.append(dict(iter=s.iter, f=s.f, x=np.array(s.x))), **NOSTOP)
        s.solve()
        assert_follows_lbfgsb(tr, ref)
