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
