"""Device-resident decisions of the native driver (csrc/opkb_chain.cuh): in steady state the
last CTA of the fused launch takes the host's decisions (src/quasinewton.jl:437-468, :533-594)
and solves the two-loop coefficients (:716-779) of the launch the host has enqueued ahead.

The gate: with the chain on, a run is the SAME BITS as with the chain off (iterates, counts, f,
|g|), launches are really adopted while in flight, every way out of the steady state (another
step length, convergence, maxiter / maxeval, run boundaries) hands over to the host cleanly, and
the trajectory agrees with the oracle like every other host."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import opkb200 as opk
from helpers import gpu_bounds, problem, relerr


CASES = [
    # kind, n, mem, kw
    ("rosen_unc", 20000, 5, {}),                 # C1's workload: L-BFGS, More-Thuente
    ("rosen_box", 30000, 5, {}),                 # C2's workload: VMLMB, array bounds, More-Toraldo
    ("quad_box", 25001, 8, {}),                  # 8 pairs, odd n (partial tile)
    ("quad_box", 4096, 3, {}),                   # the memory fills up after 3 iterations
    ("rosen_lower0", 5000, 7, {}),               # scalar bound
    ("quad_upper_scalar", 10000, 4, {}),
    # tiny problems run to (and past) convergence with the tolerances off: pairs with rho <= 0, zero steps,
    # everything on a bound -- the device must never say "go" where the host takes another path
    ("quad_box", 7, 3, {}),
    ("rosen_unc", 2, 2, {}),
    ("rosen_box", 6, 5, {}),
    ("quad_box", 64, 8, {}),
]


def _run(pb, mem, chain, plan, **kw):
    """NativeVMLMB run following `plan` (iterations per opkb_solver_run call); -> summary."""
    os.environ["OPKB_CHAIN_CHECK"] = "1"
    try:
        lo, hi = gpu_bounds(pb)
        s = opk.NativeVMLMB(pb["gpu_obj"], pb["x0"].copy(), lower=lo, upper=hi, mem=mem, chain=chain, **kw)
        assert s.chain_enabled == bool(chain)
        trace = []
        for k in plan:
            alive = s.iterate(k)
            trace.append((s.iter, s.eval, s.rejects, s.f, s.gnorm, s.stp))
            if not alive:
                break
        out = dict(trace=trace, x=np.array(s.result()), reason=s.reason, carry=s.carry_stats(),
                   chain=s.chain_stats(), finished=s.finished)
        s.close()
        return out
    finally:
        os.environ.pop("OPKB_CHAIN_CHECK", None)


def test_chain_switches():
    """The keyword, the environment switch, and solvers that cannot chain (not an error)."""
    pb = problem("quad_box", 4096, np.float64)
    lo, hi = gpu_bounds(pb)
    s = opk.NativeVMLMB(pb["gpu_obj"], pb["x0"].copy(), lower=lo, upper=hi, mem=5)
    assert not s.chain_enabled                      # opt-in
    s.close()
    os.environ["OPKB_CHAIN"] = "1"
    try:
        s = opk.NativeVMLMB(pb["gpu_obj"], pb["x0"].copy(), lower=lo, upper=hi, mem=5)
        assert s.chain_enabled
        s.close()
    finally:
        os.environ.pop("OPKB_CHAIN", None)
    for kw in (dict(mem=9), dict(mem=5, blmvm=True), dict(mem=5, twoloop="sequential")):
        s = opk.NativeVMLMB(pb["gpu_obj"], pb["x0"].copy(), lower=lo, upper=hi, chain=True, **kw)
        assert not s.chain_enabled, kw
        s.iterate(12)
        assert s.chain_stats() == (0, 0, 0)
        s.close()
    pb32 = problem("quad_box", 4096, np.float32)
    s = opk.NativeVMLMB(pb32["gpu_obj"], pb32["x0"].copy(), lower=lo.astype(np.float32), upper=hi.astype(np.float32),
                        mem=5, chain=True, twoloop="gram")
    assert not s.chain_enabled
    s.close()


@pytest.mark.parametrize("kind,n,mem,kw", CASES)
def test_chain_is_bit_identical(kind, n, mem, kw):
    pb = problem(kind, n, np.float64)
    tol = dict(xtol=(0, 0), ftol=(0, 0), gtol=(0, 0))
    for plan in ([30], [mem + 3, 1, 1, 9, 2, 14], [1] * 12):
        a = _run(pb, mem, True, plan, **tol, **kw)
        b = _run(pb, mem, False, plan, **tol, **kw)
        assert a["trace"] == b["trace"], (plan, a["trace"], b["trace"])     # bit for bit
        assert np.array_equal(a["x"], b["x"]), plan
        assert b["chain"] == (0, 0, 0)
        if plan == [30] and n >= 1000:
            # launches were adopted while already in flight (otherwise this test proves nothing)
            assert a["chain"][1] >= 5, a["chain"]
        if plan == [1] * 12:
            assert a["chain"][1] == 0, a["chain"]      # nothing is enqueued beyond a run


@pytest.mark.parametrize("kind,n,mem", [("rosen_unc", 2000, 5), ("rosen_box", 3000, 5), ("quad_box", 5000, 8)])
def test_chain_to_convergence_and_limits(kind, n, mem):
    pb = problem(kind, n, np.float64)
    # default tolerances: the convergence tests end the chain
    a = _run(pb, mem, True, [10 ** 6])
    b = _run(pb, mem, False, [10 ** 6])
    assert a["finished"] and b["finished"] and a["reason"] == b["reason"], (a["reason"], b["reason"])
    assert a["trace"] == b["trace"] and np.array_equal(a["x"], b["x"])
    assert a["chain"][1] > 0
    # maxiter / maxeval in the middle of the steady state
    for lim in (dict(maxiter=17), dict(maxeval=23)):
        tol = dict(xtol=(0, 0), ftol=(0, 0), gtol=(0, 0))
        a = _run(pb, mem, True, [10 ** 6], **tol, **lim)
        b = _run(pb, mem, False, [10 ** 6], **tol, **lim)
        assert a["finished"] and a["reason"] == b["reason"], (lim, a["reason"], b["reason"])
        assert a["trace"] == b["trace"] and np.array_equal(a["x"], b["x"]), lim


@pytest.mark.parametrize("kind,n,mem", [("rosen_unc", 20000, 5), ("rosen_box", 30000, 5), ("quad_box", 25001, 8)])
def test_chain_follows_the_oracle(oracle, kind, n, mem):
    pb = problem(kind, n, np.float64)
    tol = dict(xtol=(0, 0), ftol=(0, 0), gtol=(0, 0))
    a = _run(pb, mem, True, [10 ** 6], maxiter=20, **tol)
    xo, ro, _ = oracle.vmlmb(pb["oracle_obj"], pb["x0"].copy(), pb["lower"], pb["upper"], mem=mem, maxiter=20,
                             xtol=(0, 0), ftol=(0, 0), gtol=(0, 0))
    it, ev, rej, f, gn, _ = a["trace"][-1]
    assert (it, ev, rej) == (ro["iter"], ro["eval"], ro["rejects"]), ((it, ev, rej), ro)
    assert abs(f - ro["f"]) <= 1e-10 * abs(ro["f"])
    assert relerr(a["x"], xo) <= 1e-10
    assert a["chain"][1] >= 3, a["chain"]


def test_chain_self_check_is_clean():
    """OPKB_CHAIN_CHECK: whenever the host asks for the next direction right after a chained launch,
    the record of that launch's last CTA holds the same decision and the same coefficients."""
    import ctypes as C

    from opkb200 import _lib
    os.environ["OPKB_CHAIN"] = "1"
    os.environ["OPKB_CHAIN_CHECK"] = "1"
    try:
        for kind, n, mem in (("rosen_box", 30000, 5), ("rosen_unc", 20000, 5), ("quad_box", 25001, 8)):
            pb = problem(kind, n, np.float64)
            lo, hi = gpu_bounds(pb)
            s = opk.NativeVMLMB(pb["gpu_obj"], pb["x0"].copy(), lower=lo, upper=hi, mem=mem, xtol=(0, 0),
                                ftol=(0, 0), gtol=(0, 0))
            for _ in range(40):
                s.iterate(1)          # head launches only: every record is checked by the next call
            a, b = _lib.i64(), _lib.i64()
            _lib.check(_lib.lib().opkb_vmlmb_chain_check_stats(s._ws, C.byref(a), C.byref(b)))
            s.close()
            assert a.value >= 10 and b.value == 0, (kind, a.value, b.value)
    finally:
        os.environ.pop("OPKB_CHAIN", None)
        os.environ.pop("OPKB_CHAIN_CHECK", None)
