"""SPG parity on the GPU: fused one-pass kernels and the one-kernel-per-call
path against the oracle's `_spg!` (src/spg.jl:193-403)."""
import json
import math
import os

import numpy as np
import pytest

from helpers import problem, relerr

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def opk():
    import opkb200
    return opkb200


def run_gpu(opk, pb, m, mode, **kw):
    lo = -math.inf if pb["lower"] is None else pb["lower"]
    hi = math.inf if pb["upper"] is None else pb["upper"]
    trace = []

    def obs(s):
        trace.append(dict(iter=s.iter, fcnt=s.fcnt, pcnt=s.pcnt, f=s.f, pgtwon=s.pgtwon, pginfn=s.pginfn))
    ws = opk.Info()
    s = opk.SPG(pb["gpu_obj"], opk.BoxProjection(lo, hi), pb["x0"], m, mode=mode, observer=obs, ws=ws, **kw)
    s.solve()
    x = s.result()
    s.close()
    return x, ws, trace


CASES = [
    ("rosen_lower0", 20, np.float64, 10),
    ("rosen_box", 1000, np.float64, 10),
    ("quad_box", 1003, np.float64, 10),
    ("quad_box", 100001, np.float64, 1),
    ("quad_upper_scalar", 5000, np.float64, 5),
    ("quad_box", 100001, np.float32, 10),
    ("rosen_box", 1000, np.float32, 10),
]


@pytest.mark.parametrize("mode", ["fused", "primitive"])
@pytest.mark.parametrize("kind,n,dtype,m", CASES)
def test_spg_trajectory(opk, oracle, kind, n, dtype, m, mode):
    pb = problem(kind, n, dtype)
    xo, io, tro = oracle.spg(pb["oracle_obj"], pb["x0"], m, pb["lower"], pb["upper"], trace=True, maxit=25)
    xg, ws, trg = run_gpu(opk, pb, m, mode, maxit=25)
    rtol = 1e-10 if dtype == np.float64 else 1e-4
    assert len(trg) == len(tro)
    for g, w in zip(trg, tro):
        if w["iter"] > 20:
            break
        assert (g["iter"], g["fcnt"], g["pcnt"]) == (w["iter"], w["fcnt"], w["pcnt"])
        assert abs(g["f"] - w["f"]) <= rtol * abs(w["f"]) + 1e-300, (w["iter"], g["f"], w["f"])
        assert abs(g["pgtwon"] - w["pgtwon"]) <= rtol * abs(w["pgtwon"]) + 1e-300
        assert abs(g["pginfn"] - w["pginfn"]) <= rtol * abs(w["pginfn"]) + 1e-300
    assert (ws.iter, ws.fcnt, ws.pcnt, ws.status) == (io["iter"], io["fcnt"], io["pcnt"], io["status"])
    assert relerr(xg, xo) <= rtol


@pytest.mark.parametrize("mode", ["fused", "primitive"])
def test_spg_converges_like_reference_test(opk, oracle, mode):
    # test/rosenbrock.jl:119-122 runs spg(rosenbrock_fg!, nonnegative!, x0, 10) (never asserted)
    x0 = opk.Rosenbrock.x0(20, np.float64)
    ws = opk.Info()
    x = opk.spg(opk.Rosenbrock(), opk.BoxProjection(0.0, math.inf), x0, 10, ws=ws, mode=mode)
    assert np.abs(x - 1).max() < 1e-3
    assert (ws.iter, ws.fcnt, ws.pcnt, ws.status) == (64, 82, 130, opk.INFNORM_CONVERGENCE)
    assert opk.getreason(ws) == "Convergence with projected gradient infinite-norm"


def test_spg_user_projection_and_objective(opk):
    ctx = opk.default_context()
    n = 500
    t = ctx.from_numpy(np.linspace(-1, 2, n))

    def fg(x, g):
        opk.vcombine_(g, 2, x, -2, t)
        r = opk.vcombine_(x.similar(), 1, x, -1, t)
        return opk.vdot(r, r)

    def nonnegative(dst, src):                       # test/rosenbrock.jl:50-58
        return opk.project_variables_(dst, src, 0.0, None)

    x = opk.spg(fg, nonnegative, np.full(n, 3.0), 10)
    assert np.abs(x - np.clip(np.linspace(-1, 2, n), 0, None)).max() < 1e-5
    ws = opk.Info()
    opk.spg(fg, nonnegative, np.full(n, 3.0), 10, maxfc=2, eps1=0, eps2=0, ws=ws)
    assert ws.status == opk.TOO_MANY_EVALUATIONS and ws.fcnt == 2


def test_spg_golden(opk):
    with open(os.path.join(GOLDEN, "spg_traces.json")) as fh:
        gold = json.load(fh)
    for case in gold["cases"]:
        dtype = np.dtype(case["dtype"])
        pb = problem(case["kind"], case["n"], dtype)
        rtol = 1e-10 if dtype == np.float64 else 1e-4
        for mode in ("fused", "primitive"):
            xg, ws, trg = run_gpu(opk, pb, case["m"], mode, maxit=case["maxit"])
            for g, w in zip(trg, case["trace"]):
                if w["iter"] > 20:
                    break
                assert (g["iter"], g["fcnt"]) == (w["iter"], w["fcnt"])
                assert abs(g["f"] - w["f"]) <= rtol * abs(w["f"]) + 1e-300
                assert abs(g["pgtwon"] - w["pgtwon"]) <= rtol * abs(w["pgtwon"]) + 1e-300


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_spg_generic_convex_set(opk, oracle, dtype):
    """`prj!` is ANY projector (src/spg.jl:64-78), not only a box: the Euclidean ball
    {|x| <= R}, written with the `v*` operations on device Vectors, against the oracle running
    the same projector on numpy arrays.  Objective: the built-in quadratic."""
    n, R, m = 4001, 40.0, 10
    a, c = opk.Quadratic.synthetic(n, dtype, kappa=1e3)
    x0 = np.full(n, 0.5, dtype=dtype)

    def ball(dst, src):                               # device Vectors
        nrm = opk.vnorm2(src)
        opk.vcopy_(dst, src)
        if nrm > R:
            opk.vscale_(dst, R / nrm)
        return dst

    def ball_np(dst, src):                            # numpy arrays, the oracle's vector operations
        nrm = oracle.vnorm2(src)
        oracle.vcopy(dst, src)
        if nrm > R:
            oracle.vscale(dst, R / nrm)

    xo, io, tro = oracle.spg(("quadratic", a, c), x0, m, prj=ball_np, maxit=30, trace=True)
    trace = []
    ws = opk.Info()
    s = opk.SPG(opk.Quadratic(a, c), ball, x0, m, maxit=30, ws=ws,
                observer=lambda s: trace.append((s.iter, s.fcnt, s.f, s.pgtwon)))
    s.solve()
    x = s.result()
    s.close()
    rtol = 1e-10 if dtype == np.float64 else 1e-4
    assert (ws.iter, ws.fcnt, ws.pcnt) == (io["iter"], io["fcnt"], io["pcnt"])
    for g, w in zip(trace, tro):
        if w["iter"] > 20:
            break
        assert (g[0], g[1]) == (w["iter"], w["fcnt"])
        assert abs(g[2] - w["f"]) <= rtol * abs(w["f"])
        if w["iter"] <= 10:      # (|pg| is a difference of O(R) terms: its rounding noise grows as it shrinks)
            assert abs(g[3] - w["pgtwon"]) <= 100 * rtol * w["pgtwon"]
    assert relerr(x, xo) <= 10 * rtol
    assert np.linalg.norm(x.astype(np.float64)) <= R * (1 + 1e-6)          # feasible
    assert np.linalg.norm(c.astype(np.float64)) > R                        # and the constraint is active


@pytest.mark.parametrize("kind,n,dtype,m", CASES)
def test_native_spg_driver_equals_python_driver(opk, oracle, kind, n, dtype, m):
    """`opkb_spg_solver_*` (the loop of `_spg!` in the library) takes the decisions of the Python
    host bit for bit: same kernels, same double arithmetic on the same scalars."""
    pb = problem(kind, n, dtype)
    lo = -math.inf if pb["lower"] is None else pb["lower"]
    hi = math.inf if pb["upper"] is None else pb["upper"]
    res = []
    for cls in (opk.SPG, opk.NativeSPG):
        tr = []
        ws = opk.Info()
        s = cls(pb["gpu_obj"], opk.BoxProjection(lo, hi), pb["x0"], m, maxit=30, ws=ws,
                observer=lambda s: tr.append((s.iter, s.fcnt, s.pcnt, s.f, s.fbest)))
        if cls is opk.NativeSPG:
            while s.iterate(1):
                pass
        else:
            s.solve()
        res.append((s.result().copy(), (ws.iter, ws.fcnt, ws.pcnt, ws.status, ws.f, ws.fbest, ws.pgtwon, ws.pginfn), tr))
        s.close()
    assert res[0][1] == res[1][1]
    assert np.array_equal(res[0][0], res[1][0])
    # the native observer runs after an iteration, the Python one at its printer site: the
    # (iter, fcnt) sequences coincide from the second entry on
    assert [t[:2] for t in res[1][2][:-1]] == [t[:2] for t in res[0][2][1:len(res[1][2])]]
    xo, io, _ = oracle.spg(pb["oracle_obj"], pb["x0"], m, pb["lower"], pb["upper"], maxit=30)
    assert (res[1][1][0], res[1][1][1], res[1][1][3]) == (io["iter"], io["fcnt"], io["status"])
    assert relerr(res[1][0], xo) <= (1e-10 if dtype == np.float64 else 1e-4)


def test_native_spg_driver_api(opk):
    a, c = opk.Quadratic.synthetic(5000, np.float64, kappa=1e2)
    box = opk.BoxProjection(0.0, 1.0)
    x0 = np.full(5000, 0.5)
    x1 = opk.spg(opk.Quadratic(a, c), box, x0, 10, maxit=40)
    x2 = opk.spg(opk.Quadratic(a, c), box, x0, 10, maxit=40, driver="native")
    assert np.array_equal(x1, x2)
    ws = opk.Info()
    s = opk.NativeSPG(opk.Quadratic(a, c), box, x0, 10, maxfc=3, eps1=0, eps2=0, ws=ws)
    s.solve()
    assert ws.status == opk.TOO_MANY_EVALUATIONS and ws.fcnt == 3
    s.close()
    s = opk.NativeSPG(opk.Quadratic(a, c), box, x0, 10, eps1=0, eps2=0)
    assert s.iterate(7) and s.iter == 7 and s.iterate(5) and s.iter == 12    # resumable
    s.close()
    with pytest.raises(TypeError):
        opk.NativeSPG(lambda x, g: 0.0, box, x0, 10)                          # user objective: Python loop
    with pytest.raises(opk.OpkbError):
        from opkb200 import _lib
        import ctypes as C
        sp = opk.SPG(opk.Quadratic(a, c), box, x0, 10)
        h = _lib.vp()
        try:
            _lib.check(_lib.lib().opkb_spg_solver_create(sp._h, 0, 10, 10, 0.0, 0.0, 1.0, C.byref(h)))  # m = 0
        finally:
            sp.close()
