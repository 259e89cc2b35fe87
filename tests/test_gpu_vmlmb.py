"""VMLMB parity on the GPU (north_star gate): same iterate and active-set
sequence as the oracle for the first 20 iterations, f, |g| and x within 1e-10
relative in Float64 (1e-4 in Float32), active-set mask bit-exact; both for the
one-kernel-per-call path (mode='primitive') and for the fused P1/P3/P4 path."""
import json
import math
import os

import numpy as np
import pytest

from helpers import compare_traces, gpu_bounds, problem, record_vmlmb, relerr

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

F64_RTOL = 1e-10   # north_star tolerance, Float64
F32_RTOL = 1e-4    # north_star tolerance, Float32


@pytest.fixture(scope="module")
def opk():
    import opkb200
    return opkb200


def run_gpu(opk, pb, mode, **kw):
    lo, hi = gpu_bounds(pb)
    trace = []
    s = opk.VMLMB(pb["gpu_obj"], pb["x0"], lower=lo, upper=hi, mode=mode,
                  observer=record_vmlmb(trace), **kw)
    s.solve()
    x = s.result()
    info = dict(iter=s.iter, eval=s.eval, rejects=s.rejects, f=s.f, gnorm=s.gnorm, stage=s.stage,
                reason=s.reason)
    s.close()
    return x, info, trace


def run_oracle(oracle, pb, **kw):
    return oracle.vmlmb(pb["oracle_obj"], pb["x0"], pb["lower"], pb["upper"], trace=True, **kw)


CASES = [
    # (problem kind, n, dtype, solver keywords)
    ("rosen_unc", 20, np.float64, dict()),
    ("rosen_unc", 20, np.float64, dict(mem=15)),
    ("rosen_unc", 1000, np.float64, dict()),
    ("rosen_lower0", 20, np.float64, dict()),
    ("rosen_lower0", 20, np.float64, dict(blmvm=True)),
    ("rosen_box", 1000, np.float64, dict()),
    ("rosen_box", 100000, np.float64, dict(mem=7)),
    ("rosen_box", 1000, np.float64, dict(blmvm=True)),
    ("quad_box", 1003, np.float64, dict(mem=10)),
    ("quad_box", 200001, np.float64, dict(mem=10)),
    ("quad_upper_scalar", 5000, np.float64, dict(mem=3)),
    ("rosen_unc", 20, np.float32, dict()),
    ("rosen_lower0", 20, np.float32, dict()),
    ("rosen_box", 1000, np.float32, dict(mem=7)),
    ("quad_box", 100001, np.float32, dict(mem=5)),
]


@pytest.mark.parametrize("kind,n,dtype,kw", CASES)
def test_trajectory_primitive(opk, oracle, kind, n, dtype, kw):
    pb = problem(kind, n, dtype)
    kw = dict(kw, maxiter=25)
    xo, ro, tro = run_oracle(oracle, pb, **kw)
    xg, rg, trg = run_gpu(opk, pb, "primitive", **kw)
    rtol = F64_RTOL if dtype == np.float64 else F32_RTOL
    compare_traces(trg, tro, 20, rtol)


@pytest.mark.parametrize("kind,n,dtype,kw", CASES)
def test_trajectory_fused(opk, oracle, kind, n, dtype, kw):
    pb = problem(kind, n, dtype)
    kw = dict(kw, maxiter=25)
    xo, ro, tro = run_oracle(oracle, pb, **kw)
    xg, rg, trg = run_gpu(opk, pb, "fused", **kw)
    rtol = F64_RTOL if dtype == np.float64 else F32_RTOL
    compare_traces(trg, tro, 20, rtol)


@pytest.mark.parametrize("kind,n,dtype,kw", [c for c in CASES if c[2] == np.float64])
def test_trajectory_fused_sequential_float64(opk, oracle, kind, n, dtype, kw):
    """The fused step-by-step two-loop recursion (default for Float32) in Float64."""
    pb = problem(kind, n, dtype)
    kw = dict(kw, maxiter=25)
    xo, ro, tro = run_oracle(oracle, pb, **kw)
    xg, rg, trg = run_gpu(opk, pb, "fused", twoloop="sequential", **kw)
    compare_traces(trg, tro, 20, F64_RTOL)


@pytest.mark.parametrize("kind,n,dtype,kw", [c for c in CASES if c[2] == np.float32])
def test_trajectory_fused_gram_float32(opk, oracle, kind, n, dtype, kw):
    """Float32 with the Gram-form two-loop (opt-in): the quadratic stays within
    1e-4 (a variable in 10^4 may sit on the other side of its bound for an
    iteration: the roundings of the intermediate vector are missing); on the
    chaotic Rosenbrock they grow to ~2e-3 by iteration 20 (documented in
    DESIGN.md), hence the parity-grade default for Float32 is twoloop='sequential'."""
    pb = problem(kind, n, dtype)
    kw = dict(kw, maxiter=25, twoloop="gram")
    xo, ro, tro = run_oracle(oracle, pb, **{k: v for k, v in kw.items() if k != "twoloop"})
    xg, rg, trg = run_gpu(opk, pb, "fused", **kw)
    if kind.startswith("quad"):
        compare_traces(trg, tro, 20, F32_RTOL, mask_slack=1e-4)
    else:
        compare_traces(trg, tro, 20, 1e-2, check_mask=False)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kw", [dict(), dict(mem=15), dict(lower=0.0)], ids=["default", "mem15", "lower0"])
@pytest.mark.parametrize("mode", ["fused", "primitive"])
def test_reference_assertions(opk, dtype, kw, mode):
    # test/rosenbrock.jl:76-79, :89-92, :101-104
    x0 = opk.Rosenbrock.x0(20, dtype)
    x = opk.vmlmb(opk.Rosenbrock(), x0, mode=mode, **kw)
    assert x.dtype == dtype and np.abs(x - 1).max() < 1e-3
    assert x0[0] == dtype(-1.2)  # x0 left unchanged (src/quasinewton.jl:215)


@pytest.mark.parametrize("mode", ["fused", "primitive"])
def test_converged_solution_and_counts(opk, oracle, mode):
    pb = problem("rosen_unc", 20, np.float64)
    xo, ro, _ = run_oracle(oracle, pb)
    xg, rg, _ = run_gpu(opk, pb, mode)
    assert (rg["iter"], rg["eval"], rg["rejects"]) == (ro["iter"], ro["eval"], ro["rejects"]) == (35, 49, 0)
    assert rg["reason"] == ro["reason"] and rg["stage"] == 3
    assert np.abs(xg - xo).max() < 1e-6
    pb = problem("quad_box", 2000, np.float64)
    # gtol within reach of Float64 (the objective's rounding noise stalls the line
    # search near |p| ~ 2e-5 on this problem): a proper convergence, then every
    # variable is within |p| / min(a) of the solution clip(c, 0, 1) (a >= 1)
    kw = dict(mem=10, ftol=(0, 0), xtol=(0, 0), gtol=(0, 1e-8), maxiter=400)
    xg, rg, _ = run_gpu(opk, pb, mode, **kw)
    a, c = pb["oracle_obj"][1:]
    assert rg["stage"] == 3 and rg["reason"] == "projected gradient norm sufficiently small"
    assert np.abs(xg - np.clip(c, 0, 1)).max() <= 1.001 * rg["gnorm"] < 2e-4


def test_golden_traces(opk):
    """Committed oracle traces (tests/golden/make_golden.py) vs the GPU."""
    with open(os.path.join(GOLDEN, "vmlmb_traces.json")) as fh:
        gold = json.load(fh)
    for case in gold["cases"]:
        dtype = np.dtype(case["dtype"])
        pb = problem(case["kind"], case["n"], dtype)
        for mode in ("primitive", "fused"):
            xg, rg, trg = run_gpu(opk, pb, mode, maxiter=case["maxiter"], **case["kw"])
            rtol = F64_RTOL if dtype == np.float64 else F32_RTOL
            for g, w in zip(trg, case["trace"]):
                if w["iter"] > 20:
                    break
                assert (g["iter"], g["eval"]) == (w["iter"], w["eval"])
                assert abs(g["f"] - w["f"]) <= rtol * abs(w["f"]) + 1e-300
                assert abs(g["gnorm"] - w["gnorm"]) <= rtol * abs(w["gnorm"]) + 1e-300
                assert int(g["mask"].sum()) == w["nfree"]
                assert relerr(g["x"][:8], w["x_head"]) <= rtol


def test_user_objective_on_device(opk):
    """fg(x, g) on device Vectors: the reference's fg! contract (src/quasinewton.jl:388-392)."""
    ctx = opk.default_context()
    n = 1000
    t = ctx.from_numpy(np.linspace(-1, 2, n))

    def fg(x, g):
        opk.vcombine_(g, 2, x, -2, t)          # g = 2 (x - t)
        r = opk.vcombine_(x.similar(), 1, x, -1, t)
        return opk.vdot(r, r)

    for mode in ("fused", "primitive"):
        x = opk.vmlmb(fg, np.zeros(n), mode=mode)
        assert np.abs(x - np.linspace(-1, 2, n)).max() < 1e-5
        x = opk.vmlmb(fg, np.zeros(n), lower=opk.BoundedSet(0.0, 1.0), mode=mode)
        assert np.abs(x - np.clip(np.linspace(-1, 2, n), 0, 1)).max() < 1e-5


def test_linesearch_options_and_limits(opk, oracle):
    pb = problem("rosen_lower0", 20, np.float64)
    for ls, code in ((opk.ArmijoLineSearch(ftol=1e-4), oracle.LS_ARMIJO),
                     (opk.MoreThuenteLineSearch(ftol=1e-3, gtol=0.9, xtol=0.1), oracle.LS_MORE_THUENTE)):
        kw = dict(lnsrch=code)
        if code == oracle.LS_ARMIJO:
            kw["ls_ftol"] = 1e-4
        xo, ro, tro = oracle.vmlmb("rosenbrock", pb["x0"], 0.0, None, trace=True, maxiter=20, **kw)
        trace = []
        s = opk.VMLMB(opk.Rosenbrock(), pb["x0"], lower=0.0, lnsrch=ls, maxiter=20,
                      observer=record_vmlmb(trace))
        s.solve()
        compare_traces(trace, tro, 20, F64_RTOL)
        assert s.reason == ro["reason"]
    # maxeval / maxiter stops and the revert to the best point (:485-497)
    xo, ro, _ = oracle.vmlmb("rosenbrock", pb["x0"], maxeval=7)
    s = opk.VMLMB(opk.Rosenbrock(), pb["x0"], maxeval=7)
    s.solve()
    assert (s.eval, s.iter, s.reason) == (ro["eval"], ro["iter"], ro["reason"])
    assert relerr(s.result(), xo) < 1e-12
    # any `mem` (src/quasinewton.jl:227, :328): beyond the 20 pairs of the Gram-form kernels the fused
    # path runs the step-by-step recursion (tests/test_gpu_api.py::test_mem_above_the_gram_limit)
    for kw in (dict(), dict(mode="primitive"), dict(driver="native")):
        x = opk.vmlmb(opk.Rosenbrock(), opk.Rosenbrock.x0(100), mem=64, **kw)
        assert np.abs(x - 1).max() < 1e-3


@pytest.mark.parametrize("history", [0, 1], ids=["stored-differences", "iterate-history"])
def test_phase_kernels_single_step(opk, oracle, history, monkeypatch):
    """P3/P4 from an identical state: Gram block vs numpy, direction vs the
    oracle's apply_lbfgs! on the free set (tight tolerance, no chaos) -- for both
    representations of the pairs (include/opkb200.h, opkb_vmlmb_history): the stored
    differences of the reference, and the iterate history whose differences the ring
    kernels form on the fly."""
    import ctypes as C
    from opkb200 import _lib
    monkeypatch.setenv("OPKB_HISTORY", str(2 * history))  # (read at workspace creation; 2: also for Float32)
    rng = np.random.default_rng(3)
    ctx = opk.default_context()
    L = _lib.lib()
    for dtype, n, mem, m in ((np.float64, 5003, 10, 10), (np.float64, 777, 7, 4), (np.float32, 4099, 7, 7),
                             (np.float64, 3000, 20, 13), (np.float64, 300, 3, 3)):
        lo = np.zeros(n, dtype=dtype)
        hi = np.ones(n, dtype=dtype)
        a, c = opk.Quadratic.synthetic(n, dtype)
        s = opk.VMLMB(opk.Quadratic(a, c), rng.random(n).astype(dtype), lower=lo, upper=hi, mem=mem)
        ops = s.ops
        x = rng.uniform(0, 1, n).astype(dtype)
        x[rng.random(n) < 0.2] = 0
        x[rng.random(n) < 0.2] = 1
        g = rng.standard_normal(n).astype(dtype)
        p = oracle.project_direction(np.empty_like(g), x, lo, hi, -1, g)
        ops.vector("x").upload(x)
        ops.vector("g").upload(g)
        ops.vector("p").upload(p)
        S = [rng.standard_normal(n).astype(dtype) for _ in range(mem)]
        Y = [(1.5 * S[k] + 0.3 * rng.standard_normal(n)).astype(dtype) for k in range(mem)]
        mark = m - 1
        x0 = (x + S[mark]).astype(dtype)
        g0 = (g + Y[mark]).astype(dtype)
        if history:
            # slots 0 .. m-1 hold the ITERATES of m consecutive line searches, the current (x, g)
            # ends the newest pair: the pairs are the rounded differences of neighbours
            X = [None] * m + [x]
            Gr = [None] * m + [g]
            for k in range(m - 1, -1, -1):
                X[k] = (X[k + 1] + S[k]).astype(dtype)
                Gr[k] = (Gr[k + 1] + Y[k]).astype(dtype)
            for k in range(m):
                ops.vector("S", k).upload(X[k])
                ops.vector("Y", k).upload(Gr[k])
                S[k] = X[k] - X[k + 1]
                Y[k] = Gr[k] - Gr[k + 1]
            x0, g0 = X[mark], Gr[mark]
        else:
            for k in range(mem):
                ops.vector("S", k).upload(x0 if k == mark else S[k])
                ops.vector("Y", k).upload(g0 if k == mark else Y[k])
            S[mark] = x0 - x
            Y[mark] = g0 - g
        slots = list(range(m))
        G = (C.c_double * (2 * m * (m + 1)))()
        _lib.check(L.opkb_vmlmb_gram(s._ws, mark, m, (C.c_int * m)(*slots), G))
        G = np.array(G).reshape(2 * m, m + 1)
        assert s.history == history
        if history:      # nothing is written back: the slot still holds (x0, g0)
            assert np.array_equal(ops.vector("S", mark).download(), x0)
            assert np.array_equal(ops.vector("Y", mark).download(), g0)
        else:
            assert np.array_equal(ops.vector("S", mark).download(), S[mark])
            assert np.array_equal(ops.vector("Y", mark).download(), Y[mark])
        free = p != 0
        # Float64: exact products, double accumulation.  Float32: the products of one
        # stage (<= 32 per lane) are summed with FFMA before the double accumulation
        gtol = dict(rel=1e-12, abs=1e-9) if dtype == np.float64 else dict(rel=2e-6, abs=2e-3)
        Sd = [v.astype(np.float64) for v in S]
        Yd = [v.astype(np.float64) for v in Y]
        pd = p.astype(np.float64)
        for a in range(m):
            assert G[2 * a, 0] == pytest.approx(Sd[a][free] @ pd[free], **gtol)
            assert G[2 * a + 1, 0] == pytest.approx(Yd[a][free] @ pd[free], **gtol)
            for b in range(a, m):
                assert G[2 * a, 1 + b] == pytest.approx(Sd[a][free] @ Yd[b][free], **gtol)
                assert G[2 * a + 1, 1 + b] == pytest.approx(Yd[a][free] @ Yd[b][free], **gtol)
        # direction from the coefficients vs the oracle's sequential two-loop
        alpha, c, gamma, change = ops._solve(list(G.reshape(-1)), m, slots)
        assert change
        rho = np.zeros(mem)
        v = p.copy()
        sel = np.flatnonzero(free)
        modif, _ = oracle.apply_lbfgs_sel([S[k] for k in range(mem)], [Y[k] for k in range(mem)], rho, m, mark, v, sel)
        assert modif
        want = oracle.project_direction(np.empty_like(v), x, lo, hi, -1, v)
        out = (C.c_double * 4)()
        mark_next = (mark + 1) % mem
        _lib.check(L.opkb_vmlmb_direction(s._ws, m, (C.c_int * m)(*slots), (C.c_double * m)(*alpha),
                                          (C.c_double * m)(*c), gamma, mark_next, out))
        d = ops.vector("d").download()
        tol = 1e-11 if dtype == np.float64 else 2e-5
        assert relerr(d, want) < tol
        assert np.array_equal(d != 0, want != 0)
        assert out[0] == pytest.approx(float(g.astype(np.float64) @ d.astype(np.float64)), rel=1e-12, abs=1e-9)
        assert out[1] == pytest.approx(float(d.astype(np.float64) @ d.astype(np.float64)), rel=1e-12)
        smin, smax = oracle.step_limits(x, lo, hi, -1, d)
        assert (out[2], out[3]) == (smin, smax)
        assert np.array_equal(ops.vector("S", mark_next).download(), x)
        assert np.array_equal(ops.vector("Y", mark_next).download(), g)
        s.close()


@pytest.mark.parametrize("dtype,n,mem", [(np.float64, 1 << 24, 10), (np.float32, 1 << 24, 7)])
def test_large_properties(opk, dtype, n, mem):
    """Size-independent properties at a size the oracle is not run at: the
    iterates stay feasible, f decreases monotonically (Armijo), the run is
    bit-reproducible, and fused == primitive to rounding."""
    ctx = opk.default_context()
    obj = opk.Quadratic.synthetic_device(ctx, n, dtype, kappa=1e4)
    lo = ctx.vector(n, dtype).fill(0.0)
    hi = ctx.vector(n, dtype).fill(1.0)
    x0 = ctx.vector(n, dtype).fill(0.5)
    runs = []
    for mode in ("fused", "fused", "primitive"):
        fs = []
        s = opk.VMLMB(obj, opk.vcopy(x0), lower=lo, upper=hi, mem=mem, maxiter=12, mode=mode,
                      xtol=(0, 0), ftol=(0, 0), gtol=(0, 0), observer=lambda s: fs.append((s.iter, s.f, s.gnorm)))
        s.solve()
        x = s.result()
        assert opk.vnorminf(x) <= 1.0
        below = opk.vcombine_(x.similar(), 1, x, 0, x)
        opk.project_variables_(below, x, 0.0, 1.0)
        assert opk.vnorm2(opk.vcombine_(below, 1, below, -1, x)) == 0.0   # feasible
        assert all(b[1] < a[1] for a, b in zip(fs, fs[1:]))
        runs.append(fs)
        s.close()
    assert runs[0] == runs[1]                                             # deterministic
    rtol = 1e-9 if dtype == np.float64 else 1e-3
    for a, b in zip(runs[0], runs[2]):
        assert a[0] == b[0] and abs(a[1] - b[1]) <= rtol * abs(b[1])


# ---------------------------------------------------------------------------
# native reverse-communication driver (opkb_solver_*)
@pytest.mark.parametrize("kind,n,dtype,kw", CASES)
def test_native_driver_trajectory(opk, oracle, kind, n, dtype, kw):
    """The driver loop inside libopkb200 gives the oracle's trajectory (same
    gates as the Python twin); iterate(1) stops at every iteration boundary."""
    pb = problem(kind, n, dtype)
    kw = dict(kw, maxiter=25)
    xo, ro, tro = run_oracle(oracle, pb, **kw)
    lo, hi = gpu_bounds(pb)
    s = opk.NativeVMLMB(pb["gpu_obj"], pb["x0"], lower=lo, upper=hi, **kw)
    rtol = F64_RTOL if dtype == np.float64 else F32_RTOL
    by_iter = {t["iter"]: t for t in tro}
    seen = 0
    while s.iterate(1):
        w = by_iter.get(s.iter)
        if w is None or s.iter > 20:
            continue
        assert abs(s.f - w["f"]) <= rtol * abs(w["f"]) + 1e-300, (s.iter, s.f, w["f"])
        assert abs(s.gnorm - w["gnorm"]) <= rtol * abs(w["gnorm"]) + 1e-300
        xcur = s.current().download()
        assert relerr(xcur, w["x"]) <= rtol
        seen += 1
    assert seen >= 15
    assert (s.iter, s.eval, s.rejects, s.reason) == (ro["iter"], ro["eval"], ro["rejects"], ro["reason"])
    assert relerr(s.result(), xo) <= 10 * rtol
    s.close()


def test_native_driver_options_and_user_objective(opk, oracle):
    pb = problem("rosen_lower0", 20, np.float64)
    for kw in (dict(maxeval=7), dict(maxiter=3), dict(fmin=0.0), dict(epsilon=0.01), dict(gtol=(1e-3, 0.0)),
               dict(xtol=(0.0, 1e-2)), dict(ftol=(1e-4, 0.0)), dict(fmin=300.0), dict(blmvm=True), dict(mem=3)):
        xo, ro, _ = oracle.vmlmb("rosenbrock", pb["x0"], 0.0, None, **kw)
        x = opk.Rosenbrock.x0(20)
        s = opk.NativeVMLMB(opk.Rosenbrock(), x, lower=0.0, **kw)
        s.solve()
        assert (s.iter, s.eval, s.reason, s.stage) == (ro["iter"], ro["eval"], ro["reason"], ro["stage"]), kw
        assert relerr(s.result(), xo) <= 1e-9
        s.close()
    for ls, code in ((opk.ArmijoLineSearch(ftol=1e-4), oracle.LS_ARMIJO),
                     (opk.MoreThuenteLineSearch(ftol=1e-3, gtol=0.9, xtol=0.1), oracle.LS_MORE_THUENTE)):
        okw = dict(lnsrch=code, ls_ftol=ls.ftol)
        xo, ro, _ = oracle.vmlmb("rosenbrock", pb["x0"], 0.0, None, **okw)
        x = opk.vmlmb(opk.Rosenbrock(), pb["x0"], lower=0.0, lnsrch=ls, driver="native")
        assert relerr(x, xo) <= 1e-9
    # reverse communication with a user objective on device vectors
    ctx = opk.default_context()
    n = 1000
    t = ctx.from_numpy(np.linspace(-1, 2, n))
    calls = []

    def fg(x, g):
        calls.append(1)
        opk.vcombine_(g, 2, x, -2, t)
        r = opk.vcombine_(x.similar(), 1, x, -1, t)
        return opk.vdot(r, r)

    x = opk.vmlmb(fg, np.zeros(n), lower=opk.BoundedSet(0.0, 1.0), driver="native")
    assert np.abs(x - np.clip(np.linspace(-1, 2, n), 0, 1)).max() < 1e-5
    ncalls = len(calls)
    x2 = opk.vmlmb(fg, np.zeros(n), lower=opk.BoundedSet(0.0, 1.0))
    assert np.array_equal(x, x2) and len(calls) == 2 * ncalls    # same path as the Python driver


@pytest.mark.parametrize("driver", ["python", "native"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_large_memory_and_tiny_problems(opk, oracle, dtype, driver):
    """mem = 18 reaches the 10-role Gram decomposition (4 blocks of pairs); n = 1, 2, 3
    exercise the partial-tile paths of the ring kernels; mem = 1."""
    cls = opk.NativeVMLMB if driver == "native" else opk.VMLMB
    rtol = F64_RTOL if dtype == np.float64 else F32_RTOL
    for kind, n, kw in (("quad_box", 3001, dict(mem=18, maxiter=25)),
                        ("quad_box", 3001, dict(mem=18, maxiter=25, twoloop="sequential")),
                        ("quad_box", 1, dict(mem=3, maxiter=10)),
                        ("quad_box", 3, dict(mem=1, maxiter=10)),
                        ("rosen_lower0", 2, dict(mem=2, maxiter=15)),
                        ("rosen_box", 6, dict(mem=4, maxiter=20, twoloop="gram"))):
        if dtype == np.float32 and kw.get("twoloop") == "gram":
            continue
        pb = problem(kind, n, dtype)
        okw = {k: v for k, v in kw.items() if k != "twoloop"}
        okw["mem"] = min(okw["mem"], n)
        xo, ro, tro = run_oracle(oracle, pb, **okw)
        lo, hi = gpu_bounds(pb)
        if dtype == np.float32 and kw.get("mem") == 18 and "twoloop" not in kw:
            kw = dict(kw, twoloop="gram")          # exercise the 10-role Float32 kernel too
            rtol_case = 1e-3
        else:
            rtol_case = rtol
        s = cls(pb["gpu_obj"], pb["x0"], lower=lo, upper=hi, **kw)
        s.solve()
        assert (s.iter, s.eval) == (ro["iter"], ro["eval"]), (kind, n, kw, s.iter, s.eval, ro["iter"], ro["eval"])
        assert abs(s.f - ro["f"]) <= 100 * rtol_case * abs(ro["f"]) + 1e-20, (kind, n, kw, s.f, ro["f"])
        assert relerr(s.result(), xo) <= 100 * rtol_case + 1e-300, (kind, n, kw)
        s.close()


def _free_gib():
    import torch
    import opkb200
    opkb200.default_context().trim()   # blocks the context keeps for reuse count as free
    free, total = torch.cuda.mem_get_info()
    return free / 2 ** 30


@pytest.mark.parametrize("dtype,log2n,mem", [(np.float64, 28, 10), (np.float32, 28, 7)])
def test_full_size_properties(opk, dtype, log2n, mem):
    """BASELINE.json's full per-GPU sizes (C4: n = 2^28 Float64 m = 10; the C5 shard:
    n = 2^28 Float32 m = 7), where the oracle is not run: size-independent properties --
    feasibility of every iterate, monotone decrease (Armijo), bit reproducibility,
    agreement of the fused path with the one-kernel-per-call path, and of the
    native driver with the Python driver."""
    n = 1 << log2n
    need = (2 * mem + 12) * n * np.dtype(dtype).itemsize / 2 ** 30
    if _free_gib() < need + 4:
        pytest.skip("needs %.0f GiB of free device memory" % need)
    ctx = opk.default_context()
    obj = opk.Quadratic.synthetic_device(ctx, n, dtype, kappa=1e4)
    lo = ctx.vector(n, dtype).fill(0.0)
    hi = ctx.vector(n, dtype).fill(1.0)
    x0 = ctx.vector(n, dtype).fill(0.5)
    kw = dict(lower=lo, upper=hi, mem=mem, maxiter=12, xtol=(0, 0), ftol=(0, 0), gtol=(0, 0))
    runs = {}
    for name, cls, extra in (("fused", opk.VMLMB, {}), ("fused2", opk.VMLMB, {}),
                             ("native", opk.NativeVMLMB, {}), ("primitive", opk.VMLMB, dict(mode="primitive"))):
        fs = []
        s = cls(obj, opk.vcopy(x0), observer=lambda s: fs.append((s.iter, s.eval, s.f, s.gnorm)), **kw, **extra)
        s.solve()
        x = s.x
        assert opk.vnorminf(x) <= 1.0                                       # x <= hi
        neg = opk.project_variables_(x.similar(), x, None, 0.0)             # min(x, 0)
        assert opk.vnorminf(neg) == 0.0                                     # x >= lo
        assert all(b[2] < a[2] for a, b in zip(fs, fs[1:]))                 # monotone (Armijo)
        assert (s.iter, s.stage) == (12, 4)
        runs[name] = fs
        s.close()
        del s, x, neg
    assert runs["fused"] == runs["fused2"]                                  # bit reproducible
    assert runs["native"][-1] == runs["fused"][-1]                          # same numbers, native loop
    rtol = 1e-9 if dtype == np.float64 else 1e-3
    for a, b in zip(runs["fused"], runs["primitive"]):
        assert a[:2] == b[:2] and abs(a[2] - b[2]) <= rtol * abs(b[2]) and abs(a[3] - b[3]) <= rtol * abs(b[3])


def test_full_size_spg(opk):
    """C3: SPG on the box-constrained quadratic, n = 10^8 Float32: feasibility, the
    nonmonotone Armijo condition (f <= max of the last m values), reproducibility,
    fused == one-kernel-per-call."""
    n, dtype, m = 10 ** 8, np.float32, 10
    if _free_gib() < 12:
        pytest.skip("needs 12 GiB of free device memory")
    ctx = opk.default_context()
    obj = opk.Quadratic.synthetic_device(ctx, n, dtype, kappa=1e3)
    lo = ctx.vector(n, dtype).fill(0.0)
    hi = ctx.vector(n, dtype).fill(1.0)
    x0 = ctx.vector(n, dtype).fill(0.5)
    runs = {}
    for name, mode in (("fused", "fused"), ("fused2", "fused"), ("primitive", "primitive")):
        fs = []
        s = opk.SPG(obj, opk.BoxProjection(lo, hi), opk.vcopy(x0), m, maxit=15, eps1=0, eps2=0, mode=mode,
                    observer=lambda s: fs.append((s.iter, s.fcnt, s.f, s.pgtwon, s.pginfn)))
        s.solve()
        x = s.solution()
        assert opk.vnorminf(x) <= 1.0
        assert opk.vnorminf(opk.project_variables_(x.similar(), x, None, 0.0)) == 0.0
        for k in range(1, len(fs)):
            assert fs[k][2] <= max(f[2] for f in fs[max(0, k - m):k])
        runs[name] = fs
        s.close()
    assert runs["fused"] == runs["fused2"]
    for a, b in zip(runs["fused"], runs["primitive"]):
        assert a[:2] == b[:2] and abs(a[2] - b[2]) <= 1e-4 * abs(b[2])


@pytest.mark.parametrize("mode,native", [("fused", False), ("primitive", False), ("fused", True)])
def test_unconstrained_follows_scipy_lbfgsb(opk, mode, native):
    """The CUDA path against an implementation that is NOT the oracle: without bounds
    `vmlmb` is L-BFGS with `dcsrch`, and so is SciPy's L-BFGS-B (same line-search
    constants, same first step, same H0 scaling; tests/test_pins_scipy.py).  The
    iterates agree to the north_star tolerance over the first 20 iterations."""
    from test_pins_scipy import NOSTOP, assert_follows_lbfgsb, lbfgsb_cases, lbfgsb_trajectory

    for name, fg, oobj, gobj, x0, mem in lbfgsb_cases():
        ref = lbfgsb_trajectory(fg, x0, mem, 25)
        tr = []
        cls = opk.NativeVMLMB if native else opk.VMLMB
        # (between calls of the native run loop `x` already holds the first trial point of
        # the next line search: `current()` is the iterate)
        kw = dict(mem=mem, maxiter=25, **NOSTOP,
                  observer=lambda s: tr.append(dict(iter=s.iter, f=s.f,
                                                    x=(s.current() if native else s.x).download())))
        if not native:
            kw["mode"] = mode
        s = cls(gobj, x0, **kw)
        if native:
            while s.iterate(1):     # the native driver reports once per call
                pass
        else:
            s.solve()
        s.close()
        assert_follows_lbfgsb(tr, ref)
