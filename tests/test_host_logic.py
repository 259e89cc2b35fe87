"""CPU tests of the product's host logic (no GPU, no CUDA calls): the C-ABI
library loads and exports every declared symbol; the Python line searches match
the oracle's C state machines step for step; the VMLMB/SPG stage machines and
the coefficient-space two-loop solver reproduce the oracle driver when run on
a test-only numpy back end (tests/cpu_backend.py)."""
import math
import os
import re

import numpy as np
import pytest

import opkb200 as opk
from cpu_backend import NumpyGramOps, NumpyOps
from helpers import compare_traces, gpu_bounds, problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol():
    """The shared library loads without a GPU and exports what include/opkb200.h declares."""
    from opkb200 import _lib
    L = _lib.lib()
    header = open(os.path.join(ROOT, "include", "opkb200.h")).read()
    declared = set(re.findall(r"\b(opkb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in opkb200.h but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    assert L.opkb_version() == 100


def test_no_cpu_fallback():
    """Without a CUDA device the product fails loudly (no silent CPU path)."""
    import ctypes
    from opkb200 import _lib
    L = _lib.lib()
    h = ctypes.c_void_p()
    rc = L.opkb_context_create(ctypes.byref(h), -1)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = rc == 0
    if has_gpu:
        assert rc == 0
        L.opkb_context_destroy(h)
    else:
        assert rc != 0 and b"no CPU fallback" in L.opkb_last_error()
        with pytest.raises(opk.OpkbError):
            opk.vmlmb(opk.Rosenbrock(), np.zeros(4))


def test_cabi_argument_validation_without_gpu():
    """Status codes and opkb_last_error for invalid arguments (no CUDA call is reached)."""
    import ctypes as C
    from opkb200 import _lib
    L = _lib.lib()
    r = C.c_double()
    assert L.opkb_vdot(None, None, None, C.byref(r)) == -1
    assert b"NULL context" in L.opkb_last_error()
    assert L.opkb_vector_create(None, 1, 10, C.byref(C.c_void_p())) == -1
    assert L.opkb_vmlmb_create(None, 1, 10, 5, 2, None, 0.0, None, 1.0, C.byref(C.c_void_p())) == -1
    assert L.opkb_vmlmb_trial(None, 0, 1.0, 0, (C.c_double * 8)()) == -1
    assert L.opkb_spg_start(None, 1.0, (C.c_double * 8)()) == -1
    assert L.opkb_solver_create(None, None, C.byref(C.c_void_p())) == -1
    assert L.opkb_vector_length(None) == -1 and L.opkb_vector_destroy(None) == 0
    opt = _lib.VmlmbOptions()
    L.opkb_vmlmb_default_options(C.byref(opt))   # defaults of src/quasinewton.jl:227-242
    assert (opt.xrtol, opt.frtol, opt.grtol, opt.epsilon) == (1e-7, 1e-8, 1e-6, 0.0)
    assert opt.fmin == -math.inf and opt.maxiter == 2 ** 63 - 1 and opt.speculate == 1


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "optimpacknextgen.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "opk_oracle" not in txt and "libopkoracle" not in txt, f


@pytest.mark.parametrize("seed", range(6))
def test_linesearches_match_oracle(oracle, seed):
    """Random 1-D functions: the Python state machines and the C restatement
    return the same task and step at every call."""
    rng = np.random.default_rng(seed)
    O = oracle
    names = {O.TASK_SEARCH: "SEARCH", O.TASK_CONVERGENCE: "CONVERGENCE", O.TASK_WARNING: "WARNING"}
    for trial in range(40):
        a, b, c, w = rng.uniform(0.1, 3), rng.uniform(0.5, 5), rng.uniform(-1, 1), rng.uniform(0.5, 8)
        phi = lambda t: -a * t + b * t * t + c * math.sin(w * t) + 0.3 * t ** 3
        dphi = lambda t: -a + 2 * b * t + c * w * math.cos(w * t) + 0.9 * t * t
        if dphi(0) >= 0:
            continue
        stp0 = float(10 ** rng.uniform(-2, 1.5))
        stpmax = stp0 * float(10 ** rng.uniform(0, 3))
        for kind, mk in ((O.LS_ARMIJO, lambda: opk.ArmijoLineSearch(ftol=1e-4)),
                         (O.LS_MORE_TORALDO, lambda: opk.MoreToraldoLineSearch(ftol=1e-3, gamma=(0.1, 0.5))),
                         (O.LS_MORE_TORALDO, lambda: opk.MoreToraldoLineSearch(ftol=1e-3, strict=True)),
                         (O.LS_MORE_THUENTE, lambda: opk.MoreThuenteLineSearch(ftol=1e-3, gtol=0.9, xtol=0.1)),
                         (O.LS_MORE_THUENTE, lambda: opk.MoreThuenteLineSearch(ftol=1e-4, gtol=0.1, xtol=1e-3))):
            py = mk()
            kw = {}
            if kind == O.LS_ARMIJO:
                kw = dict(ftol=py.ftol)
            elif kind == O.LS_MORE_TORALDO:
                kw = dict(ftol=py.ftol, gamma=(py.gamma1, py.gamma2), strict=py.strict)
            else:
                kw = dict(ftol=py.ftol, gtol=py.gtol, xtol=py.xtol)
            cc = O.LineSearch(kind, **kw)
            t1 = py.start(phi(0), dphi(0), stp0, stpmin=1e-20 * stp0, stpmax=stpmax)
            t2 = cc.start(phi(0), dphi(0), stp0, 1e-20 * stp0, stpmax)
            assert t1 == names[t2] == "SEARCH"
            assert opk.usederivatives(py) == cc.usederivatives
            for _ in range(60):
                stp = py.getstep()
                assert stp == cc.stp
                t1 = py.iterate(stp, phi(stp), dphi(stp))
                t2 = cc.iterate(stp, phi(stp), dphi(stp))
                assert t1 == names[t2]
                if t1 != "SEARCH":
                    assert opk.getreason(py) == cc.reason()
                    break


def test_linesearch_errors():
    ls = opk.MoreToraldoLineSearch(ftol=1e-3)
    for args, msg in (((1.0, -1.0, 1.0, 2.0, 1.0), "STPMAX < STPMIN"), ((1.0, -1.0, 1.0, -1.0, 2.0), "STPMIN < 0"),
                      ((1.0, -1.0, 3.0, 0.0, 2.0), "STP > STPMAX"), ((1.0, -1.0, 0.5, 1.0, 2.0), "STP < STPMIN"),
                      ((1.0, 0.0, 1.0, 0.0, 2.0), "Not a descent direction")):
        with pytest.raises(ValueError, match=msg):
            ls.start(args[0], args[1], args[2], stpmin=args[3], stpmax=args[4])
    with pytest.raises(ValueError, match="descent condition violated"):
        opk.cstep(False, 0.0, 10.0, 0.0, 1.0, +1.0, 0.0, 1.0, 1.0, 1.0, 0.5, 0.2)
    ls = opk.ArmijoLineSearch()
    ls.start(1.0, -1.0, 1.0, stpmin=0.0, stpmax=2.0)
    with pytest.raises(AssertionError):
        ls.iterate(0.7, 1.0, 0.0)
    assert opk.getreason(opk.ArmijoLineSearch()) == "Algorithm not yet started"


CASES = [
    ("rosen_unc", 20, np.float64, dict()),
    ("rosen_unc", 20, np.float64, dict(mem=15)),
    ("rosen_unc", 1000, np.float64, dict()),
    ("rosen_lower0", 20, np.float64, dict()),
    ("rosen_lower0", 20, np.float64, dict(blmvm=True)),
    ("rosen_box", 1000, np.float64, dict(mem=7)),
    ("rosen_box", 1000, np.float64, dict(blmvm=True)),
    ("quad_box", 1003, np.float64, dict(mem=10)),
    ("quad_upper_scalar", 500, np.float64, dict(mem=3)),
    ("rosen_unc", 20, np.float32, dict()),
    ("rosen_box", 1000, np.float32, dict(mem=7)),
    ("quad_box", 1003, np.float32, dict(mem=5)),
]


def _run_host(pb, factory, **kw):
    lo, hi = gpu_bounds(pb)
    trace = []

    def obs(s):
        trace.append(dict(iter=s.iter, eval=s.eval, rejects=s.rejects, f=s.f, gnorm=s.gnorm, stp=s.stp,
                          x=np.array(s.x), mask=s.free_mask()))
    s = opk.VMLMB(pb["gpu_obj"], pb["x0"], lower=lo, upper=hi, observer=obs,
                  mode=lambda solver, x0: factory(solver, x0, pb["oracle_obj"]), **kw)
    s.solve()
    return s, trace


@pytest.mark.parametrize("kind,n,dtype,kw", CASES)
def test_stage_machine_equals_oracle_driver(oracle, kind, n, dtype, kw):
    """Python `_vmlmb!` + numpy/oracle vector ops == oracle C driver, bit for bit."""
    pb = problem(kind, n, dtype)
    xo, ro, tro = oracle.vmlmb(pb["oracle_obj"], pb["x0"], pb["lower"], pb["upper"], trace=True, **kw)
    s, tr = _run_host(pb, NumpyOps, **kw)
    assert (s.iter, s.eval, s.rejects, s.reason) == (ro["iter"], ro["eval"], ro["rejects"], ro["reason"])
    assert len(tr) == len(tro)
    for g, w in zip(tr, tro):
        assert (g["iter"], g["eval"], g["f"], g["gnorm"], g["stp"]) == (w["iter"], w["eval"], w["f"], w["gnorm"], w["stp"])
        assert np.array_equal(g["x"], w["x"]) and np.array_equal(g["mask"], w["mask"])
    assert np.array_equal(s.result(), xo)


@pytest.mark.parametrize("kind,n,dtype,kw", CASES)
def test_gram_form_two_loop_within_tolerance(oracle, kind, n, dtype, kw):
    """The coefficient-space two-loop (Gram block + host solve + assembly with
    the reference's rounding sequence) keeps the first 20 iterations within the
    north_star tolerance of the sequential recursion."""
    pb = problem(kind, n, dtype)
    kw = dict(kw, maxiter=25)
    xo, ro, tro = oracle.vmlmb(pb["oracle_obj"], pb["x0"], pb["lower"], pb["upper"], trace=True, **kw)
    s, tr = _run_host(pb, NumpyGramOps, **kw)
    if dtype == np.float64:
        compare_traces(tr, tro, 20, 1e-10)
    elif kind.startswith("quad"):
        compare_traces(tr, tro, 20, 1e-4)
    else:
        # Float32 + chaotic Rosenbrock: the roundings of the intermediate vector
        # (absent in coefficient space) are amplified to ~2e-3 by iteration 20,
        # which is why twoloop="sequential" is the Float32 default
        compare_traces(tr, tro, 20, 1e-2, check_mask=False)


def test_vmlmb_options(oracle):
    pb = problem("rosen_lower0", 20, np.float64)
    for kw in (dict(maxeval=7), dict(maxiter=3), dict(fmin=0.0), dict(epsilon=0.01), dict(gtol=(1e-3, 0.0)),
               dict(xtol=(0.0, 1e-2)), dict(ftol=(1e-4, 0.0)), dict(fmin=300.0)):
        xo, ro, tro = oracle.vmlmb("rosenbrock", pb["x0"], 0.0, None, **kw)
        s, tr = _run_host(pb, NumpyOps, **kw)
        assert (s.iter, s.eval, s.reason, s.stage) == (ro["iter"], ro["eval"], ro["reason"], ro["stage"]), kw
        assert np.array_equal(s.result(), xo)
    with pytest.raises(NotImplementedError):
        opk.VMLMB(opk.Rosenbrock(), pb["x0"], autodiff=True, mode=lambda s, x: None)
    with pytest.raises(AssertionError):
        opk.VMLMB(opk.Rosenbrock(), pb["x0"], epsilon=1.5, mode=lambda s, x: None)
    # print_iteration format (src/quasinewton.jl:649-660)
    import io
    buf = io.StringIO()
    opk.print_iteration(buf, 0, 1, 0, 242.0, 736.39, 0.0)
    assert buf.getvalue().splitlines()[2] == "     0      1      0    2.4200000000000000E+02  7.36E+02  0.00E+00"


def test_shard_ranges():
    from opkb200.dist import shard_range
    for n in (1 << 28, 1 << 31, 10 ** 7, 1000003, 10):
        for world in (1, 2, 4, 8):
            tot, prev = 0, 0
            for r in range(world):
                first, cnt = shard_range(n, r, world)
                assert first == prev and cnt >= 0
                if first + cnt < n:
                    assert cnt % 4 == 0
                prev = first + cnt
                tot += cnt
            assert tot == n


def test_fastmin_fastmax_fastclamp():
    """The scalar helpers SimpleBounds exports (src/bounds.jl:41, :53, :65-67), NaN rules included."""
    nan = float("nan")
    assert opk.fastmin(1.0, 2.0) == 1.0 and opk.fastmin(2.0, 1.0) == 1.0
    assert opk.fastmax(1.0, 2.0) == 2.0 and opk.fastmax(2.0, 1.0) == 2.0
    assert math.isnan(opk.fastmin(nan, 1.0)) and opk.fastmin(1.0, nan) == 1.0      # "or x otherwise"
    assert math.isnan(opk.fastmax(nan, 1.0)) and opk.fastmax(1.0, nan) == 1.0
    assert opk.fastclamp(0.5, 0.0, 1.0) == 0.5 and opk.fastclamp(-1.0, 0.0, 1.0) == 0.0
    assert opk.fastclamp(2.0, 0.0, 1.0) == 1.0 and opk.fastclamp(2.0, None, 1.0) == 1.0
    assert opk.fastclamp(-2.0, 0.0, None) == 0.0 and opk.fastclamp(7.0, None, None) == 7.0
    assert math.isnan(opk.fastclamp(nan, 0.0, 1.0))


def _build_c_host(tmpdir):
    """gcc examples/c_host.c against the C ABI: a host with no Python and no torch in it."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None or not os.path.exists(opk.SO_PATH):
        pytest.skip("needs gcc and the built library")
    exe = os.path.join(str(tmpdir), "c_host")
    libdir = os.path.dirname(opk.SO_PATH)
    r = subprocess.run([gcc, "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "c_host.c"), "-o", exe, "-L", libdir, "-lopkb200",
                        "-Wl,-rpath," + libdir, "-lm"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_host_builds_and_fails_loudly_without_gpu(tmp_path):
    """The C ABI is usable from plain C (examples/c_host.c compiles and links with gcc alone);
    without a CUDA device the very first call reports it: there is no CPU fallback."""
    import subprocess
    import torch
    exe = _build_c_host(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: tests/test_gpu_api.py::test_c_host_runs covers the run")
    r = subprocess.run([exe, "1000"], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


def test_python_twin_accepts_the_reference_keywords():
    """Drop-in surface: every keyword of the reference's `vmlmb!` (src/quasinewton.jl:226-242) and `spg!`
    (src/spg.jl:168-178) is a keyword of the twin's solvers, with the reference's defaults where Python
    has the same value (tests/test_julia_static.py does the same for the Julia module)."""
    import inspect
    import math

    import opkb200 as opk
    v = inspect.signature(opk.VMLMB.__init__).parameters
    for kw in ("mem", "lower", "upper", "autodiff", "blmvm", "fmin", "maxiter", "maxeval", "xtol", "ftol", "gtol",
               "epsilon", "verb", "printer", "output", "lnsrch"):
        assert kw in v and v[kw].kind is inspect.Parameter.KEYWORD_ONLY, kw
    assert (v["mem"].default, v["lower"].default, v["upper"].default) == (5, -math.inf, math.inf)
    assert (v["xtol"].default, v["ftol"].default, v["gtol"].default) == ((0.0, 1e-7), (0.0, 1e-8), (0.0, 1e-6))
    assert v["autodiff"].default is False and v["blmvm"].default is False and v["epsilon"].default == 0.0
    s = inspect.signature(opk.SPG.__init__).parameters
    for kw in ("autodiff", "ws", "maxit", "maxfc", "eps1", "eps2", "eta", "printer", "verb", "io"):
        assert kw in s and s[kw].kind is inspect.Parameter.KEYWORD_ONLY, kw
    assert (s["eps1"].default, s["eps2"].default, s["eta"].default) == (1e-6, 1e-6, 1.0)
    # the names north_star lists
    for name in ("vmlmb", "vmlmb_", "spg", "spg_", "BoundedSet", "MoreThuenteLineSearch", "ArmijoLineSearch",
                 "MoreToraldoLineSearch", "conjgrad", "conjgrad_"):
        assert hasattr(opk, name), name


def _random_case(seed):
    """A random small problem + option set in both vocabularies (oracle keywords, twin keywords)."""
    rng = np.random.default_rng(1000 + seed)
    dtype = np.float64 if rng.random() < 0.7 else np.float32
    rosen = rng.random() < 0.5
    n = int(rng.choice([2, 6, 50, 200])) if rosen else int(rng.choice([1, 3, 7, 51, 333]))
    if rosen:
        x0 = opk.Rosenbrock.x0(n, dtype, perturb_seed=int(rng.integers(1, 99)) if n > 2 else None)
        obj_o, obj_t = "rosenbrock", opk.Rosenbrock()
    else:
        a, c = opk.Quadratic.synthetic(n, dtype, kappa=float(rng.choice([10.0, 1e3])))
        x0 = rng.random(n).astype(dtype)
        obj_o, obj_t = ("quadratic", a, c), opk.Quadratic(a, c)
    # bounds: none | scalar / array, one- or two-sided, around the start
    def bound(kind, val):
        if kind == 0:
            return None
        if kind == 1:
            return float(val)
        return (np.full(n, val) + 0.2 * rng.random(n)).astype(dtype)
    lk, uk = int(rng.integers(0, 3)), int(rng.integers(0, 3))
    base = float(x0.min()) if rosen else 0.0
    top = float(x0.max()) if rosen else 1.0
    lower = bound(lk, base - 0.3 + 0.5 * rng.random())
    upper = bound(uk, top + 0.1 + 0.8 * rng.random())
    bounded = lower is not None or upper is not None
    okw, tkw = {}, {}
    mem = int(rng.choice([1, 2, 3, 5, 8]))
    okw["mem"] = tkw["mem"] = mem
    if bounded and rng.random() < 0.35:
        okw["blmvm"] = tkw["blmvm"] = True
    if rng.random() < 0.3:
        okw["epsilon"] = tkw["epsilon"] = 0.01
    if rng.random() < 0.2:
        okw["fmin"] = tkw["fmin"] = -1.0 if rosen else -1e6
    r = rng.random()
    if r < 0.2:
        ft = float(rng.choice([1e-4, 1e-2]))
        okw.update(lnsrch=1, ls_ftol=ft); tkw["lnsrch"] = opk.ArmijoLineSearch(ftol=ft)
    elif r < 0.45:
        ft, gt = float(rng.choice([1e-3, 1e-2])), float(rng.choice([0.9, 0.5]))
        okw.update(lnsrch=3, ls_ftol=ft, ls_gtol=gt, ls_xtol=0.1)
        tkw["lnsrch"] = opk.MoreThuenteLineSearch(ftol=ft, gtol=gt, xtol=0.1)
    elif r < 0.65:
        ft, strict = float(rng.choice([1e-3, 1e-4])), bool(rng.random() < 0.5)
        g1 = float(rng.choice([0.1, 0.25]))
        okw.update(lnsrch=2, ls_ftol=ft, ls_gamma=(g1, 0.5), ls_strict=strict)
        tkw["lnsrch"] = opk.MoreToraldoLineSearch(ftol=ft, strict=strict, gamma=(g1, 0.5))
    if rng.random() < 0.5:
        okw["maxiter"] = tkw["maxiter"] = int(rng.integers(3, 40))
    else:
        okw["maxeval"] = tkw["maxeval"] = int(rng.integers(3, 60))
    pb = dict(x0=x0, lower=lower, upper=upper, oracle_obj=obj_o, gpu_obj=obj_t)
    return pb, okw, tkw


@pytest.mark.parametrize("seed", range(80))
def test_stage_machine_equals_oracle_driver_on_random_options(oracle, seed):
    """The twin's `_vmlmb!` (the same Python that drives the GPU phases) against the oracle's C driver on
    random combinations of what src/quasinewton.jl:226-293 accepts -- element type, scalar / array / one-sided
    bounds, mem, BLMVM, epsilon, fmin, the three line searches with their parameters, maxiter / maxeval --:
    same counts, same reason, the same bits at every iteration."""
    pb, okw, tkw = _random_case(seed)
    try:
        xo, ro, tro = oracle.vmlmb(pb["oracle_obj"], pb["x0"], pb["lower"], pb["upper"], trace=True, **okw)
        oerr = None
    except Exception as e:      # (a line search that refuses its start: the twin must refuse too)
        oerr = e
    if oerr is not None:
        with pytest.raises(Exception):
            _run_host(pb, NumpyOps, **tkw)
        return
    s, tr = _run_host(pb, NumpyOps, **tkw)
    assert (s.iter, s.eval, s.rejects, s.reason) == (ro["iter"], ro["eval"], ro["rejects"], ro["reason"]), (okw, pb["lower"] is None, pb["upper"] is None)
    assert len(tr) == len(tro)
    for g, w in zip(tr, tro):
        assert (g["iter"], g["eval"], g["f"], g["gnorm"], g["stp"]) == (w["iter"], w["eval"], w["f"], w["gnorm"], w["stp"]), (g["iter"], okw)
        assert np.array_equal(g["x"], w["x"]) and np.array_equal(g["mask"], w["mask"])
    assert np.array_equal(s.result(), xo)
