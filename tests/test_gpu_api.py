"""API-surface behaviour of the host twin on the GPU: in-place variants, printing
(`verb`, `printer`, `output`), BoundedSet, status objects -- the keyword contract
of src/quasinewton.jl:226-242 and src/spg.jl:168-178."""
import io
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def opk():
    import opkb200
    return opkb200


def test_inplace_and_copy_semantics(opk):
    ctx = opk.default_context()
    x0 = opk.Rosenbrock.x0(20)
    # vmlmb leaves x0 unchanged (src/quasinewton.jl:215); vmlmb! overwrites (:226)
    x = opk.vmlmb(opk.Rosenbrock(), x0)
    assert x is not x0 and x0[0] == -1.2 and np.abs(x - 1).max() < 1e-3
    y = x0.copy()
    out = opk.vmlmb_(opk.Rosenbrock(), y)
    assert out is y and np.array_equal(y, x)
    v = ctx.from_numpy(x0)
    w = opk.vmlmb(opk.Rosenbrock(), v)
    assert isinstance(w, opk.Vector) and w is not v and np.array_equal(v.download(), x0)
    assert np.array_equal(w.download(), x)
    out = opk.vmlmb_(opk.Rosenbrock(), v, mode="primitive")
    assert out is v and np.abs(v.download() - 1).max() < 1e-3
    # Float32 in, Float32 out; integers are promoted to Float64
    assert opk.vmlmb(opk.Rosenbrock(), x0.astype(np.float32)).dtype == np.float32
    assert opk.vmlmb(opk.Quadratic(np.ones(4), np.full(4, 2.0)), np.zeros(4, dtype=np.int64)).dtype == np.float64
    # spg / spg!
    z = opk.spg(opk.Rosenbrock(), opk.BoxProjection(0.0, math.inf), x0, 10)
    assert x0[0] == -1.2 and np.abs(z - 1).max() < 1e-3
    y = x0.copy()
    assert opk.spg_(opk.Rosenbrock(), opk.BoxProjection(0.0, math.inf), y, 10) is y and np.array_equal(y, z)


def test_printing_and_reasons(opk):
    x0 = opk.Rosenbrock.x0(20)
    buf = io.StringIO()
    opk.vmlmb(opk.Rosenbrock(), x0, verb=10, output=buf)
    lines = buf.getvalue().splitlines()
    # header (2 lines) + iterations 0, 10, 20, 30 + the final iterate + the reason line
    assert lines[0].startswith("# ITER   EVAL   REJECTS") and lines[1].startswith("#-----")
    its = [int(l.split()[0]) for l in lines[2:-1]]
    assert its[:4] == [0, 10, 20, 30] and lines[-1] == "# CONVERGENCE: gradient norm sufficiently small"
    seen = []
    opk.vmlmb(opk.Rosenbrock(), x0, verb=1, maxiter=3, output=io.StringIO(),
              printer=lambda out, it, ev, rej, f, gn, stp: seen.append((it, ev)))
    assert [s[0] for s in seen] == [0, 1, 2, 3]
    buf = io.StringIO()
    ws = opk.Info()
    opk.spg(opk.Rosenbrock(), opk.BoxProjection(0.0, math.inf), x0, 10, verb=20, io=buf, ws=ws)
    out = buf.getvalue().splitlines()
    assert out[0].startswith("#  ITER   EVAL   PROJ") and out[-1].startswith("# SUCCESS: Convergence with projected")
    assert ws.status == opk.INFNORM_CONVERGENCE and ws.f == ws.fbest
    assert opk.getreason(ws) == "Convergence with projected gradient infinite-norm"


def test_boundedset_and_bound_kinds(opk):
    n = 1000
    a, c = opk.Quadratic.synthetic(n, kappa=100.0)
    want = np.clip(c, 0.2, 0.8)
    kw = dict(mem=5, ftol=(0, 0), xtol=(0, 0), gtol=(0, 1e-10), maxiter=300)
    ctx = opk.default_context()
    lo_a, hi_a = np.full(n, 0.2), np.full(n, 0.8)
    for lower, upper in ((0.2, 0.8), (lo_a, 0.8), (0.2, hi_a), (lo_a, hi_a),
                         (ctx.from_numpy(lo_a), ctx.from_numpy(hi_a)), (opk.BoundedSet(0.2, hi_a), None)):
        if upper is None:
            x = opk.vmlmb(opk.Quadratic(a, c), np.full(n, 0.5), lower=lower, **kw)
        else:
            x = opk.vmlmb(opk.Quadratic(a, c), np.full(n, 0.5), lower=lower, upper=upper, **kw)
        assert np.abs(x - want).max() < 1e-7
    x = opk.vmlmb(opk.Quadratic(a, c), np.full(n, 0.5), lower=0.2, **kw)            # one-sided
    assert np.abs(x - np.maximum(c, 0.2)).max() < 1e-7
    x = opk.vmlmb(opk.Quadratic(a, c), np.full(n, 0.5), upper=hi_a, blmvm=True, **kw)
    assert np.abs(x - np.minimum(c, 0.8)).max() < 1e-7
    with pytest.raises(TypeError):
        opk.vmlmb(opk.Quadratic(a, c), np.full(n, 0.5), lower=np.zeros(n - 1))     # "invalid lower bound type"
    with pytest.raises(NotImplementedError):
        opk.vmlmb(lambda x: 0.0, np.zeros(4), autodiff=True)


def test_context_and_profile(opk):
    ctx = opk.default_context()
    assert ctx.sm_count == 148 and ctx.nranks == 1 and ctx.rank == 0 and ctx.device >= 0
    n0 = ctx.launch_count
    v = ctx.vector(1000).fill(2.0)
    ctx.profile(True)
    assert opk.vdot(v, v) == 4000.0 and opk.vnorm2(v) == math.sqrt(4000.0)
    rep = ctx.profile_report()
    ctx.profile(False)
    assert rep["reduce"][1] == 2 and ctx.launch_count == n0 + 3
    ctx.timer_start()
    opk.vscale_(v, 0.5)
    assert ctx.timer_stop() >= 0.0 and v.download()[0] == 1.0


def test_c_host_runs(opk, tmp_path):
    """examples/c_host.c: vmlmb (native loop and reverse communication with a user objective) and
    spg (native loop) driven from plain C through the C ABI; the program checks the solutions."""
    import subprocess
    from test_host_logic import _build_c_host
    exe = _build_c_host(tmp_path)
    r = subprocess.run([exe, "200001"], capture_output=True, text=True, timeout=300)
    # every call of the ABI returned success (the program stops at the first failing status) ...
    assert "failed (" not in r.stderr and r.stdout.count("\n") == 5, r.stdout + r.stderr
    lines = r.stdout.splitlines()
    assert lines[0].startswith("vmlmb  built-in") and lines[1].startswith("vmlmb  user") and lines[2].startswith("spg")
    assert lines[3].startswith("vmlmb  group of 2 ranks")
    # the group run (one process, two ranks) takes the decisions of the one-rank run
    assert lines[3].split(":")[1].split(", f=")[0] == lines[0].split(":")[1].split(", f=")[0]
    # ... the user-objective run (f = |x - c|^2 / 2 on the box) lands on clamp(c) to rounding, and the
    # program's own sanity checks of the three solutions hold
    err_user = float(lines[1].split("|x - clamp(c)|_inf=")[1].split()[0])
    assert err_user <= 1e-12
    # ... and so do the example's own checks of the three solutions (exit status 0, "result: ok"):
    # the one end-to-end gate of the native C driver (opkb_solver_run, opkb_solver_iterate,
    # opkb_spg_solver_run) from a plain C host
    assert r.returncode == 0 and "result: ok" in lines[4], r.stdout + r.stderr


# ---------------------------------------------------------------------------
# raw device pointers, foreign streams, the device-callback objective (include/opkb200.h,
# "raw device pointers"; ADVICE r1: the data pointers of the workspace vectors MOVE)
class _Raw:
    """A raw device pointer as something torch can wrap (no copy)."""

    def __init__(self, ptr, n, dtype):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": np.dtype(dtype).str, "data": (int(ptr), False),
                                         "version": 2}


def _wrap(ptr, n, dtype=np.float64):
    import torch
    return torch.as_tensor(_Raw(ptr, n, dtype), device="cuda")


def test_user_objective_on_a_foreign_stream_with_raw_pointers(opk):
    """A user objective that (i) asks for the data pointers at every call -- they move: the saves of
    :562-563, the speculated trial point -- and (ii) runs on a stream of its own: x is complete when
    trial_begin returns, g is made visible with `wait_stream` (no host synchronisation of it)."""
    import torch
    ctx = opk.default_context()
    n = 50001
    target = np.linspace(-1, 2, n)
    t_lib = ctx.from_numpy(target)
    t_torch = torch.from_numpy(target).cuda()
    side = torch.cuda.Stream()
    seen_x, seen_g = set(), set()

    def fg_raw(x, g):
        seen_x.add(x.data_ptr)
        seen_g.add(g.data_ptr)
        xt, gt = _wrap(x.data_ptr, n), _wrap(g.data_ptr, n)
        with torch.cuda.stream(side):
            f = float(((xt - t_torch) ** 2).sum().item())        # (.item() waits for `side`)
            gt.copy_(2 * xt - 2 * t_torch)                       # ... this kernel is still in flight
        ctx.wait_stream(side.cuda_stream)                        # the library's stream waits for it
        return f

    def fg_lib(x, g):
        opk.vcombine_(g, 2, x, -2, t_lib)
        r = opk.vcombine_(x.similar(), 1, x, -1, t_lib)
        return opk.vdot(r, r)

    for cls in (opk.VMLMB, opk.NativeVMLMB):
        a = cls(fg_raw, np.full(n, 0.5), lower=opk.BoundedSet(0.0, 1.0), mem=5)
        a.solve()
        b = cls(fg_lib, np.full(n, 0.5), lower=opk.BoundedSet(0.0, 1.0), mem=5)
        b.solve()
        assert (a.iter, a.eval) == (b.iter, b.eval)
        xa, xb = a.result(), b.result()
        assert np.abs(xa - xb).max() <= 1e-12 and np.abs(xa - np.clip(target, 0, 1)).max() < 1e-5
        a.close()
        b.close()
    assert len(seen_x) > 1 and len(seen_g) > 1, "the test must see the pointers move"


def test_device_callback_objective(opk):
    """`opkb_solver_set_fg_callback`: the reverse-communication loop runs inside the library and calls
    the objective with raw device pointers and the context's stream."""
    import torch
    ctx = opk.default_context()
    n = 40000
    target = np.linspace(-1, 2, n)
    t_torch = torch.from_numpy(target).cuda()
    calls = []

    def fn(nn, xp, gp, stream):
        calls.append((xp, gp))
        assert nn == n and stream == ctx.stream
        with torch.cuda.stream(torch.cuda.ExternalStream(stream)):
            xt, gt = _wrap(xp, nn), _wrap(gp, nn)
            gt.copy_(2 * xt - 2 * t_torch)
            return float(((xt - t_torch) ** 2).sum().item())

    obj = opk.DeviceObjective(fn)
    s = opk.NativeVMLMB(obj, np.full(n, 0.5), lower=opk.BoundedSet(0.0, 1.0), mem=5)
    s.solve()
    x = s.result()
    ncb = len(calls)
    assert s.eval == ncb and np.abs(x - np.clip(target, 0, 1)).max() < 1e-5
    s.close()
    # the same object as an ordinary fg(x, g) of the Python-hosted driver: same trajectory
    s2 = opk.VMLMB(obj, np.full(n, 0.5), lower=opk.BoundedSet(0.0, 1.0), mem=5)
    s2.solve()
    assert (s2.iter, s2.eval) == (s.iter, s.eval) and np.array_equal(s2.result(), x)
    s2.close()

    # an exception in the callback surfaces as that exception, not as a crash
    def bad(nn, xp, gp, stream):
        raise KeyError("boom")

    s3 = opk.NativeVMLMB(opk.DeviceObjective(bad), np.full(8, 0.5), lower=0.0)
    with pytest.raises(KeyError):
        s3.solve()
    s3.close()


@pytest.mark.parametrize("driver", ["python", "native"])
def test_mem_above_the_gram_limit(opk, oracle, driver):
    """The reference accepts any `mem` (src/quasinewton.jl:227, :328; its own test uses 15): beyond the
    20 pairs of the Gram-form kernels the fused path runs the step-by-step recursion."""
    from helpers import gpu_bounds, problem, relerr
    pb = problem("quad_box", 5003, np.float64)
    xo, ro, _ = oracle.vmlmb(pb["oracle_obj"], pb["x0"], pb["lower"], pb["upper"], mem=27, maxiter=45)
    lo, hi = gpu_bounds(pb)
    cls = opk.NativeVMLMB if driver == "native" else opk.VMLMB
    s = cls(pb["gpu_obj"], pb["x0"], lower=lo, upper=hi, mem=27, maxiter=45)
    assert s.twoloop == "sequential"
    s.solve()
    assert (s.iter, s.eval, s.rejects) == (ro["iter"], ro["eval"], ro["rejects"])
    assert abs(s.f - ro["f"]) <= 1e-10 * abs(ro["f"]) and relerr(s.result(), xo) <= 1e-10
    s.close()
    with pytest.raises(ValueError):
        opk.VMLMB(pb["gpu_obj"], pb["x0"], lower=lo, upper=hi, mem=27, twoloop="gram")


def test_device_memory_is_recycled(opk):
    """Destroyed vectors go back to the context, the next vector of that size takes the block."""
    ctx = opk.default_context()
    ctx.trim()
    v = ctx.vector(123457)
    p = v.data_ptr
    v.free()
    assert ctx.cached_bytes >= 123457 * 8
    w = ctx.vector(123457)
    assert w.data_ptr == p and ctx.cached_bytes == 0
    w.free()
    ctx.trim()
    assert ctx.cached_bytes == 0


def test_second_context_and_foreign_current_device(opk):
    """Every entry point binds the device of its context: a context keeps working when the
    host changes the current device, and two contexts of one process do not share launch attributes."""
    import torch
    ctx2 = opk.Context(0)
    try:
        n = 300001
        a, c = opk.Quadratic.synthetic(n, np.float64, kappa=1e2)
        kw = dict(lower=np.zeros(n), upper=np.ones(n), mem=10, maxiter=15)
        x1 = opk.vmlmb(opk.Quadratic(a, c), np.full(n, 0.5), **kw)
        if torch.cuda.device_count() > 1:
            torch.cuda.set_device(1)           # the host wanders off to another GPU
        try:
            x2 = opk.vmlmb(opk.Quadratic(a, c), np.full(n, 0.5), ctx=ctx2, **kw)   # big-smem ring kernels
        finally:
            torch.cuda.set_device(0)
        assert np.array_equal(x1, x2)
    finally:
        ctx2.close()
