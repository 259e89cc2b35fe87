"""Oracle parity AT THE SIZES of BASELINE.json's configurations (C1 .. C5).

The other parity tests meet the oracle at n <= 2*10^5; here the CUDA path (fused
Python-hosted AND native driver, through the C ABI) runs beside the OpenMP oracle at
the sizes the metric is quoted on, so that the tile / grid / tail geometry of the ring
kernels (148 CTAs x several hundred thousand tiles, 64-bit offsets) is checked
against an implementation that shares none of it:

  C1  VMLMB unconstrained Rosenbrock n = 10^6 Float64 m = 5, More-Thuente
  C2  VMLMB bounded Rosenbrock (array bounds) n = 10^7 Float64 m = 5
  C3  SPG box-constrained quadratic n = 10^8 Float32
  C4  VMLMB bounded quadratic n = 2^28 Float64 m = 10 (>= 8 iterations after the
      pairs have filled up, i.e. with the incremental Gram active)
  C5  its workload -- bounded Rosenbrock Float32 m = 7 -- at the largest n the oracle
      handles in well under a minute (2^25), sequential (parity-grade) and Gram two-loop

Gates (north_star): same iteration / evaluation / rejection counts, f, |g| (and the
accepted step with the Python host, which reports it) within 1e-10 (Float64) / 1e-4
(Float32) at EVERY iteration, the free-variable mask bit-exact at every iteration
(packed, compared on the host), x within the same tolerance at checkpoints, and two
linear fingerprints of x (|x| and x.u for a fixed pseudo-random u) at every iteration.
Inputs are generated on the device and downloaded for the oracle: identical bits.
(src/quasinewton.jl:381-601, src/spg.jl:245-380)
"""
import math
import os

import numpy as np
import pytest

from helpers import relerr

pytestmark = pytest.mark.gpu

F64_RTOL = 1e-10
F32_RTOL = 1e-4


@pytest.fixture(scope="module")
def opk():
    import opkb200
    return opkb200


def _avail_gib():
    try:
        with open("/proc/meminfo") as fh:
            for line in fh:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) / 2 ** 20
    except Exception:
        pass
    return 0.0


def _free_gpu_gib():
    import torch
    import opkb200
    opkb200.default_context().trim()   # blocks the context keeps for reuse count as free
    return torch.cuda.mem_get_info()[0] / 2 ** 30


class Probe:
    """Fingerprints of the current iterate of a solver, computed on the device."""

    def __init__(self, opk, ctx, n, dtype, lo, hi, bounded):
        self.opk, self.lo, self.hi, self.bounded = opk, lo, hi, bounded
        self.u = ctx.vector(n, dtype).fill_random(99, 0, -0.5, 0.5)
        self.tmp = ctx.vector(n, dtype) if bounded else None
        self.u_host = self.u.download().astype(np.float64)

    def of(self, s, at_printer, want_x):
        opk = self.opk
        if at_printer or s.finished:
            xcur, gcur = s.x, s.ops.vector("g")
        else:   # between iterate(1) calls x holds the first trial point of the next search
            xcur, gcur = s.current(), s.ops.vector("Y", s.mark)
        rec = dict(xnorm=opk.vnorm2(xcur), xu=opk.vdot(xcur, self.u))
        if self.bounded:
            opk.project_direction_(self.tmp, xcur, self.lo, self.hi, "-", gcur)
            rec["mask"] = np.packbits(self.tmp.download() != 0)
        if want_x:
            rec["x"] = xcur.download()
        return rec


def oracle_vmlmb(oracle, obj, x0, lo, hi, niter, u_host, checkpoints, **kw):
    recs = {}

    def tick(r):
        x = r["x"]
        e = dict(iter=r["iter"], eval=r["eval"], rejects=r["rejects"], f=r["f"], gnorm=r["gnorm"], stp=r["stp"],
                 xnorm=math.sqrt(float(np.dot(x, x) if x.dtype == np.float64 else np.dot(x.astype(np.float64),
                                                                                        x.astype(np.float64)))),
                 xu=float(np.dot(x if x.dtype == np.float64 else x.astype(np.float64), u_host)))
        if r["p"] is not None:
            e["mask"] = np.packbits(r["p"] != 0)
        if r["iter"] in checkpoints:
            e["x"] = x.copy()
        recs[r["iter"]] = e

    _, res, _ = oracle.vmlmb(obj, x0, lo, hi, maxiter=niter, trace=tick, **kw)
    return res, recs


def check_record(g, w, rtol, it, what, stp=True):
    for key in ("f", "gnorm", "xnorm") + (("stp",) if stp and "stp" in g else ()):
        assert abs(g[key] - w[key]) <= rtol * max(abs(w[key]), 1e-300), (what, key, it, g[key], w[key])
    # x.u is a signed sum: its error is relative to |x| |u|, not to the (possibly small) value
    assert abs(g["xu"] - w["xu"]) <= rtol * max(abs(w["xnorm"]), 1e-300) * 0.3 * math.sqrt(w.get("n", 1.0)) + \
        rtol * abs(w["xu"]), (what, "x.u", it, g["xu"], w["xu"])
    if "mask" in w and "mask" in g:
        assert np.array_equal(g["mask"], w["mask"]), (what, "free-variable mask", it,
                                                      int(np.unpackbits(g["mask"] ^ w["mask"]).sum()))
    if "x" in w and "x" in g:
        assert relerr(g["x"], w["x"]) <= rtol, (what, "x", it, relerr(g["x"], w["x"]))


def run_python_host(opk, probe, obj, x0, lo, hi, niter, checkpoints, **kw):
    """Fused path with the Python host: the observer runs at the reference's printer site."""
    recs = {}

    def obs(s):
        r = dict(iter=s.iter, eval=s.eval, rejects=s.rejects, f=s.f, gnorm=s.gnorm, stp=s.stp)
        r.update(probe.of(s, True, s.iter in checkpoints))
        recs[s.iter] = r

    s = opk.VMLMB(obj, opk.vcopy(x0), lower=lo, upper=hi, maxiter=niter, observer=obs, **kw)
    s.solve()
    out = dict(iter=s.iter, eval=s.eval, rejects=s.rejects, f=s.f, reason=s.reason, carry=s.carry_stats())
    s.close()
    return out, recs


def run_native_host(opk, probe, obj, x0, lo, hi, niter, checkpoints, **kw):
    """The driver loop inside the library, stopped at every iteration boundary."""
    recs = {}
    s = opk.NativeVMLMB(obj, opk.vcopy(x0), lower=lo, upper=hi, maxiter=niter, **kw)
    alive = True
    while alive:
        alive = s.iterate(1)
        r = dict(iter=s.iter, eval=s.eval, rejects=s.rejects, f=s.f, gnorm=s.gnorm)
        r.update(probe.of(s, False, s.iter in checkpoints))
        recs[s.iter] = r
    out = dict(iter=s.iter, eval=s.eval, rejects=s.rejects, f=s.f, reason=s.reason, carry=s.carry_stats())
    s.close()
    return out, recs


def compare(name, res_o, recs_o, res_g, recs_g, rtol, n, native=False):
    assert (res_g["iter"], res_g["eval"], res_g["rejects"], res_g["reason"]) == \
        (res_o["iter"], res_o["eval"], res_o["rejects"], res_o["reason"]), (name, res_g, res_o)
    seen = 0
    for it in sorted(recs_o):
        if it not in recs_g:      # (the native loop reports from iteration 1 on)
            continue
        w, g = dict(recs_o[it], n=float(n)), recs_g[it]
        assert (g["iter"], g["eval"], g["rejects"]) == (w["iter"], w["eval"], w["rejects"]), (name, it, g["eval"], w["eval"])
        check_record(g, w, rtol, it, name, stp=not native)
        seen += 1
    return seen


def device_problem(opk, kind, n, dtype, kappa=1e4):
    """Inputs generated on the device (fast at 2^28) + host copies for the oracle."""
    ctx = opk.default_context()
    if kind == "quad_box":
        obj = opk.Quadratic.synthetic_device(ctx, n, dtype, kappa=kappa)
        lo = ctx.vector(n, dtype).fill(0.0)
        hi = ctx.vector(n, dtype).fill(1.0)
        x0 = ctx.vector(n, dtype).fill(0.5)
        oobj = ("quadratic", obj.a.download(), obj.c.download())
        return ctx, obj, oobj, x0, lo, hi, x0.download(), lo.download(), hi.download()
    x0h = opk.Rosenbrock.x0(n, dtype, perturb_seed=7)
    x0 = ctx.from_numpy(x0h)
    if kind == "rosen_unc":
        return ctx, opk.Rosenbrock(), "rosenbrock", x0, -math.inf, math.inf, x0h, None, None
    assert kind == "rosen_box"    # lo = 0; hi = 0.9 for even pair index, 2.0 otherwise
    loh = np.zeros(n, dtype=dtype)
    hih = np.where((np.arange(n) // 2) % 2 == 0, 0.9, 2.0).astype(dtype)
    return ctx, opk.Rosenbrock(), "rosenbrock", x0, ctx.from_numpy(loh), ctx.from_numpy(hih), x0h, loh, hih


def vmlmb_case(opk, oracle, name, kind, n, dtype, mem, niter, hosts, rtol, kappa=1e4, **kw):
    ctx, obj, oobj, x0, lo, hi, x0h, loh, hih = device_problem(opk, kind, n, dtype, kappa)
    bounded = kind != "rosen_unc"
    probe = Probe(opk, ctx, n, dtype, lo, hi, bounded)
    checkpoints = {1, niter // 2, niter}
    okw = {k: v for k, v in kw.items() if k != "twoloop"}
    res_o, recs_o = oracle_vmlmb(oracle, oobj, x0h, loh, hih, niter, probe.u_host, checkpoints, mem=mem, **okw)
    assert res_o["iter"] == niter, res_o
    out = {}
    for host in hosts:
        run = run_native_host if host == "native" else run_python_host
        res_g, recs_g = run(opk, probe, obj, x0, lo, hi, niter, checkpoints, mem=mem, **kw)
        seen = compare("%s/%s" % (name, host), res_o, recs_o, res_g, recs_g, rtol, n, native=(host == "native"))
        assert seen >= niter, (name, host, seen)
        out[host] = res_g
    return res_o, out


NOSTOP = dict(xtol=(0, 0), ftol=(0, 0), gtol=(0, 0))


def test_c1_unconstrained_rosenbrock_1e6(opk, oracle):
    """C1 (BASELINE.json configs[0]): n = 10^6 Float64 m = 5, More-Thuente, 20 iterations."""
    vmlmb_case(opk, oracle, "C1", "rosen_unc", 10 ** 6, np.float64, 5, 20, ("python", "native"), F64_RTOL, **NOSTOP)


def test_c2_bounded_rosenbrock_1e7(opk, oracle):
    """C2: n = 10^7 Float64 m = 5, array bounds; mask bit-exact at every iteration."""
    vmlmb_case(opk, oracle, "C2", "rosen_box", 10 ** 7, np.float64, 5, 20, ("python", "native"), F64_RTOL, **NOSTOP)


def test_c4_bounded_quadratic_2p28(opk, oracle):
    """C4, the configuration the metric is quoted on: n = 2^28 Float64 m = 10.  19 iterations = 11
    that fill the memory + 8 in steady state, where ONE kernel per iteration runs (direction +
    speculative trial + incremental Gram): the carry must have been used."""
    mem, niter = 10, 19
    log2n = 28
    # oracle: 4 + 2m vectors + 5 inputs; this test: x copies at 3 checkpoints, masks, u
    need = lambda l: (4 + 2 * mem + 5 + 10) * 8 * (1 << l) / 2 ** 30 * 1.1
    if _avail_gib() < need(log2n):
        log2n = 27
    if _avail_gib() < need(log2n):
        pytest.skip("needs %.0f GiB of host memory for the oracle at n = 2^27" % need(27))
    n = 1 << log2n
    if _free_gpu_gib() < (2 * mem + 16) * 8 * n / 2 ** 30:
        pytest.skip("not enough device memory")
    res_o, out = vmlmb_case(opk, oracle, "C4(2^%d)" % log2n, "quad_box", n, np.float64, mem, niter,
                            ("python", "native"), F64_RTOL, **NOSTOP)
    for host, r in out.items():
        hits, misses = r["carry"]
        assert hits >= 5, (host, "the incremental Gram was not active", hits, misses)


@pytest.mark.parametrize("twoloop", ["sequential", "gram"])
def test_c5_workload_float32(opk, oracle, twoloop):
    """C5's workload (bounded Rosenbrock, Float32, m = 7) at n = 2^25 per GPU.  `sequential` is the
    parity-grade Float32 default (1e-4, mask bit-exact); `gram` is the opt-in coefficient-space
    form, whose documented Float32 gate on the chaotic Rosenbrock problem is 1e-2 without the
    mask -- here it must at least keep that at the large size."""
    n, mem, niter = 1 << 25, 7, 20
    if twoloop == "sequential":
        vmlmb_case(opk, oracle, "C5/seq", "rosen_box", n, np.float32, mem, niter, ("python", "native"), F32_RTOL,
                   twoloop="sequential", **NOSTOP)
        return
    ctx, obj, oobj, x0, lo, hi, x0h, loh, hih = device_problem(opk, "rosen_box", n, np.float32)
    probe = Probe(opk, ctx, n, np.float32, lo, hi, True)
    res_o, recs_o = oracle_vmlmb(oracle, oobj, x0h, loh, hih, niter, probe.u_host, {niter}, mem=mem, **NOSTOP)
    res_g, recs_g = run_python_host(opk, probe, obj, x0, lo, hi, niter, {niter}, mem=mem, twoloop="gram", **NOSTOP)
    assert res_g["iter"] == res_o["iter"] == niter
    for it in range(1, niter + 1):
        g, w = recs_g[it], recs_o[it]
        assert abs(g["f"] - w["f"]) <= 1e-2 * abs(w["f"]), (it, g["f"], w["f"])
        assert abs(g["xnorm"] - w["xnorm"]) <= 1e-2 * w["xnorm"], it
    assert relerr(recs_g[niter]["x"], recs_o[niter]["x"]) <= 1e-2


def test_c3_spg_quadratic_1e8_float32(opk, oracle):
    """C3: SPG on the box-constrained quadratic, n = 10^8 Float32, m = 10, 12 iterations: f, the
    two projected-gradient norms and the counts at every iteration, x at the end (1e-4); fused
    Python-hosted loop and the native loop."""
    n, dtype, m, niter = 10 ** 8, np.float32, 10, 12
    if _avail_gib() < 12 or _free_gpu_gib() < 8:
        pytest.skip("not enough memory")
    ctx, obj, oobj, x0, lo, hi, x0h, loh, hih = device_problem(opk, "quad_box", n, dtype, kappa=1e3)
    recs_o = {}

    def tick(r):
        recs_o[r["iter"]] = dict(fcnt=r["fcnt"], pcnt=r["pcnt"], f=r["f"], pgtwon=r["pgtwon"], pginfn=r["pginfn"])

    xo, io, _ = oracle.spg(oobj, x0h, m, loh, hih, maxit=niter, eps1=0.0, eps2=0.0, trace=tick)
    for cls in (opk.SPG, opk.NativeSPG):
        recs_g = {}
        ws = opk.Info()
        s = cls(obj, opk.BoxProjection(lo, hi), opk.vcopy(x0), m, maxit=niter, eps1=0.0, eps2=0.0, ws=ws,
                observer=lambda s: recs_g.__setitem__(s.iter, dict(fcnt=s.fcnt, pcnt=s.pcnt, f=s.f,
                                                                   pgtwon=s.pgtwon, pginfn=s.pginfn)))
        if cls is opk.NativeSPG:
            while s.iterate(1):
                pass
        else:
            s.solve()
        xg = s.solution().download()
        assert (ws.iter, ws.fcnt, ws.pcnt, ws.status) == (io["iter"], io["fcnt"], io["pcnt"], io["status"])
        seen = 0
        for it, w in recs_o.items():
            g = recs_g.get(it)
            if g is None:
                continue
            assert (g["fcnt"], g["pcnt"]) == (w["fcnt"], w["pcnt"]), (cls.__name__, it)
            for key in ("f", "pgtwon", "pginfn"):
                assert abs(g[key] - w[key]) <= F32_RTOL * abs(w[key]), (cls.__name__, key, it, g[key], w[key])
            seen += 1
        assert seen >= 10, (cls.__name__, seen)
        assert relerr(xg, xo) <= F32_RTOL, (cls.__name__, relerr(xg, xo))
        s.close()


def test_two_ranks_on_hardware():
    """One process per GPU over NCCL / the shared-memory mailbox on REAL devices: tools/multigpu_parity.py
    (VMLMB, SPG, conjgrad, Smooth2D with its halo exchange -- every case against the single-process
    oracle, identical decisions on all ranks).  Skipped on a one-GPU box; the outputs of the 2-, 4- and
    8-GPU runs are kept under profiles/."""
    import subprocess
    import sys

    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541",
                        os.path.join(root, "tools", "multigpu_parity.py")],
                       capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "FAIL" not in r.stdout and r.stdout.count(" OK") >= 14, r.stdout[-3000:]
