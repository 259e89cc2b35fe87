#!/usr/bin/env python
"""Iterations/s and effective HBM GB/s of the five BASELINE.json configurations
(single GPU unless launched under torchrun); prints one JSON line per config.

    python tools/bench_configs.py [C1 C2 C3 C4 C5s]

C5 (n = 2^31 Float32) does not fit one GPU: "C5s" runs its per-GPU shard of an
8-GPU run (n = 2^28 Float32, m = 7) on one GPU; the real C5 needs torchrun."""
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import opkb200 as opk  # noqa: E402
from opkb200.dist import init_context, shard_range  # noqa: E402

PEAK = 6546.6
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def rosen_bounds(ctx, n, dtype, first):
    idx = np.arange(first, first + n)
    hi = np.where((idx // 2) % 2 == 0, 0.9, 2.0).astype(dtype)
    return ctx.vector(n, dtype).fill(0.0), ctx.from_numpy(hi)


NATIVE = bool(int(os.environ.get("OPKB_BENCH_NATIVE", "0")))   # driver loop inside the library


def run_vmlmb(ctx, name, obj, x0, lo, hi, mem, n, nglobal, K, es, walg, twoloop=None, prime=None):
    kw = dict(mem=mem, xtol=(0, 0), ftol=(0, 0), gtol=(0, 0), nglobal=nglobal, ctx=ctx, twoloop=twoloop)
    if lo is not None:
        kw.update(lower=lo, upper=hi)
    if NATIVE and not isinstance(obj, (opk.Rosenbrock, opk.Quadratic)):
        return {"config": name, "skipped": "native run loop needs a built-in objective"}
    s = (opk.NativeVMLMB if NATIVE else opk.VMLMB)(obj, x0, **kw)
    name += " [native driver]" if NATIVE else ""
    s.iterate((mem + 1 if prime is None else prime) + 3)
    ctx.synchronize()
    noprof = bool(int(os.environ.get("OPKB_BENCH_NOPROF", "0")))   # (no CUDA events around every launch)
    if not noprof:
        ctx.profile(True)
    e0, i0, l0 = s.eval, s.iter, ctx.launch_count
    ctx.timer_start()
    alive = s.iterate(K)
    ms = ctx.timer_stop()
    prof = {} if noprof else ctx.profile_report()
    ctx.profile(False)
    it = s.iter - i0
    out = {"config": name, "n": nglobal, "iterations": it, "ms_per_iteration": ms / max(it, 1),
           "iterations_per_s": it / (ms * 1e-3), "evals_per_iteration": (s.eval - e0) / max(it, 1),
           "launches_per_iteration": (ctx.launch_count - l0) / max(it, 1),
           "alg_GBps_per_gpu": walg * es * n / (ms * 1e-3 / max(it, 1)) / 1e9,
           "twoloop": s.twoloop, "stopped_early": (not alive), "f": s.f,
           "kernels_ms_per_iteration": {k: round(v[0] / max(it, 1), 4) for k, v in prof.items()},
           "kernel_launches_per_iteration": {k: round(v[1] / max(it, 1), 2) for k, v in prof.items()}}
    out["frac_of_measured_peak"] = out["alg_GBps_per_gpu"] / PEAK
    out["incremental_gram_hits_misses"] = list(s.carry_stats())
    if NATIVE:   # device-resident decisions: chains started / launches adopted in flight / cancelled by the device
        out["chain_heads_adopted_cancelled"] = list(s.chain_stats())
    s.close()
    return out


def run_one(ctx, c, rank=0, world=1, native=None):
    """One configuration line (a dict) on `ctx`; None for an unknown name."""
    global NATIVE
    if native is not None:
        NATIVE = bool(native)
    r = None
    if True:
        if c == "C1":     # VMLMB unconstrained Rosenbrock n=1e6 f64 m=5, More-Thuente
            n = 10 ** 6
            first, nl = shard_range(n, rank, world)
            x0 = ctx.from_numpy(opk.Rosenbrock.x0(nl, np.float64, perturb_seed=7, first=first))
            r = run_vmlmb(ctx, "C1 L-BFGS Rosenbrock n=1e6 f64 m=5", opk.Rosenbrock(), x0, None, None, 5,
                          nl, n, 300, 8, 4 * 5 + 13)
        elif c == "C2":   # VMLMB bound-constrained Rosenbrock n=1e7 f64 m=5
            n = 10 ** 7
            first, nl = shard_range(n, rank, world)
            x0 = ctx.from_numpy(opk.Rosenbrock.x0(nl, np.float64, perturb_seed=7, first=first))
            lo, hi = rosen_bounds(ctx, nl, np.float64, first)
            r = run_vmlmb(ctx, "C2 VMLMB Rosenbrock box n=1e7 f64 m=5", opk.Rosenbrock(), x0, lo, hi, 5,
                          nl, n, 200, 8, 4 * 5 + 20)
        elif c == "C3":   # SPG box-constrained quadratic n=1e8 f32
            n = 10 ** 8
            first, nl = shard_range(n, rank, world)
            obj = opk.Quadratic.synthetic_device(ctx, nl, np.float32, kappa=1e3, first=first)
            lo, hi = ctx.vector(nl, np.float32).fill(0.0), ctx.vector(nl, np.float32).fill(1.0)
            x0 = ctx.vector(nl, np.float32).fill(0.5)
            s = (opk.NativeSPG if NATIVE else opk.SPG)(obj, opk.BoxProjection(lo, hi), x0, 10, eps1=0, eps2=0,
                                                       ctx=ctx)
            for _ in range(5):
                s.iterate()
            ctx.synchronize()
            i0, f0c = s.iter, s.fcnt
            ctx.timer_start()
            K = 200
            if NATIVE:
                s.iterate(K)                      # one native call for the K iterations
            else:
                while s.iter < i0 + K and s.iterate():
                    pass
            ms = ctx.timer_stop()
            it = s.iter - i0
            r = {"config": "C3 SPG quadratic box n=1e8 f32 (m=10)" + (" [native driver]" if NATIVE else ""),
                 "n": n, "iterations": it,
                 "ms_per_iteration": ms / it, "iterations_per_s": it / (ms * 1e-3),
                 "evals_per_iteration": (s.fcnt - f0c) / it,
                 "alg_GBps_per_gpu": 8 * 4 * nl / (ms * 1e-3 / it) / 1e9, "f": s.f}
            r["frac_of_measured_peak"] = r["alg_GBps_per_gpu"] / PEAK
            s.close()
        elif c == "C4":
            n = 1 << 28
            first, nl = shard_range(n, rank, world)
            obj = opk.Quadratic.synthetic_device(ctx, nl, np.float64, kappa=1e4, first=first)
            lo, hi = ctx.vector(nl, np.float64).fill(0.0), ctx.vector(nl, np.float64).fill(1.0)
            x0 = ctx.vector(nl, np.float64).fill(0.5)
            r = run_vmlmb(ctx, "C4 VMLMB quadratic box n=2^28 f64 m=10", obj, x0, lo, hi, 10, nl, n, 30,
                          8, 60)
        elif c in ("C5", "C5-gram", "C5s", "C5s-gram"):   # VMLMB bound-constrained Rosenbrock f32 m=7
            n = (1 << 31) if c in ("C5", "C5-gram") else (1 << 28)
            first, nl = shard_range(n, rank, world)
            x0 = ctx.from_numpy(opk.Rosenbrock.x0(nl, np.float32, perturb_seed=7, first=first))
            lo, hi = rosen_bounds(ctx, nl, np.float32, first)
            tl = "gram" if c.endswith("gram") else None
            r = run_vmlmb(ctx, "%s VMLMB Rosenbrock box n=2^%d f32 m=7" % (c, int(math.log2(n))),
                          opk.Rosenbrock(), x0, lo, hi, 7, nl, n, 20, 4, 48, twoloop=tl)
        elif c == "S2D":   # non-separable objective with a halo exchange (user-objective path)
            rows = cols = 16384
            n = rows * cols
            r0, nr = shard_range(rows, rank, world, align=1)
            nl = nr * cols
            w = ctx.vector(nl, np.float64).fill_random(12, r0 * cols, 0.5, 1.0)
            y = ctx.vector(nl, np.float64).fill_random(11, r0 * cols, 0.0, 1.2)
            obj = opk.Smooth2D(w, y, 2.0, cols, ctx=ctx)
            x0 = ctx.vector(nl, np.float64).fill(0.5)
            # per iteration: trial_begin 5 + fg 4 + trial_end 7 + Gram sweep 2m+5 + direction 2m+6 passes
            r = run_vmlmb(ctx, "S2D VMLMB Smooth2D 16384^2 f64 m=10, box [0,1], halo exchange", obj, x0,
                          0.0, 1.0, 10, nl, n, 20, 8, 4 * 10 + 27)
        elif c in ("CG", "CGD"):   # conjgrad, stencil (halo exchange) / diagonal operator
            rows = cols = 16384
            n = rows * cols
            r0, nr = shard_range(rows, rank, world, align=1)
            nl = nr * cols
            w = ctx.vector(nl, np.float64).fill_random(12, r0 * cols, 0.5, 1.0)
            b = ctx.vector(nl, np.float64).fill_random(11, r0 * cols, 0.0, 1.2)
            if c == "CG":
                A, passes, name = opk.Smooth2DOperator(w, 2.0, cols, ctx=ctx), 11, "stencil diag(w) + 2 L, halo exchange"
            else:
                A, passes, name = opk.DiagonalOperator(w, ctx=ctx), 11, "diag(w)"
            s = opk.ConjGrad(A, b, ftol=0.0, restart=10 ** 9, nglobal=n, ctx=ctx, quiet=True)
            for _ in range(5):
                s.iterate()
            ctx.synchronize()
            ctx.profile(True)
            i0, l0, K = s.iter, ctx.launch_count, 50
            ctx.timer_start()
            while s.iter < i0 + K and s.iterate():
                pass
            ms = ctx.timer_stop()
            prof = ctx.profile_report()
            ctx.profile(False)
            it = max(s.iter - i0, 1)
            r = {"config": "%s conjgrad 16384^2 f64, %s" % (c, name), "n": n, "iterations": it,
                 "ms_per_iteration": ms / it, "iterations_per_s": it / (ms * 1e-3),
                 "launches_per_iteration": (ctx.launch_count - l0) / it, "passes_per_iteration": passes,
                 "alg_GBps_per_gpu": passes * 8 * nl / (ms * 1e-3 / it) / 1e9, "rho": s.rho,
                 "kernels_ms_per_iteration": {k: round(v[0] / it, 4) for k, v in prof.items()},
                 "kernel_launches_per_iteration": {k: round(v[1] / it, 2) for k, v in prof.items()}}
            r["frac_of_measured_peak"] = r["alg_GBps_per_gpu"] / PEAK
        else:
            return None
    r["n_gpus"] = world
    return r


def main():
    which = sys.argv[1:] or ["C1", "C2", "C3", "C4", "C5s"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = init_context(local)
    for c in which:
        r = run_one(ctx, c, rank, world)
        if r is not None and rank == 0:
            print(json.dumps(r), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
