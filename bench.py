#!/usr/bin/env python
"""bench.py -- VMLMB iterations/sec and effective HBM GB/s (BASELINE.json).

A "step" is one VMLMB iteration (an accepted line search, counter `iter` of
src/quasinewton.jl:443) of the bound-constrained convex quadratic of config C4:
n = 2^28 Float64, m = 10, array bounds [0, 1], a_i = 1e4^u, c_i = -1 + 3u,
x0 = 0.5, fused path: ONE kernel per iteration in steady state
(P4 direction + first trial of the next line search + incremental Gram block;
the P1 trial and P3 Gram-sweep kernels only run when that trial is not taken).
With N > 1 (torchrun, one rank per GPU) the SAME n is sharded in contiguous
blocks over the ranks ("strong" scaling); the only communication is one packed
all-gather of reduction partials per phase.

    python bench.py --gpus 1 --steps 50 --warmup 3
    python bench.py --impl reference --steps 5 --warmup 1     # CPU oracle port

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

# stdout carries ONE JSON line: whatever NCCL prints (e.g. "NCCL version ..." when the launcher
# exports NCCL_DEBUG=VERSION) goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "VMLMB iterations/sec at n=2^28 f64 m=10 (bound-constrained quadratic)"
UNIT = "iterations/s"
KAPPA = 1e4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2n", type=int, default=28)
    ap.add_argument("--mem", type=int, default=10)
    ap.add_argument("--cpu-log2n", type=int, default=25,
                    help="size of the bounded CPU sample of the cpu_baseline leg of the GPU arm")
    ap.add_argument("--cpu-steps", type=int, default=6)
    ap.add_argument("--ref-budget-s", type=float, default=float(os.environ.get("OPKB_REF_BUDGET_S", "270")),
                    help="wall-time cap of the reference arm's solver run (priming + warm-up + timed)")
    ap.add_argument("--driver", default="native", choices=["native", "python"],
                    help="host of the stage machine: the C++ driver inside the library, or the Python twin")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the short lines of the other BASELINE.json configurations (N = 1 only)")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def config(args, n):
    """ONE description of the workload, identical for both arms (the driver compares them)."""
    return {
        "workload": "C4: VMLMB bound-constrained quadratic n=2^%d Float64 m=%d, array bounds [0,1], "
                    "More-Toraldo line search; a step = one VMLMB iteration (accepted line search) in steady "
                    "state (all m pairs stored)" % (int(math.log2(n)), args.mem),
        "n": n, "mem": args.mem, "eltype": "Float64", "kappa": KAPPA,
        "sharding": "contiguous 1-D blocks over %d rank(s) (GPU arm); the CPU arm runs the whole n on the "
                    "host cores of rank 0" % args.gpus,
        "l2": "no flush needed: every pass streams %.3g GiB vectors (>> 126 MB L2, >> host caches)" % (n * 8 / 2 ** 30),
        "priming": "mem+1 untimed iterations before the warm-up so that all m (s,y) pairs exist",
    }


# ---------------------------------------------------------------------------
# CPU arm: the oracle port ("reference-faithful unfused" C restatement, OpenMP)
def host_ram_bytes():
    try:
        return os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES")
    except Exception:
        return 0


def cpu_vectors(mem):
    """n-vectors alive during an oracle run: a, c, lo, hi, x0 (inputs) + x, g, p, d, 2m pairs."""
    return 5 + 4 + 2 * mem


def cpu_run(args, steps, warmup, log2n, threads=None, budget_s=None):
    """Times VMLMB iterations of the oracle on the host cores at n_s = 2^log2n, in ONE solver run:
    the oracle's `printer` hook (src/quasinewton.jl:501-503) timestamps every iteration; the first
    mem+1 iterations (priming: the pairs fill up) and `warmup` more are not timed, the next `steps`
    are.  With a wall-time budget the number of timed iterations is reduced (never below 3) from a
    calibration run at a small size, and reported."""
    import opk_oracle as O
    O.build()
    O.set_sum_mode(O.SUM_ACCURATE)
    if threads is None:
        threads = int(os.environ.get("OPKB_CPU_THREADS", "0"))         # 0 = all host cores
    O.set_num_threads(threads)
    cores = O.max_threads() if threads == 0 else threads
    sys.path.insert(0, ROOT)
    import opkb200 as opk
    kw = dict(mem=args.mem, xtol=(0, 0), ftol=(0, 0), gtol=(0, 0))
    prime = args.mem + 1

    def run(log2, nwarm, nsteps):
        ns = 1 << log2
        a, c = opk.Quadratic.synthetic(ns, np.float64, kappa=KAPPA)
        lo, hi = np.zeros(ns), np.ones(ns)
        x0 = np.full(ns, 0.5)
        stamps = []

        def tick(rec):
            stamps.append((rec["iter"], rec["eval"], time.perf_counter()))

        t_begin = time.perf_counter()
        _, res, _ = O.vmlmb(("quadratic", a, c), x0, lo, hi, maxiter=prime + nwarm + nsteps, trace=tick, **kw)
        by_iter = {it: (ev, t) for it, ev, t in stamps}
        i0, i1 = prime + nwarm, prime + nwarm + nsteps
        if i0 not in by_iter or i1 not in by_iter:
            raise SystemExit("oracle stopped after %d iterations (%s)" % (res["iter"], res["reason"]))
        dt = by_iter[i1][1] - by_iter[i0][1]
        return dict(ns=ns, seconds=dt, steps=nsteps, warmup=nwarm, its=nsteps / max(dt, 1e-9),
                    evals_per_iter=(by_iter[i1][0] - by_iter[i0][0]) / nsteps,
                    total_seconds=time.perf_counter() - t_begin, f=res["f"])

    note = ""
    if budget_s is not None:
        cal_log2 = min(log2n, 22)
        cal = run(cal_log2, 1, 3)
        t_iter = (1.0 / cal["its"]) * (1 << (log2n - cal_log2))   # predicted seconds per iteration at n_s
        # priming iterations work on fewer pairs (about half the passes on average)
        avail = budget_s - t_iter * (0.6 * prime)
        k = int(avail / t_iter) - warmup
        if k < steps:
            if k < 3:
                warmup = max(1, min(warmup, int(avail / t_iter) - 3))
                k = 3
            note = ("; %d of the %d requested iterations timed to stay within the %.0f s budget "
                    "(predicted %.2f s per iteration)" % (k, steps, budget_s, t_iter))
            steps = k
    r = run(log2n, warmup, steps)
    r.update(cores=cores, note=note)
    return r


def reference_line(args):
    """`--impl reference`: the reference's CPU path for this workload (the oracle port: the reference is
    Julia and cannot run here) at the REAL n when the host has the memory for it."""
    n = 1 << args.log2n
    need = cpu_vectors(args.mem) * 8 * n * 1.08
    ram = host_ram_bytes()
    log2 = args.log2n
    while log2 > 20 and ram and cpu_vectors(args.mem) * 8 * (1 << log2) * 1.08 > 0.9 * ram:
        log2 -= 1
    r = cpu_run(args, args.steps, args.warmup, log2, budget_s=args.ref_budget_s)
    scale = r["ns"] / n
    value = r["its"] * scale
    where = ("at the full n=2^%d" % args.log2n) if log2 == args.log2n else (
        "at n_s=2^%d -- the host has %.0f GiB, the full size needs %.0f GiB -- and SCALED by n_s/n (an "
        "extrapolation, flagged in `extrapolated`)" % (log2, ram / 2 ** 30, need / 2 ** 30))
    sample = ("oracle port (C restatement of the reference's unfused CPU path, OpenMP, %d threads) %s: one "
              "solver run, every iteration timestamped at the reference's printer site; %d priming + %d "
              "warm-up iterations untimed, %d iterations timed in %.2f s (whole run %.1f s)%s; the Julia "
              "reference itself cannot run here (no Julia runtime)"
              % (r["cores"], where, args.mem + 1, r["warmup"], r["steps"], r["seconds"],
                 r["total_seconds"], r["note"]))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": 1e3 / value,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config(args, n),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "extrapolated": log2 != args.log2n,
        "requested": {"steps": args.steps, "warmup": args.warmup},
        "evals_per_iter": r["evals_per_iter"], "final": {"f": r["f"]},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def ready(self, nrows=2):
        """nvidia-smi needs a few hundred ms before its first line: True once it is streaming."""
        return self.proc is None or len(self.rows) >= nrows

    def summary(self, t0, t1, t_load0=None):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except Exception:
                self.proc.kill()
            self.t.join(timeout=3)   # no daemon thread may outlive the interpreter
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        window = "timed region"
        if t_load0 is not None and not any(t0 - 0.05 <= t <= t1 + 0.1 for t, _ in self.rows):
            # a timed region shorter than the sampling period (many GPUs): the samples taken
            # under the same load just before it (priming and warm-up iterations)
            t0, window = t_load0, "priming + warm-up + timed region (the timed region is shorter than the sampling period)"
        sm, mx, reasons = [], [], set()
        for t, line in self.rows:
            if not (t0 - 0.05 <= t <= t1 + 0.1):
                continue
            f = [s.strip() for s in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


def bind_to_gpu_numa_node(device):
    """Run this rank on the host cores next to its GPU (NVML's CPU affinity of the device): the pinned
    buffers of the e2e leg are then first-touched on that NUMA node and cross PCIe without a hop over
    the socket interconnect -- with 8 ranks the host side, not the links, bounds the uploads."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device))
        ncpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


class NumaNearGpu:
    """For the e2e leg on ONE rank: the pinned host buffers are allocated (first-touched) on the NUMA node
    the GPU hangs off -- `set_mempolicy(MPOL_PREFERRED, node)` for this thread, plus its CPU affinity when the
    container owns cores there -- and both are restored afterwards (the CPU-baseline leg wants every core).
    A buffer on the far socket crosses the socket interconnect on its way to PCIe: 25 instead of 48 GB/s
    on these boxes (profiles/r02s_bench_n1.json vs r02n_bench_n1.json).  Best effort: any failure leaves
    the defaults in place, and `report` says what was done."""

    def __init__(self, device):
        self.device, self.report, self._aff = int(device), {"numa_node": None}, None

    def __enter__(self):
        import ctypes
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.device)
            bus = pynvml.nvmlDeviceGetPciInfo(h).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            path = "/sys/bus/pci/devices/%s/numa_node" % bus.lower()[-12:]
            node = int(open(path).read().strip())
            self.report["numa_node"] = node
            if node >= 0:
                libc = ctypes.CDLL(None, use_errno=True)
                mask = ctypes.c_ulong(1 << node)
                MPOL_PREFERRED, SYS_set_mempolicy = 1, 238          # x86-64
                rc = libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, ctypes.byref(mask), 64)
                self.report["mempolicy"] = "preferred" if rc == 0 else "errno %d" % ctypes.get_errno()
                self._libc = libc if rc == 0 else None
                cpus = set()
                for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
                    a, _, b = part.partition("-")
                    cpus |= set(range(int(a), int(b or a) + 1))
                cpus &= set(os.sched_getaffinity(0))
                if cpus:
                    self._aff = os.sched_getaffinity(0)
                    os.sched_setaffinity(0, cpus)
                self.report["cpus_on_node"] = len(cpus)
        except Exception as e:
            self.report["error"] = repr(e)[:120]
        return self

    def __exit__(self, *exc):
        import ctypes
        try:
            if self._aff:
                os.sched_setaffinity(0, self._aff)
            if getattr(self, "_libc", None) is not None:
                self._libc.syscall(238, 0, None, 0)                  # MPOL_DEFAULT
        except Exception:
            pass
        return False


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def traffic_for(kernel, n_local):
    """(dram bytes per launch of `kernel`, note) from the committed ncu capture, if any.  The capture
    was taken at one size; for another shard length of the same streaming kernel the bytes are scaled
    by n_local / n_captured (every byte of traffic is per element) and the note says so."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            t = json.load(fh)
        e = t.get(kernel)
        if e:
            b = float(e["dram_bytes_per_launch"])
            if int(e["n"]) == int(n_local):
                return b, "ncu capture at this size (%s)" % e.get("source", "profiles/")
            return (b * n_local / float(e["n"]),
                    "ncu capture at n=%d (%s) scaled by n_local/n to this shard" % (int(e["n"]), e.get("source", "profiles/")))
    except Exception:
        pass
    return None, None


def ours(args):
    import torch
    import torch.distributed as dist

    import opkb200 as opk
    from opkb200.dist import init_context, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local)
    if world > 1:       # (one rank: all host cores stay available to the CPU-baseline leg)
        bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = init_context(local)

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = 1 << args.log2n
    first, nloc = shard_range(n, rank, world)
    dt = np.float64
    es = 8
    m = args.mem
    K, W = args.steps, max(args.warmup, 3)

    # ---- device-resident inputs -------------------------------------------------
    obj = opk.Quadratic.synthetic_device(ctx, nloc, dt, kappa=KAPPA, first=first)
    lo = ctx.vector(nloc, dt).fill(0.0)
    hi = ctx.vector(nloc, dt).fill(1.0)
    x0 = ctx.vector(nloc, dt).fill(0.5)
    kw = dict(lower=lo, upper=hi, mem=m, xtol=(0, 0), ftol=(0, 0), gtol=(0, 0), nglobal=n, ctx=ctx)

    Solver = opk.NativeVMLMB if args.driver == "native" else opk.VMLMB
    solver = Solver(obj, x0, **kw)
    # (nvidia-smi needs a few hundred ms before its first line: it is started before the
    # priming iterations so that it is streaming when the timed region begins)
    sampler = ClockSampler(local) if rank == 0 else None
    t_load0 = time.perf_counter()
    solver.iterate(m + 1)                               # priming: all m pairs exist (steady state)
    solver.iterate(W)                                   # warm-up (untimed)
    # With many GPUs the timed region lasts a few tens of ms, less than nvidia-smi takes to start:
    # more untimed iterations of the same load (the same number on every rank) until rank 0's
    # sampler is streaming, so that the clocks line covers the timed region or the load just before it.
    extra = 0
    while world > 1 or (sampler is not None and not sampler.ready()):
        go = 1.0 if (sampler is not None and not sampler.ready() and time.perf_counter() - t_load0 < 3.0) else 0.0
        if world > 1:
            t = torch.tensor([go], dtype=torch.float64, device="cuda")
            dist.broadcast(t, src=0)
            go = float(t.item())
        if go == 0.0:
            break
        solver.iterate(10)
        extra += 10
    ctx.profile(True)
    barrier()
    l0, e0, i0 = ctx.launch_count, solver.eval, solver.iter
    h0 = solver.carry_stats()[0]
    t0 = time.perf_counter()
    ctx.timer_start()
    alive = solver.iterate(K)                           # EXACTLY K iterations
    ms = ctx.timer_stop()
    barrier()
    t1 = time.perf_counter()
    ms = max_over_ranks(ms)
    launches = ctx.launch_count - l0
    evals = solver.eval - e0
    iters = solver.iter - i0
    prof = ctx.profile_report()
    ctx.profile(False)
    hits = solver.carry_stats()[0] - h0
    clocks = sampler.summary(t0, t1, t_load0) if sampler else None
    f_final, gnorm_final = solver.f, solver.gnorm
    history = solver.history
    solver.close()
    del solver
    if not alive or iters != K:
        raise SystemExit("solver stopped after %d of %d timed iterations" % (iters, K))
    value = K / (ms * 1e-3)

    # ---- roofline of the dominant kernel -------------------------------------------
    peak, peak_src = measured_peak()
    # algorithmic passes per launch (array bounds): P1 = 7, P3 = 2m+5, P4 = 2m+8;
    # P4 is counted with 2m+6: the two saves of :562-563 are done by buffer aliasing
    # (no traffic at all), see DESIGN.md section 4.  The iteration figure keeps the
    # survey's yardstick W_alg = 4m+20.
    # When the first trial of the next line search is fused into P4 (speculation, the
    # default with a built-in objective) that kernel also writes x, g, p: 2m+9; with the
    # incremental Gram it reads S, Y, x, g, lo, hi and writes d, x', g', s', y': 2m+9 (p' is
    # not written any more: the next launch recomputes it, OPKB_LAZY_P; 2m+10 with the write)
    # (the objective's data a, c -- 2 more vectors -- are not counted as algorithmic).
    lazy_p = 0 if os.environ.get("OPKB_LAZY_P", "1") == "0" else 1
    fused_trial = prof.get("trial", (0.0, 0))[1] < K / 2
    carried = hits >= K / 2
    # iterate history (default on the Gram path): the pairs are formed on the fly, s', y' are never
    # written (fused kernel 2m + 7 algorithmic passes) and the sweep writes nothing (2m + 3)
    hist = 1 if history == 1 else 0
    passes = {"trial": 7, "gram": 2 * m + 5 - 2 * hist,
              "direction": 2 * m + ((9 - 2 * hist + (1 - lazy_p)) if carried else ((8 + (1 - lazy_p)) if fused_trial else 6))}
    kern = {}
    for name, npass in passes.items():
        if name in prof:
            tot, cnt = prof[name]
            avg = tot / cnt
            ach = npass * es * nloc / (avg * 1e-3) / 1e9
            kern[name] = {"launches": cnt, "avg_ms": avg, "share_of_step": tot / ms,
                          "alg_bytes_per_launch": npass * es * nloc, "achieved_gbs": ach,
                          "frac": ach / peak}
    dom = max(kern, key=lambda k: kern[k]["avg_ms"] * kern[k]["launches"])
    names = {"trial": "k_trial", "gram": "k_gram2", "direction": "k_direction2"}
    traffic, traffic_note = traffic_for(dom, nloc)
    # what ONE steady-state launch of the fused kernel has to move (compulsory): read S, Y (2m), x, g,
    # lo, hi, a, c; write d, x', g', s', y' = 2m+11 passes (2m+9 "algorithmic" + the objective's a, c)
    moved = sum(v["launches"] * (v["alg_bytes_per_launch"] + (2 * es * nloc if k != "gram" else 0))
                for k, v in kern.items()) / K * world
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": kern[dom]["achieved_gbs"], "peak": peak,
                "unit": "GB/s", "frac": kern[dom]["frac"], "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": peak_src, "kernels": kern, "trial_fused_into_direction": fused_trial,
                "incremental_gram_iterations": hits, "iterate_history": bool(hist),
                "note": "k_direction2 = direction + speculative trial" +
                        (" + incremental Gram (the Gram sweep ran %d times in %d iterations)"
                         % (prof.get("gram", (0.0, 0))[1], K) if carried else ""),
                "iteration": {
                    "evals_per_iter": evals / K,
                    "compulsory_model": {
                        "passes": 2 * m + 11 - 2 * hist, "bytes": (2 * m + 11 - 2 * hist) * es * n,
                        "what": "ONE kernel per iteration: read the m pairs (2m rows), x, g, lo, hi, a, c; write d, x', g'" +
                                ("" if hist else ", s', y'") + " (p recomputed, saves by aliasing, trial fused, the pairs "
                                "cross HBM once" + ("; iterate history: the pairs are differences of stored iterates formed "
                                                    "on the fly, nothing is written back)" if hist else ")"),
                        "achieved_gbs_per_gpu": (2 * m + 11 - 2 * hist) * es * n / (ms * 1e-3 / K) / 1e9 / world,
                        "frac_per_gpu": (2 * m + 11 - 2 * hist) * es * n / (ms * 1e-3 / K) / 1e9 / world / peak},
                    "moved_bytes_per_iter": moved,
                    "survey_yardstick": {
                        "passes": 4 * m + 20, "bytes": (4 * m + 20) * es * n,
                        "what": "SURVEY 8(d) W_alg of a 3-phase iteration (P1 + P3 + P4); obsolete for this "
                                "design, which moves 2m+11 passes -- the fraction below exceeds 1 for that "
                                "reason, not because the kernel beats the HBM peak",
                        "frac_per_gpu": (4 * m + 20) * es * n / (ms * 1e-3 / K) / 1e9 / world / peak}}}

    # ---- e2e: the public call with HOST (pinned) buffers -------------------------------
    e2e = None
    if not args.no_e2e:
        host = {}
        import contextlib
        near = NumaNearGpu(local) if world == 1 else contextlib.nullcontext()   # (N > 1: bound at start-up)
        with near:
            for name, v in (("a", obj.a), ("c", obj.c), ("lo", lo), ("hi", hi), ("x0", x0)):
                t = torch.empty(nloc, dtype=torch.float64, pin_memory=True)
                arr = t.numpy()
                v.download(arr)
                host[name] = (t, arr)
            out_t = torch.empty(nloc, dtype=torch.float64, pin_memory=True)
            out_t.zero_()                                  # (first touch inside the policy)
        del obj, lo, hi, x0
        barrier()
        te0 = time.perf_counter()
        s2 = Solver(opk.Quadratic(host["a"][1], host["c"][1]), host["x0"][1], lower=host["lo"][1],
                    upper=host["hi"][1], mem=m, maxiter=K, xtol=(0, 0), ftol=(0, 0), gtol=(0, 0),
                    nglobal=n, ctx=ctx)
        ctx.synchronize()
        te1 = time.perf_counter()
        s2.solve()
        ctx.synchronize()
        te2 = time.perf_counter()
        s2.x.download(out_t.numpy())
        te3 = time.perf_counter()
        barrier()
        te = max_over_ranks(time.perf_counter() - te0)
        # PCIe floor of this call: the 5 input vectors and the result cross the link once each
        h2d_gbs = 5 * nloc * es / max(te1 - te0, 1e-9) / 1e9
        e2e = {"value": s2.iter / te, "unit": UNIT, "h2d_bytes_per_step": 5 * nloc * es * world / K,
               "d2h_bytes_per_step": nloc * es * world / K,
               "what": "opkb200.%s(Quadratic(a, c), x0, lower, upper, mem=%d, maxiter=%d).solve() "
                       "with pinned host arrays in, host array out (setup, uploads, %d iterations, "
                       "download inside the timed region)" % (Solver.__name__, m, K, s2.iter),
               "seconds": te,
               "breakdown_s": {"create_and_upload": max_over_ranks(te1 - te0), "solve": max_over_ranks(te2 - te1),
                               "download": max_over_ranks(te3 - te2),
                               "host_buffers": (near.report if world == 1 else "rank bound to the cores next to its GPU"),
                               "note": "max over ranks of each part; create_and_upload moves 5 vectors over "
                                       "PCIe (%.1f GB/s on rank 0): with K = %d iterations the transfers, not "
                                       "the iterations, bound this figure" % (h2d_gbs, K)}}
        s2.close()

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config(args, n),
        "path": "fused: direction + speculative trial + incremental Gram in ONE kernel per iteration (P1 trial and "
                "P3 Gram sweep only when the speculated step is not taken); stage machine, line search and m x m "
                "two-loop solve on the host (%s driver)" % args.driver,
        "exchange": ("none: one rank" if world == 1 else
                     ("packed partials of each phase written by the last CTA of every rank into a shared-memory "
                      "mailbox" if ctx.uses_mailbox else "packed ncclAllGather per phase")),
        "roofline": roofline,
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        "evals_per_s": evals / (ms * 1e-3), "final": {"f": f_final, "gnorm": gnorm_final},
    }
    if world == 1 and not args.no_other_configs:
        # The other configurations of BASELINE.json (parity-test cases, not bench lines): one short
        # steady-state measurement each, same method (CUDA events around K iterations after the pairs
        # have filled up), so that their rates are in the driver's record too.  C5s = the per-GPU shard
        # of C5 (n = 2^28 Float32 of the 2^31) -- sequential two-loop (parity-grade Float32 default) and
        # the Gram form; fractions are of the measured HBM peak by SURVEY 8(d)'s W_alg yardstick.
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        others = {}
        try:
            import bench_configs as BC
            ctx.trim()
            for name in ("C1", "C2", "C3", "C5s", "C5s-gram"):
                try:
                    r = BC.run_one(ctx, name, 0, 1, native=(args.driver == "native"))
                    others[name] = {k: r[k] for k in ("config", "n", "iterations", "iterations_per_s", "ms_per_iteration",
                                                      "evals_per_iteration", "frac_of_measured_peak", "twoloop",
                                                      "incremental_gram_hits_misses") if k in r}
                except Exception as e:      # never lose the headline line to a side measurement
                    others[name] = {"error": repr(e)[:200]}
                ctx.trim()
        except Exception as e:
            others = {"error": repr(e)[:200]}
        line["other_configs"] = others
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_run(args, args.cpu_steps, 1, args.cpu_log2n)
        scale = r["ns"] / n
        line["cpu_baseline"] = {
            "value": r["its"] * scale, "unit": UNIT, "cores": r["cores"], "kind": "port",
            "sample": "oracle port (C restatement of the reference's unfused CPU path, OpenMP, %d threads) on a "
                      "bounded sample: n_s=2^%d (256 MB per vector, far beyond the host caches), %d priming + 1 "
                      "warm-up iterations untimed, %d iterations timed in %.2f s, iterations/s scaled by n_s/n to "
                      "n=2^%d; `bench.py --impl reference` measures the same port at the full n"
                      % (r["cores"], args.cpu_log2n, m + 1, r["steps"], r["seconds"], args.log2n)}
        # also the scalar figure (one thread, a smaller sample) -- the Julia reference is
        # single-threaded
        l1 = min(args.cpu_log2n, 22)
        r1 = cpu_run(args, 3, 1, l1, threads=1)
        line["cpu_baseline"]["single_thread"] = {
            "value": r1["its"] * r1["ns"] / n, "unit": UNIT, "cores": 1,
            "sample": "same port, 1 thread: 3 iterations at n_s=2^%d (%.2f s) after %d priming + 1 warm-up, "
                      "scaled to n=2^%d" % (l1, r1["seconds"], m + 1, args.log2n)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    # explicit teardown, in dependency order
    obj = lo = hi = x0 = s2 = host = out_t = None
    import gc
    gc.collect()
    ctx.synchronize()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            reference_line(args)
        return
    ours(args)


if __name__ == "__main__":
    main()
