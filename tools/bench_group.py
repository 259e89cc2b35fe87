#!/usr/bin/env python
"""C4 (n = 2^28 Float64, m = 10, bounded quadratic) from ONE process on G GPUs (`opkb_group_*`):
iterations/s with CUDA-event timing on rank 0's stream, to set beside the torchrun line of bench.py.

    python tools/bench_group.py [G] [K]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import opkb200 as opk  # noqa: E402
from opkb200.dist import shard_range  # noqa: E402


def main():
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    log2n, m = int(os.environ.get("OPKB_LOG2N", "28")), 10
    n = 1 << log2n
    grp = opk.Group(G)
    G = grp.size
    solvers = []
    for r, ctx in enumerate(grp.contexts):
        first, nl = shard_range(n, r, G)
        obj = opk.Quadratic.synthetic_device(ctx, nl, np.float64, kappa=1e4, first=first)
        lo, hi = ctx.vector(nl).fill(0.0), ctx.vector(nl).fill(1.0)
        x0 = ctx.vector(nl).fill(0.5)
        solvers.append(opk.NativeVMLMB(obj, x0, lower=lo, upper=hi, mem=m, xtol=(0, 0), ftol=(0, 0), gtol=(0, 0),
                                       nglobal=n, ctx=ctx))
    ms = opk.MultiVMLMB.from_shards(grp, solvers, n)
    ms.iterate(m + 1 + 5)
    for c in grp.contexts:
        c.synchronize()
    c0 = grp.contexts[0]
    t0 = time.perf_counter()
    c0.timer_start()
    ms.iterate(K)
    dev_ms = c0.timer_stop()
    wall = time.perf_counter() - t0
    print(json.dumps({"config": "C4 from ONE process on %d GPUs (opkb_group_solver_run)" % G, "n": n, "mem": m,
                      "iterations": K, "ms_per_iteration": dev_ms / K, "iterations_per_s": K / (dev_ms * 1e-3),
                      "wall_iterations_per_s": K / wall, "f": ms.f, "iter": ms.iter, "eval": ms.eval,
                      "incremental_gram_hits_misses": list(ms.carry_stats())}), flush=True)
    ms.close()
    grp.close()


if __name__ == "__main__":
    main()
