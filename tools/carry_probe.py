"""Development aid: lockstep run of the fused path with and without the
incremental Gram; prints the largest relative deviation of the Gram block, f
and x per iteration (rounding-level deviations amplified by the trajectory)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import opkb200 as opk  # noqa: E402
from helpers import gpu_bounds, problem  # noqa: E402


def main():
    kind, n, mem = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    dtype = np.float32 if (len(sys.argv) > 4 and sys.argv[4] == "f32") else np.float64
    pb = problem(kind, n, dtype)
    lo, hi = gpu_bounds(pb)
    kw = dict(lower=lo, upper=hi, mode="fused", mem=mem, xtol=(0, 0), ftol=(0, 0), gtol=(0, 0),
              twoloop="gram")
    os.environ["OPKB_CARRY"] = "2"          # (read when the workspace is created)
    a = opk.VMLMB(pb["gpu_obj"], pb["x0"], **kw)
    os.environ["OPKB_CARRY"] = "0"
    b = opk.VMLMB(pb["gpu_obj"], pb["x0"], **kw)
    for it in range(40):
        a.iterate(1)
        b.iterate(1)
        ga, gb = a.last_gram(), b.last_gram()
        dev = -1.0
        if ga and gb and ga[0] == gb[0]:
            GA, GB = ga[1], gb[1]
            m = len(ga[0])
            scale = np.abs(GB).max()
            for r in range(2 * m):
                for c in range(m + 1):
                    if c > 0 and (c - 1) < r // 2:
                        continue
                    dev = max(dev, abs(GA[r, c] - GB[r, c]) / max(abs(GB[r, c]), 1e-4 * scale))
        xa, xb = a.x.download(), b.x.download()
        print("it %3d eval %3d/%3d  f rel %.2e  x rel %.2e  G dev %.2e  hits %s" % (
            a.iter, a.eval, b.eval, abs(a.f - b.f) / abs(b.f),
            np.linalg.norm(xa - xb) / np.linalg.norm(xb), dev, a.carry_stats()))


if __name__ == "__main__":
    main()
