#!/usr/bin/env python
"""Writes the INPUTS of the golden cases (tests/golden/make_golden.py) as raw little-endian binary
files + a manifest, so that tools/make_reference_golden.jl -- run by someone who has a Julia runtime
and the unmodified reference installed -- feeds the REAL OptimPackNextGen.jl exactly the bits the
oracle and the CUDA path are fed (a = kappa^u is not recomputed in Julia: its `^` may round the last
bit differently from libm's pow).

    python tools/make_reference_inputs.py        # -> tests/golden/reference_inputs/
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]

from helpers import problem  # noqa: E402
from make_golden import SPG_CASES, VMLMB_CASES  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "reference_inputs")


def dump(name, arr, dtype):
    np.ascontiguousarray(arr, dtype=dtype).tofile(os.path.join(OUT, name))


def bound_field(tag, b, n, dtype, case):
    """'none' | 'scalar:<value>' | 'array:<file>'"""
    if b is None:
        return "none"
    if np.isscalar(b):
        return "scalar:%r" % float(b)
    fn = "%s.%s.bin" % (case, tag)
    dump(fn, b, dtype)
    return "array:" + fn


def main():
    os.makedirs(OUT, exist_ok=True)
    lines = ["# solver case kind n dtype mem|m maxiter x0 lower upper objective  (raw little-endian arrays)"]
    for solver, cases in (("vmlmb", VMLMB_CASES), ("spg", SPG_CASES)):
        for idx, c in enumerate(cases):
            kind, n, dt = c[0], c[1], c[2]
            mem = c[3].get("mem", 5) if solver == "vmlmb" else c[3]
            maxiter = 25 if solver == "vmlmb" else 30
            case = "%s%d" % (solver, idx)
            pb = problem(kind, n, np.dtype(dt))
            dump(case + ".x0.bin", pb["x0"], dt)
            lo = bound_field("lo", pb["lower"], n, dt, case)
            hi = bound_field("hi", pb["upper"], n, dt, case)
            if isinstance(pb["oracle_obj"], tuple):
                dump(case + ".a.bin", pb["oracle_obj"][1], dt)
                dump(case + ".c.bin", pb["oracle_obj"][2], dt)
                obj = "quadratic:%s.a.bin:%s.c.bin" % (case, case)
            else:
                obj = "rosenbrock"
            lines.append(" ".join([solver, case, kind, str(n), dt, str(mem), str(maxiter), case + ".x0.bin",
                                   lo, hi, obj]))
    with open(os.path.join(OUT, "manifest.txt"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    print("wrote", OUT, "(%d cases)" % (len(lines) - 1))


if __name__ == "__main__":
    main()
