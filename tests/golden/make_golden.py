"""Generates tests/golden/*.json from the CPU oracle (canonical ACCURATE mode).

The reference (Julia) cannot run here and holds no golden vectors of its own
("parity unpinned"), so these fixtures pin the ORACLE: the CPU
suite checks the oracle still reproduces them and the GPU suite checks the CUDA
path against them without needing anything outside the repo.

    python tests/golden/make_golden.py            # vmlmb + spg
    python tests/golden/make_golden.py conjgrad   # conjgrad only
"""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]

import opk_oracle as O  # noqa: E402
from helpers import problem  # noqa: E402

VMLMB_CASES = [
    ("rosen_unc", 20, "float64", {}),
    ("rosen_unc", 20, "float64", {"mem": 15}),
    ("rosen_lower0", 20, "float64", {}),
    ("rosen_unc", 20, "float32", {}),
    ("rosen_lower0", 20, "float32", {}),
    ("rosen_box", 1000, "float64", {"mem": 5}),
    ("quad_box", 1003, "float64", {"mem": 10}),
    ("quad_box", 1003, "float32", {"mem": 7}),
]
SPG_CASES = [
    ("rosen_lower0", 20, "float64", 10),
    ("quad_box", 1003, "float64", 10),
    ("quad_box", 1003, "float32", 10),
]


CG_CASES = [
    # (operator, shape, dtype, keywords)
    ("stencil", (24, 40), "float64", {"ftol": 0.0, "maxiter": 25, "restart": 10}),
    ("stencil", (24, 40), "float32", {"ftol": 0.0, "maxiter": 25, "restart": 10}),
    ("diag", (1, 1003), "float64", {"ftol": 0.0, "maxiter": 25}),
    ("stencil", (17, 33), "float64", {}),                       # defaults: stops on the f test
]


def cg_problem(op, shape, dt):
    """-> (oracle operator, b): the data of tests (Smooth2D.synthetic: deterministic)."""
    import opkb200 as opk
    rows, cols = shape
    w, y = opk.Smooth2D.synthetic(rows, cols, np.dtype(dt))
    A = opk.Smooth2DOperator.reference(w, 2.0, cols) if op == "stencil" else ("diag", w)
    return A, (w * y).astype(dt), w


def make_conjgrad(rev):
    out = {"oracle_git": rev, "sum_mode": "accurate", "cases": []}
    for op, shape, dt, kw in CG_CASES:
        A, b, _ = cg_problem(op, shape, dt)
        x, info, tr = O.conjgrad(A, b, trace=True, **kw)
        out["cases"].append({
            "operator": op, "shape": list(shape), "dtype": dt, "kw": kw,
            "final": {k: info[k] for k in ("iter", "status", "rho", "phi")},
            "x_final_head": [float(v) for v in x[:8]],
            "trace": [{"iter": t["iter"], "rho": t["rho"], "phi": t["phi"],
                       "x_head": [float(v) for v in t["x"][:8]]} for t in tr],
        })
    with open(os.path.join(HERE, "conjgrad_traces.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", os.path.join(HERE, "conjgrad_traces.json"))


def main():
    O.build()
    O.set_sum_mode(O.SUM_ACCURATE)
    try:
        rev = subprocess.run(["git", "-C", ROOT, "rev-parse", "HEAD"], capture_output=True,
                             text=True).stdout.strip()
    except Exception:
        rev = "unknown"
    if sys.argv[1:] == ["conjgrad"]:      # only the conjgrad fixture (the others stay as committed)
        return make_conjgrad(rev)
    out = {"oracle_git": rev, "sum_mode": "accurate", "cases": []}
    for kind, n, dt, kw in VMLMB_CASES:
        pb = problem(kind, n, np.dtype(dt))
        x, res, tr = O.vmlmb(pb["oracle_obj"], pb["x0"], pb["lower"], pb["upper"], trace=True,
                             maxiter=25, **kw)
        out["cases"].append({
            "kind": kind, "n": n, "dtype": dt, "kw": kw, "maxiter": 25,
            "final": {k: res[k] for k in ("iter", "eval", "rejects", "f", "gnorm", "reason")},
            "trace": [{"iter": t["iter"], "eval": t["eval"], "rejects": t["rejects"], "f": t["f"],
                       "gnorm": t["gnorm"], "stp": t["stp"], "nfree": int(t["mask"].sum()),
                       "x_head": [float(v) for v in t["x"][:8]]} for t in tr],
        })
    with open(os.path.join(HERE, "vmlmb_traces.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    out = {"oracle_git": rev, "sum_mode": "accurate", "cases": []}
    for kind, n, dt, m in SPG_CASES:
        pb = problem(kind, n, np.dtype(dt))
        x, info, tr = O.spg(pb["oracle_obj"], pb["x0"], m, pb["lower"], pb["upper"], trace=True,
                            maxit=30)
        out["cases"].append({
            "kind": kind, "n": n, "dtype": dt, "m": m, "maxit": 30,
            "final": {k: info[k] for k in ("iter", "fcnt", "pcnt", "status", "fbest")},
            "trace": [{"iter": t["iter"], "fcnt": t["fcnt"], "f": t["f"], "pgtwon": t["pgtwon"],
                       "pginfn": t["pginfn"], "x_head": [float(v) for v in t["x"][:8]]} for t in tr],
        })
    with open(os.path.join(HERE, "spg_traces.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", os.path.join(HERE, "vmlmb_traces.json"), os.path.join(HERE, "spg_traces.json"))


if __name__ == "__main__":
    main()
