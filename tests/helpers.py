"""Shared helpers of the parity tests: seeded problems in the two vocabularies
(oracle: numpy arrays; product: opkb200 objects) and trace recording."""
import numpy as np

import opkb200 as opk


def problem(kind, n, dtype, seed=0):
    """-> dict(x0, lower, upper, oracle_obj, gpu_obj) (host arrays / scalars)."""
    dtype = np.dtype(dtype)
    rng = np.random.default_rng(seed)
    if kind == "rosen_unc":
        return dict(x0=opk.Rosenbrock.x0(n, dtype, perturb_seed=7 if n > 20 else None),
                    lower=None, upper=None, oracle_obj="rosenbrock", gpu_obj=opk.Rosenbrock())
    if kind == "rosen_lower0":
        return dict(x0=opk.Rosenbrock.x0(n, dtype, perturb_seed=7 if n > 20 else None),
                    lower=0.0, upper=None, oracle_obj="rosenbrock", gpu_obj=opk.Rosenbrock())
    if kind == "rosen_box":
        # lo = 0; hi = 0.9 for even pair index, 2.0 otherwise (arrays)
        lo = np.zeros(n, dtype=dtype)
        hi = np.where((np.arange(n) // 2) % 2 == 0, 0.9, 2.0).astype(dtype)
        return dict(x0=opk.Rosenbrock.x0(n, dtype, perturb_seed=7), lower=lo, upper=hi,
                    oracle_obj="rosenbrock", gpu_obj=opk.Rosenbrock())
    if kind == "quad_box":
        a, c = opk.Quadratic.synthetic(n, dtype, kappa=1e3)
        lo = np.zeros(n, dtype=dtype)
        hi = np.ones(n, dtype=dtype)
        return dict(x0=np.full(n, 0.5, dtype=dtype), lower=lo, upper=hi,
                    oracle_obj=("quadratic", a, c), gpu_obj=opk.Quadratic(a, c))
    if kind == "quad_upper_scalar":
        a, c = opk.Quadratic.synthetic(n, dtype, kappa=1e2)
        return dict(x0=np.full(n, 0.25, dtype=dtype) + rng.random(n).astype(dtype) * dtype.type(0.1),
                    lower=None, upper=0.75, oracle_obj=("quadratic", a, c), gpu_obj=opk.Quadratic(a, c))
    raise ValueError(kind)


def gpu_bounds(pb):
    import math
    lo = -math.inf if pb["lower"] is None else pb["lower"]
    hi = math.inf if pb["upper"] is None else pb["upper"]
    return lo, hi


def record_vmlmb(trace):
    def obs(s):
        trace.append(dict(iter=s.iter, eval=s.eval, rejects=s.rejects, f=s.f, gnorm=s.gnorm,
                          stp=s.stp, x=s.x.download(), mask=s.free_mask()))
    return obs


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = max(float(np.linalg.norm(b)), 1e-300)
    return float(np.linalg.norm(a - b)) / den


def compare_traces(got, want, niter, rtol, check_mask=True, xtol=None, mask_slack=0.0):
    """Same iterate and active-set sequence for the first `niter` iterations."""
    k = 0
    for g, w in zip(got, want):
        if w["iter"] > niter:
            break
        assert (g["iter"], g["eval"], g["rejects"]) == (w["iter"], w["eval"], w["rejects"]), (g["iter"], w["iter"], g["eval"], w["eval"])
        assert abs(g["f"] - w["f"]) <= rtol * max(abs(w["f"]), 1e-300), ("f", w["iter"], g["f"], w["f"])
        assert abs(g["gnorm"] - w["gnorm"]) <= rtol * max(abs(w["gnorm"]), 1e-300), ("gnorm", w["iter"], g["gnorm"], w["gnorm"])
        assert abs(g["stp"] - w["stp"]) <= rtol * max(abs(w["stp"]), 1e-300), ("stp", w["iter"])
        assert relerr(g["x"], w["x"]) <= (xtol if xtol is not None else rtol), ("x", w["iter"], relerr(g["x"], w["x"]))
        if check_mask and mask_slack > 0:   # opt-in, not parity-grade modes only
            assert np.count_nonzero(g["mask"] != w["mask"]) <= mask_slack * g["mask"].size, ("mask", w["iter"])
        elif check_mask:
            assert np.array_equal(g["mask"], w["mask"]), ("mask", w["iter"])
        k += 1
    assert k >= min(niter, len(want) - 1), "trace too short"
    return k
