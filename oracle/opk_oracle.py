"""ctypes binding of the CPU oracle (libopkoracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, by ``__graft_entry__.smoke()``
and by the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The
product package never imports this module.  PARITY UNPINNED: see
``opk_oracle.h`` (no golden vectors in the reference, no Julia runtime here).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SUM_SEQUENTIAL, SUM_PAIRWISE, SUM_ACCURATE = 0, 1, 2
LS_ARMIJO, LS_MORE_TORALDO, LS_MORE_THUENTE = 1, 2, 3
TASK_START, TASK_SEARCH, TASK_CONVERGENCE, TASK_WARNING, TASK_ERROR = range(5)
TASK_NAMES = {0: "START", 1: "SEARCH", 2: "CONVERGENCE", 3: "WARNING", 4: "ERROR"}

i64 = C.c_int64
f64 = C.c_double


class LineSearchStruct(C.Structure):
    _fields_ = [("kind", C.c_int), ("task", C.c_int), ("status", C.c_int),
                ("ftol", f64), ("finit", f64), ("ginit", f64), ("stp", f64),
                ("stpmin", f64), ("gamma1", f64), ("gamma2", f64),
                ("strict", C.c_int), ("gtol", f64), ("xtol", f64),
                ("stpmax", f64), ("stx", f64), ("fx", f64), ("gx", f64),
                ("sty", f64), ("fy", f64), ("gy", f64), ("lower", f64),
                ("upper", f64), ("width", f64), ("oldwidth", f64),
                ("stage", C.c_int), ("brackt", C.c_int)]


class VmlmbOptions(C.Structure):
    _fields_ = [("mem", i64), ("blmvm", C.c_int), ("fmin", f64),
                ("maxiter", i64), ("maxeval", i64), ("xatol", f64),
                ("xrtol", f64), ("fatol", f64), ("frtol", f64), ("gatol", f64),
                ("grtol", f64), ("epsilon", f64), ("lnsrch", C.c_int),
                ("ls_ftol", f64), ("ls_gtol", f64), ("ls_xtol", f64),
                ("ls_gamma1", f64), ("ls_gamma2", f64), ("ls_strict", C.c_int)]


class VmlmbResult(C.Structure):
    _fields_ = [("iter", i64), ("eval", i64), ("rejects", i64), ("f", f64),
                ("gnorm", f64), ("stp", f64), ("stage", C.c_int),
                ("reason", C.c_char * 96)]


class SpgInfo(C.Structure):
    _fields_ = [("f", f64), ("fbest", f64), ("pginfn", f64), ("pgtwon", f64),
                ("iter", i64), ("fcnt", i64), ("pcnt", i64), ("status", C.c_int)]


class CgInfo(C.Structure):
    _fields_ = [("rho", f64), ("phi", f64), ("psi", f64), ("iter", i64), ("status", C.c_int)]


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (gcc); returns the .so path."""
    so = os.path.join(_HERE, "libopkoracle.so")
    srcs = [os.path.join(_HERE, f) for f in
            ("opk_oracle.c", "opk_oracle_impl.h", "opk_oracle.h", "Makefile")]
    stale = (not os.path.exists(so) or
             any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True,
                       capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libopkoracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.opko_set_sum_mode.argtypes = [C.c_int]
        L.opko_get_sum_mode.restype = C.c_int
        L.opko_set_num_threads.argtypes = [C.c_int]
        L.opko_get_max_threads.restype = C.c_int
        L.opko_ls_reason.restype = C.c_char_p
        L.opko_ls_reason.argtypes = [C.POINTER(LineSearchStruct)]
        L.opko_ls_armijo.argtypes = [C.POINTER(LineSearchStruct), f64]
        L.opko_ls_more_toraldo.argtypes = [C.POINTER(LineSearchStruct), f64, f64, f64, C.c_int]
        L.opko_ls_more_thuente.argtypes = [C.POINTER(LineSearchStruct), f64, f64, f64]
        L.opko_ls_usederivatives.argtypes = [C.POINTER(LineSearchStruct)]
        L.opko_ls_start.argtypes = [C.POINTER(LineSearchStruct), f64, f64, f64, f64, f64]
        L.opko_ls_iterate.argtypes = [C.POINTER(LineSearchStruct), f64, f64, f64]
        L.opko_cstep.argtypes = [C.POINTER(C.c_int), f64, f64, C.POINTER(f64 * 7), f64, f64]
        L.opko_vmlmb_default_options.argtypes = [C.POINTER(VmlmbOptions)]
        for pre, ct in (("opko_d_", C.c_double), ("opko_f_", C.c_float)):
            P = C.POINTER(ct)
            PP = C.POINTER(P)
            Pi = C.POINTER(i64)
            Pd = C.POINTER(f64)

            def sig(name, restype, *args, _pre=pre):
                fn = getattr(L, _pre + name)
                fn.restype = restype
                fn.argtypes = list(args)

            sig("vdot", f64, i64, P, P)
            sig("vdot_sel", f64, i64, Pi, P, P)
            sig("vnorm2", f64, i64, P)
            sig("vnorminf", f64, i64, P)
            sig("vcopy", None, i64, P, P)
            sig("vscale", None, i64, P, f64)
            sig("vupdate", None, i64, P, f64, P)
            sig("vupdate_sel", None, P, i64, Pi, f64, P)
            sig("vcombine", None, i64, P, f64, P, f64, P)
            sig("vproduct", None, i64, P, P, P)
            sig("project_variables", None, i64, P, P, P, f64, P, f64)
            sig("project_direction", C.c_int, i64, P, P, P, f64, P, f64, C.c_int, P)
            sig("step_limits", C.c_int, i64, P, P, f64, P, f64, C.c_int, P, Pd, Pd)
            sig("get_free_variables", i64, i64, Pi, P)
            sig("get_free_variables_dir", i64, i64, Pi, P, P, f64, P, f64, C.c_int, P)
            sig("apply_lbfgs", C.c_int, i64, i64, PP, PP, Pd, f64, P, i64, i64, P, Pd)
            sig("apply_lbfgs_sel", C.c_int, i64, i64, PP, PP, Pd, i64, i64, P, Pd, i64, Pi)
            sig("rosenbrock_fg", f64, C.c_void_p, i64, P, P)
            sig("rosenbrock_init", None, i64, P)
            sig("quadratic_fg", f64, C.c_void_p, i64, P, P)
            sig("box_prj", None, C.c_void_p, i64, P, P)
            sig("vmlmb", C.c_int, C.c_void_p, C.c_void_p, i64, P, P, f64, P, f64,
                C.POINTER(VmlmbOptions), C.POINTER(VmlmbResult), C.c_void_p, C.c_void_p)
            sig("spg", C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, i64, P,
                i64, i64, i64, f64, f64, f64, C.POINTER(SpgInfo), C.c_void_p, C.c_void_p)
            sig("conjgrad", C.c_int, C.c_void_p, C.c_void_p, i64, P, P, f64, f64, f64, f64, i64, i64,
                C.POINTER(CgInfo), C.c_void_p, C.c_void_p)
            sig("diag_apply", None, C.c_void_p, i64, P, P)
        _LIB = L
    return _LIB


# --------------------------------------------------------------------------
def set_sum_mode(mode: int) -> None:
    lib().opko_set_sum_mode(mode)


def set_num_threads(n: int) -> None:
    lib().opko_set_num_threads(n)


def max_threads() -> int:
    return lib().opko_get_max_threads()


def _ct(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "opko_d_", C.c_double
    if dtype == np.float32:
        return "opko_f_", C.c_float
    raise TypeError("oracle handles float32/float64 only")


def _ptr(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def _fn(x, name):
    pre, ct = _ct(x.dtype)
    return getattr(lib(), pre + name), ct


def _bound(b, dtype, ct, default):
    """-> (array pointer or NULL, scalar value, keep-alive)"""
    if b is None:
        return None, default, None
    if np.isscalar(b):
        return None, float(b), None
    arr = np.ascontiguousarray(b, dtype=dtype)
    return _ptr(arr, ct), default, arr


# ---- vector ops (numpy in, numpy out / in place) ---------------------------
def vdot(x, y):
    fn, ct = _fn(x, "vdot")
    return fn(x.size, _ptr(x, ct), _ptr(y, ct))


def vdot_sel(sel, x, y):
    fn, ct = _fn(x, "vdot_sel")
    sel = np.ascontiguousarray(sel, dtype=np.int64)
    return fn(sel.size, sel.ctypes.data_as(C.POINTER(i64)), _ptr(x, ct), _ptr(y, ct))


def vnorm2(x):
    fn, ct = _fn(x, "vnorm2")
    return fn(x.size, _ptr(x, ct))


def vnorminf(x):
    fn, ct = _fn(x, "vnorminf")
    return fn(x.size, _ptr(x, ct))


def vcopy(dst, src):
    fn, ct = _fn(dst, "vcopy")
    fn(dst.size, _ptr(dst, ct), _ptr(src, ct))
    return dst


def vscale(x, alpha):
    fn, ct = _fn(x, "vscale")
    fn(x.size, _ptr(x, ct), alpha)
    return x


def vupdate(y, alpha, x):
    fn, ct = _fn(y, "vupdate")
    fn(y.size, _ptr(y, ct), alpha, _ptr(x, ct))
    return y


def vupdate_sel(y, sel, alpha, x):
    fn, ct = _fn(y, "vupdate_sel")
    sel = np.ascontiguousarray(sel, dtype=np.int64)
    fn(_ptr(y, ct), sel.size, sel.ctypes.data_as(C.POINTER(i64)), alpha, _ptr(x, ct))
    return y


def vcombine(dst, alpha, x, beta, y):
    fn, ct = _fn(dst, "vcombine")
    fn(dst.size, _ptr(dst, ct), alpha, _ptr(x, ct), beta, _ptr(y, ct))
    return dst


def vproduct(dst, x, y):
    fn, ct = _fn(dst, "vproduct")
    fn(dst.size, _ptr(dst, ct), _ptr(x, ct), _ptr(y, ct))
    return dst


def project_variables(dst, src, lo=None, hi=None):
    fn, ct = _fn(dst, "project_variables")
    lp, lv, k1 = _bound(lo, dst.dtype, ct, -math.inf)
    hp, hv, k2 = _bound(hi, dst.dtype, ct, math.inf)
    fn(dst.size, _ptr(dst, ct), _ptr(src, ct), lp, lv, hp, hv)
    return dst


def project_direction(dst, x, lo, hi, orient, d):
    fn, ct = _fn(dst, "project_direction")
    lp, lv, k1 = _bound(lo, dst.dtype, ct, -math.inf)
    hp, hv, k2 = _bound(hi, dst.dtype, ct, math.inf)
    rc = fn(dst.size, _ptr(dst, ct), _ptr(x, ct), lp, lv, hp, hv, int(orient), _ptr(d, ct))
    if rc == -1:
        raise ValueError("invalid orientation")
    if rc == -2:
        raise ValueError("invalid bounds")
    return dst


def step_limits(x, lo, hi, orient, d):
    fn, ct = _fn(x, "step_limits")
    lp, lv, k1 = _bound(lo, x.dtype, ct, -math.inf)
    hp, hv, k2 = _bound(hi, x.dtype, ct, math.inf)
    smin, smax = f64(), f64()
    rc = fn(x.size, _ptr(x, ct), lp, lv, hp, hv, int(orient), _ptr(d, ct),
            C.byref(smin), C.byref(smax))
    if rc == -1:
        raise ValueError("invalid orientation")
    if rc == -2:
        raise ValueError("invalid bounds")
    return smin.value, smax.value


def get_free_variables(gp):
    fn, ct = _fn(gp, "get_free_variables")
    sel = np.empty(gp.size, dtype=np.int64)
    j = fn(gp.size, sel.ctypes.data_as(C.POINTER(i64)), _ptr(gp, ct))
    return sel[:j].copy()


def get_free_variables_dir(x, lo, hi, orient, d):
    fn, ct = _fn(x, "get_free_variables_dir")
    lp, lv, k1 = _bound(lo, x.dtype, ct, -math.inf)
    hp, hv, k2 = _bound(hi, x.dtype, ct, math.inf)
    sel = np.empty(x.size, dtype=np.int64)
    j = fn(x.size, sel.ctypes.data_as(C.POINTER(i64)), _ptr(x, ct), lp, lv, hp, hv, int(orient),
           _ptr(d, ct))
    if j < 0:
        raise ValueError("invalid orientation" if j == -1 else "invalid bounds")
    return sel[:j].copy()


def _ptr_array(vecs, ct):
    arr = (C.POINTER(ct) * len(vecs))(*[_ptr(v, ct) for v in vecs])
    return arr


def apply_lbfgs(S, Y, rho, gamma, m, mark, v, diag=None):
    """0-based `mark`; returns (modif, alpha)."""
    fn, ct = _fn(v, "apply_lbfgs")
    mem = len(S)
    rho = np.ascontiguousarray(rho, dtype=np.float64)
    alpha = np.zeros(mem)
    Pd = C.POINTER(f64)
    modif = fn(v.size, mem, _ptr_array(S, ct), _ptr_array(Y, ct),
               rho.ctypes.data_as(Pd), float(gamma),
               None if diag is None else _ptr(diag, ct), m, mark, _ptr(v, ct),
               alpha.ctypes.data_as(Pd))
    return bool(modif), alpha


def apply_lbfgs_sel(S, Y, rho, m, mark, v, sel):
    """0-based `mark`; rho is overwritten; returns (modif, alpha)."""
    fn, ct = _fn(v, "apply_lbfgs_sel")
    mem = len(S)
    alpha = np.zeros(mem)
    Pd = C.POINTER(f64)
    sel = np.ascontiguousarray(sel, dtype=np.int64)
    modif = fn(v.size, mem, _ptr_array(S, ct), _ptr_array(Y, ct),
               rho.ctypes.data_as(Pd), m, mark, _ptr(v, ct),
               alpha.ctypes.data_as(Pd), sel.size,
               sel.ctypes.data_as(C.POINTER(i64)))
    return bool(modif), alpha


# ---- objectives -----------------------------------------------------------
def rosenbrock_init(n, dtype=np.float64):
    x = np.empty(n, dtype=dtype)
    fn, ct = _fn(x, "rosenbrock_init")
    fn(n, _ptr(x, ct))
    return x


def rosenbrock_fg(x, g):
    fn, ct = _fn(x, "rosenbrock_fg")
    return fn(None, x.size, _ptr(x, ct), _ptr(g, ct))


def quadratic_fg(a, c, x, g):
    fn, ct = _fn(x, "quadratic_fg")
    q = (C.c_void_p * 2)(a.ctypes.data, c.ctypes.data)
    return fn(C.cast(q, C.c_void_p), x.size, _ptr(x, ct), _ptr(g, ct))


# ---- line searches ---------------------------------------------------------
class LineSearch:
    """Thin object wrapper over the C line-search state machine."""

    def __init__(self, kind, **kw):
        self.s = LineSearchStruct()
        L = lib()
        if kind == LS_ARMIJO:
            L.opko_ls_armijo(C.byref(self.s), kw.get("ftol", 1e-4))
        elif kind == LS_MORE_TORALDO:
            g = kw.get("gamma", (0.1, 0.5))
            L.opko_ls_more_toraldo(C.byref(self.s), kw.get("ftol", 1e-4), g[0], g[1],
                                   int(kw.get("strict", False)))
        elif kind == LS_MORE_THUENTE:
            L.opko_ls_more_thuente(C.byref(self.s), kw.get("ftol", 1e-3),
                                   kw.get("gtol", 0.9), kw.get("xtol", 0.1))
        else:
            raise ValueError("unknown line search")

    def start(self, f0, g0, stp, stpmin=0.0, stpmax=math.inf):
        return lib().opko_ls_start(C.byref(self.s), f0, g0, stp, stpmin, stpmax)

    def iterate(self, stp, f, g):
        return lib().opko_ls_iterate(C.byref(self.s), stp, f, g)

    @property
    def stp(self):
        return self.s.stp

    @property
    def task(self):
        return self.s.task

    @property
    def status(self):
        return self.s.status

    @property
    def usederivatives(self):
        return bool(lib().opko_ls_usederivatives(C.byref(self.s)))

    def reason(self):
        return lib().opko_ls_reason(C.byref(self.s)).decode()


# ---- drivers ---------------------------------------------------------------
def _objective(obj, dtype, ct):
    """obj: 'rosenbrock' | ('quadratic', a, c) | python callable f = fg(x, g).
    -> (function pointer as c_void_p, ctx c_void_p, keep-alive list)"""
    pre, _ = _ct(dtype)
    L = lib()
    P = C.POINTER(ct)
    if isinstance(obj, str) and obj == "rosenbrock":
        return C.cast(getattr(L, pre + "rosenbrock_fg"), C.c_void_p), None, []
    if isinstance(obj, tuple) and obj[0] == "quadratic":
        a = np.ascontiguousarray(obj[1], dtype=dtype)
        c = np.ascontiguousarray(obj[2], dtype=dtype)
        q = (C.c_void_p * 2)(a.ctypes.data, c.ctypes.data)
        return (C.cast(getattr(L, pre + "quadratic_fg"), C.c_void_p),
                C.cast(q, C.c_void_p), [a, c, q])
    if callable(obj):
        FG = C.CFUNCTYPE(f64, C.c_void_p, i64, P, P)

        def _cb(_ctx, n, xp, gp):
            x = np.ctypeslib.as_array(xp, shape=(n,))
            g = np.ctypeslib.as_array(gp, shape=(n,))
            return float(obj(x, g))

        cb = FG(_cb)
        return C.cast(cb, C.c_void_p), None, [cb]
    raise TypeError("unknown objective")


def vmlmb(obj, x0, lower=None, upper=None, *, mem=5, blmvm=False, fmin=-math.inf,
          maxiter=None, maxeval=None, xtol=(0.0, 1e-7), ftol=(0.0, 1e-8),
          gtol=(0.0, 1e-6), epsilon=0.0, lnsrch=0, ls_ftol=1e-3, ls_gtol=0.9,
          ls_xtol=0.1, ls_gamma=(0.1, 0.5), ls_strict=False, trace=False):
    """Oracle `vmlmb` (src/quasinewton.jl:215).  Returns (x, result dict,
    trace list).  Each trace entry = dict(iter, eval, rejects, f, gnorm, stp,
    x, mask) recorded at the reference's `printer` call site (:501-503).
    `trace` may also be a callable: it is then called at that site with a dict of
    the scalars and VIEWS of x and p (no copies; valid during the call only)."""
    x = np.array(x0, copy=True)
    dtype = x.dtype
    pre, ct = _ct(dtype)
    L = lib()
    P = C.POINTER(ct)
    n = x.size
    opt = VmlmbOptions()
    L.opko_vmlmb_default_options(C.byref(opt))
    opt.mem = min(mem, n) if n > 0 else mem
    opt.blmvm = int(blmvm)
    opt.fmin = fmin
    if maxiter is not None:
        opt.maxiter = maxiter
    if maxeval is not None:
        opt.maxeval = maxeval
    opt.xatol, opt.xrtol = xtol
    opt.fatol, opt.frtol = ftol
    opt.gatol, opt.grtol = gtol
    opt.epsilon = epsilon
    opt.lnsrch = lnsrch
    opt.ls_ftol, opt.ls_gtol, opt.ls_xtol = ls_ftol, ls_gtol, ls_xtol
    opt.ls_gamma1, opt.ls_gamma2 = ls_gamma
    opt.ls_strict = int(ls_strict)
    res = VmlmbResult()
    fgp, fgctx, keep = _objective(obj, dtype, ct)
    lp, lv, k1 = _bound(lower, dtype, ct, -math.inf)
    hp, hv, k2 = _bound(upper, dtype, ct, math.inf)
    records = []
    TR = C.CFUNCTYPE(None, C.c_void_p, i64, i64, i64, f64, f64, f64, i64, P, P, P)

    def _tr(_ctx, it, ev, rej, f, gnorm, stp, nn, xp, gp, pp):
        if callable(trace):
            # light trace (large problems, timing): scalars + VIEWS valid during the call only
            trace(dict(iter=it, eval=ev, rejects=rej, f=f, gnorm=gnorm, stp=stp,
                       x=np.ctypeslib.as_array(xp, shape=(nn,)),
                       p=np.ctypeslib.as_array(pp, shape=(nn,)) if pp else None))
            return
        xx = np.ctypeslib.as_array(xp, shape=(nn,)).copy()
        if pp:
            mask = np.ctypeslib.as_array(pp, shape=(nn,)) != 0
            pv = np.ctypeslib.as_array(pp, shape=(nn,)).copy()
        else:
            mask = np.ones(nn, dtype=bool)
            pv = None
        records.append(dict(iter=it, eval=ev, rejects=rej, f=f, gnorm=gnorm,
                            stp=stp, x=xx, mask=mask, p=pv))

    trcb = TR(_tr) if trace else None
    rc = getattr(L, pre + "vmlmb")(
        fgp, fgctx, n, _ptr(x, ct), lp, lv, hp, hv, C.byref(opt), C.byref(res),
        C.cast(trcb, C.c_void_p) if trcb else None, None)
    out = dict(rc=rc, iter=res.iter, eval=res.eval, rejects=res.rejects, f=res.f,
               gnorm=res.gnorm, stp=res.stp, stage=res.stage,
               reason=res.reason.decode())
    return x, out, records


def spg(obj, x0, m, lower=None, upper=None, prj=None, *, maxit=None, maxfc=None,
        eps1=1e-6, eps2=1e-6, eta=1.0, trace=False):
    """Oracle `spg` (src/spg.jl:155) with the box projection (lower/upper) or a
    python `prj(dst, src)`.  Returns (x, info dict, trace list)."""
    x = np.array(x0, copy=True)
    dtype = x.dtype
    pre, ct = _ct(dtype)
    L = lib()
    P = C.POINTER(ct)
    n = x.size
    fgp, fgctx, keep = _objective(obj, dtype, ct)
    if prj is None:
        class Box(C.Structure):
            _fields_ = [("lo", C.c_void_p), ("lo_val", f64), ("hi", C.c_void_p),
                        ("hi_val", f64)]
        lp, lv, k1 = _bound(lower, dtype, ct, -math.inf)
        hp, hv, k2 = _bound(upper, dtype, ct, math.inf)
        box = Box(C.cast(lp, C.c_void_p) if lp else None, lv,
                  C.cast(hp, C.c_void_p) if hp else None, hv)
        prjp = C.cast(getattr(L, pre + "box_prj"), C.c_void_p)
        prjctx = C.cast(C.pointer(box), C.c_void_p)
    else:
        PRJ = C.CFUNCTYPE(None, C.c_void_p, i64, P, P)

        def _p(_ctx, nn, dp, sp):
            prj(np.ctypeslib.as_array(dp, shape=(nn,)),
                np.ctypeslib.as_array(sp, shape=(nn,)))

        pcb = PRJ(_p)
        prjp, prjctx = C.cast(pcb, C.c_void_p), None
    info = SpgInfo()
    records = []
    TR = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(SpgInfo), i64, P, P)

    def _tr(_ctx, nfo, nn, xp, pgp):
        nf = nfo.contents
        if callable(trace):   # light trace: scalars + a VIEW of x (valid during the call only)
            trace(dict(iter=nf.iter, fcnt=nf.fcnt, pcnt=nf.pcnt, f=nf.f, fbest=nf.fbest,
                       pgtwon=nf.pgtwon, pginfn=nf.pginfn,
                       x=np.ctypeslib.as_array(xp, shape=(nn,))))
            return
        records.append(dict(iter=nf.iter, fcnt=nf.fcnt, pcnt=nf.pcnt, f=nf.f,
                            fbest=nf.fbest, pgtwon=nf.pgtwon, pginfn=nf.pginfn,
                            x=np.ctypeslib.as_array(xp, shape=(nn,)).copy()))

    trcb = TR(_tr) if trace else None
    big = np.iinfo(np.int64).max
    rc = getattr(L, pre + "spg")(
        fgp, fgctx, prjp, prjctx, n, _ptr(x, ct), m,
        big if maxit is None else maxit, big if maxfc is None else maxfc,
        eps1, eps2, eta, C.byref(info),
        C.cast(trcb, C.c_void_p) if trcb else None, None)
    out = dict(rc=rc, f=info.f, fbest=info.fbest, pginfn=info.pginfn,
               pgtwon=info.pgtwon, iter=info.iter, fcnt=info.fcnt,
               pcnt=info.pcnt, status=info.status)
    return x, out, records


def conjgrad(A, b, x0=None, *, ftol=1e-8, gtol=(0.0, 0.0), xtol=0.0, maxiter=None, restart=None,
             trace=False):
    """Oracle `conjgrad` (LazyAlgebra, re-exported by src/OptimPackNextGen.jl:18-19, :36;
    call site test/conjgrad-tests.jl:30).  A: ('diag', a) | python callable A(dst, src) on
    numpy arrays.  Returns (x, info dict, trace list of dict(iter, phi, rho, x))."""
    b = np.ascontiguousarray(b)
    dtype = b.dtype
    pre, ct = _ct(dtype)
    L = lib()
    P = C.POINTER(ct)
    n = b.size
    x = np.zeros(n, dtype=dtype) if x0 is None else np.array(x0, dtype=dtype, copy=True).reshape(-1)
    keep = []
    if isinstance(A, tuple) and A[0] == "diag":
        a = np.ascontiguousarray(A[1], dtype=dtype)
        keep.append(a)
        app, actx = C.cast(getattr(L, pre + "diag_apply"), C.c_void_p), C.c_void_p(a.ctypes.data)
    else:
        OP = C.CFUNCTYPE(None, C.c_void_p, i64, P, P)

        def _a(_ctx, nn, dp, sp):
            A(np.ctypeslib.as_array(dp, shape=(nn,)), np.ctypeslib.as_array(sp, shape=(nn,)))

        cb = OP(_a)
        keep.append(cb)
        app, actx = C.cast(cb, C.c_void_p), None
    records = []
    TR = C.CFUNCTYPE(None, C.c_void_p, i64, f64, f64, i64, P, P)

    def _tr(_ctx, k, phi, rho, nn, xp, rp):
        records.append(dict(iter=k, phi=phi, rho=rho, x=np.ctypeslib.as_array(xp, shape=(nn,)).copy()))

    trcb = TR(_tr) if trace else None
    info = CgInfo()
    big = np.iinfo(np.int64).max
    rc = getattr(L, pre + "conjgrad")(
        app, actx, n, _ptr(x, ct), _ptr(b.reshape(-1), ct), ftol, gtol[0], gtol[1], xtol,
        big if maxiter is None else maxiter, min(50, max(n, 1)) if restart is None else restart,
        C.byref(info), C.cast(trcb, C.c_void_p) if trcb else None, None)
    if rc != 0:
        raise ValueError("invalid conjgrad option")
    return x, dict(iter=info.iter, status=info.status, rho=info.rho, phi=info.phi, psi=info.psi), records
