"""`spg` / `spg!`: host driver of the reference's SPG2 (src/spg.jl:155-403).

* `prj` = `BoxProjection(lower, upper)` and a built-in objective: the fused
  path, ONE kernel per function evaluation (`opkb_spg_step` / `_backtrack`);
* any callable `prj(dst, src)` / `fg(x, g)` on device Vectors: the reference's
  own sequence of `v*` operations, one kernel each (Tier 1).
"""
from __future__ import annotations

import ctypes as C
import math
import sys

import numpy as np

from . import _lib
from ._lib import check, f64, vp
from .objectives import is_builtin
from .vectors import (Context, Vector, default_context, project_variables_, vcombine_, vcopy_, vdot,
                      vnorm2, vnorminf, _eltype)
from .vmlmb import BoundedSet, _builtin_as_callable, verbose

# src/spg.jl:31-36
SEARCHING = 0
INFNORM_CONVERGENCE = 1
TWONORM_CONVERGENCE = 2
FUNCTION_CONVERGENCE = 3
TOO_MANY_ITERATIONS = -1
TOO_MANY_EVALUATIONS = -2

# src/spg.jl:158-164
REASON = {
    SEARCHING: "Work in progress",
    INFNORM_CONVERGENCE: "Convergence with projected gradient infinite-norm",
    TWONORM_CONVERGENCE: "Convergence with projected gradient 2-norm",
    FUNCTION_CONVERGENCE: "Function does not change in the last `m` iterations",
    TOO_MANY_ITERATIONS: "Too many iterations",
    TOO_MANY_EVALUATIONS: "Too many function evaluations",
}


class Info:
    """SPG.Info (src/spg.jl:38-49)."""

    def __init__(self):
        self.f = 0.0
        self.fbest = 0.0
        self.pginfn = 0.0
        self.pgtwon = 0.0
        self.iter = 0
        self.fcnt = 0
        self.pcnt = 0
        self.status = 0


def getreason(ws: Info) -> str:
    return REASON.get(ws.status, "unknown status")


def default_printer(io, nfo: Info):
    """src/spg.jl:405-414"""
    if nfo.iter == 0:
        io.write("# %s\n# %s\n" % (" ITER   EVAL   PROJ             F(x)              ‖PG(X)‖_2 ‖PG(X)‖_∞",
                                   "---------------------------------------------------------------------"))
    io.write(" %6d %6d %6d %3s %24.17e %9.2e %9.2e\n" % (
        nfo.iter, nfo.fcnt, nfo.pcnt, "(*)" if nfo.f <= nfo.fbest else "   ", nfo.f, nfo.pgtwon,
        nfo.pginfn))


class BoxProjection(BoundedSet):
    """The box `prj!` (for lower = 0, upper = +Inf: `nonnegative!` of
    test/rosenbrock.jl:50-58).  Callable on device Vectors like any `prj!`."""

    def __call__(self, dst: Vector, src: Vector) -> Vector:
        return project_variables_(dst, src, self._lo, self._hi)

    def bind(self, ctx: Context, n: int, dtype):
        self._keep = []

        def conv(b):
            if isinstance(b, Vector):
                return b
            if isinstance(b, (int, float, np.floating, np.integer)):
                return float(b)
            v = ctx.from_numpy(np.ascontiguousarray(b, dtype=dtype).reshape(-1))
            if v.n != n:
                raise ValueError("bound of the wrong length")
            self._keep.append(v)
            return v

        self._lo, self._hi = conv(self.lower), conv(self.upper)
        return self._lo, self._hi


def _clamp(x, lo, hi):
    return hi if x > hi else (lo if x < lo else x)


class SPG:
    """Steppable `_spg!` (src/spg.jl:193-403): `iterate()` is one trip of the
    main loop; `solve()` runs to termination."""

    def __init__(self, fg, prj, x0, m, *, autodiff=False, ws: Info = None, maxit=None, maxfc=None,
                 eps1=1e-6, eps2=1e-6, eta=1.0, printer=default_printer, verb=0, io=None,
                 ctx: Context = None, mode="fused", observer=None):
        if autodiff:
            raise NotImplementedError("autodiff is a host-side hook of the reference (src/autodiff.jl)")
        self.ctx = ctx or (x0.ctx if isinstance(x0, Vector) else default_context())
        self._numpy_in = not isinstance(x0, Vector)
        if self._numpy_in:
            x0 = np.ascontiguousarray(x0)
            if x0.dtype not in (np.float32, np.float64):
                x0 = x0.astype(np.float64)
            self._shape = x0.shape
            x0 = x0.reshape(-1)
        n, dtype = len(x0), np.dtype(x0.dtype)
        self.n, self.dtype = n, dtype
        self.m = int(m)
        big = 2 ** 63 - 1
        self.maxit = big if maxit is None else int(maxit)
        self.maxfc = big if maxfc is None else int(maxfc)
        self.eps1, self.eps2, self.eta = float(eps1), float(eps2), float(eta)
        assert self.m >= 1 and self.eps1 >= 0 and self.eps2 >= 0 and self.eta > 0   # :197-200
        self.ws = ws if ws is not None else Info()
        self.printer, self.verb = printer, int(verb)
        self.io = io if io is not None else sys.stdout
        self.observer = observer
        # :201-211
        self.lmin, self.lmax, self.ftol, self.amin, self.amax = 1e-30, 1e+30, 1e-4, 0.1, 0.9
        self.iter = self.fcnt = self.pcnt = 0
        self.status = SEARCHING
        self.pgtwon = self.pginfn = math.inf
        self.lastfv = [-math.inf] * self.m                                    # :222
        self.finished = False
        self._started = False
        self.f = self.f0 = self.fbest = 0.0

        fused = (mode == "fused" and isinstance(prj, BoundedSet) and is_builtin(fg))
        self.fused = fused
        self._h = None
        L = _lib.lib()
        if isinstance(prj, BoundedSet) and not isinstance(prj, BoxProjection):
            prj = BoxProjection(prj.lower, prj.upper)
        if isinstance(prj, BoxProjection):
            lo, hi = prj.bind(self.ctx, n, dtype)
        self.prj = prj
        if fused:
            h = vp()
            lo_h = lo.handle if isinstance(lo, Vector) else None
            hi_h = hi.handle if isinstance(hi, Vector) else None
            check(L.opkb_spg_create(self.ctx.handle, _eltype(dtype), n, lo_h,
                                    lo if lo_h is None else -math.inf, hi_h,
                                    hi if hi_h is None else math.inf, C.byref(h)))
            self._h = h
            a, c = fg.vectors(self.ctx, n, dtype)
            check(L.opkb_spg_set_objective(h, fg.kind, a.handle if a is not None else None, c.handle if c is not None else None))
            xv = self._vec("x")
            if isinstance(x0, Vector):
                vcopy_(xv, x0)
            else:
                xv.upload(x0)
            self._out = (f64 * 8)()
            self._lambda = 0.0
        else:
            self.fg = _builtin_as_callable(fg, self.ctx, n, dtype) if is_builtin(fg) else fg
            self.x = x0 if isinstance(x0, Vector) else self.ctx.from_numpy(x0)
            # :223-227 (d = pg, s = x0, y = g0 share storage in the reference too)
            self.g = self.x.similar()
            self.pg = self.x.similar()
            self.x0 = self.x.similar()
            self.g0 = self.x.similar()
            self.xbest = self.x.similar()

    def _vec(self, which) -> Vector:
        return Vector.borrowed(self.ctx, _lib.lib().opkb_spg_vector(self._h, ord(which)), self)

    def close(self):
        if self._h:
            if self.ctx is not None and self.ctx.handle:     # (not after the context itself)
                _lib.lib().opkb_spg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- device operations of the two back ends ------------------------------
    def _start(self):
        L = _lib.lib()
        if self.fused:
            check(L.opkb_spg_start(self._h, self.eta, self._out))             # :230-241
            self.f = self._out[0]
        else:
            self.prj(self.x, self.x)                                          # :230
            vcopy_(self.x0, self.x)                                           # :231
            self.f = float(self.fg(self.x, self.g))                           # :236
            vcopy_(self.xbest, self.x)                                        # :241
        self.pcnt += 1
        self.fcnt += 1
        self.fbest = self.f
        self._started = True

    def _pgnorms(self):
        if self.fused:
            return math.sqrt(self._out[2]), self._out[3]
        eta = self.eta
        # :259: pg = (x - prj(x - eta*g))/eta
        vcombine_(self.pg, 1, self.x, -eta, self.g)
        self.prj(self.pg, self.pg)
        vcombine_(self.pg, 1 / eta, self.x, -1 / eta, self.pg)
        return vnorm2(self.pg), vnorminf(self.pg)                             # :261-262

    def _sy(self):
        if self.fused:
            return self._out[4], self._out[5]
        vcombine_(self.x0, 1, self.x, -1, self.x0)                            # s = x - x0 :309
        vcombine_(self.g0, 1, self.g, -1, self.g0)                            # y = g - g0 :310
        sty = vdot(self.x0, self.g0)                                          # :311
        sts = vdot(self.x0, self.x0) if sty > 0 else 0.0                      # :314
        return sty, sts

    def _spectral_step(self, lam):
        """:322-330 and the first evaluation :336; returns (f, delta)."""
        if self.fused:
            self._lambda = lam
            check(_lib.lib().opkb_spg_step(self._h, lam, self.eta, self._out))
            return self._out[0], self._out[1]
        vcopy_(self.x0, self.x)                                               # :322
        vcopy_(self.g0, self.g)                                               # :323
        vcombine_(self.x, 1, self.x0, -lam, self.g0)                          # :327
        self.prj(self.x, self.x)
        vcombine_(self.pg, 1, self.x, -1, self.x0)                            # d = x - x0 :329
        delta = vdot(self.g0, self.pg)                                        # :330
        return float(self.fg(self.x, self.g)), delta                          # :336

    def _backtrack(self, stp):
        if self.fused:
            check(_lib.lib().opkb_spg_backtrack(self._h, self._lambda, stp, self.eta, self._out))
            return self._out[0]
        vcombine_(self.x, 1, self.x0, stp, self.pg)                           # :368
        return float(self.fg(self.x, self.g))                                 # :336

    def _mark_best(self):
        if self.fused:
            check(_lib.lib().opkb_spg_mark_best(self._h))
        else:
            vcopy_(self.xbest, self.x)                                        # :344

    # ---- main loop -------------------------------------------------------------
    def iterate(self) -> bool:
        """One trip of the main loop (:245-380).  Returns False once finished."""
        if self.finished:
            return False
        if not self._started:
            self._start()
        m = self.m
        if m > 1:                                                             # :249-255
            self.lastfv[self.iter % m] = self.f
            fmin, fmax = min(self.lastfv), max(self.lastfv)
            fconst = not (fmin < fmax)
        else:
            fmin = fmax = self.f
            fconst = False
        self.pgtwon, self.pginfn = self._pgnorms()                            # :259-262
        self.pcnt += 1
        if verbose(self.verb, self.iter):                                     # :265-275
            self._fill(self.f)
            self.printer(self.io, self.ws)
        if self.observer is not None:
            self.observer(self)
        # :278-302
        if self.pginfn <= self.eps1:
            return self._stop(INFNORM_CONVERGENCE)
        if self.pgtwon <= self.eps2:
            return self._stop(TWONORM_CONVERGENCE)
        if fconst:
            return self._stop(FUNCTION_CONVERGENCE)
        if self.iter >= self.maxit:
            return self._stop(TOO_MANY_ITERATIONS)
        if self.fcnt >= self.maxfc:
            return self._stop(TOO_MANY_EVALUATIONS)
        # :305-319
        if self.iter == 0:
            lam = _clamp(1 / self.pginfn, self.lmin, self.lmax)
        else:
            sty, sts = self._sy()
            lam = _clamp(sts / sty, self.lmin, self.lmax) if sty > 0 else self.lmax
        self.f0 = self.f                                                      # :324
        self.f, delta = self._spectral_step(lam)                              # :327-336
        self.pcnt += 1
        stp = 1.0                                                             # :333
        while True:
            self.fcnt += 1                                                    # :337
            if self.f < self.fbest:                                           # :342-345
                self.fbest = self.f
                self._mark_best()
            if self.f <= fmax + stp * self.ftol * delta:                      # :348
                break
            if self.fcnt >= self.maxfc:                                       # :352-356
                self.status = TOO_MANY_EVALUATIONS
                break
            q = -delta * (stp * stp)                                          # :359-365
            r = (self.f - self.f0 - stp * delta) * 2
            if r > 0 and self.amin * r <= q <= self.amax * stp * r:
                stp = q / r
            else:
                stp /= 2
            self.f = self._backtrack(stp)                                     # :368, :336
        if self.status != SEARCHING:                                          # :371-375
            return self._stop(self.status)
        self.iter += 1                                                        # :378
        return True

    def _fill(self, f):
        ws = self.ws
        ws.f, ws.fbest, ws.pgtwon, ws.pginfn = f, self.fbest, self.pgtwon, self.pginfn
        ws.iter, ws.fcnt, ws.pcnt, ws.status = self.iter, self.fcnt, self.pcnt, self.status

    def _stop(self, status):
        self.status = status
        self.finished = True
        self._fill(self.fbest)                                                # :383-390
        if verbose(self.verb, 0):                                             # :391-398
            reason = getreason(self.ws)
            if status < 0:
                sys.stderr.write("# WARNING: " + reason + "\n")
            else:
                self.io.write("# SUCCESS: " + reason + "\n")
        return False

    def solve(self):
        while self.iterate():
            pass
        return self.solution()

    def solution(self) -> Vector:
        """:401 -- the best point found."""
        if self.fused:
            return self._vec("b") if self.fbest < self.f else self._vec("x")
        return self.xbest if self.fbest < self.f else self.x

    def result(self):
        v = self.solution()
        if self._numpy_in:
            return v.download().reshape(self._shape)
        return v


class NativeSPG(SPG):
    """`spg!` with the loop inside libopkb200 (`opkb_spg_solver_*`): same keywords and results
    as `SPG` for a built-in objective and a box (`BoundedSet`), one native call per `iterate(k)` /
    `solve()`.  `printer` / `verb` are host-side hooks of the Python loop and are not called here
    (use `observer`, invoked after every `iterate` call)."""

    def __init__(self, fg, prj, x0, m, **kw):
        kw.pop("mode", None)
        super().__init__(fg, prj, x0, m, mode="fused", **kw)
        if not self.fused:
            raise TypeError("the native SPG loop needs a built-in objective and a BoundedSet")
        self._solver = vp()
        big = 2 ** 63 - 1
        check(_lib.lib().opkb_spg_solver_create(self._h, self.m, min(self.maxit, big), min(self.maxfc, big),
                                                self.eps1, self.eps2, self.eta, C.byref(self._solver)))
        self._best_is_x = True

    def close(self):
        if getattr(self, "_solver", None):
            _lib.lib().opkb_spg_solver_destroy(self._solver)
            self._solver = None
        super().close()

    def _refresh(self):
        f, fb, pi, p2 = f64(), f64(), f64(), f64()
        it, fc, pc = _lib.i64(), _lib.i64(), _lib.i64()
        st, bx = C.c_int(), C.c_int()
        check(_lib.lib().opkb_spg_solver_info(self._solver, C.byref(f), C.byref(fb), C.byref(pi), C.byref(p2),
                                              C.byref(it), C.byref(fc), C.byref(pc), C.byref(st), C.byref(bx)))
        self.f, self.fbest, self.pginfn, self.pgtwon = f.value, fb.value, pi.value, p2.value
        self.iter, self.fcnt, self.pcnt, self.status = it.value, fc.value, pc.value, st.value
        self._best_is_x = bool(bx.value)

    def iterate(self, niter: int = 1) -> bool:
        if self.finished:
            return False
        alive = C.c_int()
        check(_lib.lib().opkb_spg_solver_run(self._solver, int(min(niter, 2 ** 62)), C.byref(alive)))
        self._refresh()
        if self.observer is not None:
            self.observer(self)
        if not alive.value:
            self.finished = True
            self._fill(self.fbest)
        return bool(alive.value)

    def solve(self):
        while self.iterate(2 ** 62):
            pass
        return self.solution()

    def solution(self) -> Vector:
        return self._vec("x") if self._best_is_x else self._vec("b")


def spg_(fg, prj, x, m, driver="python", **kwds):
    """`spg!(fg!, prj!, x, m; kwds...) -> x` (src/spg.jl:168-191).  `driver="native"` runs the
    loop inside the library (built-in objective and box only)."""
    solver = (NativeSPG if driver == "native" else SPG)(fg, prj, x, m, **kwds)
    solver.solve()
    if isinstance(x, Vector):
        sol = solver.solution()
        if sol is not x:
            vcopy_(x, sol)
    else:
        x[...] = solver.result().astype(x.dtype, copy=False)
    solver.close()
    return x


def spg(fg, prj, x0, m, **kwds):
    """`spg(fg!, prj!, x0, m; kwds...) -> x` (src/spg.jl:155-156)."""
    if isinstance(x0, Vector):
        x = vcopy_(x0.similar(), x0)
    else:
        x = np.array(x0, copy=True)
        if x.dtype not in (np.float32, np.float64):
            x = x.astype(np.float64)
    return spg_(fg, prj, x, m, **kwds)
