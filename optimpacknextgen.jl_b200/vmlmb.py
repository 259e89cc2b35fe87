"""`vmlmb` / `vmlmb!`: host driver of the reference (src/quasinewton.jl:215-612).

The stage machine, the line searches and the convergence tests run here on
Float64 scalars; every n-vector stays on the GPU.  Two device back ends exist:

* ``mode="fused"`` (default) -- three kernels per iteration:
  P1 trial (`opkb_vmlmb_trial`), P3 Gram sweep (`opkb_vmlmb_gram`) and P4
  direction (`opkb_vmlmb_direction`); the two-loop recursion is solved on the
  host in coefficient space from the Gram block.
* ``mode="primitive"`` -- one kernel per reference call (Tier 1), i.e. the
  reference's own sequence of `v*` / `SimpleBounds` operations; this is the
  parity-grade path and the one a LazyAlgebra-style plugin would use.

A Julia host does the same through `ccall`.
"""
from __future__ import annotations

import ctypes as C
import math
import sys

import numpy as np

from . import _lib
from ._lib import check, f64, vp
from .linesearches import LineSearch, MoreThuenteLineSearch, MoreToraldoLineSearch
from .objectives import is_builtin
from .vectors import (Context, Vector, default_context, project_direction_, project_variables_,
                      step_limits, vcombine_, vcopy_, vdot, vnorm2, vscale_, vupdate_, _eltype)

EMULATE_BLMVM = 1  # src/quasinewton.jl:40


class BoundedSet:
    """`BoundedSet(lower, upper)`: the box of `lower=` / `upper=` keywords of
    `vmlmb` (src/quasinewton.jl:228-229) as one object.  (The reference only
    sketches it, doc/ConvexSets.md:7-10; semantics are those of src/bounds.jl.)"""

    def __init__(self, lower=-math.inf, upper=math.inf):
        self.lower = lower
        self.upper = upper


def _jl_max(a, b):
    if a != a:
        return a
    if b != b:
        return b
    return a if a > b else b


def _jl_min(a, b):
    if a != a:
        return a
    if b != b:
        return b
    return a if a < b else b


def verbose(verb: int, it: int) -> bool:
    """src/quasinewton.jl:641-642"""
    return verb > 0 and it % verb == 0


def print_iteration(io, it, ev, rejects, f, gnorm, step):
    """src/quasinewton.jl:649-660"""
    if it == 0:
        io.write("#%s%s\n#%s%s\n" % (" ITER   EVAL   REJECTS", "          F(X)           ||G(X)||    STEP",
                                     "----------------------",
                                     "-------------------------------------------"))
    io.write(" %5d  %5d  %5d  %24.16E %9.2E %9.2E\n" % (it, ev, rejects, f, gnorm, step))


def sufficient_descent(gd, eps, gnorm, dnorm) -> bool:
    """src/quasinewton.jl:631-632"""
    return (gd <= -eps * dnorm * gnorm) if eps > 0 else (gd < 0)


# ---------------------------------------------------------------------------
# apply_lbfgs! on device vectors, one kernel per reference call
def apply_lbfgs_(S, Y, rho, H0, m, mark, v, alpha):
    """src/quasinewton.jl:716-745.  `H0` is a scalar gamma (:703-708), a Vector
    of diagonal terms (:710-714, vproduct!) or a callable `H0(v)`.  0-based mark."""
    from .vectors import vproduct_
    mem = min(len(S), len(Y), len(rho), len(alpha))
    assert 1 <= m <= mem and 0 <= mark < mem
    modif = False
    k = mark + 1
    for _ in range(m):
        k = k - 1 if k > 0 else mem - 1
        if rho[k] > 0:
            alpha[k] = vdot(S[k], v) / rho[k]
            vupdate_(v, -alpha[k], Y[k])
            modif = True
    if modif:
        if callable(H0):
            H0(v)
        elif isinstance(H0, Vector):
            vproduct_(v, H0, v)
        else:
            assert H0 > 0
            vscale_(v, H0)
        for _ in range(m):
            if rho[k] > 0:
                beta = vdot(Y[k], v) / rho[k]
                vupdate_(v, alpha[k] - beta, S[k])
            k = k + 1 if k < mem - 1 else 0
    return modif


def apply_lbfgs_masked_(S, Y, rho, m, mark, v, alpha, w):
    """src/quasinewton.jl:747-779; the index list `sel` is the mask w != 0."""
    mem = min(len(S), len(Y), len(rho), len(alpha))
    assert 1 <= m <= mem and 0 <= mark < mem
    gamma = 0.0
    k = mark + 1
    for _ in range(m):
        k = k - 1 if k > 0 else mem - 1
        rho[k] = vdot(w, Y[k], S[k])
        if rho[k] > 0:
            alpha[k] = vdot(w, S[k], v) / rho[k]
            vupdate_(v, w, -alpha[k], Y[k])
            if gamma == 0:
                gamma = rho[k] / vdot(w, Y[k], Y[k])
    if gamma != 0:
        vscale_(v, gamma)
        for _ in range(m):
            if rho[k] > 0:
                beta = vdot(w, Y[k], v) / rho[k]
                vupdate_(v, w, alpha[k] - beta, S[k])
            k = k + 1 if k < mem - 1 else 0
    return gamma != 0


# ---------------------------------------------------------------------------
class _PrimitiveOps:
    """Device back end made of Tier-1 calls only: the reference's sequence."""

    def __init__(self, solver, x: Vector):
        self.s = solver
        self.x = x
        self.g = x.similar()
        self.p = x.similar() if solver.method > 0 else None
        self.d = x.similar()
        self.sv = x.similar()
        self.S = [x.similar() for _ in range(solver.mem)]
        self.Y = [x.similar() for _ in range(solver.mem)]
        self.rho = [0.0] * solver.mem
        self.alpha = [0.0] * solver.mem
        self._mark = 0

    def evaluate(self, mark, stp, initial):
        s = self.s
        if not initial:
            vcombine_(self.x, 1, self.S[mark], -stp, self.d)           # :599
        if s.method > 0:
            project_variables_(self.x, self.x, s.lo, s.hi)              # :386
        f = float(s.fg(self.x, self.g))                                 # :391
        if s.method > 0:
            project_direction_(self.p, self.x, s.lo, s.hi, "-", self.g)  # :396
        self._mark = mark
        return f

    def gnorm(self):
        return vnorm2(self.p if self.s.method > 0 else self.g)          # :399, :495

    def effective_step(self):
        vcombine_(self.sv, 1, self.x, -1, self.S[self._mark])           # :427

    def gd_search(self, stp):
        if self.s.method > 0:
            return vdot(self.g, self.sv) / stp                          # :432
        return -vdot(self.g, self.d)                                    # :434

    def xnorm(self):
        return vnorm2(self.x)                                           # :446

    def snorm(self):
        return vnorm2(self.sv)                                          # :448

    def nfree(self):
        return None

    def revert(self, mark, stp):
        vcombine_(self.x, 1, self.S[mark], -stp, self.d)                # :490
        if self.s.method > 0:
            project_variables_(self.x, self.x, self.s.lo, self.s.hi)    # :492

    def new_direction(self):
        s = self.s
        mark = s.mark
        vupdate_(self.S[mark], -1, self.x)                              # :512
        vupdate_(self.Y[mark], -1, self.p if s.method == 1 else self.g)  # :513
        if s.method < 2:
            self.rho[mark] = vdot(self.Y[mark], self.S[mark])           # :515
            if self.rho[mark] > 0:
                s.gamma = self.rho[mark] / vdot(self.Y[mark], self.Y[mark])  # :518
        s.m = min(s.m + 1, s.mem)                                       # :521
        if s.method < 2:
            vcopy_(self.d, self.g)                                      # :525
            change = apply_lbfgs_(self.S, self.Y, self.rho, s.gamma, s.m, mark, self.d, self.alpha)
        else:
            vcopy_(self.d, self.p)                                      # :528
            change = apply_lbfgs_masked_(self.S, self.Y, self.rho, s.m, mark, self.d, self.alpha,
                                         self.p)                        # sel = {p != 0} :530
        if not change:
            return False, 0.0, 0.0
        if s.method > 0:
            project_direction_(self.d, self.x, s.lo, s.hi, "-", self.d)  # :535
        gd = -vdot(self.g, self.d)                                      # :537
        dnorm = vnorm2(self.d)                                          # :629
        return True, gd, dnorm

    def steepest(self):
        vcopy_(self.d, self.p if self.s.method > 0 else self.g)         # :554

    def begin_search(self, mark):
        s = self.s
        vcopy_(self.S[mark], self.x)                                    # :562
        vcopy_(self.Y[mark], self.p if s.method == 1 else self.g)       # :563
        if s.method > 0:
            return step_limits(self.x, s.lo, s.hi, -1, self.d)          # :577
        return math.inf, math.inf

    def vector(self, which, slot=0):
        if which == "S":
            return self.S[slot]
        if which == "Y":
            return self.Y[slot]
        return {"x": self.x, "g": self.g, "p": self.p, "d": self.d}[which]


class _FusedOps:
    """Device back end made of the fused phases P1 / P3 / P4 (Tier 2)."""

    def __init__(self, solver, ws, x: Vector):
        self.s = solver
        self.ws = ws
        self.L = _lib.lib()
        self.x = x
        self.r = [0.0] * 8
        self._out8 = (f64 * 8)()
        self._out4 = (f64 * 4)()
        self._out3 = (f64 * 3)()
        self._out2 = (f64 * 2)()
        self._alpha = [0.0] * solver.mem
        self._limits = (math.inf, math.inf)
        self._spec8 = (f64 * 8)()
        self._spec = None        # results of a speculative first trial (stp = 1) fused into P4
        self._spec_mark = -1
        self._saved_for = -1
        self._need_steepest = False
        self.rho = [0.0] * solver.mem    # per slot, for L-BFGS / BLMVM (:515)
        self.g = self._vec("g")
        self.p = self._vec("p") if solver.method > 0 else None
        self.d = self._vec("d")
        self.builtin = is_builtin(solver.fg)

    def _vec(self, which, slot=0):
        h = self.L.opkb_vmlmb_vector(self.ws, ord(which), slot)
        return Vector.borrowed(self.s.ctx, h, self)

    def vector(self, which, slot=0):
        return self._vec(which, slot)

    def evaluate(self, mark, stp, initial):
        if self._spec is not None:
            spec, self._spec = self._spec, None
            if not initial and mark == self._spec_mark and stp == 1.0:
                self.r = spec    # the trial point was evaluated by the direction kernel
                return spec[0]
            # another first step: the trial below simply overwrites x, g, p
        if self.builtin:
            check(self.L.opkb_vmlmb_trial(self.ws, mark, stp, int(initial), self._out8))
        else:
            check(self.L.opkb_vmlmb_trial_begin(self.ws, mark, stp, int(initial)))
            f = float(self.s.fg(self.x, self.g))
            check(self.L.opkb_vmlmb_trial_end(self.ws, mark, f, int(initial), self._out8))
        self.r = list(self._out8)
        return self.r[0]

    def gnorm(self):
        return math.sqrt(self.r[1])

    def effective_step(self):
        pass

    def gd_search(self, stp):
        if self.s.method > 0:
            return self.r[2] / stp
        return -self.r[3]

    def xnorm(self):
        return math.sqrt(self.r[4])

    def snorm(self):
        return math.sqrt(self.r[5])

    def nfree(self):
        return int(self.r[6])

    def revert(self, mark, stp):
        check(self.L.opkb_vmlmb_revert(self.ws, mark, stp))

    def _solve(self, G, m, slots):
        """Two-loop recursion in coefficient space (apply_lbfgs!, :716-779) from
        the Gram block; pairs j = 0 (oldest) .. m-1 (newest)."""
        s = self.s
        W = m + 1
        sv = lambda a: G[2 * a * W]                       # s_a . v0
        yv = lambda a: G[(2 * a + 1) * W]                 # y_a . v0
        SY = lambda a, b: G[2 * a * W + 1 + b]            # s_a . y_b, b >= a
        YY = lambda a, b: (G[(2 * a + 1) * W + 1 + b] if b >= a else G[(2 * b + 1) * W + 1 + a])
        nan = math.nan
        if s.method == 2:
            rho = [SY(j, j) for j in range(m)]            # :758
        else:
            newest = m - 1
            self.rho[slots[newest]] = SY(newest, newest)  # :515
            if self.rho[slots[newest]] > 0:
                s.gamma = self.rho[slots[newest]] / YY(newest, newest)  # :518
            rho = [self.rho[slots[j]] for j in range(m)]
        alpha = [nan] * m
        c = [0.0] * m
        gamma = 0.0
        for k in range(m - 1, -1, -1):                    # :725-732 / :756-766
            if rho[k] > 0:
                acc = sv(k)
                for j in range(m - 1, k, -1):
                    if rho[j] > 0:
                        acc -= alpha[j] * SY(k, j)
                alpha[k] = acc / rho[k]
                if gamma == 0:
                    gamma = rho[k] / YY(k, k)             # :762-764
        modif = any(r > 0 for r in rho)
        if s.method < 2:
            gamma = s.gamma                               # :526 (H0 = gamma*I)
        if not modif or gamma == 0:
            return alpha, c, gamma, False
        for k in range(m):                                # :735-741 / :769-775
            if rho[k] > 0:
                acc = yv(k)
                for j in range(m - 1, -1, -1):
                    if rho[j] > 0:
                        acc -= alpha[j] * YY(k, j)
                acc *= gamma
                for j in range(k):
                    if rho[j] > 0:
                        acc += c[j] * SY(j, k)
                beta = acc / rho[k]
                c[k] = alpha[k] - beta
        return alpha, c, gamma, True

    def _new_direction_sequential(self):
        """Parity-grade variant: the reference's sequential two-loop recursion
        (apply_lbfgs!, :716-779) with the same roundings of the intermediate
        vector in the element type, one fused kernel per loop iteration: the
        pending vupdate! of the previous iteration, the vscale! between the
        loops and the dot products of the current pair share one pass."""
        s = self.s
        L, ws = self.L, self.ws
        mark, mem, method = s.mark, s.mem, s.method
        out2, out3 = self._out2, self._out3
        check(L.opkb_vmlmb_pair_update(ws, mark, out2))                  # :512-513
        rho, alpha = self.rho, self._alpha
        if method < 2:
            rho[mark] = out2[0]                                          # :515
            if rho[mark] > 0:
                s.gamma = rho[mark] / out2[1]                            # :518
        s.m = min(s.m + 1, mem)                                          # :521
        m = s.m
        S, Y = ord("S"), ord("Y")
        first = 1            # d not written yet: v is still p (g for L-BFGS / BLMVM), :525/:528
        pend = (0, 0, 0.0)   # pending vupdate!: (which, slot, coefficient)
        gamma = 0.0
        modif = False
        k = mark + 1
        for _ in range(m):                                               # :725-732 / :756-766
            k = k - 1 if k > 0 else mem - 1
            if method < 2 and not rho[k] > 0:
                continue
            check(L.opkb_vmlmb_twoloop_step(ws, first, pend[0], pend[1], pend[2], 0, 1.0, S, k,
                                            int(method == 2), out3))
            if pend[0]:
                first = 0
            pend = (0, 0, 0.0)
            if method == 2:
                rho[k] = out3[1]                                         # :758
            if rho[k] > 0:
                alpha[k] = out3[0] / rho[k]                              # :728 / :760
                pend = (Y, k, -alpha[k])                                 # :729 / :761
                modif = True
                if method == 2 and gamma == 0:
                    gamma = rho[k] / out3[2]                             # :762-764
        self._saved_for = -1
        if method < 2:
            gamma = s.gamma
            change = modif
        else:
            change = gamma != 0
        if not change:
            return False, 0.0, 0.0
        scale = 1                                                        # vscale!(v, gamma) :734 / :768
        for _ in range(m):                                               # :735-741 / :769-775
            if rho[k] > 0:
                check(L.opkb_vmlmb_twoloop_step(ws, first, pend[0], pend[1], pend[2], scale, gamma,
                                                Y, k, 0, out3))
                first, scale = 0, 0
                beta = out3[0] / rho[k]                                  # :737 / :771
                pend = (S, k, alpha[k] - beta)                           # :738 / :772
            k = k + 1 if k < mem - 1 else 0
        mark_next = mark + 1 if mark < mem - 1 else 0
        # last vupdate! + :535, :537, :629, :562-563, :577
        check(L.opkb_vmlmb_finish_direction(ws, mark_next, pend[0], pend[1], pend[2], self._out4))
        self._limits = (self._out4[2], self._out4[3])
        self._saved_for = mark_next
        self._need_steepest = False
        return True, -self._out4[0], math.sqrt(self._out4[1])

    def new_direction(self):
        s = self.s
        if s.twoloop == "sequential":
            return self._new_direction_sequential()
        mark = s.mark
        m = min(s.m + 1, s.mem)                           # :521
        slots = [(mark - (m - 1) + j) % s.mem for j in range(m)]
        cslots = (C.c_int * m)(*slots)
        G = (f64 * (2 * m * (m + 1)))()
        check(self.L.opkb_vmlmb_gram(self.ws, mark, m, cslots, G))
        s.m = m
        alpha, c, gamma, change = self._solve(G, m, slots)
        self._saved_for = -1
        if not change:
            return False, 0.0, 0.0
        mark_next = mark + 1 if mark < s.mem - 1 else 0
        if s.speculate and self.builtin and s.method != 1:
            # P4 + P1: the first trial of the next search (stp = 1, :567) in the same pass
            check(self.L.opkb_vmlmb_direction_trial(self.ws, m, cslots, (f64 * m)(*alpha),
                                                    (f64 * m)(*c), gamma, mark_next, self._out4,
                                                    self._spec8))
            self._spec = list(self._spec8) if self._spec8[7] == 1.0 else None
            self._spec_mark = mark_next
        else:
            check(self.L.opkb_vmlmb_direction(self.ws, m, cslots, (f64 * m)(*alpha), (f64 * m)(*c),
                                              gamma, mark_next, self._out4))
        gd = -self._out4[0]
        dnorm = math.sqrt(self._out4[1])
        self._limits = (self._out4[2], self._out4[3])
        self._saved_for = mark_next
        self._need_steepest = False
        return True, gd, dnorm

    def steepest(self):
        if self._spec is not None:   # the direction was rejected: drop its speculative trial
            check(self.L.opkb_vmlmb_discard_trial(self.ws, self._spec_mark))
            self._spec = None
        self._need_steepest = True

    def begin_search(self, mark):
        if self._need_steepest or self._saved_for != mark:
            check(self.L.opkb_vmlmb_steepest(self.ws, mark, self._out4))
            self._limits = (self._out4[2], self._out4[3])
            self._need_steepest = False
        self._saved_for = -1
        return self._limits


# ---------------------------------------------------------------------------
class VMLMB:
    """Steppable `_vmlmb!` (src/quasinewton.jl:296-612).

    `step()` performs one trip of the reference's `while true` loop (one
    evaluation of the objective); `iterate(k)` runs until `iter` has advanced by
    k; `solve()` runs to termination.  After termination `x` holds the best
    point (as `vmlmb!` returns it)."""

    def __init__(self, fg, x0, *, mem=5, lower=-math.inf, upper=math.inf, autodiff=False,
                 blmvm=False, fmin=-math.inf, maxiter=None, maxeval=None, xtol=(0.0, 1e-7),
                 ftol=(0.0, 1e-8), gtol=(0.0, 1e-6), epsilon=0.0, verb=0, printer=print_iteration,
                 output=None, lnsrch=None, ctx: Context = None, mode="fused", twoloop=None,
                 speculate=True, observer=None, nglobal=None):
        if autodiff:
            raise NotImplementedError("autodiff is a host-side hook of the reference "
                                      "(src/autodiff.jl); pass fg(x, g) -> f")
        if isinstance(lower, BoundedSet):
            lower, upper = lower.lower, lower.upper
        self.fg = fg
        # `mode` may also be a factory `ops = mode(solver, x0)` of a device back
        # end: the seam the CPU tests of this host logic use (tests/cpu_backend.py)
        external = callable(mode)
        self.ctx = None if external else (ctx or (x0.ctx if isinstance(x0, Vector) else default_context()))
        self._numpy_in = not isinstance(x0, Vector)
        if self._numpy_in:
            x0 = np.ascontiguousarray(x0)
            if x0.dtype not in (np.float32, np.float64):
                x0 = x0.astype(np.float64)
            self._shape = x0.shape
            x0 = x0.reshape(-1)
        n = len(x0)
        dtype = x0.dtype
        self.n = n
        self.dtype = np.dtype(dtype)

        # vmlmb! :244-293 -------------------------------------------------------
        bounds = 0
        self._keep = []
        lo, hi = lower, upper
        if isinstance(lo, (int, float, np.floating, np.integer)):
            lo = float(lo)
            if lo > -math.inf:
                bounds |= 1
        else:
            lo = lo if external else self._as_vector(lo, n, dtype, "lower")
            bounds |= 1
        if isinstance(hi, (int, float, np.floating, np.integer)):
            hi = float(hi)
            if hi < math.inf:
                bounds |= 2
        else:
            hi = hi if external else self._as_vector(hi, n, dtype, "upper")
            bounds |= 2
        self.lo, self.hi = lo, hi
        self.method = 0 if bounds == 0 else (1 if blmvm else 2)            # :273
        if lnsrch is None:                                                 # :276-284
            if self.method == 0:
                lnsrch = MoreThuenteLineSearch(ftol=1e-3, gtol=0.9, xtol=0.1)
            else:
                lnsrch = MoreToraldoLineSearch(ftol=1e-3, gamma=(0.1, 0.5))
        if not isinstance(lnsrch, LineSearch):
            raise TypeError("lnsrch must be a LineSearch")
        self.lnsrch = lnsrch

        # _vmlmb! :310-379 --------------------------------------------------------
        big = 2 ** 63 - 1
        self.maxiter = big if maxiter is None else int(maxiter)
        self.maxeval = big if maxeval is None else int(maxeval)
        self.xatol, self.xrtol = float(xtol[0]), float(xtol[1])
        self.fatol, self.frtol = float(ftol[0]), float(ftol[1])
        self.gatol, self.grtol = float(gtol[0]), float(gtol[1])
        self.epsilon = float(epsilon)
        self.fmin = float(fmin)
        mem = int(mem)
        assert mem >= 1 and self.maxiter >= 0 and self.maxeval >= 1
        assert self.xatol >= 0 and self.xrtol >= 0 and self.fatol >= 0 and self.frtol >= 0
        assert self.gatol >= 0 and self.grtol >= 0 and 0 <= self.epsilon < 1
        self.mem = min(mem, int(nglobal) if nglobal is not None else n)   # :328
        self.mem = max(self.mem, 1)
        self.fminset = (not math.isnan(self.fmin)) and self.fmin > -math.inf  # :330
        self.verb = int(verb)
        self.printer = printer
        self.output = output if output is not None else sys.stdout
        self.observer = observer
        self.reason = ""
        self.mark = 0
        self.m = 0
        self.iter = 0
        self.eval = 0
        self.rejects = 0
        self.f = self.f0 = self.gd = self.gd0 = self.stp = 0.0
        self.stpmin = self.stpmax = self.smin = self.smax = 0.0
        self.gamma = 1.0
        self.gnorm = self.gtest = 0.0
        self.beststp = self.bestf = self.bestgnorm = 0.0
        self.stage = 0
        self.finished = False

        # The fused path evaluates the two-loop recursion either in coefficient
        # space from one Gram sweep ("gram": 4m+3 passes) or as the reference's
        # sequence of dots and axpys ("sequential": parity-grade).  In Float32
        # the roundings of the intermediate vector are part of the result, and
        # the chaotic Rosenbrock trajectory amplifies a 1-ulp change beyond the
        # 1e-4 gate within 20 iterations, so "sequential" is the Float32 default.
        # The reference accepts any `mem` (src/quasinewton.jl:227, :328): the Gram-form kernels
        # hold MEM_MAX pairs, beyond that the step-by-step recursion (no such limit) runs.
        if twoloop is None:
            twoloop = "gram" if (self.dtype == np.float64 and self.mem <= _lib.MEM_MAX) else "sequential"
        if twoloop not in ("gram", "sequential"):
            raise ValueError("twoloop must be 'gram' or 'sequential'")
        if twoloop == "gram" and self.mem > _lib.MEM_MAX and mode == "fused":
            raise ValueError(f"twoloop='gram' supports mem <= {_lib.MEM_MAX}; the default "
                             "(twoloop='sequential' beyond) has no limit")
        self.twoloop = twoloop
        # With a built-in objective the fused path evaluates the first trial point
        # of the next line search (stp = 1) inside the direction kernel; the result
        # is used iff the driver indeed takes that step (same numbers either way).
        self.speculate = bool(speculate)

        # device storage (:358-371) ----------------------------------------------
        self.mode = mode
        self._ws = None
        if external:
            self.ops = mode(self, x0)
            self.x = self.ops.x
            return
        L = _lib.lib()
        if mode == "fused":
            if self.mem > _lib.SLOTS_MAX:
                raise ValueError(f"mode='fused' stores at most {_lib.SLOTS_MAX} pairs; use mode='primitive'")
            ws = vp()
            lo_h = lo.handle if isinstance(lo, Vector) else None
            hi_h = hi.handle if isinstance(hi, Vector) else None
            check(L.opkb_vmlmb_create(self.ctx.handle, _eltype(dtype), n, self.mem, self.method,
                                      lo_h, lo if lo_h is None else -math.inf,
                                      hi_h, hi if hi_h is None else math.inf, C.byref(ws)))
            self._ws = ws
            self.x = Vector.borrowed(self.ctx, L.opkb_vmlmb_vector(ws, ord("x"), 0), self)
            if isinstance(x0, Vector):
                vcopy_(self.x, x0)
            else:
                self.x.upload(x0)
            if is_builtin(fg):
                a, c = fg.vectors(self.ctx, n, dtype)
                check(L.opkb_vmlmb_set_objective(ws, fg.kind, a.handle if a is not None else None,
                                                 c.handle if c is not None else None))
            else:
                check(L.opkb_vmlmb_set_objective(ws, _lib.OBJ_USER, None, None))
            self.ops = _FusedOps(self, ws, self.x)
        elif mode == "primitive":
            if isinstance(x0, Vector):
                self.x = x0                       # vmlmb! works in place
            else:
                self.x = self.ctx.from_numpy(x0)
            if is_builtin(fg):
                self.fg = _builtin_as_callable(fg, self.ctx, n, dtype)
            self.ops = _PrimitiveOps(self, self.x)
        else:
            raise ValueError("mode must be 'fused' or 'primitive'")

    def _as_vector(self, b, n, dtype, name):
        if isinstance(b, Vector):
            if b.n != n or b.dtype != dtype:
                raise TypeError(f"invalid {name} bound type")
            return b
        arr = np.ascontiguousarray(b, dtype=dtype).reshape(-1)
        if arr.size != n:
            raise TypeError(f"invalid {name} bound type")
        v = self.ctx.from_numpy(arr)
        self._keep.append(v)
        return v

    def close(self):
        if self._ws:
            # never call into the library for a workspace whose context is gone (e.g. a Group that
            # was closed first): its device memory went with the context
            if self.ctx is not None and self.ctx.handle:
                _lib.lib().opkb_vmlmb_destroy(self._ws)
            self._ws = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -----------------------------------------------------------------------
    def step(self) -> bool:
        """One trip of `while true` (:381-601).  Returns False once finished."""
        if self.finished:
            return False
        o = self.ops
        ls = self.lnsrch
        # :385-397 (and :599 of the previous trip, fused into the evaluation)
        self.f = o.evaluate(self.mark, self.stp, self.eval == 0)
        self.eval += 1
        if self.eval == 1 or self.f < self.bestf:                          # :398-419
            self.gnorm = o.gnorm()
            self.beststp = self.stp
            self.bestf = self.f
            self.bestgnorm = self.gnorm
            if self.eval == 1:
                self.gtest = _jl_max(self.gatol, self.grtol * self.gnorm)
            if self.gnorm <= self.gtest:
                self.stage = 3
                if self.gnorm == 0:
                    self.reason = "a stationary point has been found"
                elif self.method > 0:
                    self.reason = "projected gradient norm sufficiently small"
                else:
                    self.reason = "gradient norm sufficiently small"
                if self.eval > 1:
                    self.iter += 1
        if self.fminset and self.f < self.fmin:                            # :420-423
            self.stage = 4
            self.reason = "f < fmin"

        if self.stage == 1:                                                # :425-474
            o.effective_step()
            if ls.usederivatives:
                self.gd = o.gd_search(self.stp)
            task = ls.iterate(self.stp, self.f, self.gd)
            if task == "SEARCH":
                self.stp = ls.getstep()
            elif task == "CONVERGENCE":
                self.iter += 1
                xtest = _jl_max(self.xatol, 0.0)
                if self.xrtol > 0:
                    xtest = _jl_max(xtest, self.xrtol * o.xnorm())
                if o.snorm() <= xtest:
                    self.stage = 3
                    self.reason = "X test satisfied"
                else:
                    delta = _jl_max(abs(self.f - self.f0), self.stp * abs(self.gd0))
                    if delta <= self.fatol:
                        self.stage = 3
                        self.reason = "fatol test satisfied"
                    elif delta <= self.frtol * abs(self.f0):
                        self.stage = 3
                        self.reason = "frtol test satisfied"
                    elif self.iter >= self.maxiter:
                        self.stage = 4
                        self.reason = "too many iterations"
                    elif self.eval >= self.maxeval:
                        self.stage = 4
                        self.reason = "too many evaluations"
                    else:
                        self.stage = 2
            else:
                self.stage = 4
                self.reason = ls.getreason()
        if self.stage < 2 and self.eval >= self.maxeval:                   # :475-478
            self.stage = 4
            self.reason = "too many evaluations"

        if self.stage != 1:                                                # :480-596
            if self.stp != self.beststp:                                   # :485-497
                if self.stage >= 3:
                    self.stp = self.beststp
                    self.f = self.bestf
                    self.gnorm = self.bestgnorm
                    o.revert(self.mark, self.stp)
                else:
                    self.gnorm = o.gnorm()
            if verbose(self.verb, self.iter):                              # :501-503
                self.printer(self.output, self.iter, self.eval, self.rejects, self.f, self.gnorm,
                             self.stp)
            if self.observer is not None:
                self.observer(self)
            if self.stage >= 3:                                            # :504-506
                self._finish()
                return False

            if self.stage == 2:                                            # :509-550
                change, gd, dnorm = o.new_direction()
                if change:
                    self.gd = gd
                    reject = not sufficient_descent(self.gd, self.epsilon, self.gnorm, dnorm)
                else:
                    reject = True
                if reject:
                    self.stage = 0
                    self.rejects += 1
                self.mark = self.mark + 1 if self.mark < self.mem - 1 else 0  # :549
            if self.stage == 0:                                            # :551-556
                o.steepest()
                self.gd = -self.gnorm ** 2

            self.f0 = self.f                                               # :560-563
            self.gd0 = self.gd
            self.smin, self.smax = o.begin_search(self.mark)

            if self.stage == 2:                                            # :566-574
                self.stp = 1.0
            elif self.fminset and self.fmin < self.f0:
                self.stp = 2 * (self.fmin - self.f0) / self.gd0
            else:
                self.stp = 1 / self.gnorm
            STPMIN, STPMAX = 1e-20, 1e+20                                  # :321-322
            if self.method > 0:                                            # :575-584
                self.stp = _jl_min(self.stp, self.smax)
                self.stpmin = STPMIN * self.stp
                self.stpmax = _jl_min(STPMAX * self.stp, self.smax)
            else:
                self.stpmin = STPMIN * self.stp
                self.stpmax = STPMAX * self.stp
            # :587-594 (start! throws ArgumentError -> ValueError here)
            task = ls.start(self.f0, self.gd0, self.stp, stpmin=self.stpmin, stpmax=self.stpmax)
            if task != "SEARCH":
                self.stage = 4
                self.reason = ls.getreason()
                self._finish()
                return False
            self.stage = 1
        return True

    def _finish(self):
        self.finished = True
        if verbose(self.verb, 0):                                          # :604-610
            prefix = "WARNING: " if self.stage > 3 else "CONVERGENCE: "
            self.output.write("# " + prefix + self.reason + "\n")

    def iterate(self, niter: int = 1) -> bool:
        """Run until `iter` has advanced by `niter` (or termination)."""
        target = self.iter + niter
        while self.iter < target:
            if not self.step():
                return False
        return True

    def solve(self):
        while self.step():
            pass
        return self.x

    def current(self):
        """The current iterate (the last accepted point) as a device Vector, valid
        between calls of `step()` / `iterate()`: during a line search `x` holds the
        trial point and slot `mark` of S holds the iterate (:562)."""
        if self.stage == 1 and not self.finished:
            return self.ops.vector("S", self.mark)
        return self.x

    def result(self):
        """The solution as the caller's kind of array."""
        if self._numpy_in:
            x = self.x.download() if isinstance(self.x, Vector) else np.array(self.x)
            return x.reshape(self._shape)
        return self.x

    def carry_stats(self):
        """(hits, misses) of the incremental Gram of the fused path (opkb200.h,
        opkb_vmlmb_carry_stats): Gram blocks assembled without a sweep / sweeps that
        ran although a carry was armed."""
        if not self._ws:
            return (0, 0)
        h, m = _lib.i64(), _lib.i64()
        check(_lib.lib().opkb_vmlmb_carry_stats(self._ws, C.byref(h), C.byref(m)))
        return (h.value, m.value)

    def chain_stats(self):
        """(heads, adopted, cancelled) of the device-resident decisions of the native driver
        (opkb200.h, opkb_vmlmb_chain_stats): fused launches that started a chain, launches adopted
        while already in flight, launches enqueued ahead that the device cancelled."""
        if not self._ws:
            return (0, 0, 0)
        a, b, c = _lib.i64(), _lib.i64(), _lib.i64()
        check(_lib.lib().opkb_vmlmb_chain_stats(self._ws, C.byref(a), C.byref(b), C.byref(c)))
        return (a.value, b.value, c.value)

    @property
    def history(self) -> int:
        """1: the workspace keeps the iterate history (pairs formed on the fly, nothing written
        back); 0: the reference's stored differences; -1: not decided yet (opkb_vmlmb_history)."""
        return _lib.lib().opkb_vmlmb_history(self._ws) if self._ws else 0

    def last_gram(self):
        """(slots, G) of the last Gram block (2m x (m+1), row-major) or None."""
        if not self._ws:
            return None
        m = C.c_int()
        slots = (C.c_int * _lib.MEM_MAX)()
        G = (f64 * (2 * _lib.MEM_MAX * (_lib.MEM_MAX + 1)))()
        check(_lib.lib().opkb_vmlmb_last_gram(self._ws, C.byref(m), slots, G))
        if m.value == 0:
            return None
        k = m.value
        return list(slots[:k]), np.array(G[:2 * k * (k + 1)]).reshape(2 * k, k + 1)

    def free_mask(self) -> np.ndarray:
        """Free-variable mask {i : p[i] != 0} of this rank (all True for L-BFGS)."""
        if self.method == 0:
            return np.ones(self.n, dtype=bool)
        p = self.ops.vector("p")
        return (p.download() if isinstance(p, Vector) else np.asarray(p)) != 0


class NativeVMLMB(VMLMB):
    """`vmlmb!` with the driver loop inside libopkb200 (`opkb_solver_*`, the
    reverse-communication API in the style of the reference's own C FFI,
    src/wrappers.jl:1831-1859): same keywords and results as `VMLMB`, one native
    call per objective evaluation (a user `fg`) or per run (built-in objective)."""

    def __init__(self, fg, x0, **kw):
        from .linesearches import ArmijoLineSearch
        ls = kw.get("lnsrch")
        kw.pop("mode", None)
        # chain=True: device-resident decisions of the run loop (opkb_solver_set_chain; same results,
        # opt-in, see DESIGN.md "Chained launches"); None: the library's default (OPKB_CHAIN)
        chain = kw.pop("chain", None)
        super().__init__(fg, x0, mode="fused", **kw)
        L = _lib.lib()
        opt = _lib.VmlmbOptions()
        L.opkb_vmlmb_default_options(C.byref(opt))
        opt.fmin = self.fmin
        opt.maxiter, opt.maxeval = self.maxiter, self.maxeval
        opt.xatol, opt.xrtol = self.xatol, self.xrtol
        opt.fatol, opt.frtol = self.fatol, self.frtol
        opt.gatol, opt.grtol = self.gatol, self.grtol
        opt.epsilon = self.epsilon
        opt.twoloop = 0 if self.twoloop == "gram" else 1
        opt.speculate = int(self.speculate)
        ls = self.lnsrch if ls is None else ls
        if isinstance(ls, MoreThuenteLineSearch):
            opt.lnsrch, opt.ls_ftol, opt.ls_gtol, opt.ls_xtol = 3, ls.ftol, ls.gtol, ls.xtol
        elif isinstance(ls, MoreToraldoLineSearch):
            opt.lnsrch, opt.ls_ftol, opt.ls_strict = 2, ls.ftol, int(ls.strict)
            opt.ls_gamma1, opt.ls_gamma2 = ls.gamma1, ls.gamma2
        elif isinstance(ls, ArmijoLineSearch):
            opt.lnsrch, opt.ls_ftol = 1, ls.ftol
        else:
            raise TypeError("the native driver knows the three line searches of the reference")
        self._solver = vp()
        check(L.opkb_solver_create(self._ws, C.byref(opt), C.byref(self._solver)))
        if chain is not None:
            check(L.opkb_solver_set_chain(self._solver, int(bool(chain))))
        self._task = C.c_int(_lib.TASK_START)
        self._g = self.ops.vector("g")
        self._builtin = is_builtin(fg)
        from .objectives import DeviceObjective
        self._callback = isinstance(fg, DeviceObjective)
        if self._callback:   # the reverse-communication loop runs inside the library
            check(L.opkb_solver_set_fg_callback(self._solver, C.cast(fg._cb, vp), fg.user, int(fg.partial)))

    def close(self):
        if getattr(self, "_solver", None):
            _lib.lib().opkb_solver_destroy(self._solver)
            self._solver = None
        super().close()

    @property
    def chain_enabled(self) -> bool:
        """Whether the run loop uses the device-resident decisions (opkb_solver_chain_enabled)."""
        return bool(self._solver) and bool(_lib.lib().opkb_solver_chain_enabled(self._solver))

    def _refresh(self):
        L = _lib.lib()
        it, ev, rej = _lib.i64(), _lib.i64(), _lib.i64()
        f, gn, stp = f64(), f64(), f64()
        stage, newx = C.c_int(), C.c_int()
        check(L.opkb_solver_status(self._solver, C.byref(it), C.byref(ev), C.byref(rej), C.byref(f),
                                   C.byref(gn), C.byref(stp), C.byref(stage), C.byref(newx)))
        self.iter, self.eval, self.rejects = it.value, ev.value, rej.value
        self.f, self.gnorm, self.stp, self.stage = f.value, gn.value, stp.value, stage.value
        self.reason = L.opkb_solver_reason(self._solver).decode()
        self.mark = L.opkb_solver_mark(self._solver)
        self.finished = self._task.value in (_lib.TASK_FINAL_X, _lib.TASK_WARNING)
        return bool(newx.value)

    def step(self):
        raise NotImplementedError("the native driver advances by iterations: use iterate()")

    def iterate(self, niter: int = 1) -> bool:
        L = _lib.lib()
        if self.finished:
            return False
        if self._builtin or self._callback:
            try:
                check(L.opkb_solver_run(self._solver, int(min(niter, 2 ** 62)), C.byref(self._task)))
            except _lib.OpkbError as e:
                if self._callback and self.fg.error is not None:
                    raise self.fg.error from e
                if "line search" in str(e):
                    raise ValueError(str(e)) from e
                raise
            if self._refresh() and self.observer is not None:
                self.observer(self)
            return not self.finished
        target = self.iter + niter
        if self._task.value == _lib.TASK_START:
            check(L.opkb_solver_start(self._solver, C.byref(self._task)))
        while self._task.value == _lib.TASK_COMPUTE_FG and self.iter < target:
            f = float(self.fg(self.x, self._g))
            try:
                check(L.opkb_solver_iterate(self._solver, f, C.byref(self._task)))
            except _lib.OpkbError as e:
                if "line search" in str(e):   # ArgumentError thrown by start!/cstep in the reference
                    raise ValueError(str(e)) from e
                raise
            if self._refresh() and self.observer is not None:
                self.observer(self)
        return not self.finished

    def solve(self):
        while self.iterate(2 ** 62):
            pass
        if verbose(self.verb, 0):
            self._finish()
        return self.x


def _builtin_as_callable(obj, ctx, n, dtype):
    """Stand-alone `fg!(x, g)` of a built-in objective (Tier 1), for mode='primitive'."""
    L = _lib.lib()
    a, c = obj.vectors(ctx, n, dtype)
    r = f64()

    if obj.kind == _lib.OBJ_ROSENBROCK:
        def fg(x, g):
            check(L.opkb_rosenbrock_fg(ctx.handle, x.handle, g.handle, C.byref(r)))
            return r.value
    else:
        def fg(x, g):
            check(L.opkb_quadratic_fg(ctx.handle, a.handle, c.handle, x.handle, g.handle, C.byref(r)))
            return r.value
    fg._keep = (a, c)
    return fg


def vmlmb_(fg, x, **kwds):
    """`vmlmb!(fg!, x; kwds...) -> x` (src/quasinewton.jl:226): the best solution
    is stored in `x` (a device Vector or a writeable host array) and returned."""
    driver = kwds.pop("driver", "python")
    if driver not in ("python", "native"):
        raise ValueError("driver must be 'python' or 'native'")
    ngpus, group = kwds.pop("ngpus", None), kwds.pop("group", None)
    if ngpus is not None or group is not None:
        # ONE process, several GPUs: the native driver of every rank on its own host thread
        from .multi import MultiVMLMB
        if isinstance(x, Vector):
            raise TypeError("ngpus= shards a HOST array over the GPUs")
        ms = MultiVMLMB(fg, x, group=group, ngpus=ngpus or 0, **kwds)
        x[...] = ms.solve().astype(x.dtype, copy=False)
        ms.close()
        return x
    solver = (NativeVMLMB if driver == "native" else VMLMB)(fg, x, **kwds)
    solver.solve()
    if isinstance(x, Vector):
        if solver.x is not x:
            vcopy_(x, solver.x)
    else:
        x[...] = solver.result().astype(x.dtype, copy=False)
    solver.close()
    return x


def vmlmb(fg, x0, **kwds):
    """`vmlmb(fg!, x0; kwds...) -> x` (src/quasinewton.jl:215): x0 is left unchanged."""
    if isinstance(x0, Vector):
        x = vcopy_(x0.similar(), x0)
    else:
        x = np.array(x0, copy=True)
        if x.dtype not in (np.float32, np.float64):
            x = x.astype(np.float64)
    return vmlmb_(fg, x, **kwds)
