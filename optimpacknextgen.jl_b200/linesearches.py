"""Line searches of the reference (src/linesearches.jl), host side.

Scalar Float64 state machines: they stay on the host (north_star) and are
driven by the scalars the fused kernels return.  Same names, arguments and
error behaviour as the reference: `start!` -> `start`, `iterate!` -> `iterate`,
`getstep`, `gettask`, `getstatus`, `getreason`, `usederivatives`; tasks and
statuses are the reference's symbols as strings; Julia's `ArgumentError` is
`ValueError`, a failed `@assert` is `AssertionError`.
"""
from __future__ import annotations

import math

# src/linesearches.jl:33-51
REASONS = {
    "NOT_STARTED": "Algorithm not yet started",
    "F_LE_FMIN": "F(X) ≤ FMIN",
    "COMPUTE_FG": "Caller must compute f(x) and g(x)",
    "NEW_X": "A new iterate is available for examination",
    "FINAL_X": "A solution has been found within tolerances",
    "FATOL_TEST_SATISFIED": "FATOL test satisfied",
    "FRTOL_TEST_SATISFIED": "FRTOL test satisfied",
    "GATOL_TEST_SATISFIED": "GATOL test satisfied",
    "GRTOL_TEST_SATISFIED": "GRTOL test satisfied",
    "FIRST_WOLFE_SATISFIED": "First Wolfe condition satisfied",
    "STRONG_WOLFE_SATISFIED": "Strong Wolfe conditions both satisfied",
    "STP_EQ_STPMIN": "Step at lower bound",
    "STP_EQ_STPMAX": "Step at upper bound",
    "XTOL_TEST_SATISFIED": "XTOL test satisfied",
    "ROUNDING_ERRORS": "Rounding errors prevent progress",
    "NOT_DESCENT": "Search direction is not a descent direction",
    "INIT_STP_LE_ZERO": "Initial step ≤ 0",
}

XTRAPL = 1.1   # src/linesearches.jl:456
XTRAPU = 4.0   # src/linesearches.jl:457
STPMAX = 1e20  # src/linesearches.jl:458


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


def _clamp(x, lo, hi):
    return hi if x > hi else (lo if x < lo else x)


def _div(a, b):
    """IEEE division (Julia semantics): x/0 is +-Inf or NaN, never an exception."""
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0:
            return math.nan
        neg = (a < 0) != (math.copysign(1.0, b) < 0)
        return -math.inf if neg else math.inf


class LineSearch:
    """abstract type LineSearch{T} (src/linesearches.jl:96); T is Float64."""

    usederivatives = False

    def __init__(self):
        self.task = "START"
        self.status = "NOT_STARTED"
        self.stp = 0.0

    # src/linesearches.jl:104-133
    def getstep(self):
        return self.stp

    def gettask(self):
        return self.task

    def getstatus(self):
        return self.status

    def getreason(self):
        # the reference reports :FIRST_WOLFE_CONDITION_SATISFIED (:239, :364),
        # a key missing from REASONS; the text of :FIRST_WOLFE_SATISFIED is used
        key = "FIRST_WOLFE_SATISFIED" if self.status == "FIRST_WOLFE_CONDITION_SATISFIED" else self.status
        return REASONS[key]

    def _report(self, task, status):
        self.status = status
        self.task = task
        return task

    def start(self, f0, g0, stp, stpmin=0.0, stpmax=math.inf):
        raise NotImplementedError

    def iterate(self, stp, f, g):
        raise NotImplementedError


def _start_backtracking(ls, f0, g0, stp, stpmin, stpmax):
    # src/linesearches.jl:331-358
    f0, g0, stp, stpmin, stpmax = map(float, (f0, g0, stp, stpmin, stpmax))
    if stpmax < stpmin:
        raise ValueError("STPMAX < STPMIN")
    if stpmin < 0:
        raise ValueError("STPMIN < 0")
    if stp > stpmax:
        raise ValueError("STP > STPMAX")
    if stp < stpmin:
        raise ValueError("STP < STPMIN")
    if g0 >= 0:
        raise ValueError("Not a descent direction")
    ls.finit = f0
    ls.ginit = g0
    ls.stp = stp
    ls.stpmin = stpmin
    return ls._report("SEARCH", "COMPUTE_FG")


class ArmijoLineSearch(LineSearch):
    """src/linesearches.jl:185-247."""

    usederivatives = False

    def __init__(self, ftol=1e-4):
        super().__init__()
        assert 0 < ftol <= 0.5
        self.ftol = float(ftol)
        self.finit = self.ginit = self.stpmin = 0.0

    def start(self, f0, g0, stp, stpmin=0.0, stpmax=math.inf):
        return _start_backtracking(self, f0, g0, stp, stpmin, stpmax)

    def iterate(self, stp, f, g=0.0):
        assert self.task == "SEARCH"
        assert stp == self.stp
        if f <= self.finit + self.ftol * stp * self.ginit:
            return self._report("CONVERGENCE", "FIRST_WOLFE_CONDITION_SATISFIED")
        elif stp > self.stpmin:
            self.stp = _jl_max(stp / 2, self.stpmin)
            return self._report("SEARCH", "COMPUTE_FG")
        else:
            self.stp = self.stpmin
            return self._report("WARNING", "STP_EQ_STPMIN")


class MoreToraldoLineSearch(LineSearch):
    """src/linesearches.jl:251-381 (note: `strict` defaults to false, :318)."""

    usederivatives = False

    def __init__(self, ftol=1e-4, strict=False, gamma=(0.1, 0.5)):
        super().__init__()
        assert 0 < ftol <= 0.5
        assert 0 < gamma[0] < gamma[1] < 1
        self.gamma1, self.gamma2 = float(gamma[0]), float(gamma[1])
        self.ftol = float(ftol)
        self.strict = bool(strict)
        self.finit = self.ginit = self.stpmin = 0.0

    def start(self, f0, g0, stp, stpmin=0.0, stpmax=math.inf):
        return _start_backtracking(self, f0, g0, stp, stpmin, stpmax)

    def iterate(self, stp, f, g=0.0):
        assert self.task == "SEARCH"
        assert stp == self.stp
        if f <= self.finit + self.ftol * stp * self.ginit:
            return self._report("CONVERGENCE", "FIRST_WOLFE_CONDITION_SATISFIED")
        elif stp > self.stpmin:
            q = -self.ginit * stp
            r = f - (self.finit - q)
            if (r > 0 and q > 0) if self.strict else (r != 0):
                stp *= _clamp(_div(q, r + r), self.gamma1, self.gamma2)
            else:
                stp /= 2
            self.stp = _jl_max(stp, self.stpmin)
            return self._report("SEARCH", "COMPUTE_FG")
        else:
            self.stp = self.stpmin
            return self._report("WARNING", "STP_EQ_STPMIN")


class MoreThuenteLineSearch(LineSearch):
    """src/linesearches.jl:413-647."""

    usederivatives = True

    def __init__(self, ftol=0.001, gtol=0.9, xtol=0.1):
        super().__init__()
        assert 0 < ftol < 1 and 0 < gtol < 1 and 0 < xtol < 1
        self.ftol, self.gtol, self.xtol = float(ftol), float(gtol), float(xtol)
        self.stpmin = self.stpmax = self.finit = self.ginit = 0.0
        self.stx = self.fx = self.gx = self.sty = self.fy = self.gy = 0.0
        self.lower = self.upper = self.width = self.oldwidth = 0.0
        self.stage = 0
        self.brackt = False

    def start(self, f0, g0, stp, stpmin=0.0, stpmax=None, ftol=None, gtol=None, xtol=None):
        # src/linesearches.jl:460-529
        f0, g0, stp, stpmin = map(float, (f0, g0, stp, stpmin))
        stpmax = stp * STPMAX if stpmax is None else float(stpmax)
        ftol = self.ftol if ftol is None else float(ftol)
        gtol = self.gtol if gtol is None else float(gtol)
        xtol = self.xtol if xtol is None else float(xtol)
        if stpmax < stpmin:
            raise ValueError("STPMAX < STPMIN")
        if stpmin < 0:
            raise ValueError("STPMIN < 0")
        if xtol < 0:
            raise ValueError("XTOL < 0")
        if ftol <= 0:
            raise ValueError("FTOL ≤ 0")
        if gtol <= 0:
            raise ValueError("GTOL ≤ 0")
        if g0 >= 0:
            raise ValueError("Not a descent direction")
        if stp > stpmax:
            raise ValueError("STP > STPMAX")
        if stp < stpmin:
            raise ValueError("STP < STPMIN")
        self.ftol, self.gtol, self.xtol = ftol, gtol, xtol
        self.stpmin, self.stpmax = stpmin, stpmax
        self.finit, self.ginit = f0, g0
        self.stp = stp
        self.stx, self.fx, self.gx = 0.0, f0, g0
        self.sty, self.fy, self.gy = 0.0, f0, g0
        self.lower = 0.0
        self.upper = stp + stp * XTRAPU
        self.width = stpmax - stpmin
        self.oldwidth = 2 * (stpmax - stpmin)
        self.stage = 1
        self.brackt = False
        return self._report("SEARCH", "COMPUTE_FG")

    def iterate(self, stp, f, g):
        # src/linesearches.jl:536-647
        stp, f, g = float(stp), float(f), float(g)
        assert self.task == "SEARCH"
        assert stp == self.stp
        gtest = self.ftol * self.ginit
        ftest = self.finit + stp * gtest
        if self.stage == 1 and f <= ftest and g >= 0:
            self.stage = 2
        if f <= ftest and abs(g) <= -self.gtol * self.ginit:
            return self._report("CONVERGENCE", "STRONG_WOLFE_SATISFIED")
        if stp == self.stpmin and (f > ftest or g >= gtest):
            return self._report("WARNING", "STP_EQ_STPMIN")
        if stp == self.stpmax and f <= ftest and g <= gtest:
            return self._report("WARNING", "STP_EQ_STPMAX")
        if self.brackt:
            if self.upper - self.lower <= self.xtol * self.upper:
                return self._report("WARNING", "XTOL_TEST_SATISFIED")
            if stp <= self.lower or stp >= self.upper:
                return self._report("WARNING", "ROUNDING_ERRORS")

        if self.stage == 1 and f <= self.fx and f > ftest:
            # modified function (:577-593)
            (info, self.brackt, self.stx, fxm, gxm, self.sty, fym, gym, stp) = cstep(
                self.brackt, self.lower, self.upper,
                self.stx, self.fx - self.stx * gtest, self.gx - gtest,
                self.sty, self.fy - self.sty * gtest, self.gy - gtest,
                stp, f - stp * gtest, g - gtest)
            self.fx = fxm + self.stx * gtest
            self.fy = fym + self.sty * gtest
            self.gx = gxm + gtest
            self.gy = gym + gtest
        else:
            (info, self.brackt, self.stx, self.fx, self.gx, self.sty, self.fy, self.gy, stp) = cstep(
                self.brackt, self.lower, self.upper,
                self.stx, self.fx, self.gx, self.sty, self.fy, self.gy, stp, f, g)

        if self.brackt:  # :609-616
            wcur = abs(self.sty - self.stx)
            if wcur >= 0.66 * self.oldwidth:
                stp = self.stx + 0.5 * (self.sty - self.stx)
            self.oldwidth = self.width
            self.width = wcur

        if self.brackt:  # :619-630
            if self.stx <= self.sty:
                self.lower, self.upper = self.stx, self.sty
            else:
                self.lower, self.upper = self.sty, self.stx
        else:
            self.lower = stp + XTRAPL * (stp - self.stx)
            self.upper = stp + XTRAPU * (stp - self.stx)

        stp = _clamp(stp, self.stpmin, self.stpmax)  # :633
        if self.brackt and (stp <= self.lower or stp >= self.upper or
                            self.upper - self.lower <= self.xtol * self.upper):
            stp = self.stx  # :637-640
        self.stp = stp
        return self._report("SEARCH", "COMPUTE_FG")


def _sqrt(x):
    if x < 0:
        raise ValueError("sqrt of a negative number (DomainError in the reference)")
    return math.sqrt(x)


def cstep(brackt, stpmin, stpmax, stx, fx, dx, sty, fy, dy, stp, fp, dp):
    """Safeguarded cubic step, src/linesearches.jl:709-870.
    -> (info, brackt, stx, fx, dx, sty, fy, dy, stp)"""
    if brackt and (stp <= _jl_min(stx, sty) or stp >= _jl_max(stx, sty)):
        raise ValueError("STP outside bracket (STX,STY)")
    if dx * (stp - stx) >= 0:
        raise ValueError("descent condition violated")
    if stpmax < stpmin:
        raise ValueError("STPMAX < STPMIN")

    opposite = (dx < 0 and dp > 0) or (dx > 0 and dp < 0)

    if fp > fx:
        info = 1
        theta = 3 * (fx - fp) / (stp - stx) + dx + dp
        s = _jl_max(_jl_max(abs(theta), abs(dx)), abs(dp))
        gamma = s * _sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s))
        if stp < stx:
            gamma = -gamma
        p = (gamma - dx) + theta
        q = ((gamma - dx) + gamma) + dp
        r = _div(p, q)
        stpc = stx + r * (stp - stx)
        stpq = stx + (_div(dx, (fx - fp) / (stp - stx) + dx) / 2) * (stp - stx)
        if abs(stpc - stx) < abs(stpq - stx):
            stpf = stpc
        else:
            stpf = stpc + (stpq - stpc) / 2
        brackt = True
    elif opposite:
        info = 2
        theta = 3 * (fx - fp) / (stp - stx) + dx + dp
        s = _jl_max(_jl_max(abs(theta), abs(dx)), abs(dp))
        gamma = s * _sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s))
        if stp > stx:
            gamma = -gamma
        p = (gamma - dp) + theta
        q = ((gamma - dp) + gamma) + dx
        r = _div(p, q)
        stpc = stp + r * (stx - stp)
        stpq = stp + (dp / (dp - dx)) * (stx - stp)
        stpf = stpc if abs(stpc - stp) > abs(stpq - stp) else stpq
        brackt = True
    elif abs(dp) < abs(dx):
        info = 3
        theta = 3 * (fx - fp) / (stp - stx) + dx + dp
        s = _jl_max(_jl_max(abs(theta), abs(dx)), abs(dp))
        t = (theta / s) * (theta / s) - (dx / s) * (dp / s)
        gamma = s * math.sqrt(t) if t > 0 else 0.0
        if stp > stx:
            gamma = -gamma
        p = (gamma - dp) + theta
        q = (gamma + (dx - dp)) + gamma
        r = _div(p, q)
        if r < 0 and gamma != 0:
            stpc = stp + r * (stx - stp)
        elif stp > stx:
            stpc = stpmax
        else:
            stpc = stpmin
        stpq = stp + (dp / (dp - dx)) * (stx - stp)
        if brackt:
            stpf = stpc if abs(stpc - stp) < abs(stpq - stp) else stpq
            t = stp + 0.66 * (sty - stp)
            if (stpf > t) if stp > stx else (stpf < t):
                stpf = t
        else:
            stpf = stpc if abs(stpc - stp) > abs(stpq - stp) else stpq
            stpf = _jl_max(stpmin, _jl_min(stpf, stpmax))
    else:
        info = 4
        if brackt:
            theta = 3 * (fp - fy) / (sty - stp) + dy + dp
            s = _jl_max(_jl_max(abs(theta), abs(dy)), abs(dp))
            gamma = s * _sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s))
            if stp > sty:
                gamma = -gamma
            p = (gamma - dp) + theta
            q = ((gamma - dp) + gamma) + dy
            r = _div(p, q)
            stpc = stp + r * (sty - stp)
            stpf = stpc
        elif stp > stx:
            stpf = stpmax
        else:
            stpf = stpmin

    if fp > fx:
        sty, fy, dy = stp, fp, dp
    else:
        if opposite:
            sty, fy, dy = stx, fx, dx
        stx, fx, dx = stp, fp, dp
    stp = stpf
    return info, brackt, stx, fx, dx, sty, fy, dy, stp


# free functions with the reference's names
def start_(ls, f0, g0, stp, stpmin=0.0, stpmax=math.inf):
    return ls.start(f0, g0, stp, stpmin=stpmin, stpmax=stpmax)


def iterate_(ls, stp, f, g):
    return ls.iterate(stp, f, g)


def getstep(ls):
    return ls.getstep()


def gettask(ls):
    return ls.gettask()


def getstatus(ls):
    return ls.getstatus()


def getreason(ls):
    return ls.getreason()


def usederivatives(ls):
    return bool(ls.usederivatives)
