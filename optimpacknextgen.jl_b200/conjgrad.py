"""`conjgrad` / `conjgrad!`: linear conjugate gradient.

The reference re-exports both names from LazyAlgebra (src/OptimPackNextGen.jl:18-19,
:36); LazyAlgebra is not vendored and not version pinned (Project.toml:8), so the
host loop below restates its published algorithm (LazyAlgebra `src/conjgrad.jl`, 0.2.x:
Hestenes-Stiefel recurrences with beta = rho/oldrho, periodic restart on the true
residuals, the `ftol` / `gtol` / `xtol` tests), with the keywords the reference's own
call site uses (test/conjgrad-tests.jl:30: `conjgrad(A!, b; verb, ftol=1e-16,
maxiter=length(b))`).

    conjgrad(A, b, x0=None; ftol=1e-8, gtol=(0, 0), xtol=0, maxiter=typemax,
             restart=min(50, length(b)), verb=False, io=stdout, quiet=False,
             strict=True) -> x

minimises x'.A.x/2 - b'.x, i.e. solves A.x = b for a symmetric positive definite A.
`A` is any callable `A(dst, src)` on device Vectors (the reference's `A!(dst, src)`),
or one of the built-in operators whose kernels return p.q and p.p from the same pass
as q = A.p:

* `DiagonalOperator(a)`    -- `vproduct!`; direction + product + curvature = one kernel;
* `Smooth2DOperator(w, mu, cols)` -- diag(w) + mu * (5-point Laplacian, free boundary),
  the Hessian of `Smooth2D`; row-sharded over the ranks with the halo exchange.

Per iteration the vectors cross HBM 11 (diagonal, stencil) / 14 (user operator, plus the
operator) times instead of the reference's 15-17.
"""
from __future__ import annotations

import ctypes as C
import math
import sys

import numpy as np

from . import _lib
from ._lib import check, f64
from .vectors import Context, Vector, default_context, vcombine_, vcopy_, vdot, vupdate_

# status of the last run (`ConjGrad.status`)
CG_SEARCHING = 0
CG_GTEST = 1
CG_FTEST = 2
CG_XTEST = 3
CG_TOO_MANY_ITERATIONS = -1
CG_NOT_POSITIVE_DEFINITE = -2

CG_REASON = {
    CG_SEARCHING: "Work in progress",
    CG_GTEST: "Convergence (gtest satisfied)",
    CG_FTEST: "Convergence (ftest satisfied)",
    CG_XTEST: "Convergence (xtest satisfied)",
    CG_TOO_MANY_ITERATIONS: "Too many iteration(s)",
    CG_NOT_POSITIVE_DEFINITE: "Operator is not positive definite",
}


class DiagonalOperator:
    """A = diag(a): `A(dst, src)` is `vproduct!(dst, a, src)`."""

    def __init__(self, a, ctx: Context = None):
        self.ctx = ctx or (a.ctx if isinstance(a, Vector) else default_context())
        self.a = a if isinstance(a, Vector) else self.ctx.from_numpy(np.ascontiguousarray(a).reshape(-1))

    def __call__(self, dst: Vector, src: Vector) -> Vector:
        check(_lib.lib().opkb_vproduct(self.ctx.handle, dst.handle, self.a.handle, src.handle))
        return dst

    def direction_apply(self, p: Vector, r: Vector, q: Vector, beta: float, restart: bool):
        """p = r | beta*p + r; q = A.p -> (p.q, p.p): one kernel."""
        out = (f64 * 2)()
        check(_lib.lib().opkb_cg_diag_step(self.ctx.handle, self.a.handle, p.handle, r.handle, q.handle,
                                           float(beta), int(restart), out))
        return out[0], out[1]


class Smooth2DOperator:
    """A = diag(w) + mu*L on a rows x cols grid (L: graph Laplacian of the horizontal and
    vertical neighbour pairs): the Hessian of `Smooth2D`, so that `conjgrad(A, w.*y)` is
    its unconstrained minimiser.  Row-sharded like `Smooth2D` (`opkb_smooth2d_apply`)."""

    def __init__(self, w, mu: float, cols: int, ctx: Context = None):
        self.ctx = ctx or (w.ctx if isinstance(w, Vector) else default_context())
        self.w = w if isinstance(w, Vector) else self.ctx.from_numpy(np.ascontiguousarray(w).reshape(-1))
        self.mu = float(mu)
        self.cols = int(cols)
        if self.w.n % self.cols:
            raise ValueError("w must hold whole rows of `cols` elements")
        self.up = self.dn = None
        if self.ctx.nranks > 1:
            self.up = self.ctx.vector(self.cols, self.w.dtype)
            self.dn = self.ctx.vector(self.cols, self.w.dtype)
        self._out = (f64 * 2)()
        self._scratch = None
        self.up_r = self.dn_r = self.up_p = self.dn_p = None

    def direction_apply(self, p: Vector, r: Vector, q: Vector, beta: float, restart: bool):
        """p = r | beta*p + r; q = A.p -> (p.q, p.p): one kernel, 5 passes
        (`opkb_cg_smooth2d_step`; the neighbours of an element must see the OLD direction
        during the pass, so the new one goes to a scratch vector and the buffers are swapped)."""
        if self._scratch is None or self._scratch.n != p.n:
            self._scratch = self.ctx.vector(p.n, p.dtype)
            if self.ctx.nranks > 1:
                # the neighbours' boundary rows of r (received at every call) and of the
                # direction (kept from call to call: they follow the same recurrence)
                self.up_r, self.dn_r, self.up_p, self.dn_p = (self.ctx.vector(self.cols, p.dtype)
                                                              for _ in range(4))
        h = lambda v: v.handle if v is not None else None
        check(_lib.lib().opkb_cg_smooth2d_step(
            self.ctx.handle, self.w.handle, self.mu, self.cols, p.handle, self._scratch.handle, r.handle,
            q.handle, h(self.up_p), h(self.dn_p), h(self.up_r), h(self.dn_r), float(beta), int(restart),
            self._out))
        return self._out[0], self._out[1]

    def apply_dot(self, q: Vector, p: Vector):
        """q = A.p -> (p.q, p.p), reduced in the same pass."""
        check(_lib.lib().opkb_smooth2d_apply(
            self.ctx.handle, self.w.handle, self.mu, self.cols, p.handle,
            self.up.handle if self.up is not None else None, self.dn.handle if self.dn is not None else None, q.handle, self._out))
        return self._out[0], self._out[1]

    def __call__(self, dst: Vector, src: Vector) -> Vector:
        self.apply_dot(dst, src)
        return dst

    @staticmethod
    def reference(w, mu, cols):
        """The same operator as a numpy `A(dst, src)` on the WHOLE grid, same order of the
        element-wise operations (bit-identical product); for the CPU oracle."""
        w2 = np.ascontiguousarray(w).reshape(-1, cols)
        mu_t = w2.dtype.type(mu)

        def A(dst, src):
            x2 = src.reshape(-1, cols)
            du = np.zeros_like(x2)
            dd = np.zeros_like(x2)
            dl = np.zeros_like(x2)
            dr = np.zeros_like(x2)
            du[1:, :] = x2[1:, :] - x2[:-1, :]
            dd[:-1, :] = x2[:-1, :] - x2[1:, :]
            dl[:, 1:] = x2[:, 1:] - x2[:, :-1]
            dr[:, :-1] = x2[:, :-1] - x2[:, 1:]
            dst.reshape(-1, cols)[...] = w2 * x2 + mu_t * (((du + dd) + dl) + dr)
            return dst
        return A


def _cg_update(ctx, x, r, p, q, alpha):
    out = (f64 * 2)()
    check(_lib.lib().opkb_cg_update(ctx.handle, x.handle, r.handle, p.handle, q.handle, float(alpha), out))
    return out[0], out[1]


def _cg_curvature(ctx, p, q):
    out = (f64 * 2)()
    check(_lib.lib().opkb_cg_curvature(ctx.handle, p.handle, q.handle, out))
    return out[0], out[1]


class ConjGrad:
    """Steppable `conjgrad!`.  `iterate()` is one trip of the loop; `solve()` runs to
    termination and returns x.  After termination: `status`, `iter`, `rho` (squared norm
    of the residuals), `phi` (decrease of the objective since x0)."""

    def __init__(self, A, b, x0=None, *, p: Vector = None, q: Vector = None, r: Vector = None,
                 ftol=1e-8, gtol=(0.0, 0.0), xtol=0.0, maxiter=None, restart=None, verb=False, io=None,
                 quiet=False, strict=True, ctx: Context = None, observer=None, nglobal=None, ops=None):
        external = ops is not None      # a test-only vector back end (tests/cpu_backend.py)
        self._numpy_in = not external and not isinstance(b, Vector)
        if external:
            self.ctx = None
            self.b = b
            self.x = x0
            self.n = len(b)
        else:
            self.ctx = ctx or (b.ctx if isinstance(b, Vector) else default_context())
            if self._numpy_in:
                b = np.ascontiguousarray(b)
                if b.dtype not in (np.float32, np.float64):
                    b = b.astype(np.float64)
                self._shape = b.shape
                self.b = self.ctx.from_numpy(b.reshape(-1))
                if x0 is None:
                    self.x = self.ctx.vector(self.b.n, self.b.dtype).fill(0.0)
                else:
                    self.x = self.ctx.from_numpy(np.ascontiguousarray(x0, dtype=b.dtype).reshape(-1))
            else:
                self.b = b
                if x0 is None:
                    self.x = self.ctx.vector(b.n, b.dtype).fill(0.0)
                elif isinstance(x0, Vector):
                    self.x = x0                   # conjgrad! works in place
                else:
                    self.x = self.ctx.from_numpy(np.ascontiguousarray(x0, dtype=b.dtype).reshape(-1))
            if self.x.n != self.b.n or self.x.dtype != self.b.dtype:
                raise ValueError("x0 and b must belong to the same space")
            self.n = self.b.n
        self.A = A
        self.ftol = float(ftol)
        self.gatol, self.grtol = float(gtol[0]), float(gtol[1])
        self.xtol = float(xtol)
        if not (self.ftol >= 0 and self.gatol >= 0 and self.grtol >= 0 and self.xtol >= 0):
            raise ValueError("tolerances must be nonnegative")
        self.maxiter = (2 ** 63 - 1) if maxiter is None else int(maxiter)
        nglob = int(nglobal) if nglobal is not None else self.n
        self.restart = min(50, max(nglob, 1)) if restart is None else int(restart)
        if self.restart < 1:
            raise ValueError("restart must be at least 1")
        self.verb, self.io, self.quiet, self.strict = bool(verb), io or sys.stdout, bool(quiet), bool(strict)
        self.observer = observer
        if external:
            self.ops = ops
            self.p, self.q, self.r = ops.workspace(self.b)
        else:
            self.ops = _DeviceOps(self.ctx, A)
            mk = lambda v: v if v is not None else self.ctx.vector(self.n, self.b.dtype)
            self.p, self.q, self.r = mk(p), mk(q), mk(r)
        self.iter = 0
        self.status = CG_SEARCHING
        self.rho = self.oldrho = self.phi = self.psi = self.psimax = 0.0
        self.gtest = 0.0
        self.finished = False
        self._started = False

    # ---------------------------------------------------------------------------------
    def _start(self):
        o = self.ops
        if o.dot(self.x, self.x) > 0:          # r = b - A.x (x0 may be non-zero)
            o.apply(self.A, self.r, self.x)
            o.combine(self.r, 1.0, self.b, -1.0, self.r)
        else:                                   # spare one product when x0 = 0
            o.copy(self.r, self.b)
        self.rho = o.dot(self.r, self.r)
        self.gtest = max(self.gatol, self.grtol * math.sqrt(self.rho))
        self._started = True

    def _stop(self, status):
        self.status = status
        self.finished = True
        if status < 0:
            msg = CG_REASON[status]
            if status == CG_NOT_POSITIVE_DEFINITE and self.strict:
                raise ArithmeticError(msg)
            if not self.quiet:
                print("# WARNING: " + msg, file=sys.stderr)
        return False

    def iterate(self) -> bool:
        """One trip of the loop.  Returns False once finished."""
        if self.finished:
            return False
        if not self._started:
            self._start()
        o, k = self.ops, self.iter
        if self.verb:
            if k == 0:
                self.io.write("# %s\n# %s\n" % ("  ITER       PHI          ‖R‖", "-" * 36))
            self.io.write(" %6d %12.4e %12.4e\n" % (k, self.phi, math.sqrt(self.rho)))
        if self.observer is not None:
            self.observer(self)
        if math.sqrt(self.rho) <= self.gtest:
            return self._stop(CG_GTEST)
        if k >= self.maxiter:
            return self._stop(CG_TOO_MANY_ITERATIONS)

        # search direction, q = A.p, gamma = p.q (and p.p for the x test)
        first = (k % self.restart == 0)
        if first and k > 0:                     # restart on the true residuals
            o.apply(self.A, self.r, self.x)
            o.combine(self.r, 1.0, self.b, -1.0, self.r)
        beta = 0.0 if first else self.rho / self.oldrho
        gamma, pp = o.direction_apply(self.A, self.p, self.r, self.q, beta, first)
        if not gamma > 0:
            return self._stop(CG_NOT_POSITIVE_DEFINITE)
        alpha = self.rho / gamma

        # x += alpha*p, r -= alpha*q, rho = r.r: one pass (the two tests of the reference
        # sit between the two updates; phi's test only needs host scalars, so it is decided
        # first and r is left alone as in the reference)
        self.psi = alpha * self.rho / 2          # phi(x_k) - phi(x_{k+1})
        self.phi -= self.psi
        self.psimax = max(self.psi, self.psimax)
        if self.psi <= self.ftol * self.psimax:
            o.update_x(self.x, alpha, self.p)
            self.iter = k + 1
            return self._stop(CG_FTEST)
        rho, xx = o.update(self.x, self.r, self.p, self.q, alpha)
        self.iter = k + 1
        if self.xtol > 0 and alpha * math.sqrt(pp) <= self.xtol * math.sqrt(xx):
            return self._stop(CG_XTEST)          # (r is one update ahead of the reference here)
        self.oldrho, self.rho = self.rho, rho
        return True

    def solve(self):
        while self.iterate():
            pass
        return self.x

    def result(self):
        if self._numpy_in:
            return self.x.download().reshape(self._shape)
        return self.x


class _DeviceOps:
    """The vector operations of one iteration on the GPU (C ABI), fused where the
    operator allows it."""

    def __init__(self, ctx, A):
        self.ctx = ctx
        self._fused_dir = hasattr(A, "direction_apply")
        self._fused_dot = hasattr(A, "apply_dot")

    def dot(self, x, y):
        return vdot(x, y)

    def copy(self, dst, src):
        vcopy_(dst, src)

    def combine(self, dst, a, x, b, y):
        vcombine_(dst, a, x, b, y)

    def apply(self, A, dst, src):
        A(dst, src)

    def direction_apply(self, A, p, r, q, beta, first):
        if self._fused_dir:
            return A.direction_apply(p, r, q, beta, first)
        if first:
            vcopy_(p, r)
        else:
            vcombine_(p, beta, p, 1.0, r)
        if self._fused_dot:
            return A.apply_dot(q, p)
        A(q, p)
        return _cg_curvature(self.ctx, p, q)

    def update_x(self, x, alpha, p):
        vupdate_(x, alpha, p)

    def update(self, x, r, p, q, alpha):
        return _cg_update(self.ctx, x, r, p, q, alpha)


def conjgrad_(x: Vector, A, b: Vector, x0: Vector = None, p: Vector = None, q: Vector = None,
              r: Vector = None, **kwds) -> Vector:
    """`conjgrad!(x, A, b, x0, p, q, r; ...)`: the solution overwrites x (a device Vector);
    p, q, r are optional workspace vectors."""
    if x0 is not None and x0 is not x:
        vcopy_(x, x0)
    elif x0 is None:
        x.fill(0.0)
    s = ConjGrad(A, b, x, p=p, q=q, r=r, **kwds)
    s.solve()
    return x


def conjgrad(A, b, x0=None, **kwds):
    """`conjgrad(A, b, x0=vzeros(b); ...) -> x` (host array in -> host array out; device
    Vector in -> device Vector out, x0 is not modified)."""
    if isinstance(b, Vector) and isinstance(x0, Vector):
        x = b.ctx.vector(b.n, b.dtype)
        vcopy_(x, x0)
        x0 = x
    s = ConjGrad(A, b, x0, **kwds)
    s.solve()
    return s.result()
