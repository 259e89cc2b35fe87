"""TEST-ONLY device back end for the host drivers: the same `ops` interface as
optimpacknextgen.jl_b200/vmlmb.py's _PrimitiveOps / _FusedOps, implemented with
the CPU oracle's vector operations on numpy arrays.  It lets the CPU suite
exercise the Python stage machine, the coefficient-space two-loop solver and
the sharded (world_size > 1, gloo) reduction logic without a GPU.  It is never
imported by the product package."""
import math

import numpy as np

import opk_oracle as O


class NumpyOps:
    """1:1 with _PrimitiveOps (reference call sequence)."""

    def __init__(self, solver, x0, obj):
        self.s = solver
        self.obj = obj
        self.x = np.array(x0, copy=True)
        n = self.x.size
        self.g = np.empty_like(self.x)
        self.p = np.empty_like(self.x) if solver.method > 0 else None
        self.d = np.empty_like(self.x)
        self.sv = np.empty_like(self.x)
        self.S = [np.empty_like(self.x) for _ in range(solver.mem)]
        self.Y = [np.empty_like(self.x) for _ in range(solver.mem)]
        self.rho = np.zeros(solver.mem)
        self._mark = 0

    def _fg(self, x, g):
        if callable(self.obj):                       # fg(x, g) -> this rank's part of f
            return self.obj(x, g)
        if self.obj == "rosenbrock":
            return O.rosenbrock_fg(x, g)
        return O.quadratic_fg(self.obj[1], self.obj[2], x, g)

    def _b(self):
        lo, hi = self.s.lo, self.s.hi
        lo = None if (np.isscalar(lo) and lo == -math.inf) else lo
        hi = None if (np.isscalar(hi) and hi == math.inf) else hi
        return lo, hi

    def evaluate(self, mark, stp, initial):
        s = self.s
        lo, hi = self._b()
        if not initial:
            O.vcombine(self.x, 1, self.S[mark], -stp, self.d)
        if s.method > 0:
            O.project_variables(self.x, self.x, lo, hi)
        f = self._fg(self.x, self.g)
        if s.method > 0:
            O.project_direction(self.p, self.x, lo, hi, -1, self.g)
        self._mark = mark
        return f

    def gnorm(self):
        return O.vnorm2(self.p if self.s.method > 0 else self.g)

    def effective_step(self):
        O.vcombine(self.sv, 1, self.x, -1, self.S[self._mark])

    def gd_search(self, stp):
        if self.s.method > 0:
            return O.vdot(self.g, self.sv) / stp
        return -O.vdot(self.g, self.d)

    def xnorm(self):
        return O.vnorm2(self.x)

    def snorm(self):
        return O.vnorm2(self.sv)

    def revert(self, mark, stp):
        lo, hi = self._b()
        O.vcombine(self.x, 1, self.S[mark], -stp, self.d)
        if self.s.method > 0:
            O.project_variables(self.x, self.x, lo, hi)

    def new_direction(self):
        s = self.s
        mark = s.mark
        lo, hi = self._b()
        O.vupdate(self.S[mark], -1, self.x)
        O.vupdate(self.Y[mark], -1, self.p if s.method == 1 else self.g)
        if s.method < 2:
            self.rho[mark] = O.vdot(self.Y[mark], self.S[mark])
            if self.rho[mark] > 0:
                s.gamma = self.rho[mark] / O.vdot(self.Y[mark], self.Y[mark])
        s.m = min(s.m + 1, s.mem)
        if s.method < 2:
            self.d[:] = self.g
            change, _ = O.apply_lbfgs(self.S, self.Y, self.rho, s.gamma, s.m, mark, self.d)
        else:
            self.d[:] = self.p
            sel = O.get_free_variables(self.d)
            change, _ = O.apply_lbfgs_sel(self.S, self.Y, self.rho, s.m, mark, self.d, sel)
        if not change:
            return False, 0.0, 0.0
        if s.method > 0:
            O.project_direction(self.d, self.x, lo, hi, -1, self.d)
        return True, -O.vdot(self.g, self.d), O.vnorm2(self.d)

    def steepest(self):
        self.d[:] = self.p if self.s.method > 0 else self.g

    def begin_search(self, mark):
        s = self.s
        lo, hi = self._b()
        self.S[mark][:] = self.x
        self.Y[mark][:] = self.p if s.method == 1 else self.g
        if s.method > 0:
            return O.step_limits(self.x, lo, hi, -1, self.d)
        return math.inf, math.inf

    def vector(self, which, slot=0):
        if which == "S":
            return self.S[slot]
        if which == "Y":
            return self.Y[slot]
        return {"x": self.x, "g": self.g, "p": self.p, "d": self.d}[which]


class NumpyGramOps(NumpyOps):
    """Same phases as _FusedOps (Gram block -> host solve -> assembly), with the
    Gram block and the assembly computed by numpy in place of the kernels; the
    coefficient-space solver is the product's own (_FusedOps._solve)."""

    def __init__(self, solver, x0, obj):
        super().__init__(solver, x0, obj)
        from opkb200.vmlmb import _FusedOps
        self._solve = lambda G, m, slots: _FusedOps._solve(self, G, m, slots)
        self.rho = [0.0] * solver.mem

    def new_direction(self):
        s = self.s
        mark = s.mark
        lo, hi = self._b()
        T = self.x.dtype.type
        O.vupdate(self.S[mark], -1, self.x)
        O.vupdate(self.Y[mark], -1, self.p if s.method == 1 else self.g)
        m = min(s.m + 1, s.mem)
        s.m = m
        slots = [(mark - (m - 1) + j) % s.mem for j in range(m)]
        v0 = self.p if s.method == 2 else self.g
        free = (v0 != 0) if s.method == 2 else np.ones(v0.size, dtype=bool)
        W = m + 1
        G = np.zeros((2 * m, W))
        cols = [v0] + [self.Y[k] for k in slots]
        for a, k in enumerate(slots):
            for ci, cv in enumerate(cols):
                G[2 * a, ci] = O.vdot_sel(np.flatnonzero(free), self.S[k], cv)
                G[2 * a + 1, ci] = O.vdot_sel(np.flatnonzero(free), self.Y[k], cv)
        alpha, c, gamma, change = self._solve(list(G.reshape(-1)), m, slots)
        if not change:
            return False, 0.0, 0.0
        # assembly with the reference's rounding sequence (what k_direction does)
        v = v0.copy()
        for j in range(m - 1, -1, -1):
            if alpha[j] == alpha[j]:
                v[free] = v[free] + T(-alpha[j]) * self.Y[slots[j]][free]
        v = T(gamma) * v
        for j in range(m):
            if alpha[j] == alpha[j]:
                v[free] = v[free] + T(c[j]) * self.S[slots[j]][free]
        self.d[:] = v
        if s.method > 0:
            O.project_direction(self.d, self.x, lo, hi, -1, self.d)
        return True, -O.vdot(self.g, self.d), O.vnorm2(self.d)


class ShardedNumpyOps(NumpyOps):
    """The N > 1 data path on CPU: every rank holds a contiguous shard, each
    reduction produces per-rank partials that are all-gathered (gloo) and
    combined in fixed rank order by opkb200.dist.combine_partials -- the same
    scheme libopkb200 uses over NCCL (opkb_finish_reduction)."""

    def __init__(self, solver, x0, obj):
        super().__init__(solver, x0, obj)
        self.alpha = [0.0] * solver.mem
        self.rho = [0.0] * solver.mem

    def _allreduce(self, vals, kinds=None):
        import torch
        import torch.distributed as dist
        from opkb200.dist import combine_partials
        t = torch.tensor(vals, dtype=torch.float64)
        out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(out, t)
        return combine_partials([o.tolist() for o in out], kinds)

    def _dot(self, x, y, sel=None):
        part = O.vdot(x, y) if sel is None else O.vdot_sel(sel, x, y)
        return self._allreduce([part])[0]

    def evaluate(self, mark, stp, initial):
        f = super().evaluate(mark, stp, initial)
        return self._allreduce([f])[0]

    def gnorm(self):
        v = self.p if self.s.method > 0 else self.g
        return math.sqrt(self._dot(v, v))

    def gd_search(self, stp):
        if self.s.method > 0:
            return self._dot(self.g, self.sv) / stp
        return -self._dot(self.g, self.d)

    def xnorm(self):
        return math.sqrt(self._dot(self.x, self.x))

    def snorm(self):
        return math.sqrt(self._dot(self.sv, self.sv))

    def new_direction(self):
        s = self.s
        mark, mem = s.mark, s.mem
        lo, hi = self._b()
        S, Y, rho, alpha = self.S, self.Y, self.rho, self.alpha
        O.vupdate(S[mark], -1, self.x)
        O.vupdate(Y[mark], -1, self.p if s.method == 1 else self.g)
        if s.method < 2:
            rho[mark] = self._dot(Y[mark], S[mark])
            if rho[mark] > 0:
                s.gamma = rho[mark] / self._dot(Y[mark], Y[mark])
        s.m = min(s.m + 1, mem)
        v = self.d
        v[:] = self.p if s.method == 2 else self.g
        sel = O.get_free_variables(self.p) if s.method == 2 else None
        gamma = 0.0 if s.method == 2 else s.gamma
        modif = False
        k = mark + 1
        for _ in range(s.m):                     # apply_lbfgs!, src/quasinewton.jl:716-779
            k = k - 1 if k > 0 else mem - 1
            if s.method == 2:
                rho[k] = self._dot(Y[k], S[k], sel)
            if rho[k] > 0:
                alpha[k] = self._dot(S[k], v, sel) / rho[k]
                if sel is None:
                    O.vupdate(v, -alpha[k], Y[k])
                else:
                    O.vupdate_sel(v, sel, -alpha[k], Y[k])
                modif = True
                if s.method == 2 and gamma == 0:
                    gamma = rho[k] / self._dot(Y[k], Y[k], sel)
        change = modif and gamma != 0
        if change:
            O.vscale(v, gamma)
            for _ in range(s.m):
                if rho[k] > 0:
                    beta = self._dot(Y[k], v, sel) / rho[k]
                    if sel is None:
                        O.vupdate(v, alpha[k] - beta, S[k])
                    else:
                        O.vupdate_sel(v, sel, alpha[k] - beta, S[k])
                k = k + 1 if k < mem - 1 else 0
        if not change:
            return False, 0.0, 0.0
        if s.method > 0:
            O.project_direction(self.d, self.x, lo, hi, -1, self.d)
        return True, -self._dot(self.g, self.d), math.sqrt(self._dot(self.d, self.d))

    def begin_search(self, mark):
        smin, smax = super().begin_search(mark)
        return tuple(self._allreduce([smin, smax], [2, 1]))


class ShardedSmooth2D:
    """CPU twin of `opkb_smooth2d_fg` on one rank of a row-sharded grid: the last
    row of the previous rank and the first row of the next one are exchanged
    (gloo isend / irecv), the vertical edge between two ranks is counted by the
    lower one, and the call returns THIS rank's part of f (the caller sums the
    parts in rank order).  Same element-wise arithmetic as Smooth2D.reference."""

    def __init__(self, w, y, mu, cols):
        self.w = np.ascontiguousarray(w).reshape(-1, cols)
        self.y = np.ascontiguousarray(y).reshape(-1, cols)
        self.mu, self.cols = mu, cols

    def __call__(self, x, g):
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        x2 = x.reshape(-1, self.cols)
        up = torch.zeros(self.cols, dtype=torch.float64)
        dn = torch.zeros(self.cols, dtype=torch.float64)
        reqs = []
        if rank > 0:
            reqs += [dist.isend(torch.from_numpy(x2[0].copy()), rank - 1), dist.irecv(up, rank - 1)]
        if rank < world - 1:
            reqs += [dist.isend(torch.from_numpy(x2[-1].copy()), rank + 1), dist.irecv(dn, rank + 1)]
        for r in reqs:
            r.wait()
        xx = np.vstack(([up.numpy()] if rank > 0 else []) + [x2] + ([dn.numpy()] if rank < world - 1 else []))
        o = 1 if rank > 0 else 0                      # offset of the first own row in xx
        rows = x2.shape[0]
        r = x2 - self.y
        t = self.w * r
        du, dd, dl, dr = (np.zeros_like(x2) for _ in range(4))
        lo = 0 if rank > 0 else 1                     # own rows that have a row above
        du[lo:, :] = xx[o + lo:o + rows, :] - xx[o + lo - 1:o + rows - 1, :]
        hi = rows if rank < world - 1 else rows - 1   # own rows that have a row below
        dd[:hi, :] = xx[o:o + hi, :] - xx[o + 1:o + hi + 1, :]
        dl[:, 1:] = x2[:, 1:] - x2[:, :-1]
        dr[:, :-1] = x2[:, :-1] - x2[:, 1:]
        g.reshape(-1, self.cols)[...] = t + self.mu * (((du + dd) + dl) + dr)
        return 0.5 * float(np.sum(t * r)) + 0.5 * self.mu * float(np.sum(du * du) + np.sum(dl * dl))


class NumpyCgOps:
    """Vector back end of `opkb200.ConjGrad` (same interface as conjgrad._DeviceOps) on
    numpy arrays with the oracle's vector operations; `comm` (optional) sums the partial
    dot products over the ranks in rank order, like the library's packed all-gather."""

    def __init__(self, comm=None):
        self.comm = comm

    def _sum(self, *vals):
        if self.comm is None:
            return vals if len(vals) > 1 else vals[0]
        out = self.comm(np.array(vals, dtype=np.float64))
        return tuple(out) if len(vals) > 1 else float(out[0])

    def workspace(self, b):
        return np.empty_like(b), np.empty_like(b), np.empty_like(b)

    def dot(self, x, y):
        return self._sum(O.vdot(x, y))

    def copy(self, dst, src):
        O.vcopy(dst, src)

    def combine(self, dst, a, x, b, y):
        O.vcombine(dst, a, x, b, y)

    def apply(self, A, dst, src):
        A(dst, src)

    def direction_apply(self, A, p, r, q, beta, first):
        if first:
            O.vcopy(p, r)
        else:
            O.vcombine(p, beta, p, 1.0, r)
        A(q, p)
        return self._sum(O.vdot(p, q), O.vdot(p, p))

    def update_x(self, x, alpha, p):
        O.vupdate(x, alpha, p)

    def update(self, x, r, p, q, alpha):
        O.vupdate(x, alpha, p)
        O.vupdate(r, -alpha, q)
        return self._sum(O.vdot(r, r), O.vdot(x, x))
