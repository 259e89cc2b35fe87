"""ONE process, several GPUs (include/opkb200.h `opkb_group_*`).

`Group(ngpus)` owns one context per device, the ranks of an in-process communicator.
`MultiVMLMB` / `vmlmb(..., ngpus=G)` shard the host arrays of a problem in contiguous blocks
over them (as `opkb200.dist.shard_range` does for the ranks of a torchrun job), create the
native driver of every rank and run them concurrently, one host thread per GPU inside the
library (`opkb_group_solver_run`): the same kernels, the same rank-order combination of the
packed partials, hence bit for bit the trajectory of the G-process run.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import check, vp
from .dist import shard_range
from .objectives import Quadratic, Rosenbrock
from .vectors import Context, Vector


class Group:
    """`opkb_group_create`: contexts[r] is rank r of `ngpus` (0 / None: every device)."""

    def __init__(self, ngpus: int = 0, devices=None):
        self._h = vp()
        n = int(ngpus or 0)
        dev = None
        if devices is not None:
            n = len(devices)
            dev = (C.c_int * n)(*[int(d) for d in devices])
        check(_lib.lib().opkb_group_create(C.byref(self._h), n, dev))
        self.size = _lib.lib().opkb_group_size(self._h)
        self.contexts = [Context.borrowed(_lib.lib().opkb_group_context(self._h, r), self) for r in range(self.size)]

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            # the contexts die with the group: mark them first, so that whatever still refers to them
            # (vectors, workspaces, solvers) only drops its handles from now on
            for c in self.contexts:
                c._h = vp()
            _lib.lib().opkb_group_destroy(self._h)
            self._h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def _shard(a, first, count, dtype=None):
    if a is None or np.isscalar(a):
        return a
    out = np.ascontiguousarray(a).reshape(-1)[first:first + count]
    return out if dtype is None else np.ascontiguousarray(out, dtype=dtype)


class MultiVMLMB:
    """`vmlmb!` of ONE host process on the G GPUs of a `Group`: keywords of `NativeVMLMB`; `fg` is a
    built-in objective (`Rosenbrock()`, `Quadratic(a, c)` with host arrays of the whole problem) or
    a callable `make_fg(rank, ctx, first, count) -> fg` returning this rank's `DeviceObjective`
    (a separable objective: `partial=True`) -- the loop of every rank stays inside the library."""

    def __init__(self, fg, x0, *, group: Group = None, ngpus: int = 0, lower=-math.inf, upper=math.inf, **kw):
        from .vmlmb import BoundedSet, NativeVMLMB
        if isinstance(lower, BoundedSet):
            lower, upper = lower.lower, lower.upper
        self._own_group = group is None
        self.group = group if group is not None else Group(ngpus)
        x0 = np.ascontiguousarray(x0)
        if x0.dtype not in (np.float32, np.float64):
            x0 = x0.astype(np.float64)
        self._shape, self.dtype = x0.shape, x0.dtype
        x0 = x0.reshape(-1)
        n, G = x0.size, self.group.size
        self.n = n
        self.ranges = [shard_range(n, r, G) for r in range(G)]
        self.solvers = []
        for r, (first, cnt) in enumerate(self.ranges):
            ctx = self.group.contexts[r]
            if isinstance(fg, Quadratic):
                obj = Quadratic(_shard(fg.a, first, cnt, self.dtype), _shard(fg.c, first, cnt, self.dtype))
            elif isinstance(fg, Rosenbrock):
                obj = fg
            elif callable(fg):
                obj = fg(r, ctx, first, cnt)
            else:
                raise TypeError("fg must be a built-in objective or a per-rank factory of DeviceObjective")
            self.solvers.append(NativeVMLMB(obj, x0[first:first + cnt].copy(), lower=_shard(lower, first, cnt, self.dtype),
                                            upper=_shard(upper, first, cnt, self.dtype), nglobal=n, ctx=ctx, **kw))
        self._bind()

    @classmethod
    def from_shards(cls, group: Group, solvers, n: int = None):
        """The same driver over per-rank `NativeVMLMB` solvers the caller has built on
        `group.contexts[r]` (e.g. with inputs generated on the devices)."""
        self = cls.__new__(cls)
        self._own_group = False
        self.group = group
        self.solvers = list(solvers)
        assert len(self.solvers) == group.size
        self.dtype = self.solvers[0].dtype
        self.ranges, first = [], 0
        for s in self.solvers:
            self.ranges.append((first, s.n))
            first += s.n
        self.n = first if n is None else n
        self._shape = (self.n,)
        self._bind()
        return self

    def _bind(self):
        G = self.group.size
        self._handles = (vp * G)(*[s._solver for s in self.solvers])
        self._tasks = (C.c_int * G)()
        self.finished = False

    def _refresh(self):
        for s, t in zip(self.solvers, self._tasks):
            s._task.value = t
            s._refresh()
        s0 = self.solvers[0]
        for name in ("iter", "eval", "rejects", "f", "gnorm", "stp", "stage", "reason"):
            setattr(self, name, getattr(s0, name))
            # identical decisions on every rank: the scalars they are made from are combined in rank order
            assert all(getattr(s, name) == getattr(s0, name) for s in self.solvers), name
        self.finished = s0.finished

    def iterate(self, niter: int = 1) -> bool:
        if self.finished:
            return False
        check(_lib.lib().opkb_group_solver_run(self.group.handle, self._handles, int(min(niter, 2 ** 62)), self._tasks))
        self._refresh()
        return not self.finished

    def solve(self):
        while self.iterate(2 ** 62):
            pass
        return self.result()

    def result(self) -> np.ndarray:
        out = np.empty(self.n, dtype=self.dtype)
        for s, (first, cnt) in zip(self.solvers, self.ranges):
            x = s.x if s.finished else s.current()
            x.download(out[first:first + cnt])
        return out.reshape(self._shape)

    def carry_stats(self):
        return self.solvers[0].carry_stats()

    def close(self):
        for s in self.solvers:
            s.close()
        self.solvers = []
        if self._own_group:
            self.group.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
