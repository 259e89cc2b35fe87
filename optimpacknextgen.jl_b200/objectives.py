"""Built-in objectives with fused f + gradient kernels, and the deterministic
synthetic inputs of the benchmarks.

* `Rosenbrock`  -- test/rosenbrock.jl:16-29 (pairs (x[2i-1], x[2i])).
* `Quadratic`   -- f = 1/2 sum a_i (x_i - c_i)^2 (BASELINE.json north_star).

* `Smooth2D`    -- smoothness-regularised weighted least squares on a 2-D grid:
                   non-separable, row-sharded, with a halo exchange over NVLink
                   (SURVEY.md 8(f)-2); used through the user-objective path.

A user objective is any callable `fg(x: Vector, g: Vector) -> float` working on
device vectors (the reference's `fg!(x, g)`).
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .vectors import Context, Vector

_MASK = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(seed: int, idx) -> np.ndarray:
    """Counter-based generator: the same integers as opkb_splitmix64 in the library."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (np.asarray(idx, dtype=np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform(seed: int, first: int, n: int) -> np.ndarray:
    """u(seed, i) in [0, 1) for i = first .. first+n-1 (53-bit, Float64)."""
    out = np.empty(n, dtype=np.float64)
    step = 1 << 22      # in chunks: the temporaries of a 2^28 vector would be several times its size
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        z = splitmix64(seed, np.arange(first + lo, first + hi, dtype=np.uint64))
        out[lo:hi] = (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return out


class Rosenbrock:
    """f = sum (1 - x1)^2 + sum (10 (x2 - x1^2))^2 over pairs (test/rosenbrock.jl:16-29)."""

    kind = _lib.OBJ_ROSENBROCK
    a = None
    c = None

    @staticmethod
    def x0(n: int, dtype=np.float64, perturb_seed=None, first: int = 0) -> np.ndarray:
        """rosenbrock_init! (test/rosenbrock.jl:10-14); with `perturb_seed` the
        SURVEY 8(d) "pert" variant adds 0.1*(2u - 1) to break the pair degeneracy."""
        x = np.empty(n, dtype=np.float64)
        x[0::2] = -1.2
        x[1::2] = 1.0
        if perturb_seed is not None:
            x += 0.1 * (2.0 * uniform(perturb_seed, first, n) - 1.0)
        return x.astype(dtype)

    def vectors(self, ctx: Context, n: int, dtype):
        return None, None


class Quadratic:
    """f = 1/2 sum a_i (x_i - c_i)^2.  `a`, `c` are numpy arrays or device Vectors
    (this rank's shard)."""

    kind = _lib.OBJ_QUADRATIC

    def __init__(self, a, c):
        self.a = a
        self.c = c
        self._va = self._vc = None

    def vectors(self, ctx: Context, n: int, dtype):
        if self._va is None:
            self._va = self.a if isinstance(self.a, Vector) else ctx.from_numpy(np.asarray(self.a, dtype=dtype))
            self._vc = self.c if isinstance(self.c, Vector) else ctx.from_numpy(np.asarray(self.c, dtype=dtype))
        if self._va.n != n or self._vc.n != n:
            raise ValueError("a, c must have the length of the variables")
        return self._va, self._vc

    @staticmethod
    def synthetic(n: int, dtype=np.float64, kappa: float = 1e4, first: int = 0, seed_a: int = 3,
                  seed_c: int = 5):
        """a_i = kappa^u(3,i), c_i = -1 + 3 u(5,i) (host arrays)."""
        a = uniform(seed_a, first, n)
        c = uniform(seed_c, first, n)
        step = 1 << 22
        for lo in range(0, n, step):   # in place, chunk by chunk (same values as the array expressions)
            a[lo:lo + step] = kappa ** a[lo:lo + step]
            c[lo:lo + step] = -1.0 + 3.0 * c[lo:lo + step]
        return a.astype(dtype, copy=False), c.astype(dtype, copy=False)

    @classmethod
    def synthetic_device(cls, ctx: Context, n: int, dtype=np.float64, kappa: float = 1e4,
                         first: int = 0, seed_a: int = 3, seed_c: int = 5) -> "Quadratic":
        """Same distribution generated on the device (for the full-size benchmark;
        a_i uses exp2/log2 of the GPU, so it is not bit-identical to `synthetic`)."""
        a = ctx.vector(n, dtype).fill_random(seed_a, first, 1.0, kappa, log_scale=True)
        c = ctx.vector(n, dtype).fill_random(seed_c, first, -1.0, 2.0)
        return cls(a, c)


class Smooth2D:
    """f(x) = 1/2 sum w (x - y)^2 + mu/2 sum_edges (x_i - x_j)^2 on a rows x cols grid
    (row-major; horizontal and vertical neighbours; free boundary).  A callable
    `fg(x, g) -> f` on device Vectors, i.e. a user objective for `vmlmb` / `spg`.

    With several ranks (`ctx.nranks > 1`) every rank holds a block of whole rows of
    x, w, y, g; the boundary rows are exchanged point to point (one ncclSend/ncclRecv
    group) while the inner rows are processed (`opkb_smooth2d_fg`)."""

    def __init__(self, w, y, mu: float, cols: int, ctx: Context = None):
        from .vectors import default_context
        self.ctx = ctx or (w.ctx if isinstance(w, Vector) else default_context())
        self.w = w if isinstance(w, Vector) else self.ctx.from_numpy(np.ascontiguousarray(w).reshape(-1))
        self.y = y if isinstance(y, Vector) else self.ctx.from_numpy(
            np.ascontiguousarray(y, dtype=self.w.dtype).reshape(-1))
        self.mu = float(mu)
        self.cols = int(cols)
        if self.w.n != self.y.n or self.w.n % self.cols:
            raise ValueError("w, y must hold whole rows of `cols` elements")
        self.up = self.dn = None
        if self.ctx.nranks > 1:
            self.up = self.ctx.vector(self.cols, self.w.dtype)
            self.dn = self.ctx.vector(self.cols, self.w.dtype)
        self._f = _lib.f64()

    def __call__(self, x: Vector, g: Vector) -> float:
        import ctypes as C
        _lib.check(_lib.lib().opkb_smooth2d_fg(
            self.ctx.handle, self.w.handle, self.y.handle, self.mu, self.cols, x.handle,
            self.up.handle if self.up is not None else None, self.dn.handle if self.dn is not None else None, g.handle,
            C.byref(self._f)))
        return self._f.value

    @staticmethod
    def reference(w, y, mu, cols):
        """The same objective as a numpy `fg(x, g) -> f` on the WHOLE grid (same
        order of the elementwise operations: g is bit-identical); for parity
        checks against the CPU oracle, which takes python callables."""
        w = np.ascontiguousarray(w)
        dt = w.dtype
        w2 = w.reshape(-1, cols)
        y2 = np.ascontiguousarray(y, dtype=dt).reshape(-1, cols)
        mu_t = dt.type(mu)

        def fg(x, g):
            x2 = x.reshape(-1, cols)
            r = x2 - y2
            t = w2 * r
            du = np.zeros_like(x2)
            dd = np.zeros_like(x2)
            dl = np.zeros_like(x2)
            dr = np.zeros_like(x2)
            du[1:, :] = x2[1:, :] - x2[:-1, :]
            dd[:-1, :] = x2[:-1, :] - x2[1:, :]
            dl[:, 1:] = x2[:, 1:] - x2[:, :-1]
            dr[:, :-1] = x2[:, :-1] - x2[:, 1:]
            s = ((du + dd) + dl) + dr
            g.reshape(-1, cols)[...] = t + mu_t * s
            fd = float(np.sum((t * r).astype(np.float64)))
            fe = float(np.sum((du * du).astype(np.float64)) + np.sum((dl * dl).astype(np.float64)))
            return 0.5 * fd + (0.5 * float(mu)) * fe
        return fg

    @staticmethod
    def synthetic(rows: int, cols: int, dtype=np.float64, first_row: int = 0, seed: int = 11):
        """Deterministic data for the benchmarks: y = a smooth image + noise in [0, 1.2],
        w in [0.5, 1.5] (this rank's rows first_row .. first_row + rows - 1)."""
        i = (np.arange(first_row, first_row + rows, dtype=np.float64))[:, None]
        j = np.arange(cols, dtype=np.float64)[None, :]
        u = uniform(seed, first_row * cols, rows * cols).reshape(rows, cols)
        v = uniform(seed + 1, first_row * cols, rows * cols).reshape(rows, cols)
        y = 0.6 + 0.4 * np.sin(i / 37.0) * np.cos(j / 53.0) + 0.4 * (u - 0.5)
        w = 0.5 + v
        return w.astype(dtype).reshape(-1), y.astype(dtype).reshape(-1)


class DeviceObjective:
    """A user objective as a DEVICE CALLBACK for the native driver (`NativeVMLMB`,
    `opkb_solver_set_fg_callback`): `fn(n, x_ptr, g_ptr, stream) -> f` receives raw device
    pointers (integers) to this rank's x and g and the context's cudaStream_t; it must write the
    gradient with kernels submitted to that stream (or completed before it returns) and return
    f(x).  The whole reverse-communication loop then runs inside the library.  `partial=True`:
    the returned value is this rank's share of a separable objective (summed over the ranks by the
    library).  `fn` may also be a C function pointer of type `_lib.FG_CALLBACK` (e.g. the address
    of a compiled CUDA objective), with `user` passed through."""

    def __init__(self, fn, partial: bool = False, user=None):
        self.partial = bool(partial)
        self.user = user
        self.error = None
        if isinstance(fn, _lib.FG_CALLBACK):
            self._cb = fn
        else:
            def _trampoline(_user, n, xp, gp, stream, fout):
                try:
                    fout[0] = float(fn(int(n), int(xp or 0), int(gp or 0), int(stream or 0)))
                    return 0
                except BaseException as e:   # never unwind through the C frames
                    self.error = e
                    return 1
            self._cb = _lib.FG_CALLBACK(_trampoline)

    def __call__(self, x: Vector, g: Vector) -> float:
        """Also usable as an ordinary `fg(x, g)` (Python-hosted drivers): same callback, pointers
        asked for at every call -- they move between solver calls."""
        import ctypes as C
        f = C.c_double()
        if self._cb(self.user, x.n, x.data_ptr, g.data_ptr, x.ctx.stream, C.byref(f)) != 0:
            raise self.error or RuntimeError("the objective callback failed")
        return f.value


def is_builtin(fg) -> bool:
    return isinstance(fg, (Rosenbrock, Quadratic))
