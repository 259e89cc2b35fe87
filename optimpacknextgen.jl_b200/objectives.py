"""Built-in objectives with fused f + gradient kernels, and the deterministic
synthetic inputs of the benchmarks.

* `Rosenbrock`  -- test/rosenbrock.jl:16-29 (pairs (x[2i-1], x[2i])).
* `Quadratic`   -- f = 1/2 sum a_i (x_i - c_i)^2 (BASELINE.json north_star).

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
    z = splitmix64(seed, np.arange(first, first + n, dtype=np.uint64))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


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
        a = kappa ** uniform(seed_a, first, n)
        c = -1.0 + 3.0 * uniform(seed_c, first, n)
        return a.astype(dtype), c.astype(dtype)

    @classmethod
    def synthetic_device(cls, ctx: Context, n: int, dtype=np.float64, kappa: float = 1e4,
                         first: int = 0, seed_a: int = 3, seed_c: int = 5) -> "Quadratic":
        """Same distribution generated on the device (for the full-size benchmark;
        a_i uses exp2/log2 of the GPU, so it is not bit-identical to `synthetic`)."""
        a = ctx.vector(n, dtype).fill_random(seed_a, first, 1.0, kappa, log_scale=True)
        c = ctx.vector(n, dtype).fill_random(seed_c, first, -1.0, 2.0)
        return cls(a, c)


def is_builtin(fg) -> bool:
    return isinstance(fg, (Rosenbrock, Quadratic))
