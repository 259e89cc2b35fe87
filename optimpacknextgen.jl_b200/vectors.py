"""Device contexts and n-vectors, plus the LazyAlgebra / SimpleBounds operations
of the reference (doc/algebra.md:84-116, src/bounds.jl:17-26) on them.

Every function is one call of the C ABI (one CUDA kernel); the names and the
argument order are the reference's, with the `!` dropped."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import OpkbError, check, f64, i64, vp

_DEFAULT_CTX = None


def _eltype(dtype) -> int:
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return _lib.DOUBLE
    if dtype == np.float32:
        return _lib.FLOAT
    raise TypeError("only Float32 and Float64 variables are supported")


def _dtype(eltype: int):
    return np.float64 if eltype == _lib.DOUBLE else np.float32


class Context:
    """One CUDA device + stream + reduction scratch (+ NCCL communicator)."""

    def __init__(self, device: int = -1):
        self._h = vp()
        self._owner = None
        check(_lib.lib().opkb_context_create(C.byref(self._h), int(device)))

    @classmethod
    def borrowed(cls, handle, owner) -> "Context":
        """A context that belongs to `owner` (a `Group`): never destroyed from here."""
        self = cls.__new__(cls)
        self._h = vp(handle)
        self._owner = owner
        return self

    @property
    def handle(self):
        return self._h

    @property
    def device(self) -> int:
        return _lib.lib().opkb_context_device(self._h)

    @property
    def sm_count(self) -> int:
        return _lib.lib().opkb_context_sm_count(self._h)

    @property
    def launch_count(self) -> int:
        return _lib.lib().opkb_context_launch_count(self._h)

    @property
    def rank(self) -> int:
        return _lib.lib().opkb_comm_rank(self._h)

    @property
    def nranks(self) -> int:
        return _lib.lib().opkb_comm_size(self._h)

    def synchronize(self) -> None:
        check(_lib.lib().opkb_context_synchronize(self._h))

    def timer_start(self) -> None:
        check(_lib.lib().opkb_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = f64()
        check(_lib.lib().opkb_timer_stop(self._h, C.byref(ms)))
        return ms.value

    KERNEL_CLASSES = ("other", "trial", "gram", "direction", "spg", "reduce", "elementwise",
                      "bounds", "objective")

    def profile(self, on: bool) -> None:
        check(_lib.lib().opkb_profile_enable(self._h, int(bool(on))))
        if on:
            check(_lib.lib().opkb_profile_reset(self._h))

    def profile_report(self) -> dict:
        """{kernel class: (total ms, launches)} since the profile was enabled."""
        out = {}
        for k, name in enumerate(self.KERNEL_CLASSES):
            ms, cnt = f64(), i64()
            check(_lib.lib().opkb_profile_get(self._h, k, C.byref(ms), C.byref(cnt)))
            if cnt.value:
                out[name] = (ms.value, cnt.value)
        return out

    def flush_l2(self) -> None:
        check(_lib.lib().opkb_flush_l2(self._h))

    @property
    def stream(self) -> int:
        """The cudaStream_t every kernel of this context runs on (as an integer), for callers that
        run their own kernels on the vectors (e.g. `torch.cuda.ExternalStream(ctx.stream)`)."""
        return _lib.lib().opkb_context_stream(self._h) or 0

    def wait_stream(self, user_stream: int) -> None:
        """Everything submitted so far to `user_stream` (a cudaStream_t as an integer) completes
        before anything this context launches afterwards (no host synchronisation)."""
        check(_lib.lib().opkb_context_wait_stream(self._h, vp(int(user_stream))))

    def trim(self) -> None:
        """Release the device memory of destroyed vectors that the context keeps for reuse."""
        check(_lib.lib().opkb_context_trim(self._h))

    @property
    def cached_bytes(self) -> int:
        return _lib.lib().opkb_context_cached_bytes(self._h)

    @property
    def uses_mailbox(self) -> bool:
        """True when the reduced scalars of a phase travel through the node's shared-memory
        mailbox (written by the last CTA of every rank) rather than a packed ncclAllGather."""
        return bool(_lib.lib().opkb_comm_uses_mailbox(self._h))

    def comm_init(self, unique_id: bytes, rank: int, nranks: int) -> None:
        buf = C.create_string_buffer(bytes(unique_id), 128)
        check(_lib.lib().opkb_comm_init(self._h, buf, int(rank), int(nranks)))

    def close(self) -> None:
        if self._h and getattr(self, "_owner", None) is None:
            _lib.lib().opkb_context_destroy(self._h)
        self._h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # vcreate
    def vector(self, n: int, dtype=np.float64) -> "Vector":
        return Vector(self, n, dtype)

    def from_numpy(self, a) -> "Vector":
        a = np.ascontiguousarray(a)
        v = Vector(self, a.size, a.dtype)
        v.upload(a)
        return v


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    check(_lib.lib().opkb_comm_unique_id(buf))
    return buf.raw


def default_context() -> Context:
    global _DEFAULT_CTX
    if _DEFAULT_CTX is None:
        _DEFAULT_CTX = Context()
    return _DEFAULT_CTX


class Vector:
    """A device n-vector (this rank's shard).  `vcreate(x)` -> `x.similar()`."""

    def __init__(self, ctx: Context, n: int, dtype=np.float64, _borrowed=None, _owner=None):
        self.ctx = ctx
        self.dtype = np.dtype(dtype)
        self._owner = _owner
        if _borrowed is not None:
            self._h = vp(_borrowed)
            self._own = False
        else:
            self._h = vp()
            check(_lib.lib().opkb_vector_create(ctx.handle, _eltype(dtype), int(n), C.byref(self._h)))
            self._own = True
        self.n = int(_lib.lib().opkb_vector_length(self._h))

    @classmethod
    def borrowed(cls, ctx, handle, owner) -> "Vector":
        if not handle:
            raise OpkbError(-1, "NULL vector handle")
        et = _lib.lib().opkb_vector_eltype(vp(handle))
        return cls(ctx, 0, _dtype(et), _borrowed=handle, _owner=owner)

    @property
    def handle(self):
        return self._h

    def __len__(self):
        return self.n

    @property
    def data_ptr(self) -> int:
        return int(_lib.lib().opkb_vector_data(self._h) or 0)

    def similar(self) -> "Vector":
        return Vector(self.ctx, self.n, self.dtype)

    def upload(self, a, offset: int = 0) -> "Vector":
        a = np.ascontiguousarray(a, dtype=self.dtype)
        check(_lib.lib().opkb_vector_upload(self._h, a.ctypes.data, int(offset), a.size))
        return self

    def download(self, out=None, offset: int = 0, count=None):
        count = self.n - offset if count is None else count
        if out is None:
            out = np.empty(count, dtype=self.dtype)
        check(_lib.lib().opkb_vector_download(self._h, out.ctypes.data, int(offset), int(count)))
        return out

    def fill(self, value: float) -> "Vector":
        check(_lib.lib().opkb_vector_fill(self._h, float(value)))
        return self

    def fill_random(self, seed: int, first: int, lo: float, hi: float, log_scale=False) -> "Vector":
        check(_lib.lib().opkb_vector_fill_random(self._h, int(seed), int(first), float(lo),
                                                 float(hi), int(bool(log_scale))))
        return self

    def free(self) -> None:
        # never call into the library for a vector whose context is gone
        if self._own and self._h and self.ctx is not None and self.ctx.handle:
            _lib.lib().opkb_vector_destroy(self._h)
        self._h = vp()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ---------------------------------------------------------------------------
def _bound(b, x: Vector, default: float):
    """-> (vector handle or None, scalar value)"""
    if b is None:
        return None, default
    if isinstance(b, Vector):
        return b.handle, default
    return None, float(b)


def fastmin(x, y):
    """src/bounds.jl:41: the least of x and y, or x if one of them is a NaN."""
    return y if x > y else x


def fastmax(x, y):
    """src/bounds.jl:53: the greatest of x and y, or x if one of them is a NaN."""
    return y if x < y else x


def fastclamp(x, lo, hi):
    """src/bounds.jl:65-67: x subject to the bounds (None = no bound); NaNs are not special."""
    if hi is not None:
        x = fastmin(x, hi)
    if lo is not None:
        x = fastmax(x, lo)
    return x


def vcreate(x: Vector) -> Vector:
    return x.similar()


def vcopy(x: Vector) -> Vector:
    return vcopy_(x.similar(), x)


def vcopy_(dst: Vector, src: Vector) -> Vector:
    check(_lib.lib().opkb_vcopy(dst.ctx.handle, dst.handle, src.handle))
    return dst


def vfill_(x: Vector, alpha: float) -> Vector:
    return x.fill(alpha)


def vdot(*args) -> float:
    """vdot(x, y) or vdot(w, x, y) with the free set {i : w[i] != 0}."""
    r = f64()
    if len(args) == 2:
        x, y = args
        check(_lib.lib().opkb_vdot(x.ctx.handle, x.handle, y.handle, C.byref(r)))
    else:
        w, x, y = args
        check(_lib.lib().opkb_vdot_masked(x.ctx.handle, w.handle, x.handle, y.handle, C.byref(r)))
    return r.value


def vnorm2(x: Vector) -> float:
    r = f64()
    check(_lib.lib().opkb_vnorm2(x.ctx.handle, x.handle, C.byref(r)))
    return r.value


def vnorminf(x: Vector) -> float:
    r = f64()
    check(_lib.lib().opkb_vnorminf(x.ctx.handle, x.handle, C.byref(r)))
    return r.value


def vnorm1(x: Vector) -> float:
    """doc/algebra.md:50"""
    r = f64()
    check(_lib.lib().opkb_vnorm1(x.ctx.handle, x.handle, C.byref(r)))
    return r.value


def vswap_(x: Vector, y: Vector) -> None:
    """vswap!(x, y) (doc/algebra.md:31): exchange the contents (the buffers, when both
    vectors own theirs)."""
    check(_lib.lib().opkb_vswap(x.ctx.handle, x.handle, y.handle))


def vscale_(x: Vector, alpha: float, src: Vector = None) -> Vector:
    """vscale!(x, alpha) or vscale!(dst, alpha, src) (= vcombine!(dst, alpha, src, 0, src),
    doc/algebra.md:99-101)."""
    if src is not None and src is not x:
        check(_lib.lib().opkb_vcombine(x.ctx.handle, x.handle, float(alpha), src.handle, 0.0, src.handle))
        return x
    check(_lib.lib().opkb_vscale(x.ctx.handle, x.handle, float(alpha)))
    return x


def vupdate_(y: Vector, *args) -> Vector:
    """vupdate!(y, alpha, x) or vupdate!(y, w, alpha, x) (masked by w != 0)."""
    if len(args) == 2:
        alpha, x = args
        check(_lib.lib().opkb_vupdate(y.ctx.handle, y.handle, float(alpha), x.handle))
    else:
        w, alpha, x = args
        check(_lib.lib().opkb_vupdate_masked(y.ctx.handle, y.handle, w.handle, float(alpha), x.handle))
    return y


def vcombine_(dst: Vector, alpha: float, x: Vector, beta: float, y: Vector, gamma: float = None,
              z: Vector = None) -> Vector:
    """vcombine!(dst, alpha, x, beta, y) or the three-term form (doc/algebra.md:60-67: vcombine!
    then update!, one kernel here with the same roundings)."""
    if z is not None:
        check(_lib.lib().opkb_vcombine3(dst.ctx.handle, dst.handle, float(alpha), x.handle, float(beta),
                                        y.handle, float(gamma), z.handle))
        return dst
    check(_lib.lib().opkb_vcombine(dst.ctx.handle, dst.handle, float(alpha), x.handle, float(beta),
                                   y.handle))
    return dst


def vproduct_(dst: Vector, *args) -> Vector:
    """vproduct!(dst, x, y), vproduct!(dst, src) = vproduct!(dst, dst, src), or
    vproduct!(dst, sel, x, y) with sel = {w != 0} (doc/algebra.md:103, :116)."""
    L = _lib.lib()
    if len(args) == 1:
        check(L.opkb_vproduct(dst.ctx.handle, dst.handle, dst.handle, args[0].handle))
    elif len(args) == 2:
        check(L.opkb_vproduct(dst.ctx.handle, dst.handle, args[0].handle, args[1].handle))
    else:
        w, x, y = args
        check(L.opkb_vproduct_masked(dst.ctx.handle, dst.handle, w.handle, x.handle, y.handle))
    return dst


def project_variables_(dst: Vector, src: Vector, lo=None, hi=None) -> Vector:
    lh, lv = _bound(lo, dst, -math.inf)
    hh, hv = _bound(hi, dst, math.inf)
    check(_lib.lib().opkb_project_variables(dst.ctx.handle, dst.handle, src.handle, lh, lv, hh, hv))
    return dst


def _orient(o) -> int:
    if o in ("+", 1, 1.0):
        return _lib.FORWARD
    if o in ("-", -1, -1.0):
        return _lib.BACKWARD
    if isinstance(o, (int, float)) and o != 0 and o == o:
        return _lib.FORWARD if o > 0 else _lib.BACKWARD
    raise ValueError("invalid orientation")


def project_direction_(dst: Vector, x: Vector, lo, hi, orient, d: Vector) -> Vector:
    lh, lv = _bound(lo, dst, -math.inf)
    hh, hv = _bound(hi, dst, math.inf)
    try:
        check(_lib.lib().opkb_project_direction(dst.ctx.handle, dst.handle, x.handle, lh, lv, hh, hv,
                                                _orient(orient), d.handle))
    except OpkbError as e:
        if e.code == -5:
            raise ValueError("invalid bounds") from e
        raise
    return dst


def project_gradient_(dst: Vector, x: Vector, lo, hi, d: Vector) -> Vector:
    return project_direction_(dst, x, lo, hi, "-", d)


def step_limits(x: Vector, lo, hi, orient, d: Vector):
    lh, lv = _bound(lo, x, -math.inf)
    hh, hv = _bound(hi, x, math.inf)
    smin, smax = f64(), f64()
    try:
        check(_lib.lib().opkb_step_limits(x.ctx.handle, x.handle, lh, lv, hh, hv, _orient(orient),
                                          d.handle, C.byref(smin), C.byref(smax)))
    except OpkbError as e:
        if e.code == -5:
            raise ValueError("invalid bounds") from e
        raise
    return smin.value, smax.value


def get_free_variables(*args, want_mask: bool = False):
    """`get_free_variables(gp)` (src/bounds.jl:701-714) or
    `get_free_variables(x, lo, hi, orient, d)` (:716-844): the free set as a mask
    (no index list is materialised).  Returns the count, or (count, bool mask)."""
    if len(args) == 5:
        x, lo, hi, orient, d = args
        if (lo is None or (not isinstance(lo, Vector) and float(lo) == -math.inf)) and \
                (hi is None or (not isinstance(hi, Vector) and float(hi) == math.inf)):
            # no bounds: the reference lists every variable (:762-767)
            return (x.n, np.ones(x.n, dtype=bool)) if want_mask else x.n
        # can_change(x, lo, hi, o, d) <=> projdir(x, lo, hi, o, d) != 0 (:306-311, :638-648)
        gp = project_direction_(x.similar(), x, lo, hi, orient, d)
    else:
        (gp,) = args
    cnt = i64()
    mask = np.empty(gp.n, dtype=np.uint8) if want_mask else None
    check(_lib.lib().opkb_free_variables(gp.ctx.handle, gp.handle, C.byref(cnt),
                                         mask.ctypes.data if want_mask else None))
    if want_mask:
        return cnt.value, mask.astype(bool)
    return cnt.value
