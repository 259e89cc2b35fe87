"""Tier-1 parity on the GPU: every LazyAlgebra / SimpleBounds primitive of the
C ABI against the CPU oracle on the same seeded inputs.  Element-wise results
are bit-exact; reductions agree to rounding of a double sum."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DTYPES = (np.float64, np.float32)
SIZES = (1, 2, 7, 1000, 100003, 1 << 20)


@pytest.fixture(scope="module")
def opk():
    import opkb200
    return opkb200


@pytest.fixture(scope="module")
def ctx(opk):
    return opk.default_context()


def _vecs(ctx, rng, n, dtype, k):
    hs = [rng.standard_normal(n).astype(dtype) for _ in range(k)]
    return hs, [ctx.from_numpy(h) for h in hs]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", SIZES)
def test_reductions(opk, ctx, oracle, dtype, n):
    rng = np.random.default_rng(n)
    (x, y, w), (vx, vy, vw) = _vecs(ctx, rng, n, dtype, 3)
    w[rng.random(n) < 0.4] = 0
    vw.upload(w)
    scale = float(np.abs(x.astype(np.float64) * y.astype(np.float64)).sum()) + 1e-300
    assert abs(opk.vdot(vx, vy) - oracle.vdot(x, y)) <= 4e-16 * scale * math.log2(n + 2)
    sel = np.flatnonzero(w != 0)
    assert abs(opk.vdot(vw, vx, vy) - oracle.vdot_sel(sel, x, y)) <= 4e-16 * scale * math.log2(n + 2)
    assert opk.vnorm2(vx) == pytest.approx(oracle.vnorm2(x), rel=1e-14)
    assert opk.vnorminf(vx) == oracle.vnorminf(x)
    # deterministic: same bits on every call
    assert opk.vdot(vx, vy) == opk.vdot(vx, vy)
    cnt, mask = opk.get_free_variables(vw, want_mask=True)
    assert cnt == sel.size and np.array_equal(np.flatnonzero(mask), oracle.get_free_variables(w))


def test_empty_vectors(opk, ctx):
    v = ctx.vector(0, np.float64)
    assert opk.vdot(v, v) == 0.0 and opk.vnorm2(v) == 0.0 and opk.get_free_variables(v) == 0
    assert v.download().size == 0


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", (1, 5, 1000, 100003))
def test_elementwise_bit_exact(opk, ctx, oracle, dtype, n):
    rng = np.random.default_rng(n + 1)
    (a, b, w), (va, vb, vw) = _vecs(ctx, rng, n, dtype, 3)
    w[rng.random(n) < 0.5] = 0
    vw.upload(w)
    sel = np.flatnonzero(w != 0)
    for alpha in (0.0, 1.0, -1.0, math.pi, -0.3):
        got = opk.vscale_(ctx.from_numpy(a), alpha).download()
        assert np.array_equal(got, oracle.vscale(a.copy(), alpha))
        got = opk.vupdate_(ctx.from_numpy(a), alpha, vb).download()
        assert np.array_equal(got, oracle.vupdate(a.copy(), alpha, b))
        got = opk.vupdate_(ctx.from_numpy(a), vw, alpha, vb).download()
        assert np.array_equal(got, oracle.vupdate_sel(a.copy(), sel, alpha, b))
    for alpha in (0.0, 1.0, -1.0, math.pi):
        for beta in (0.0, 1.0, -1.0, -math.e):
            got = opk.vcombine_(ctx.vector(n, dtype), alpha, va, beta, vb).download()
            want = oracle.vcombine(np.empty_like(a), alpha, a, beta, b)
            assert np.array_equal(got, want), (alpha, beta)
    # in place (dst aliases an operand), as the reference uses it (src/spg.jl:259)
    v = ctx.from_numpy(a)
    opk.vcombine_(v, 1, vb, -0.5, v)
    assert np.array_equal(v.download(), oracle.vcombine(np.empty_like(a), 1, b, -0.5, a))
    assert np.array_equal(opk.vproduct_(ctx.vector(n, dtype), va, vb).download(), a * b)
    assert np.array_equal(opk.vcopy(va).download(), a)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", (4, 1001, 65536 + 3))
@pytest.mark.parametrize("kinds", ["ss", "as", "sa", "aa", "ns", "sn", "nn", "na", "an"])
def test_bounds_bit_exact(opk, ctx, oracle, dtype, n, kinds):
    rng = np.random.default_rng(n)
    x = rng.uniform(-0.5, 1.5, n).astype(dtype)
    d = rng.standard_normal(n).astype(dtype)
    d[rng.random(n) < 0.1] = 0
    lo_a = rng.uniform(-0.2, 0.3, n).astype(dtype)
    hi_a = (lo_a + rng.uniform(0.0, 1.0, n)).astype(dtype)
    x[::7] = lo_a[::7]
    x[3::11] = hi_a[3::11]

    def pick(k, arr, val):
        return {"s": val, "a": arr, "n": None}[k]
    lo, hi = pick(kinds[0], lo_a, 0.25), pick(kinds[1], hi_a, 0.75)
    if kinds == "ss":
        x[::7], x[3::11] = 0.25, 0.75
    vx, vd = ctx.from_numpy(x), ctx.from_numpy(d)
    vlo = ctx.from_numpy(lo) if isinstance(lo, np.ndarray) else lo
    vhi = ctx.from_numpy(hi) if isinstance(hi, np.ndarray) else hi
    got = opk.project_variables_(ctx.vector(n, dtype), vx, vlo, vhi).download()
    assert np.array_equal(got, oracle.project_variables(np.empty_like(x), x, lo, hi))
    xp = got
    vxp = ctx.from_numpy(xp)
    for orient in (-1, +1):
        got = opk.project_direction_(ctx.vector(n, dtype), vxp, vlo, vhi, orient, vd).download()
        assert np.array_equal(got, oracle.project_direction(np.empty_like(x), xp, lo, hi, orient, d))
        assert opk.step_limits(vxp, vlo, vhi, orient, vd) == oracle.step_limits(xp, lo, hi, orient, d)
        # get_free_variables(x, lo, hi, orient, d), src/bounds.jl:716-844
        cnt, mask = opk.get_free_variables(vxp, vlo, vhi, orient, vd, want_mask=True)
        sel = oracle.get_free_variables_dir(xp, lo, hi, orient, d)
        assert cnt == sel.size and np.array_equal(np.flatnonzero(mask), sel)


def test_bounds_errors_and_nan(opk, ctx, oracle):
    x = np.array([-2, 0.5, 3, np.nan])
    v = ctx.from_numpy(x)
    out = opk.project_variables_(ctx.vector(4), v, 0.0, 1.0).download()
    assert np.array_equal(out[:3], [0, 0.5, 1]) and np.isnan(out[3])   # NaN kept (bounds.jl:41,53)
    with pytest.raises(ValueError):
        opk.project_direction_(ctx.vector(4), v, 1.0, 0.0, "-", v)
    with pytest.raises(ValueError):
        opk.step_limits(v, 1.0, 0.0, "-", v)
    with pytest.raises(ValueError):
        opk.project_direction_(ctx.vector(4), v, 0.0, 1.0, 0, v)
    with pytest.raises(opk.OpkbError):
        opk.vdot(v, ctx.vector(5))
    with pytest.raises(opk.OpkbError):
        opk.vdot(v, ctx.vector(4, np.float32))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", (2, 6, 1000, 100002))
def test_objectives(opk, ctx, oracle, dtype, n):
    import ctypes as C
    from opkb200 import _lib
    rng = np.random.default_rng(n)
    x = rng.uniform(-1.5, 1.5, n).astype(dtype)
    vx, vg = ctx.from_numpy(x), ctx.vector(n, dtype)
    r = C.c_double()
    _lib.check(_lib.lib().opkb_rosenbrock_fg(ctx.handle, vx.handle, vg.handle, C.byref(r)))
    g = np.empty_like(x)
    f = oracle.rosenbrock_fg(x, g)
    assert np.array_equal(vg.download(), g)
    assert r.value == pytest.approx(f, rel=1e-14)
    a, c = opk.Quadratic.synthetic(n, dtype, kappa=1e3)
    va, vc = ctx.from_numpy(a), ctx.from_numpy(c)
    _lib.check(_lib.lib().opkb_quadratic_fg(ctx.handle, va.handle, vc.handle, vx.handle, vg.handle, C.byref(r)))
    f = oracle.quadratic_fg(a, c, x, g)
    assert np.array_equal(vg.download(), g)
    assert r.value == pytest.approx(f, rel=1e-14)


def test_fill_random_matches_host(opk, ctx):
    n, first = 1000, 12345
    v = ctx.vector(n).fill_random(5, first, -1.0, 2.0)
    want = -1.0 + 3.0 * opk.uniform(5, first, n)
    assert np.array_equal(v.download(), want)
    v = ctx.vector(n).fill_random(3, first, 1.0, 1e4, log_scale=True)
    want = 1e4 ** opk.uniform(3, first, n)
    assert np.allclose(v.download(), want, rtol=1e-12)


@pytest.mark.parametrize("dtype", DTYPES)
def test_apply_lbfgs_variants(opk, ctx, oracle, dtype):
    """apply_lbfgs! with gamma*I, a diagonal H0 (vproduct!, src/quasinewton.jl:710-714)
    and on a free set (:747-779), one kernel per reference call, vs the oracle."""
    rng = np.random.default_rng(11)
    n, mem, m, mark = 3001, 6, 5, 2
    S = [rng.standard_normal(n).astype(dtype) for _ in range(mem)]
    Y = [(1.3 * S[k] + 0.4 * rng.standard_normal(n)).astype(dtype) for k in range(mem)]
    Y[4] = (-Y[4]).astype(dtype)                       # a pair with rho <= 0 is skipped
    v0 = rng.standard_normal(n).astype(dtype)
    diag = rng.uniform(0.5, 2.0, n).astype(dtype)
    rho = np.array([oracle.vdot(Y[k], S[k]) for k in range(mem)])
    dS, dY = [ctx.from_numpy(a) for a in S], [ctx.from_numpy(a) for a in Y]
    tol = 1e-12 if dtype == np.float64 else 2e-5
    for H0, dH0 in ((0.7, 0.7), (diag, ctx.from_numpy(diag))):
        want = v0.copy()
        modif, _ = oracle.apply_lbfgs(S, Y, rho, H0 if np.isscalar(H0) else 1.0, m, mark, want,
                                      diag=None if np.isscalar(H0) else H0)
        got = ctx.from_numpy(v0)
        alpha = [0.0] * mem
        assert opk.apply_lbfgs_(dS, dY, list(rho), dH0, m, mark, got, alpha) == modif
        assert np.linalg.norm(got.download() - want) <= tol * np.linalg.norm(want)
    w = v0.copy()
    w[rng.random(n) < 0.4] = 0
    sel = np.flatnonzero(w != 0)
    want = w.copy()
    rho2 = np.zeros(mem)
    modif, _ = oracle.apply_lbfgs_sel(S, Y, rho2, m, mark, want, sel)
    got = ctx.from_numpy(w)
    rho3 = [0.0] * mem
    assert opk.apply_lbfgs_masked_(dS, dY, rho3, m, mark, got, [0.0] * mem, ctx.from_numpy(w)) == modif
    assert np.linalg.norm(got.download() - want) <= tol * np.linalg.norm(want)
    assert np.allclose(rho3, rho2, rtol=1e-12, atol=1e-9)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", (0, 1, 5, 1000, 100003))
def test_rest_of_the_vector_space_contract(opk, ctx, dtype, n):
    """doc/algebra.md:31 vswap!, :50 vnorm1, :60-67 vcombine! with three terms (= vcombine! then
    update!: same roundings), :99-103 vscale!(dst, alpha, src) / vproduct!(dst, src), :116
    vproduct!(dst, sel, x, y): element-wise results bit-exact against the array expressions."""
    rng = np.random.default_rng(n + 7)
    x, y, z, w = (rng.standard_normal(n).astype(dtype) for _ in range(4))
    w[rng.random(n) < 0.5] = 0
    vx, vy, vz, vw = (ctx.from_numpy(v) for v in (x, y, z, w))
    T = np.dtype(dtype).type
    a, b, c = 0.37, -1.25, 2.5
    # three-term vcombine!
    dst = ctx.vector(n, dtype)
    opk.vcombine_(dst, a, vx, b, vy, c, vz)
    assert np.array_equal(dst.download(), (T(a) * x + T(b) * y) + T(c) * z)
    # vnorm1
    assert opk.vnorm1(vx) == pytest.approx(float(np.abs(x.astype(np.float64)).sum()), rel=1e-14, abs=0)
    # vproduct! on a free set, the other elements are left alone
    d0 = rng.standard_normal(n).astype(dtype)
    dst.upload(d0)
    opk.vproduct_(dst, vw, vx, vy)
    assert np.array_equal(dst.download(), np.where(w != 0, x * y, d0))
    # vproduct!(dst, src) and vscale!(dst, alpha, src)
    dst.upload(d0)
    opk.vproduct_(dst, vx)
    assert np.array_equal(dst.download(), d0 * x)
    opk.vscale_(dst, a, vx)
    assert np.array_equal(dst.download(), T(a) * x)
    # vswap!: two owned vectors (buffers exchanged) and a borrowed one (element-wise)
    opk.vswap_(vx, vy)
    assert np.array_equal(vx.download(), y) and np.array_equal(vy.download(), x)
    if n > 0:
        s = opk.VMLMB(opk.Rosenbrock(), np.zeros(2 * ((n + 1) // 2), dtype=dtype), lower=0.0, mem=2)
        g = s.ops.vector("g")                       # belongs to the workspace
        other = ctx.from_numpy(np.arange(g.n, dtype=dtype))
        g.upload(np.full(g.n, 3, dtype=dtype))
        opk.vswap_(g, other)
        assert np.array_equal(g.download(), np.arange(g.n, dtype=dtype))
        assert np.array_equal(other.download(), np.full(g.n, 3, dtype=dtype))
        s.close()
    with pytest.raises(opk.OpkbError):
        opk.vswap_(vx, ctx.vector(n + 1, dtype))
