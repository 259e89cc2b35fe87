"""The non-separable objective with a neighbour exchange (SURVEY 8(f)-2,
`opkb_smooth2d_fg`): gradient bit-identical to the array expression, f to
rounding, and VMLMB / SPG trajectories on it equal to the oracle's (which runs
the numpy twin of the objective as a callback)."""
import numpy as np
import pytest

from helpers import compare_traces, record_vmlmb

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def opk():
    import opkb200
    return opkb200


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(37, 53), (1, 300), (300, 1), (2, 2), (64, 1000)])
def test_fg_matches_array_expression(opk, dtype, shape):
    rows, cols = shape
    rng = np.random.default_rng(5)
    w, y = opk.Smooth2D.synthetic(rows, cols, dtype)
    x = rng.random(rows * cols).astype(dtype)
    mu = 0.7
    ctx = opk.default_context()
    obj = opk.Smooth2D(w, y, mu, cols)
    xv, gv = ctx.from_numpy(x), ctx.vector(rows * cols, dtype)
    f = obj(xv, gv)
    g_ref = np.empty_like(x)
    f_ref = opk.Smooth2D.reference(w, y, mu, cols)(x, g_ref)
    assert np.array_equal(gv.download(), g_ref)              # element-wise: bit-exact
    assert abs(f - f_ref) <= 1e-13 * abs(f_ref)              # reduction: to rounding of a double sum
    assert np.array_equal(xv.download(), x)                  # x untouched


def test_fg_rejects_bad_shapes(opk):
    ctx = opk.default_context()
    w, y = opk.Smooth2D.synthetic(4, 5)
    with pytest.raises(ValueError):
        opk.Smooth2D(w, y, 1.0, 3)
    obj = opk.Smooth2D(w, y, 1.0, 5)
    x = ctx.from_numpy(np.zeros(20))
    with pytest.raises(opk.OpkbError):
        obj(x, x)                                            # g must not alias x
    with pytest.raises(opk.OpkbError):
        obj(ctx.from_numpy(np.zeros(15)), ctx.vector(15))    # another space than w, y


@pytest.mark.parametrize("mode", ["fused", "primitive"])
@pytest.mark.parametrize("dtype,shape,mem", [(np.float64, (40, 64), 5), (np.float64, (129, 77), 10),
                                             (np.float32, (40, 64), 5)])
def test_vmlmb_trajectory_on_smooth2d(opk, oracle, mode, dtype, shape, mem):
    rows, cols = shape
    n = rows * cols
    w, y = opk.Smooth2D.synthetic(rows, cols, dtype)
    mu = 2.0
    x0 = np.full(n, 0.5, dtype=dtype)
    lo, hi = np.zeros(n, dtype=dtype), np.ones(n, dtype=dtype)       # y exceeds 1 in places: active bounds
    xo, ro, tro = oracle.vmlmb(opk.Smooth2D.reference(w, y, mu, cols), x0, lo, hi, mem=mem, maxiter=25,
                               trace=True)
    trace = []
    s = opk.VMLMB(opk.Smooth2D(w, y, mu, cols), x0, lower=lo, upper=hi, mem=mem, maxiter=25, mode=mode,
                  observer=record_vmlmb(trace))
    s.solve()
    assert (s.iter, s.eval) == (ro["iter"], ro["eval"])
    # This well-conditioned problem is nearly solved after 20 iterations (|p| has dropped by
    # 1e-4..1e-5: g is then a difference of O(1) terms and its rounding noise reaches 1e-10
    # relative), so the 1e-10 gate is applied to the first 12 iterations, 1e-8 up to 20.
    compare_traces(trace, tro, 12, 1e-10 if dtype == np.float64 else 1e-4)
    compare_traces(trace, tro, 20, 1e-8 if dtype == np.float64 else 1e-4)
    assert 0 < np.count_nonzero(~trace[-1]["mask"]) < n              # some variables sit on a bound
    s.close()


def test_spg_on_smooth2d(opk, oracle):
    rows, cols = 48, 80
    n = rows * cols
    w, y = opk.Smooth2D.synthetic(rows, cols)
    mu = 1.5
    x0 = np.full(n, 0.5)
    lo, hi = np.zeros(n), np.ones(n)
    xo, io, _ = oracle.spg(opk.Smooth2D.reference(w, y, mu, cols), x0, 10, lo, hi, maxit=30)
    ws = opk.Info()
    xg = opk.spg(opk.Smooth2D(w, y, mu, cols), opk.BoxProjection(lo, hi), x0, 10, maxit=30, ws=ws)
    assert (ws.iter, ws.fcnt) == (io["iter"], io["fcnt"])
    assert np.linalg.norm(xg - xo) <= 1e-10 * np.linalg.norm(xo)


def test_large_grid_properties(opk):
    """2048 x 2048 (the oracle would take minutes): monotone decrease, feasibility,
    run-to-run bit reproducibility, and the solution is smoother than the data."""
    rows = cols = 2048
    n = rows * cols
    ctx = opk.default_context()
    w, y = opk.Smooth2D.synthetic(rows, cols)
    res = []
    for _ in range(2):
        fs = []
        s = opk.VMLMB(opk.Smooth2D(w, y, 4.0, cols), np.full(n, 0.5), lower=0.0, upper=1.0, mem=7,
                      maxiter=30, observer=lambda s: fs.append(s.f))
        s.solve()
        x = s.result()
        res.append((s.f, x))
        s.close()
        assert all(b <= a for a, b in zip(fs, fs[1:]))
        assert x.min() >= 0.0 and x.max() <= 1.0
    assert res[0][0] == res[1][0] and np.array_equal(res[0][1], res[1][1])
    x2, y2 = res[0][1].reshape(rows, cols), y.reshape(rows, cols)
    assert np.abs(np.diff(x2, axis=0)).mean() < 0.25 * np.abs(np.diff(y2, axis=0)).mean()
