"""Maximum sizes: vectors longer than 2^31 elements (64-bit indexing in every
kernel: Tier 1, the fused phases, the shared-memory rings and their tails)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = (1 << 31) + 7 * 448 * 5 + 6      # > 2^31, whole ring tiles + a partial tile + a 2-element tail


def _free_gib():
    import torch
    import opkb200
    opkb200.default_context().trim()   # blocks the context keeps for reuse count as free
    free, total = torch.cuda.mem_get_info()
    return free / 2 ** 30


def test_primitives_beyond_int32():
    import opkb200 as opk
    if _free_gib() < 3 * N * 4 / 2 ** 30 + 4:
        pytest.skip("needs 30 GiB of free device memory")
    ctx = opk.default_context()
    x = ctx.vector(N, np.float32).fill(1.0)
    y = ctx.vector(N, np.float32).fill(2.0)
    assert opk.vdot(x, y) == 2.0 * N                       # exact in Float64
    assert opk.vnorm2(x) == pytest.approx(np.sqrt(float(N)), rel=1e-15)
    z = opk.vcombine_(x.similar(), 1, x, 1, y)
    assert opk.vnorminf(z) == 3.0
    assert np.array_equal(z.download(offset=N - 8, count=8), np.full(8, 3.0, np.float32))
    tail = np.arange(8, dtype=np.float32) - 3.0            # -3 .. 4 in the last 8 elements
    z.upload(tail, offset=N - 8)
    opk.project_variables_(z, z, 0.0, 3.5)                 # clamp to [0, 3.5]
    assert np.array_equal(z.download(offset=N - 8, count=8), np.clip(tail, 0.0, 3.5))
    assert opk.vnorminf(z) == 3.5
    assert opk.vdot(z, x) == 3.0 * (N - 8) + float(np.clip(tail, 0.0, 3.5).sum())


def test_fused_vmlmb_beyond_int32():
    import opkb200 as opk
    if _free_gib() < 14 * N * 4 / 2 ** 30 + 6:
        pytest.skip("needs 120 GiB of free device memory")
    ctx = opk.default_context()
    obj = opk.Quadratic.synthetic_device(ctx, N, np.float32, kappa=1e2)
    x0 = ctx.vector(N, np.float32).fill(0.5)
    old = os.environ.get("OPKB_CARRY")
    runs = {}
    try:
        os.environ["OPKB_CARRY"] = "2"                      # the incremental Gram also for Float32
        for name, kw in (("gram", dict(twoloop="gram")), ("sequential", dict(twoloop="sequential"))):
            fs = []
            s = opk.VMLMB(obj, opk.vcopy(x0), lower=0.0, upper=1.0, mem=2, maxiter=5, xtol=(0, 0), ftol=(0, 0),
                          gtol=(0, 0), observer=lambda s: fs.append((s.iter, s.eval, s.f, s.gnorm, s.ops.nfree())),
                          **kw)
            s.solve()
            assert opk.vnorminf(s.x) <= 1.0
            assert opk.vnorminf(opk.project_variables_(s.x.similar(), s.x, None, 0.0)) == 0.0
            assert all(b[2] < a[2] for a, b in zip(fs, fs[1:]))
            if name == "gram":
                assert s.carry_stats()[0] >= 2
            runs[name] = fs
            s.close()
            del s
    finally:
        if old is None:
            os.environ.pop("OPKB_CARRY", None)
        else:
            os.environ["OPKB_CARRY"] = old
    for a, b in zip(runs["gram"], runs["sequential"]):
        assert a[:2] == b[:2]
        assert abs(a[2] - b[2]) <= 1e-4 * abs(b[2]) and abs(a[3] - b[3]) <= 1e-3 * abs(b[3])
        assert abs(a[4] - b[4]) <= 1e-5 * N                 # size of the free set
    assert 0 < runs["gram"][-1][4] < N
