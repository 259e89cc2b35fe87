"""ONE process, several GPUs (`opkb_group_*`, `opk.Group`, `opk.MultiVMLMB`, `vmlmb(..., ngpus=G)`;
SURVEY 8(e) "One process, G devices").  The ranks of a group are contexts of one process driven by
one host thread each inside the library; when the box has a single GPU the ranks share device 0 --
the code path (in-process mailbox, rank-order combination, threads) is the same."""
import numpy as np
import pytest

from helpers import gpu_bounds, problem, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def opk():
    import opkb200
    return opkb200


def _devices(G):
    import torch
    have = torch.cuda.device_count()
    return [r % have for r in range(G)]


@pytest.mark.parametrize("kind,n,dtype,kw", [
    ("quad_box", 200001, np.float64, dict(mem=10)),
    ("rosen_box", 100000, np.float64, dict(mem=7)),
    ("rosen_unc", 100000, np.float64, dict()),
    ("rosen_box", 100000, np.float32, dict(mem=7)),
])
@pytest.mark.parametrize("G", [2, 3])
def test_group_follows_the_oracle(opk, oracle, kind, n, dtype, kw, G):
    pb = problem(kind, n, dtype)
    xo, ro, tro = oracle.vmlmb(pb["oracle_obj"], pb["x0"], pb["lower"], pb["upper"], trace=True, maxiter=25, **kw)
    lo, hi = gpu_bounds(pb)
    rtol = 1e-10 if dtype == np.float64 else 1e-4
    by_iter = {t["iter"]: t for t in tro}
    with opk.Group(devices=_devices(G)) as grp:
        assert grp.size == G and [c.rank for c in grp.contexts] == list(range(G))
        s = opk.MultiVMLMB(pb["gpu_obj"], pb["x0"], group=grp, lower=lo, upper=hi, maxiter=25, **kw)
        seen = 0
        alive = True
        while alive:
            alive = s.iterate(1)
            w = by_iter.get(s.iter)
            if w is None or s.iter > 20:
                continue
            assert (s.eval, s.rejects) == (w["eval"], w["rejects"]), s.iter
            assert abs(s.f - w["f"]) <= rtol * abs(w["f"]) and abs(s.gnorm - w["gnorm"]) <= rtol * abs(w["gnorm"])
            assert relerr(s.result(), w["x"]) <= rtol, s.iter
            seen += 1
        assert seen >= 15
        assert (s.iter, s.eval, s.rejects, s.reason) == (ro["iter"], ro["eval"], ro["rejects"], ro["reason"])
        assert relerr(s.result(), xo) <= 10 * rtol
        s.close()


def test_vmlmb_ngpus_keyword_and_one_rank_group(opk):
    """`vmlmb(fg, x0; ..., ngpus=G)` from one process; a group of ONE rank is the plain solver, bit for bit."""
    n = 50001
    a, c = opk.Quadratic.synthetic(n, np.float64, kappa=1e2)
    kw = dict(lower=np.zeros(n), upper=np.ones(n), mem=5, maxiter=30)
    x1 = opk.vmlmb(opk.Quadratic(a, c), np.full(n, 0.5), driver="native", **kw)
    with opk.Group(devices=[0]) as g1:
        x1g = opk.vmlmb(opk.Quadratic(a, c), np.full(n, 0.5), group=g1, **kw)
    assert np.array_equal(x1, x1g)
    x2 = opk.vmlmb(opk.Quadratic(a, c), np.full(n, 0.5), group=opk.Group(devices=_devices(2)), **kw)
    assert relerr(x2, x1) <= 1e-10


def test_group_with_per_rank_device_callbacks(opk):
    """A separable user objective on a group: every rank's DeviceObjective returns ITS share of f
    (`partial=True`), the library sums the shares in rank order (opkb_solver_set_fg_callback)."""
    import torch
    from test_gpu_api import _wrap
    n = 60000
    target = np.linspace(-1, 2, n)

    def make_fg(rank, ctx, first, count):
        dev = ctx.device
        t = torch.from_numpy(target[first:first + count]).to("cuda:%d" % dev)

        def fn(nn, xp, gp, stream):
            with torch.cuda.device(dev), torch.cuda.stream(torch.cuda.ExternalStream(stream, device=dev)):
                xt, gt = _wrap(xp, nn), _wrap(gp, nn)
                gt.copy_(2 * xt - 2 * t)
                return float(((xt - t) ** 2).sum().item())
        return opk.DeviceObjective(fn, partial=True)

    with opk.Group(devices=_devices(2)) as grp:
        s = opk.MultiVMLMB(make_fg, np.full(n, 0.5), group=grp, lower=opk.BoundedSet(0.0, 1.0), mem=5)
        x = s.solve()
        assert np.abs(x - np.clip(target, 0, 1)).max() < 1e-5 and s.stage == 3
        s.close()


def test_group_with_an_empty_shard(opk, oracle):
    """A rank whose shard is EMPTY (n = 8 over 3 ranks in blocks of 4: 4 + 4 + 0) runs the same kernels on
    no element, reduces the same slots and takes the same decisions as the others (the carry
    decision must not depend on rank-local quantities)."""
    pb = problem("quad_box", 8, np.float64)
    xo, ro, _ = oracle.vmlmb(pb["oracle_obj"], pb["x0"], pb["lower"], pb["upper"], mem=3, maxiter=12)
    lo, hi = gpu_bounds(pb)
    with opk.Group(devices=_devices(3)) as grp:
        s = opk.MultiVMLMB(pb["gpu_obj"], pb["x0"], group=grp, lower=lo, upper=hi, mem=3, maxiter=12)
        assert [cnt for _, cnt in s.ranges] == [4, 4, 0]
        x = s.solve()
        assert (s.iter, s.eval, s.rejects, s.reason) == (ro["iter"], ro["eval"], ro["rejects"], ro["reason"])
        assert relerr(x, xo) <= 1e-10 and abs(s.f - ro["f"]) <= 1e-10 * abs(ro["f"])
        s.close()
