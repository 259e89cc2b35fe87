"""The N > 1 path on CPU (gloo, world_size 2): sharded variables, per-phase
all-gather of packed reduction partials, fixed rank-order combination.  Both
ranks run the product's host driver on a test-only numpy back end and must (i)
take identical decisions and (ii) reproduce the single-rank oracle trajectory
within the north_star tolerance (the summation tree differs between 1 and 2
ranks, so agreement is to rounding, not bit-exact: SURVEY.md 8(e))."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, kind, n, kw, q):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist

    import opk_oracle as O
    import opkb200 as opk
    from cpu_backend import ShardedNumpyOps
    from helpers import gpu_bounds, problem
    from opkb200.dist import shard_range

    O.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    pb = problem(kind, n, np.float64)
    first, cnt = shard_range(n, rank, world)
    sl = slice(first, first + cnt)
    lo, hi = gpu_bounds(pb)
    lo = lo[sl] if isinstance(lo, np.ndarray) else lo
    hi = hi[sl] if isinstance(hi, np.ndarray) else hi
    obj = pb["oracle_obj"]
    if isinstance(obj, tuple):
        obj = (obj[0], obj[1][sl].copy(), obj[2][sl].copy())
    trace = []

    def obs(s):
        trace.append((s.iter, s.eval, s.f, s.gnorm, s.stp, int(s.free_mask().sum())))
    s = opk.VMLMB(pb["gpu_obj"], pb["x0"][sl].copy(), lower=lo, upper=hi, observer=obs, nglobal=n,
                  mode=lambda solver, x0: ShardedNumpyOps(solver, x0, obj), **kw)
    s.solve()
    q.put((rank, first, cnt, np.array(s.x), trace, (s.iter, s.eval, s.reason)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind,n,kw", [
    ("rosen_box", 1000, dict(mem=7, maxiter=25)),
    ("quad_box", 1003, dict(mem=10, maxiter=25)),
    ("rosen_unc", 1000, dict(maxiter=25)),
    ("rosen_box", 1000, dict(blmvm=True, maxiter=25)),
])
def test_two_ranks_match_single_rank_oracle(oracle, kind, n, kw):
    import torch.multiprocessing as mp
    from helpers import problem
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, n, kw, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # (i) identical decisions on every rank (same scalars, bit for bit)
    assert [t[:5] for t in res[0][4]] == [t[:5] for t in res[1][4]] and res[0][5] == res[1][5]
    # (ii) the single-rank oracle trajectory within tolerance
    pb = problem(kind, n, np.float64)
    xo, ro, tro = oracle.vmlmb(pb["oracle_obj"], pb["x0"], pb["lower"], pb["upper"], trace=True, **kw)
    x = np.concatenate([r[3] for r in res])
    assert sum(r[2] for r in res) == n and res[0][1] == 0 and res[1][1] == res[0][2]
    got = res[0][4]
    for g, w in zip(got, tro):
        if w["iter"] > 20:
            break
        assert (g[0], g[1]) == (w["iter"], w["eval"])
        assert abs(g[2] - w["f"]) <= 1e-10 * abs(w["f"])
        assert abs(g[3] - w["gnorm"]) <= 1e-10 * abs(w["gnorm"])
    # the active set: local free counts add up to the oracle's, iteration by iteration
    for t0, t1, w in zip(res[0][4], res[1][4], tro):
        if w["iter"] > 20:
            break
        assert t0[5] + t1[5] == int(w["mask"].sum())
    assert np.linalg.norm(x - xo) <= 1e-9 * np.linalg.norm(xo)


def _smooth_worker(rank, world, port, rows, cols, q):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist

    import opk_oracle as O
    import opkb200 as opk
    from cpu_backend import ShardedNumpyOps, ShardedSmooth2D
    from opkb200.dist import shard_range

    O.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    n = rows * cols
    w, y = opk.Smooth2D.synthetic(rows, cols)
    r0, nr = shard_range(rows, rank, world, align=1)
    sl = slice(r0 * cols, (r0 + nr) * cols)
    obj = ShardedSmooth2D(w[sl], y[sl], 2.0, cols)
    trace = []
    s = opk.VMLMB(obj, np.full(nr * cols, 0.5), lower=0.0, upper=1.0, mem=5, maxiter=25, nglobal=n,
                  observer=lambda s: trace.append((s.iter, s.eval, s.f, s.gnorm, s.stp)),
                  mode=lambda solver, x0: ShardedNumpyOps(solver, x0, obj))
    s.solve()
    q.put((rank, np.array(s.x), trace))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_halo_exchange_smooth2d(oracle):
    """The non-separable objective on two ranks: rows sharded, boundary
    rows exchanged point to point, edge between the ranks counted once."""
    import torch.multiprocessing as mp

    import opkb200 as opk
    world, rows, cols = 2, 31, 40
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_smooth_worker, args=(r, world, port, rows, cols, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][2] == res[1][2]                                   # identical decisions
    w, y = opk.Smooth2D.synthetic(rows, cols)
    xo, ro, tro = oracle.vmlmb(opk.Smooth2D.reference(w, y, 2.0, cols), np.full(rows * cols, 0.5), 0.0, 1.0,
                               mem=5, maxiter=25, trace=True)
    for g, wt in zip(res[0][2], tro):
        if wt["iter"] > 20:
            break
        assert (g[0], g[1]) == (wt["iter"], wt["eval"])
        assert abs(g[2] - wt["f"]) <= 1e-10 * abs(wt["f"])
        assert abs(g[3] - wt["gnorm"]) <= 1e-10 * abs(wt["gnorm"])
    x = np.concatenate([r[1] for r in res])
    assert np.linalg.norm(x - xo) <= 1e-9 * np.linalg.norm(xo)


def test_combine_partials_fixed_order():
    from opkb200.dist import combine_partials
    parts = [[1.0, 5.0, 3.0], [1e-17, 7.0, -2.0], [2.0, 6.0, 9.0]]
    out = combine_partials(parts, [0, 1, 2])
    assert out == [(1.0 + 1e-17) + 2.0, 7.0, -2.0]
    assert combine_partials([[1.0]], None) == [1.0]


def _cg_worker(rank, world, port, rows, cols, q):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist

    import opk_oracle as O
    import opkb200 as opk
    from cpu_backend import NumpyCgOps, ShardedSmooth2D
    from opkb200.dist import combine_partials, shard_range

    O.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    w, y = opk.Smooth2D.synthetic(rows, cols)
    r0, nr = shard_range(rows, rank, world, align=1)
    sl = slice(r0 * cols, (r0 + nr) * cols)
    stencil = ShardedSmooth2D(w[sl], np.zeros(nr * cols), 2.0, cols)     # y = 0: g = A.x

    def A(dst, src):
        stencil(src, dst)
        return dst

    def comm(vals):        # packed all-gather + fixed rank-order sum (opkb_finish_reduction)
        t = torch.from_numpy(vals)
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return np.array(combine_partials([o.tolist() for o in out], None))

    trace = []
    s = opk.ConjGrad(A, (w * y)[sl].copy(), np.zeros(nr * cols), ops=NumpyCgOps(comm), ftol=0.0, maxiter=25,
                     restart=10, quiet=True, nglobal=rows * cols,
                     observer=lambda s: trace.append((s.iter, s.rho, s.phi)))
    s.solve()
    q.put((rank, np.array(s.x), trace, (s.iter, s.status)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_conjgrad_on_the_stencil_operator(oracle):
    """`conjgrad` with the row-sharded stencil operator and its halo
    exchange on two ranks: identical decisions, the single-rank oracle's iterates."""
    import torch.multiprocessing as mp

    import opkb200 as opk
    world, rows, cols = 2, 31, 40
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cg_worker, args=(r, world, port, rows, cols, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][2] == res[1][2] and res[0][3] == res[1][3]          # identical decisions
    w, y = opk.Smooth2D.synthetic(rows, cols)
    xo, info, tro = oracle.conjgrad(opk.Smooth2DOperator.reference(w, 2.0, cols), w * y, ftol=0.0, maxiter=25,
                                    restart=10, trace=True)
    assert res[0][3] == (info["iter"], info["status"])
    for g, wt in zip(res[0][2], tro):
        assert g[0] == wt["iter"] and abs(g[1] - wt["rho"]) <= 1e-9 * wt["rho"]
    x = np.concatenate([r[1] for r in res])
    assert np.linalg.norm(x - xo) <= 1e-10 * np.linalg.norm(xo)
