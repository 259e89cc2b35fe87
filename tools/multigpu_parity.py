#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun on N GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/multigpu_parity.py

Every rank holds a contiguous shard on its own GPU; the reductions are combined
by one packed NCCL all-gather per phase inside libopkb200.  Rank 0 compares the
gathered iterates with the single-process CPU oracle (first 20 iterations within
1e-10, identical decisions on every rank)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist

    import opk_oracle as O
    import opkb200 as opk
    from helpers import gpu_bounds, problem
    from opkb200.dist import init_context, shard_range

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = init_context(local)
    assert ctx.nranks == world and ctx.rank == rank
    if rank == 0:
        print(f"[{world} GPUs] reduced scalars travel through: "
              f"{'shared-memory mailbox' if ctx.uses_mailbox else 'packed ncclAllGather'}", flush=True)
    ok = True
    for kind, n, dtype, kw in (("quad_box", 200001, np.float64, dict(mem=10)),
                               ("rosen_box", 100000, np.float64, dict(mem=7)),
                               ("rosen_unc", 100000, np.float64, dict()),
                               ("rosen_box", 100000, np.float32, dict(mem=7)),
                               ("quad_box", 200001, np.float64, dict(mem=10, mode="primitive"))):
        pb = problem(kind, n, dtype)
        first, cnt = shard_range(n, rank, world)
        sl = slice(first, first + cnt)
        lo, hi = gpu_bounds(pb)
        lo = lo[sl] if isinstance(lo, np.ndarray) else lo
        hi = hi[sl] if isinstance(hi, np.ndarray) else hi
        obj = pb["gpu_obj"]
        if isinstance(obj, opk.Quadratic):
            obj = opk.Quadratic(obj.a[sl].copy(), obj.c[sl].copy())
        trace = []
        s = opk.VMLMB(obj, pb["x0"][sl].copy(), lower=lo, upper=hi, maxiter=25, nglobal=n, ctx=ctx,
                      observer=lambda s: trace.append((s.iter, s.eval, s.f, s.gnorm, s.stp)), **kw)
        s.solve()
        xloc = s.result()
        s.close()
        # identical decisions on all ranks
        t = torch.tensor([v for row in trace for v in row], dtype=torch.float64, device="cuda")
        tmax, tmin = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        same = bool(torch.equal(tmax, tmin))
        # gather x on rank 0
        per = shard_range(n, 0, world)[1]
        buf = torch.zeros(per, dtype=torch.float64, device="cuda")
        buf[:cnt] = torch.from_numpy(xloc.astype(np.float64)).cuda()
        bufs = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(bufs, buf)
        if rank == 0:
            x = np.concatenate([b.cpu().numpy()[:shard_range(n, r, world)[1]] for r, b in enumerate(bufs)])
            okw = {k: v for k, v in kw.items() if k != "mode"}
            xo, ro, tro = O.vmlmb(pb["oracle_obj"], pb["x0"], pb["lower"], pb["upper"], trace=True,
                                  maxiter=25, **okw)
            rtol = 1e-10 if dtype == np.float64 else 1e-4
            worst = 0.0
            for g, w in zip(trace, tro):
                if w["iter"] > 20:
                    break
                assert (g[0], g[1]) == (w["iter"], w["eval"]), (g, w["iter"], w["eval"])
                worst = max(worst, abs(g[2] - w["f"]) / abs(w["f"]), abs(g[3] - w["gnorm"]) / w["gnorm"])
            ex = np.linalg.norm(x - xo) / np.linalg.norm(xo)
            good = same and worst <= rtol and ex <= 10 * rtol
            ok = ok and good
            print(f"[{world} GPUs] {kind} n={n} {np.dtype(dtype).name} {kw}: same_decisions={same} "
                  f"worst_rel(f,|g|)={worst:.2e} rel(x)={ex:.2e} {'OK' if good else 'FAIL'}", flush=True)
    # Smooth2D: the non-separable objective, rows sharded, boundary rows exchanged over NVLink
    for dtype, (rows, cols), mem in ((np.float64, (301, 257), 7), (np.float64, (world, 1000), 3),
                                     (np.float32, (128, 96), 5)):
        n = rows * cols
        w, y = opk.Smooth2D.synthetic(rows, cols, dtype)
        r0, nr = shard_range(rows, rank, world, align=1)
        sl = slice(r0 * cols, (r0 + nr) * cols)
        x0 = np.full(n, 0.5, dtype=dtype)
        trace = []
        s = opk.VMLMB(opk.Smooth2D(w[sl].copy(), y[sl].copy(), 2.0, cols, ctx=ctx), x0[sl].copy(), lower=0.0,
                      upper=1.0, mem=mem, maxiter=25, nglobal=n, ctx=ctx,
                      observer=lambda s: trace.append((s.iter, s.eval, s.f, s.gnorm, s.stp)))
        s.solve()
        xloc = s.result()
        s.close()
        t = torch.tensor([v for row in trace for v in row], dtype=torch.float64, device="cuda")
        tmax, tmin = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        same = bool(torch.equal(tmax, tmin))
        per = shard_range(rows, 0, world, align=1)[1] * cols
        buf = torch.zeros(per, dtype=torch.float64, device="cuda")
        buf[:nr * cols] = torch.from_numpy(xloc.astype(np.float64)).cuda()
        bufs = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(bufs, buf)
        if rank == 0:
            x = np.concatenate([b.cpu().numpy()[:shard_range(rows, r, world, align=1)[1] * cols]
                                for r, b in enumerate(bufs)])
            xo, ro, tro = O.vmlmb(opk.Smooth2D.reference(w, y, 2.0, cols), x0, 0.0, 1.0, mem=mem, maxiter=25,
                                  trace=True)
            rtol = 1e-10 if dtype == np.float64 else 1e-4
            worst = 0.0
            for g, wt in zip(trace, tro):
                if wt["iter"] > 12:      # (nearly solved afterwards: see tests/test_gpu_smooth2d.py)
                    break
                assert (g[0], g[1]) == (wt["iter"], wt["eval"]), (g, wt["iter"], wt["eval"])
                worst = max(worst, abs(g[2] - wt["f"]) / abs(wt["f"]), abs(g[3] - wt["gnorm"]) / wt["gnorm"])
            ex = np.linalg.norm(x - xo) / np.linalg.norm(xo)
            good = same and worst <= rtol and ex <= 100 * rtol
            ok = ok and good
            print(f"[{world} GPUs] Smooth2D {rows}x{cols} {np.dtype(dtype).name} mem={mem} (halo exchange): "
                  f"same_decisions={same} worst_rel(f,|g|)={worst:.2e} rel(x)={ex:.2e} "
                  f"{'OK' if good else 'FAIL'}", flush=True)
    # SPG: the same sharding, one packed all-gather per function evaluation
    for kind, n, dtype, m in (("quad_box", 200001, np.float64, 10), ("rosen_box", 100000, np.float64, 10)):
        pb = problem(kind, n, dtype)
        first, cnt = shard_range(n, rank, world)
        sl = slice(first, first + cnt)
        obj = pb["gpu_obj"]
        if isinstance(obj, opk.Quadratic):
            obj = opk.Quadratic(obj.a[sl].copy(), obj.c[sl].copy())
        ws = opk.Info()
        s = opk.SPG(obj, opk.BoxProjection(pb["lower"][sl].copy(), pb["upper"][sl].copy()), pb["x0"][sl].copy(),
                    m, maxit=25, ws=ws, ctx=ctx)
        s.solve()
        xloc = s.result()
        s.close()
        per = shard_range(n, 0, world)[1]
        buf = torch.zeros(per, dtype=torch.float64, device="cuda")
        buf[:cnt] = torch.from_numpy(xloc.astype(np.float64)).cuda()
        bufs = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(bufs, buf)
        if rank == 0:
            x = np.concatenate([b.cpu().numpy()[:shard_range(n, r, world)[1]] for r, b in enumerate(bufs)])
            xo, io, _ = O.spg(pb["oracle_obj"], pb["x0"], m, pb["lower"], pb["upper"], maxit=25)
            ex = np.linalg.norm(x - xo) / np.linalg.norm(xo)
            good = (ws.iter, ws.fcnt, ws.status) == (io["iter"], io["fcnt"], io["status"]) and ex <= 1e-9
            ok = ok and good
            print(f"[{world} GPUs] SPG {kind} n={n}: iter/fcnt={ws.iter}/{ws.fcnt} rel(x)={ex:.2e} "
                  f"{'OK' if good else 'FAIL'}", flush=True)
    # conjgrad: the row-sharded stencil operator (halo exchange) and the
    # diagonal one, sums combined by the packed all-gather
    for dtype, (rows, cols), which in ((np.float64, (301, 257), "stencil"), (np.float64, (world, 1000), "stencil"),
                                       (np.float32, (128, 96), "stencil"), (np.float64, (400, 500), "diag")):
        n = rows * cols
        w, y = opk.Smooth2D.synthetic(rows, cols, dtype)
        r0, nr = shard_range(rows, rank, world, align=1)
        sl = slice(r0 * cols, (r0 + nr) * cols)
        b = (w * y).astype(dtype)
        A = (opk.Smooth2DOperator(w[sl].copy(), 2.0, cols, ctx=ctx) if which == "stencil"
             else opk.DiagonalOperator(w[sl].copy(), ctx=ctx))
        trace = []
        s = opk.ConjGrad(A, b[sl].copy(), ftol=0.0, maxiter=25, restart=10, nglobal=n, ctx=ctx, quiet=True,
                         observer=lambda s: trace.append((s.iter, s.rho, s.phi)))
        s.solve()
        xloc = s.result()
        t = torch.tensor([v for row in trace for v in row], dtype=torch.float64, device="cuda")
        tmax, tmin = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        same = bool(torch.equal(tmax, tmin))
        per = shard_range(rows, 0, world, align=1)[1] * cols
        buf = torch.zeros(per, dtype=torch.float64, device="cuda")
        buf[:nr * cols] = torch.from_numpy(xloc.astype(np.float64)).cuda()
        bufs = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(bufs, buf)
        if rank == 0:
            x = np.concatenate([bb.cpu().numpy()[:shard_range(rows, r, world, align=1)[1] * cols]
                                for r, bb in enumerate(bufs)])
            Ao = opk.Smooth2DOperator.reference(w, 2.0, cols) if which == "stencil" else ("diag", w)
            xo, info, tro = O.conjgrad(Ao, b, ftol=0.0, maxiter=25, restart=10, trace=True)
            rtol = 1e-10 if dtype == np.float64 else 1e-4
            worst = max(abs(g[1] - wt["rho"]) / wt["rho"] for g, wt in list(zip(trace, tro))[:21])
            ex = np.linalg.norm(x - xo) / np.linalg.norm(xo)
            good = same and (s.iter, s.status) == (info["iter"], info["status"]) and worst <= 100 * rtol and ex <= rtol
            ok = ok and good
            print(f"[{world} GPUs] conjgrad {which} {rows}x{cols} {np.dtype(dtype).name}: same_decisions={same} "
                  f"worst_rel(rho)={worst:.2e} rel(x)={ex:.2e} {'OK' if good else 'FAIL'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
