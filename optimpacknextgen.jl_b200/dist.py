"""Multi-GPU plumbing: one process per GPU, every n-vector sharded in
contiguous blocks, no data-path communication except one small exchange of the
packed reduction partials per phase: through a shared-memory
mailbox the kernels write into when the ranks share a node, through a packed
ncclAllGather otherwise (include/opkb200.h, opkb_comm_uses_mailbox).
`torch.distributed` is used only to broadcast the NCCL unique id."""
from __future__ import annotations

from .vectors import Context, comm_unique_id


def shard_range(n: int, rank: int, nranks: int, align: int = 4):
    """Contiguous block [first, first + count) of rank `rank`.  Block sizes are
    multiples of `align` (even => Rosenbrock pairs never straddle a boundary;
    4 => 16-byte packs of Float32) except possibly the last one."""
    per = -(-n // nranks)
    per = -(-per // align) * align
    first = min(rank * per, n)
    last = min(first + per, n)
    return first, last - first


def combine_partials(per_rank, kinds=None):
    """Fixed rank-order combination of packed per-rank partials, exactly what
    `opkb_finish_reduction` does after the all-gather (kinds: 0 sum, 1 max, 2 min)."""
    nred = len(per_rank[0])
    out = list(per_rank[0])
    for r in range(1, len(per_rank)):
        for k in range(nred):
            kind = kinds[k] if kinds else 0
            b = per_rank[r][k]
            if kind == 0:
                out[k] = out[k] + b
            elif kind == 1:
                out[k] = b if b > out[k] else out[k]
            else:
                out[k] = b if b < out[k] else out[k]
    return out


def init_context(device=None) -> Context:
    """Create this rank's context and, when torch.distributed is initialised
    with more than one rank, its NCCL communicator."""
    import torch.distributed as dist

    ctx = Context(-1 if device is None else device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init(box[0], rank, world)
    return ctx
