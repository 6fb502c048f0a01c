"""Multi-GPU layout of the detection path: streams (SDRs / recorded channels) are independent analyzers
(reference: one process per device, radiotracking/__main__.py:94-140), so they shard by stream with no
collective on the data path.  One process per GPU owns a contiguous range of streams and its own engine;
the only exchange is the host-side gather of the (small) per-stream Signal lists, in deterministic
(stream, bin, time) order -- the order `SignalMatcher` (match.py:54-82) would see from a single process.
"""
from typing import List, Optional, Sequence


def stream_range(n_streams: int, world_size: int, rank: int) -> range:
    """Contiguous, balanced share of `n_streams` for `rank` (the first `n_streams % world_size` ranks get one more)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(n_streams, world_size)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def gather_signals(local: Sequence[list], n_streams: int, group=None, dst: int = 0) -> Optional[List[list]]:
    """Collect the per-stream Signal lists of every rank on `dst`: returns `n_streams` lists in global stream
    order there, None elsewhere.  Without an initialised process group the local lists are returned as they are."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return [list(x) for x in local]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = stream_range(n_streams, world, rank)
    if len(local) != len(mine):
        raise ValueError(f"rank {rank} owns {len(mine)} streams, got {len(local)} lists")
    buckets = [None] * world if rank == dst else None
    dist.gather_object([list(x) for x in local], buckets, dst=dst, group=group)
    if rank != dst:
        return None
    out: List[list] = []
    for r in range(world):
        assert len(buckets[r]) == len(stream_range(n_streams, world, r))
        out.extend(buckets[r])
    return out
