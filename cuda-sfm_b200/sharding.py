"""Host-side sharding logic for the two multi-GPU modes.  The
reference is single-GPU (no collective anywhere); this is new functionality
named by BASELINE.json's north_star.

 * hypothesis sharding (one pair, many hypotheses): rank r generates and scores
   hypotheses [lo, hi) of H_total against the replicated correspondences, then
   ONE all-reduce(MAX) of the 8-byte packed (count, index) key per pair picks the
   winner; every rank regenerates the winning E from its index (deterministic),
   so no second collective is needed.
 * pair sharding (many pairs): rank r owns pairs [lo, hi); no data-path collective.
"""
from __future__ import annotations


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [lo, hi): the first (total % world) ranks get one extra."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_key(count: int, index: int) -> int:
    """(count << 32) | (0xFFFFFFFF - index): MAX picks the highest count and, on
    ties, the LOWEST index - thrust::max_element semantics (sfm.cu:136).  Fits a
    signed int64 for count < 2^31, so torch.int64 MAX reductions are exact."""
    return (int(count) << 32) | (0xFFFFFFFF - int(index))


def unpack_key(key: int) -> tuple[int, int]:
    key = int(key)
    return key >> 32, 0xFFFFFFFF - (key & 0xFFFFFFFF)


def allreduce_best(best_i64, group=None):
    """In-place MAX all-reduce of the packed winners (torch int64 tensor [pairs])."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(best_i64, op=dist.ReduceOp.MAX, group=group)
    return best_i64


def estimate_e_sharded(handle, H_total: int, seed: int, thr: float, rank: int, world: int, d_idx=None, group=None):
    """Hypothesis-sharded estimateE on an already-ingested handle (all ranks hold
    the same correspondences).  Returns (lo, hi) this rank scored."""
    lo, hi = shard_range(H_total, rank, world)
    handle.estimate_e(hi - lo, seed, thr, d_idx=d_idx, H_total=H_total, h_begin=lo)
    if world > 1:
        # the handle enqueues on torch's current stream, and torch orders the
        # collective against that stream, so no host synchronisation is needed
        allreduce_best(handle.best_buffer(), group)
        handle.adopt_best(H_total, seed, d_idx)
    return lo, hi
