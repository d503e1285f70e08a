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


def connect_peers(handle, rank: int, world: int, group=None, timeout_ms: int | None = None):
    """One-time set-up of the peer-memory exchange (csrc/mg.cu): every rank exports an opaque blob (the CUDA IPC
    handle of its exchange buffer + its device UUID), the blobs are all-gathered (plumbing, once; works over nccl
    and gloo groups) and every rank maps its peers' buffers.  sfmb200_mg_connect refuses peers it cannot reach with
    native P2P atomics."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    lib = handle.lib
    nbytes = lib.raw("sfmb200_mg_handle_bytes")()
    mine = (C.c_ubyte * nbytes)()
    lib.call("sfmb200_mg_init", handle._h, rank, world, C.cast(mine, C.c_void_p))
    if timeout_ms is not None:
        lib.call("sfmb200_mg_set_timeout_ms", handle._h, int(timeout_ms))
    on_gpu = world > 1 and dist.get_backend(group) == "nccl"
    t = torch.tensor(list(bytes(mine)), dtype=torch.uint8, device="cuda" if on_gpu else "cpu")
    table = [torch.empty_like(t) for _ in range(world)]
    if world > 1:
        dist.all_gather(table, t, group=group)
    else:
        table[0] = t
    blob = b"".join(bytes(x.cpu().numpy().tobytes()) for x in table)
    buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
    lib.call("sfmb200_mg_connect", handle._h, C.cast(buf, C.c_void_p))


def disconnect_peers(handle):
    """Unmaps the peers' buffers and frees this rank's (sfmb200_mg_close); connect_peers may be called again."""
    handle.lib.call("sfmb200_mg_close", handle._h)


def estimate_e_p2p(handle, H_total: int, seed: int, thr: float, d_idx=None):
    """Hypothesis-sharded estimateE whose exchange step runs over peer memory (NVLink P2P atomics), no collective
    call: same result as estimate_e_sharded / a single-GPU estimate over all H_total hypotheses, bit for bit."""
    import ctypes as C

    dptr = C.c_void_p(d_idx.data_ptr()) if d_idx is not None else C.c_void_p(0)
    handle.lib.call("sfmb200_estimate_e_mg", handle._h, dptr, H_total, C.c_uint64(seed), C.c_float(thr))
    handle.H = H_total


def p2p_timeouts(handle) -> int:
    import ctypes as C

    v = C.c_int32(0)
    handle.lib.call("sfmb200_mg_status", handle._h, C.byref(v))
    return v.value
