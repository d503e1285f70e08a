#!/usr/bin/env python
"""Peer-memory exchange (csrc/mg.cu) against the NCCL all-reduce path, under torchrun with >= 2 ranks:
same winner, same E bits; device time of both for a config-2-sized pair and a config-3-style pair."""
import json, os, sys
import numpy as np, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = entry.load_package()
S, sh = pkg.synthetic, pkg.sharding
K, Kinv = S.reference_K()


def timed(fn, reps):
    ms = []
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms.append(float(t.item()))
    return min(ms)


for n, H, pairs in ((10000, 65536, 1), (200000, 262144, 1), (4096, 4096, 3)):
    px = np.stack([S.synthetic_pair(n, seed=1234 + b)["px"] for b in range(pairs)])
    lo, hi = sh.shard_range(H, rank, world)
    h = pkg.BatchedPairs(K, Kinv, pairs, n, hi - lo)
    h.set_points_xy(torch.from_numpy(px).cuda())
    sh.connect_peers(h, rank, world)
    sh.estimate_e_sharded(h, H, 1237, 1e-6, rank, world)
    ref = (h.get_best()[0].copy(), h.get_best()[1].copy(), h.get_E().copy())
    for _ in range(3):                      # several calls: both slot parities, slot reuse
        sh.estimate_e_p2p(h, H, 1237, 1e-6)
        got = (h.get_best()[0].copy(), h.get_best()[1].copy(), h.get_E().copy())
        assert all(np.array_equal(a, b) for a, b in zip(ref, got)), (rank, ref[:2], got[:2])
    assert sh.p2p_timeouts(h) == 0
    t_nccl = timed(lambda: sh.estimate_e_sharded(h, H, 1237, 1e-6, rank, world), 20)
    t_p2p = timed(lambda: sh.estimate_e_p2p(h, H, 1237, 1e-6), 20)
    if world > 1:
        E = torch.from_numpy(h.get_E()).cuda()
        Es = [torch.empty_like(E) for _ in range(world)]
        dist.all_gather(Es, E)
        assert all(torch.equal(Es[0], e) for e in Es)
    if rank == 0:
        print(json.dumps(dict(n=n, H=H, pairs=pairs, world=world, winner=[int(ref[0][0]), int(ref[1][0])], ms_nccl_allreduce=t_nccl,
                              ms_peer_memory=t_p2p, timeouts=0)), flush=True)
    h.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
