#!/usr/bin/env python
"""BASELINE.json configs 3, 4, 5 (run directly for 1 GPU or under torchrun for N):
  c3: one pair, 1M correspondences x 1M hypotheses, HYPOTHESES sharded across ranks,
      one 8-byte all-reduce(MAX) of the packed (count, index) key, winner regenerated locally
  c4: 4,096 pairs x 4k correspondences x 4,096 hypotheses, PAIRS sharded across ranks, no collective
  c5: full path with triangulation of 1M points on each rank (replicas)
Device timing: CUDA events on the current stream, max over ranks.  One JSON line per config (rank 0)."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry

THR = 1e-6


def timed(fn, reps, world):
    ms = []
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms.append(float(t.item()))
    return ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="+", choices=["c3", "c4", "c5"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the configs (tests)")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--variant", type=int, default=-1, help="scoring kernel variant (-1 = auto)")
    ap.add_argument("--exchange", choices=["nccl", "p2p"], default="nccl",
                    help="c3: exchange of the packed winners by all_reduce(MAX) or by the library's peer-memory push (csrc/mg.cu)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = entry.load_package()
    O = pkg.synthetic          # scenes + intrinsics only: no checker code in this tool
    K, Kinv = O.reference_K()

    def emit(**kw):
        if rank == 0:
            print(json.dumps(kw, default=float), flush=True)

    if "c3" in args.which:
        n = int((1 << 20) * args.scale)
        H = int((1 << 20) * args.scale)
        sc = O.synthetic_pair(n, seed=1234)                       # replicated: every rank builds the same pair
        d_px = torch.from_numpy(sc["px"]).cuda()
        lo, hi = pkg.sharding.shard_range(H, rank, world)
        h = pkg.BatchedPairs(K, Kinv, 1, n, hi - lo)
        h.set_option(2, args.variant)
        h.set_points_xy(d_px)

        if args.exchange == "p2p":
            pkg.sharding.connect_peers(h, rank, world)

        def step():
            if args.exchange == "p2p":
                pkg.sharding.estimate_e_p2p(h, H, 1237, THR)
            else:
                pkg.sharding.estimate_e_sharded(h, H, 1237, THR, rank, world)
        step()
        ms = timed(step, args.reps, world)
        idx, cnt = h.get_best()
        E = torch.from_numpy(h.get_E()).cuda()
        if world > 1:                                                # every rank must hold the same winner
            Es = [torch.empty_like(E) for _ in range(world)]
            dist.all_gather(Es, E)
            assert all(torch.equal(Es[0], e) for e in Es), "ranks disagree on the selected E"
        t = min(ms)
        emit(config="c3", n_gpus=world, n=n, H=H, sharding="hypotheses", ms=t, ms_all=ms, evals_per_s=n * H / (t * 1e-3),
             best_index=int(idx[0]), inliers=int(cnt[0]),
             collective="one all_reduce(MAX) of 8 bytes" if args.exchange == "nccl" else "none: keys pushed into peer memory (NVLink P2P atomics)",
             plan=h.score_plan())
        h.close()

    if "c4" in args.which:
        pairs, n, H = int(4096 * args.scale), 4096, 4096
        lo, hi = pkg.sharding.shard_range(pairs, rank, world)
        mine = hi - lo
        base = [O.synthetic_pair(n, seed=500 + i)["px"] for i in range(8)]
        px = np.stack([base[(lo + b) % 8] for b in range(mine)])
        d_px = torch.from_numpy(px).cuda()
        h = pkg.BatchedPairs(K, Kinv, mine, n, H)

        def step():
            h.run_device(d_px, H, 99 + lo, THR, n=n)
        step()
        ms = timed(step, args.reps, world)
        t = min(ms)
        emit(config="c4", n_gpus=world, pairs=pairs, n=n, H=H, sharding="pairs", ms=t, ms_all=ms, pairs_per_s=pairs / (t * 1e-3),
             evals_per_s=pairs * n * H / (t * 1e-3), hypotheses_per_s_incl_everything=pairs * H / (t * 1e-3),
             inliers_first=[int(v) for v in h.get_best()[1][:4]], collective="none")
        h.close()

    if "c5" in args.which:
        n, H = int((1 << 20) * args.scale), 65536
        sc = O.synthetic_pair(n, seed=77)
        d_px = torch.from_numpy(sc["px"]).cuda()
        h = pkg.BatchedPairs(K, Kinv, 1, n, H)
        h.set_option(4, 1)

        def step():
            h.run_device(d_px, H, 1237, THR)
        step()
        h.set_option(4, 1)
        ms = timed(step, args.reps, world)
        st = h.stage_times().mean(axis=0)
        emit(config="c5", n_gpus=world, n=n, H=H, ms=min(ms), stage_ms=dict(zip(h.STAGES, st)),
             tri_points_per_s=n / (st[6] * 1e-3), tri_gbs=32 * n / (st[6] * 1e-3) / 1e9, evals_per_s_score=n * H / (st[2] * 1e-3))
        h.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
