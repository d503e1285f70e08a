"""CPU tier, world_size 2 over gloo: the hypothesis-sharded arg-max exchange.
Each rank scores its slice with the oracle (the CUDA kernels need a GPU), packs
(count, index) and the MAX all-reduce must reproduce the single-process winner."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, H, ret):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry

    pkg, O = entry.load_package(), entry.load_oracle()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    K, Kinv = O.reference_K()
    sc = O.synthetic_pair(600, seed=77)
    x = O.normalise_points(sc["px"], Kinv)
    lo, hi = pkg.sharding.shard_range(H, rank, world)
    idx = O.sample_indices(42, hi - lo, len(x), h0=lo)          # slice regenerated locally
    E = O.hypotheses(x, idx).reshape(-1, 9).astype(np.float32)
    cnt, _ = O.inlier_counts(E.astype(np.float64), x, 1e-6)
    keys = [pkg.sharding.pack_key(c, lo + i) for i, c in enumerate(cnt)]
    best = torch.tensor([max(keys)], dtype=torch.int64)
    pkg.sharding.allreduce_best(best)
    ret[rank] = pkg.sharding.unpack_key(int(best[0]))
    dist.destroy_process_group()


def test_sharded_argmax_matches_single_process():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry

    pkg, O = entry.load_package(), entry.load_oracle()
    H, world = 257, 2
    K, Kinv = O.reference_K()
    sc = O.synthetic_pair(600, seed=77)
    x = O.normalise_points(sc["px"], Kinv)
    idx = O.sample_indices(42, H, len(x))
    E = O.hypotheses(x, idx).reshape(-1, 9).astype(np.float32)
    cnt, _ = O.inlier_counts(E.astype(np.float64), x, 1e-6)
    want = (int(cnt.max()), int(np.argmax(cnt)))
    ret = mp.Manager().dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, H, ret), nprocs=world, join=True)
    assert ret[0] == ret[1] == want
