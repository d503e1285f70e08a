"""GPU tier: peer-memory exchange of the hypothesis-sharded estimate (csrc/mg.cu).  On one GPU the world is 1
(the rank pushes its key into its own exchange buffer): several calls exercise both slot parities and the slot
reuse, and the result must equal the plain estimate bit for bit.  Two ranks: two processes (one per GPU when the box
has two, else both on cuda:0 - CUDA IPC and system-scope atomics work between processes on one device too), incl. the
lost-peer case: bounded wait, poisoned exchange, error on the next call, reconnect."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _two_rank_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import __graft_entry__ as entry

    try:
        pkg = entry.load_package()
        sh = pkg.sharding
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        torch.cuda.set_device(rank % torch.cuda.device_count())
        K, Kinv = pkg.synthetic.reference_K()
        n, H, pairs = 3000, 5000, 2                       # 5000 hypotheses: uneven slices would need H % world != 0 -> use 5001 below too
        px = np.stack([pkg.synthetic.synthetic_pair(n, seed=60 + b)["px"] for b in range(pairs)])
        d_px = torch.from_numpy(px).cuda()
        plain = pkg.BatchedPairs(K, Kinv, pairs, n, 5001)
        plain.set_points_xy(d_px)
        h = pkg.BatchedPairs(K, Kinv, pairs, n, 5001)
        h.set_points_xy(d_px)
        sh.connect_peers(h, rank, world, timeout_ms=20000)
        out = {}
        ok = True
        for call, (Ht, seed) in enumerate(((H, 5), (5001, 6), (H, 5), (777, 9))):
            plain.estimate_e(Ht, seed, 1e-6)
            want = (plain.get_best()[0].copy(), plain.get_best()[1].copy(), plain.get_E().copy())
            sh.estimate_e_p2p(h, Ht, seed, 1e-6)
            got = (h.get_best()[0], h.get_best()[1], h.get_E())
            ok = ok and all(np.array_equal(a, b) for a, b in zip(want, got))
        out["healthy_equal"] = bool(ok)
        out["healthy_timeouts"] = sh.p2p_timeouts(h)
        dist.barrier()
        # ---- lost peer: rank 1 sits this call out ----
        h.lib.call("sfmb200_mg_set_timeout_ms", h._h, 300)
        if rank == 0:
            sh.estimate_e_p2p(h, H, 5, 1e-6)
            out["lost_timeouts"] = sh.p2p_timeouts(h)          # synchronises: the wait has expired
            out["lost_count"] = [int(v) for v in h.get_best()[1]]
            try:
                sh.estimate_e_p2p(h, H, 5, 1e-6)
                out["next_call_raises"] = False
            except pkg.SfmError:
                out["next_call_raises"] = True
        dist.barrier()
        if rank == 1:
            # the late call still finds rank 0's key of that call: complete and correct here
            plain.estimate_e(H, 5, 1e-6)
            sh.estimate_e_p2p(h, H, 5, 1e-6)
            out["late_equal"] = bool(np.array_equal(plain.get_best()[1], h.get_best()[1]) and np.array_equal(plain.get_E(), h.get_E()))
            # ... but rank 0 is poisoned and silent from now on: the next call times out here as well
            sh.estimate_e_p2p(h, H, 6, 1e-6)
            out["lost_timeouts"] = sh.p2p_timeouts(h)
        dist.barrier()
        # ---- reconnect on every rank: healthy again ----
        sh.disconnect_peers(h)
        dist.barrier()
        sh.connect_peers(h, rank, world, timeout_ms=20000)
        plain.estimate_e(H, 11, 1e-6)
        sh.estimate_e_p2p(h, H, 11, 1e-6)
        out["reconnected_equal"] = bool(np.array_equal(plain.get_best()[0], h.get_best()[0]) and np.array_equal(plain.get_E(), h.get_E()))
        out["reconnected_timeouts"] = sh.p2p_timeouts(h)
        dist.barrier()
        h.close()
        plain.close()
        ret[rank] = out
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        import traceback

        ret[rank] = {"error": f"{e!r}\n{traceback.format_exc()}"}
        raise


def test_peer_exchange_two_ranks_and_lost_peer():
    import torch
    import torch.multiprocessing as mp

    assert torch.cuda.is_available()
    ret = mp.Manager().dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_two_rank_worker, args=(2, port, ret), nprocs=2, join=True)
    r0, r1 = ret[0], ret[1]
    assert "error" not in r0 and "error" not in r1, (r0, r1)
    for r in (r0, r1):
        assert r["healthy_equal"] and r["healthy_timeouts"] == 0
        assert r["reconnected_equal"] and r["reconnected_timeouts"] == 0
    assert r0["lost_timeouts"] == 1 and r0["lost_count"] == [0, 0] and r0["next_call_raises"]
    assert r1["late_equal"] and r1["lost_timeouts"] == 1


def test_peer_exchange_world1_equals_plain_estimate(pkg, O):
    import torch

    K, Kinv = O.reference_K()
    sh = pkg.sharding
    n, H, pairs = 3000, 4096, 3
    px = np.stack([O.synthetic_pair(n, seed=90 + b)["px"] for b in range(pairs)])
    h = pkg.BatchedPairs(K, Kinv, pairs, n, H)
    h.set_points_xy(torch.from_numpy(px).cuda())
    with pytest.raises(Exception):
        sh.estimate_e_p2p(h, H, 5, 1e-6)                 # before the peers are connected
    h.estimate_e(H, 5, 1e-6)
    ref = (h.get_best()[0].copy(), h.get_best()[1].copy(), h.get_E().copy())
    sh.connect_peers(h, 0, 1)
    for seed in (5, 5, 6, 5, 5):
        sh.estimate_e_p2p(h, H, seed, 1e-6)
        if seed == 5:
            got = (h.get_best()[0], h.get_best()[1], h.get_E())
            assert all(np.array_equal(a, b) for a, b in zip(ref, got))
    assert sh.p2p_timeouts(h) == 0
    h.pose_candidates(); h.choose_pose(); h.triangulate()   # downstream stages run on the exchanged winner
    assert np.all(np.isfinite(h.get_points_host(0)))
    with pytest.raises(Exception):
        sh.connect_peers(h, 0, 1)                        # twice
    h.close()
