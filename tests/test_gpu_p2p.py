"""GPU tier: peer-memory exchange of the hypothesis-sharded estimate (csrc/mg.cu).  On one GPU the world is 1
(the rank pushes its key into its own exchange buffer): several calls exercise both slot parities and the slot
reuse, and the result must equal the plain estimate bit for bit.  With >= 2 ranks: tools/p2p_check.py under
torchrun (profiles/r01_p2p_exchange.md)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


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
