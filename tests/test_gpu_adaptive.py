"""GPU tier: RANSAC with adaptive termination (sfmb200_estimate_e_adaptive; SURVEY.md 8f
rank 2, the reference's README.md:65-69 future work).  The rounds are skipped on the
device, so the checks are: hypotheses used == the oracle's restatement of the rule applied
to the per-hypothesis counts, and the result is bit-identical to a plain estimate over the
first `used` hypotheses."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
THR = 1e-6


def _plain(h, H, seed):
    h.estimate_e(H, seed, THR)
    idx, cnt = h.get_best()
    counts = np.stack([h.get_inlier_counts(b).cpu().numpy() for b in range(h.pairs)])
    return h.get_E().copy(), idx.copy(), cnt.copy(), counts


@pytest.mark.parametrize("outliers,first,growth", [(0.3, 256, 2), (0.55, 512, 4), (0.8, 1024, 4)])
def test_adaptive_matches_rule_and_prefix_estimate(pkg, O, outliers, first, growth):
    import torch

    K, Kinv = O.reference_K()
    n, H_max, seed, conf = 5000, 32768, 11, 0.99
    sc = O.synthetic_pair(n, outlier_frac=outliers, noise_px=0.5, seed=21)
    h = pkg.BatchedPairs(K, Kinv, 1, n, H_max)
    h.set_points_xy(torch.from_numpy(sc["px"]).cuda())
    _, _, _, counts = _plain(h, H_max, seed)
    bounds = O.adaptive_rounds(H_max, first, growth)
    best = [counts[0, :b].max() for b in bounds]
    want = O.adaptive_used(best, n, float(np.float32(conf)), bounds)
    used = h.estimate_e_adaptive(H_max, seed, THR, conf, first, growth)
    E_a = h.get_E().copy()
    idx_a, cnt_a = h.get_best()
    print(f"\noutliers {outliers}: used {used} of {H_max} (rule: {want}), best count {int(cnt_a[0])}")
    assert used == want
    if outliers >= 0.8:
        assert used == H_max            # w^8 ~ 2.6e-6: the bound is far above H_max
    if outliers <= 0.3:
        assert used < H_max              # 70 % inliers: the bound is reached long before H_max
    E_p, idx_p, cnt_p, _ = _plain(h, used, seed)
    assert int(idx_a[0]) == int(idx_p[0]) and int(cnt_a[0]) == int(cnt_p[0])
    assert np.array_equal(E_a, E_p)
    # downstream stages accept the adaptive result
    h.estimate_e_adaptive(H_max, seed, THR, conf, first, growth)
    h.pose_candidates(); h.choose_pose(); h.triangulate()
    assert np.all(np.isfinite(h.get_points_host()))
    h.close()


def test_adaptive_batch_waits_for_the_hardest_pair(pkg, O):
    import torch

    K, Kinv = O.reference_K()
    n, H_max, seed, conf, first, growth = 4096, 16384, 5, 0.999, 256, 2
    fr = [0.2, 0.6, 0.4]
    px = np.stack([O.synthetic_pair(n, outlier_frac=f, noise_px=0.5, seed=60 + i)["px"] for i, f in enumerate(fr)])
    h = pkg.BatchedPairs(K, Kinv, len(fr), n, H_max)
    h.set_points_xy(torch.from_numpy(px).cuda())
    _, _, _, counts = _plain(h, H_max, seed)
    bounds = O.adaptive_rounds(H_max, first, growth)
    best = [counts[:, :b].max(axis=1) for b in bounds]
    want = O.adaptive_used(best, n, float(np.float32(conf)), bounds)
    used = h.estimate_e_adaptive(H_max, seed, THR, conf, first, growth)
    alone = [O.adaptive_used([b[i:i + 1] for b in best], n, float(np.float32(conf)), bounds) for i in range(len(fr))]
    print(f"\nbatch used {used}; each pair alone would stop at {alone}")
    assert used == want == max(alone) and min(alone) < max(alone)
    E_a = h.get_E().copy()
    idx_a, cnt_a = (v.copy() for v in h.get_best())
    E_p, idx_p, cnt_p, _ = _plain(h, used, seed)
    assert np.array_equal(idx_a, idx_p) and np.array_equal(cnt_a, cnt_p) and np.array_equal(E_a, E_p)
    h.close()


def test_adaptive_argument_errors(pkg, O):
    import torch

    K, Kinv = O.reference_K()
    h = pkg.BatchedPairs(K, Kinv, 1, 1000, 1024)
    with pytest.raises(Exception):
        h.estimate_e_adaptive(1024, 0, THR)                       # no points yet
    h.set_points_xy(torch.from_numpy(O.synthetic_pair(1000, seed=3)["px"]).cuda())
    for kw in (dict(confidence=1.0), dict(confidence=0.0), dict(growth=1), dict(first_round=0)):
        with pytest.raises(Exception):
            h.estimate_e_adaptive(1024, 0, THR, **kw)
    with pytest.raises(Exception):
        h.estimate_e_adaptive(1 << 20, 0, THR, first_round=1024, growth=2)   # a round larger than the arena
    assert h.estimate_e_adaptive(1024, 0, THR, first_round=64, growth=2) in O.adaptive_rounds(1024, 64, 2)
    with pytest.raises(Exception):
        h.get_inlier_counts()                                      # per-hypothesis getters are not available afterwards
    h.close()
