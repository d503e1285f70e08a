"""GPU tier: N-view chaining of consecutive pairs (csrc/chain.cu, sfmb200_chain_views; SURVEY.md 8f
rank 4) against the fp64 restatement (oracle.chain_*) fed with the GPU's own per-pair results, and
against the ground truth of the synthetic sequence."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
THR = 1e-6


def _reconstruct(pkg, O, views, n, seed=4321, H=16384):
    import torch

    K, Kinv = O.reference_K()
    sc = O.synthetic_sequence(views, n, seed=seed)
    h = pkg.BatchedPairs(K, Kinv, views - 1, n, H)
    h.set_option(1, 0)                                  # pose by inlier vote
    h.set_points_xy(torch.from_numpy(sc["px_pairs"]).cuda())
    h.estimate_e(H, 11, THR)
    h.refine_e(6)
    h.pose_candidates(); h.choose_pose(); h.triangulate()
    h.bundle_adjust(4, 40)
    return h, sc, Kinv


def test_chain_matches_oracle_and_ground_truth(pkg, O):
    views, n = 4, 4000
    h, sc, Kinv = _reconstruct(pkg, O, views, n)
    B = views - 1
    out = h.chain_views()
    ind = h.get_pose_index()
    Ms = [h.get_poses()[b][int(ind[b])].astype(np.float64) for b in range(B)]
    Es = h.get_E().astype(np.float64)
    Xs = [h.get_points_host(b).astype(np.float64) for b in range(B)]
    xs = [O.normalise_points(sc["px_pairs"][b], Kinv) for b in range(B)]
    valids = [O.chain_valid(xs[b], Ms[b], Xs[b], O.sampson_mask_f32(Es[b], xs[b], THR)) for b in range(B)]
    scales, used = O.chain_scales(Ms, Xs, valids)
    G, S = O.chain_cameras(Ms, scales)
    cloud, cnt = O.chain_merge(Xs, valids, G, S)
    print(f"\nscales gpu {out['scales']} oracle {scales} truth {np.r_[1, sc['baselines'][1:] / sc['baselines'][:-1]]}; used {out['used']}")
    # restatement: same links, same scales, same cameras, same cloud
    assert np.array_equal(out["used"], used)
    assert np.allclose(out["scales"], scales, rtol=3e-3)
    assert np.allclose(out["cameras"], G[:, :3, :], rtol=0, atol=3e-3 * np.abs(G[:, :3, 3]).max())
    g_cnt = out["count"].cpu().numpy()
    g_cloud = out["cloud"].cpu().numpy().astype(np.float64)
    assert np.array_equal(g_cnt, cnt)
    m = cnt > 0
    assert np.all(g_cloud[3] == 1) and np.all(g_cloud[:3, ~m] == 0)
    rel = np.linalg.norm(g_cloud[:3, m] - cloud[:3, m], axis=0) / np.linalg.norm(cloud[:3, m], axis=0)
    assert np.median(rel) < 3e-3 and rel.max() < 3e-2
    # ground truth (units of the first baseline): relative scales, camera centres, cloud
    true_ratio = sc["baselines"][1:] / sc["baselines"][:-1]
    assert np.allclose(out["scales"][1:], true_ratio, rtol=0.06)
    Gt = sc["G"].copy()
    Gt[:, :3, 3] /= sc["baselines"][0]
    for k in range(views):
        Rg, tg = out["cameras"][k][:, :3].astype(np.float64), out["cameras"][k][:, 3].astype(np.float64)
        cg, ct = -Rg.T @ tg, -Gt[k][:3, :3].T @ Gt[k][:3, 3]
        assert np.linalg.norm(Rg - Gt[k][:3, :3]) < 0.03
        assert np.linalg.norm(cg - ct) < 0.06 * max(1.0, np.linalg.norm(ct))
    Xt = (sc["X"] / sc["baselines"][0]).T
    seen_all = g_cnt == B
    err = np.linalg.norm(g_cloud[:3, seen_all] - Xt[:, seen_all], axis=0) / np.linalg.norm(Xt[:, seen_all], axis=0)
    print(f"tracks seen by all pairs: {seen_all.sum()}, median relative cloud error {np.median(err):.4f}")
    assert seen_all.sum() > 0.3 * n and np.median(err) < 0.03
    h.close()


def test_chain_single_pair_and_determinism(pkg, O):
    import torch

    views, n = 3, 3000
    h, sc, Kinv = _reconstruct(pkg, O, views, n, seed=99)
    a = h.chain_views()
    b = h.chain_views()
    for k in ("cameras", "scales", "used"):
        assert np.array_equal(a[k], b[k])
    assert torch.equal(a["cloud"], b["cloud"]) and torch.equal(a["count"], b["count"])
    assert a["scales"][0] == 1 and np.allclose(a["cameras"][0], np.eye(4)[:3])
    # camera 1 is pair 0's own camera; without a cloud the host outputs still come back
    M0 = h.get_poses()[0][int(h.get_pose_index()[0])]
    assert np.allclose(a["cameras"][1], M0[:3], atol=1e-6)
    c = h.chain_views(want_cloud=False)
    assert c["cloud"] is None and np.array_equal(c["scales"], a["scales"])
    h.close()
    # one pair: the chain is the pair itself
    K, Kinv = O.reference_K()
    s1 = O.synthetic_pair(2000, seed=3)
    h1 = pkg.BatchedPairs(K, Kinv, 1, 2000, 4096)
    h1.set_option(1, 0)
    with pytest.raises(Exception):
        h1.chain_views()                                # nothing reconstructed yet
    h1.set_points_xy(torch.from_numpy(s1["px"][None]).cuda())
    h1.estimate_e(4096, 1, THR); h1.pose_candidates(); h1.choose_pose(); h1.triangulate()
    o = h1.chain_views()
    X = h1.get_points_host(0)
    cnt = o["count"].cpu().numpy()
    assert np.array_equal(o["used"], [0]) and np.array_equal(o["scales"], [1])
    assert np.allclose(o["cloud"].cpu().numpy()[:3, cnt == 1], X[:3, cnt == 1], atol=1e-6)
    assert cnt.sum() > 0 and set(np.unique(cnt)) <= {0, 1}
    h1.close()
