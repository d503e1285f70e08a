"""GPU tier: two-view bundle adjustment with inlier re-selection (csrc/bundle.cu,
sfmb200_bundle_adjust; SURVEY.md 8f rank 4, the reference's README.md:65-69 future work)
against the fp64 restatement oracle.bundle_adjust_rounds and against ground truth."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
THR = 1e-6


def _pose_err(M, R, t):
    Rm, tm = M[:3, :3], M[:3, 3] / np.linalg.norm(M[:3, 3])
    tt = t / np.linalg.norm(t)
    return float(np.linalg.norm(Rm - R)), float(min(np.linalg.norm(tm - tt), np.linalg.norm(tm + tt)))


def _prepare(pkg, O, n, seed, noise, H=4096, pairs=1, compat=0):
    import torch

    K, Kinv = O.reference_K()
    scs = [O.synthetic_pair(n, outlier_frac=0.3, noise_px=noise, seed=seed + 7 * b) for b in range(pairs)]
    px = np.stack([sc["px"] for sc in scs])
    h = pkg.BatchedPairs(K, Kinv, pairs, n, H)
    h.set_option(1, compat)
    h.set_points_xy(torch.from_numpy(px).cuda())
    h.estimate_e(H, 3, THR)
    h.refine_e(6)
    h.pose_candidates(); h.choose_pose(); h.triangulate()
    return h, scs, Kinv


def test_bundle_adjust_matches_oracle_and_ground_truth(pkg, O):
    n = 3000
    h, scs, Kinv = _prepare(pkg, O, n, 5, 1.0)
    sc = scs[0]
    x = O.normalise_points(sc["px"], Kinv)
    E0 = h.get_E()[0].astype(np.float64)
    M0 = h.get_poses()[0][int(h.get_pose_index()[0])].astype(np.float64)
    c0 = int(h.get_best()[1][0])
    e0 = _pose_err(M0, sc["R"], sc["t"])
    st = h.bundle_adjust(5, 10)[0]
    M1 = h.get_poses()[0][int(h.get_pose_index()[0])].astype(np.float64)
    E1 = h.get_E()[0].astype(np.float64)
    c1 = int(h.get_best()[1][0])
    e1 = _pose_err(M1, sc["R"], sc["t"])
    ref = O.bundle_adjust_rounds(x, M0, E0, THR, 5, 10)
    eo = _pose_err(ref["M"], sc["R"], sc["t"])
    print(f"\ninliers {c0} -> gpu {c1} / oracle {ref['inliers']}; pose error (R, t) {e0} -> gpu {e1} / oracle {eo}; "
          f"last round: active {int(st[0])}, cost {st[1]:.3e} -> {st[2]:.3e}, accepted {int(st[3])}")
    # the same iteration in fp32 and fp64 lands on the same model
    assert abs(c1 - ref["inliers"]) <= 0.01 * ref["inliers"]
    assert np.linalg.norm(M1[:3, :3] - ref["M"][:3, :3]) < 5e-3
    assert np.linalg.norm(M1[:3, 3] - ref["M"][:3, 3]) < 5e-3
    last = ref["rounds"][-1]
    assert abs(st[0] - last["n_active"]) <= 0.01 * last["n_active"]
    assert abs(st[2] - last["cost"]) <= 0.05 * last["cost"]
    # ground truth: re-selection + adjustment recovers (nearly) every true inlier and a far better pose
    assert c1 > 1.5 * c0 and c1 > 0.9 * (~sc["is_outlier"]).sum()
    assert e1[0] < 0.34 * e0[0] and e1[1] < 0.34 * e0[1]
    # state is self-consistent
    R, t = M1[:3, :3], M1[:3, 3]
    assert np.abs(R @ R.T - np.eye(3)).max() < 1e-5 and abs(np.linalg.det(R) - 1) < 1e-5
    assert abs(np.linalg.norm(t) - 1) < 1e-5 and np.allclose(M1[3], [0, 0, 0, 1])
    assert O.e_distance(E1[None], O.essential_from_pose(M1)[None])[0] < 1e-5
    assert np.allclose(np.linalg.svd(E1)[1], [1, 1, 0], atol=1e-5)
    assert c1 == int(h.get_inlier_mask().sum()) == int(O.sampson_mask_f32(E1, x, THR).sum())
    assert st[2] <= st[1] and st[6] == c1 and st[7] == 1
    # adjusted points reproject onto their observations
    X = h.get_points_host(0).astype(np.float64)
    m = O.sampson_mask_f32(E1, x, THR).astype(bool) & (X[2] > 0)
    Y = X[:3].T @ R.T + t
    r1 = X[:2].T / X[2][:, None] - x[:, :2]
    r2 = Y[:, :2] / Y[:, 2:3] - x[:, 2:]
    rms_px = np.sqrt((r1[m] ** 2 + r2[m] ** 2).sum() / (2 * m.sum())) * 2360
    print(f"reprojection rms of the inliers: {rms_px:.3f} px")
    assert np.all(np.isfinite(X)) and rms_px < 1.5
    h.close()


def test_bundle_adjust_leaves_exact_data_alone(pkg, O):
    n = 2000
    h, scs, Kinv = _prepare(pkg, O, n, 9, 0.0)
    sc = scs[0]
    M0 = h.get_poses()[0][int(h.get_pose_index()[0])].astype(np.float64)
    e0 = _pose_err(M0, sc["R"], sc["t"])
    st = h.bundle_adjust(2, 5)[0]
    M1 = h.get_poses()[0][int(h.get_pose_index()[0])].astype(np.float64)
    e1 = _pose_err(M1, sc["R"], sc["t"])
    print(f"\nnoise-free: pose error {e0} -> {e1}, cost {st[1]:.3e} -> {st[2]:.3e}")
    assert e1[0] < 1e-3 and e1[1] < 1e-3 and st[2] <= st[1] + 1e-12
    assert int(st[6]) >= 0.99 * (~sc["is_outlier"]).sum()
    h.close()


def test_bundle_adjust_batched_equals_single_and_is_deterministic(pkg, O):
    n, B = 2500, 3

    def run(pairs):
        h, scs, _ = _prepare(pkg, O, n, 31, 0.7, pairs=pairs)
        st = h.bundle_adjust(3, 8)
        out = (h.get_poses().copy(), h.get_pose_index().copy(), h.get_E().copy(),
               np.stack([h.get_points_host(b) for b in range(pairs)]), st.copy(), h.get_best()[1].copy())
        h.close()
        return out

    a, b2, one = run(B), run(B), run(1)
    for u, v in zip(a, b2):
        assert np.array_equal(u, v)                      # run to run
    for u, v in zip(a, one):
        assert np.array_equal(u[:1], v)                  # pair 0 of the batch == the same pair alone
    st = a[4]
    print(f"\nbatched BA: active {st[:, 0]}, accepted (last round) {st[:, 3]}, inliers {st[:, 6]}")
    assert np.all(st[:, 2] <= st[:, 1]) and np.all(st[:, 0] >= 8)


def test_bundle_adjust_persistent_equals_two_kernel_path(pkg, O):
    """One cooperative launch per round (grid barriers) and two launches per iteration (tickets) share their
    phases and reduction orders: same bits."""
    n = 3000

    def run(persistent, pairs):
        h, _, _ = _prepare(pkg, O, n, 13, 0.8, pairs=pairs)
        h.set_option(6, persistent)
        l0 = h.launch_count()
        st = h.bundle_adjust(3, 12)
        launches = h.launch_count() - l0
        out = (h.get_poses().copy(), h.get_E().copy(), np.stack([h.get_points_host(b) for b in range(pairs)]), st.copy())
        h.close()
        return out, launches

    for pairs in (1, 2):
        (a, la), (b, lb) = run(1, pairs), run(0, pairs)
        for u, v in zip(a, b):
            assert np.array_equal(u, v)
        assert la == 3 * 9 and lb == 3 * (8 + 2 * 12)
        assert a[3][0, 3] >= 1 and a[3][0, 2] < a[3][0, 1]          # steps were accepted, the cost went down


def test_bundle_adjust_compat_mode_and_errors(pkg, O):
    import torch

    K, Kinv = O.reference_K()
    n = 2000
    sc = O.synthetic_pair(n, noise_px=0.5, seed=77)
    h = pkg.BatchedPairs(K, Kinv, 1, n, 2048)
    h.set_points_xy(torch.from_numpy(sc["px"]).cuda())
    with pytest.raises(Exception):
        h.bundle_adjust(1, 1)                       # nothing estimated
    h.estimate_e(2048, 3, THR)
    with pytest.raises(Exception):
        h.bundle_adjust(1, 1)                       # no pose chosen
    h.pose_candidates(); h.choose_pose(); h.triangulate()
    for bad in ((0, 5), (1, 0), (65, 5), (1, 257)):
        with pytest.raises(Exception):
            h.bundle_adjust(*bad)
    # reference-compatible pose selection (cheirality of correspondence 0 only): whatever it chose,
    # the adjustment never raises the cost and leaves a finite, consistent state
    st = h.bundle_adjust(2, 6)[0]
    assert st[2] <= st[1] and np.all(np.isfinite(h.get_points_host())) and np.all(np.isfinite(h.get_E()))
    assert int(h.get_best()[1][0]) == int(h.get_inlier_mask().sum())
    h.close()
