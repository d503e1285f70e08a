"""CPU tier: the oracle against itself (numpy fp64 vs the C restatement), against
synthetic ground truth, and against the reference's own host code / literals."""
import ctypes as C

import numpy as np
import pytest

fp = C.POINTER(C.c_float)
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)


def P(a, t=fp):
    return a.ctypes.data_as(t)


def test_sample_indices_scalar_vector_agree(O):
    idx = O.sample_indices(1237, 500, 1000)
    assert idx.shape == (500, 8) and idx.dtype == np.int32
    for h in (0, 1, 2, 77, 499):
        assert list(idx[h]) == O.sample_indices_one(1237, h, 1000)
    assert all(len(set(r)) == 8 for r in idx)
    assert idx.min() >= 0 and idx.max() < 1000
    # slices regenerate identically (multi-GPU sharding relies on it)
    assert np.array_equal(O.sample_indices(1237, 100, 1000, h0=200), idx[200:300])
    # minimum n
    tiny = O.sample_indices(5, 64, 8)
    assert all(sorted(r) == list(range(8)) for r in tiny)


def test_scene_is_consistent_with_ground_truth(O):
    K, Kinv = O.reference_K()
    sc = O.synthetic_pair(2000, outlier_frac=0.0, noise_px=0.0, seed=5)
    x = O.normalise_points(sc["px"], Kinv).astype(np.float64)
    # ground-truth E for x1^T E x2 = 0: (E_textbook)^T with E_textbook = [t]x R
    t, R = sc["t"], sc["R"]
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    E = (tx @ R).T
    n2, den = O.sampson_terms(E[None], x)
    assert (n2 / den).max() < 1e-9       # only fp32 rounding of the pixels is left
    # an 8-point hypothesis from clean data reproduces E up to sign/scale
    idx = O.sample_indices(3, 16, len(x))
    Eh = O.hypotheses(x, idx)
    d = O.e_distance(Eh, np.repeat(E[None], 16, 0))
    assert np.median(d) < 1e-3


def test_numpy_and_c_oracle_agree(O, oracle_c, scene_small):
    x = scene_small["x"]
    H = 300
    idx = O.sample_indices(11, H, len(x))
    E_np = O.hypotheses(x, idx)
    E_c = np.zeros((H, 9))
    oracle_c.oracle_hypotheses_f64(P(x), len(x), P(idx, ip), H, P(E_c, dp))
    assert O.e_distance(E_c, E_np).max() < 1e-9
    # counts: numpy fp64 == C fp64; C fp32 within the borderline band
    E32 = E_np.reshape(H, 9).astype(np.float32)
    cnt_np, amb_np = O.inlier_counts(E32.astype(np.float64), x, 1e-6)
    cnt64 = np.zeros(H, np.int32)
    amb64 = np.zeros(H, np.int32)
    cnt32 = np.zeros(H, np.int32)
    oracle_c.oracle_counts_f64(P(E32), H, P(x), len(x), C.c_double(1e-6), C.c_double(1e-6), P(cnt64, ip), P(amb64, ip))
    oracle_c.oracle_counts_f32(P(E32), H, P(x), len(x), C.c_float(1e-6), P(cnt32, ip))
    assert np.array_equal(cnt64, cnt_np)
    assert np.all(np.abs(cnt32 - cnt64) <= np.maximum(amb64, amb_np))
    assert oracle_c.oracle_argmax_first(P(cnt64, ip), H) == O.argmax_first(cnt_np)
    # the ground-truth E scores (nearly) all of the 70 % true inliers and none of
    # the 8-point hypotheses from 1-px-noisy narrow-FOV samples beats it
    t, R = scene_small["t"], scene_small["R"]
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    cnt_true, _ = O.inlier_counts((tx @ R).T[None], x, 1e-6)
    n_in = int((~scene_small["is_outlier"]).sum())
    assert 0.9 * n_in < cnt_true[0] < n_in + 0.05 * len(x)
    assert cnt_np.max() <= cnt_true[0]


def test_argmax_first_semantics(O, oracle_c):
    # literal of the reference's testThrust_max (SfM/sfm.cu:455-466): max 6 at position 5
    a = np.array([1, 2, 3, 4, 5, 6, 4, 1, 3], np.int32)
    assert O.argmax_first(a[:6]) == 5
    assert oracle_c.oracle_argmax_first(P(a, ip), 6) == 5
    ties = np.array([3, 9, 9, 2, 9], np.int32)
    assert O.argmax_first(ties) == 1 and oracle_c.oracle_argmax_first(P(ties, ip), 5) == 1


def test_triangulation_recovers_ground_truth(O, oracle_c):
    K, Kinv = O.reference_K()
    sc = O.synthetic_pair(500, outlier_frac=0.0, noise_px=0.0, seed=8)
    x = O.normalise_points(sc["px"], Kinv)
    M = np.eye(4)
    M[:3, :3], M[:3, 3] = sc["R"], sc["t"]
    pts = O.triangulate(x, M)
    assert np.abs(pts[:3].T - sc["X"]).max() < 5e-3      # fp32 pixel rounding only
    out = np.zeros((4, len(x)))
    oracle_c.oracle_triangulate_f64(P(x), len(x), P(np.ascontiguousarray(M), dp), P(out, dp))
    assert np.abs(out - pts).max() < 1e-6


def test_pose_candidates_correct_mode_contains_truth(O):
    K, Kinv = O.reference_K()
    sc = O.synthetic_pair(400, outlier_frac=0.0, noise_px=0.0, seed=9)
    x = O.normalise_points(sc["px"], Kinv).astype(np.float64)
    t, R = sc["t"], sc["R"]
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    E = (tx @ R).T
    Pc = O.pose_candidates(E, compat=False)
    ind, _ = O.choose_pose(x, Pc, compat=False)
    assert np.abs(Pc[ind][:3, :3] - R).max() < 1e-6
    assert np.abs(Pc[ind][:3, 3] - t).max() < 1e-6


def test_pose_candidates_compat_structure(O):
    rng = np.random.default_rng(0)
    E = O.project_essential(rng.normal(size=(3, 3)))
    Pc = O.pose_candidates(E, compat=True)
    U, _, V = O.svd_reference_orientation(E)
    for i in range(4):
        assert np.allclose(Pc[i][3], [0, 0, 0, 1])
        assert abs(abs(np.linalg.det(Pc[i][:3, :3])) - 1) < 1e-9
        s = -1.0 if i in (0, 2) else 1.0
        assert np.allclose(Pc[i][:3, 3], s * U[:, 2])
    # compat choosePose returns the inverses as a side effect
    x = np.array([[0.01, 0.02, 0.03, 0.01]] * 3)
    ind, Pinv = O.choose_pose(x, Pc, compat=True)
    assert 0 <= ind < 4
    for i in range(4):
        assert np.allclose(Pinv[i] @ Pc[i], np.eye(4), atol=1e-9)


def test_det_typo_matches_reference_host_code(O, ref_lib):
    """svd.h:337-341 compiled from the reference sources, run on the CPU."""
    rng = np.random.default_rng(1)
    for _ in range(200):
        a = rng.normal(size=9).astype(np.float32)
        got = ref_lib.ref_host_det(P(a))
        assert abs(got - O.det_reference_typo(a.astype(np.float64))) < 1e-4 * (1 + abs(got))


def test_reference_host_svd_contract(O, ref_lib):
    """The reference's own svd() (svd.h:311-335) pins the contract our svd3 keeps:
    U, V proper rotations, sorted singular values, a = u s v^T."""
    rng = np.random.default_rng(2)
    for _ in range(200):
        a = rng.normal(size=9).astype(np.float32)
        u, s, v = (np.zeros(9, np.float32) for _ in range(3))
        ref_lib.ref_host_svd(P(a), P(u), P(s), P(v))
        A, U, S, V = (m.reshape(3, 3).astype(np.float64) for m in (a, u, s, v))
        assert np.abs(U @ S @ V.T - A).max() < 1e-4
        assert abs(np.linalg.det(U) - 1) < 1e-4 and abs(np.linalg.det(V) - 1) < 1e-4
        sv = np.linalg.svd(A, compute_uv=False)
        assert np.abs(np.abs(np.diag(S)) - sv).max() < 1e-3
        Uo, So, Vo = O.svd_rot(A)
        # U V^T (the polar factor) is algorithm independent under the SO(3) contract;
        # its conditioning is 1/(s2+s3), and the reference's svd() runs 4 approximate
        # Jacobi sweeps, hence the loose tolerance
        if sv[1] + sv[2] > 0.5:
            assert np.abs(U @ V.T - Uo @ Vo.T).max() < 2e-2


def test_reference_svd_orientation_is_replayed(O, ref_lib, lib):
    """The one discrete freedom of the reference's SVD contract - the sign of v3 / u3, which ORDERS the four pose
    candidates - must come out as the reference's own host svd() (svd.h:311-335, compiled from the reference's
    sources) produces it, in the oracle's restatement and in the library's compat SVD alike."""
    rng = np.random.default_rng(3)
    n, agree_o, agree_l = 3000, 0, 0
    for t in range(n):
        Q1, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        Q2, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        E = Q1 @ np.diag([1.0, 1.0, 0.0]) @ Q2.T                   # essential matrices, as normalizeE leaves them
        if t % 3 == 1:
            E = E + 1e-3 * rng.normal(size=(3, 3))
        if t % 3 == 2:
            E = rng.normal(size=(3, 3))
        a = np.ascontiguousarray(E, np.float32).reshape(9)
        ur, sr, vr, u, s, v = (np.zeros(9, np.float32) for _ in range(6))
        ref_lib.ref_host_svd(P(a), P(ur), P(sr), P(vr))
        Ur, Vr = ur.reshape(3, 3), vr.reshape(3, 3)
        Uo, _, Vo = O.svd_reference_orientation(a.reshape(3, 3))
        agree_o += bool(Vo[:, 2] @ Vr[:, 2] > 0 and Uo[:, 2] @ Ur[:, 2] > 0)
        lib.raw("sfmb200_host_svd3_reference_orientation")(P(a), P(u), P(s), P(v))
        U, S, V = u.reshape(3, 3), s.reshape(3, 3), v.reshape(3, 3)
        agree_l += bool(V[:, 2] @ Vr[:, 2] > 0 and U[:, 2] @ Ur[:, 2] > 0)
        assert np.abs(U @ S @ V.T - a.reshape(3, 3)).max() < 1e-5
        assert abs(np.linalg.det(U.astype(np.float64)) - 1) < 1e-4 and abs(np.linalg.det(V.astype(np.float64)) - 1) < 1e-4
    assert agree_o == n and agree_l == n, (agree_o, agree_l, n)
    # ... and therefore the four candidates come in the reference's order: the oracle's candidates from the reference's
    # own (U, V) equal the oracle's candidates from its own oriented SVD, index by index
    for t in range(200):
        Q1, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        Q2, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        a = np.ascontiguousarray(Q1 @ np.diag([1.0, 1.0, 0.0]) @ Q2.T, np.float32).reshape(9)
        ur, sr, vr = (np.zeros(9, np.float32) for _ in range(3))
        ref_lib.ref_host_svd(P(a), P(ur), P(sr), P(vr))
        Ur, Vr = ur.reshape(3, 3).astype(np.float64), vr.reshape(3, 3).astype(np.float64)
        if O.det_reference_typo(Ur @ Vr.T) < 0:
            Vr = -Vr
        W = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1]], float)
        Pm = O.pose_candidates(a.reshape(3, 3).astype(np.float64), compat=True)
        for i in range(4):
            Rm = Ur @ (W if i < 2 else W.T) @ Vr.T
            sgn = -1.0 if i in (0, 2) else 1.0
            assert np.abs(Pm[i, :3, :3] - Rm.T).max() < 5e-3 and np.abs(Pm[i, :3, 3] - sgn * Ur[:, 2]).max() < 5e-3, (t, i)


def test_f32_mask_emulation_matches_c_port(O, oracle_c, scene_small):
    """oracle.sampson_mask_f32 (numpy emulation of the kernels' fp32 fma tree) against
    the C port that uses real fmaf: identical counts up to rare double-rounding ties."""
    x = scene_small["x"]
    idx = O.sample_indices(5, 64, len(x))
    E32 = O.hypotheses(x, idx).reshape(-1, 9).astype(np.float32)
    cnt32 = np.zeros(len(E32), np.int32)
    oracle_c.oracle_counts_f32(P(E32), len(E32), P(x), len(x), C.c_float(1e-6), P(cnt32, ip))
    mine = np.array([O.sampson_mask_f32(e, x, 1e-6).sum() for e in E32])
    assert np.abs(mine - cnt32).max() <= 1


def test_refit_on_inliers_improves_a_noisy_minimal_hypothesis(O):
    K, Kinv = O.reference_K()
    sc = O.synthetic_pair(4000, outlier_frac=0.3, noise_px=0.5, seed=12)
    x = O.normalise_points(sc["px"], Kinv)
    idx = O.sample_indices(9, 2000, len(x))
    E = O.hypotheses(x, idx)
    cnt, _ = O.inlier_counts(E.reshape(-1, 9).astype(np.float32).astype(np.float64), x, 1e-6)
    b = int(np.argmax(cnt))
    E1, c1, acc = O.refit_on_inliers(x, E[b].astype(np.float32), 1e-6, 6)
    assert acc >= 1 and c1 > cnt[b]
    t, R = sc["t"], sc["R"]
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    Etrue = (tx @ R).T[None]
    assert O.e_distance(E1[None], Etrue)[0] < O.e_distance(E[b][None], Etrue)[0]


def test_homography_oracle_recovers_planar_ground_truth(O):
    sc = O.planar_pair(2000, outlier_frac=0.3, noise_px=0.3, seed=3)
    x = sc["px"]
    idx = O.sample_indices(4, 500, len(x))[:, :4]
    Hc = O.homography_hypotheses(x, idx)
    cnt, _ = O.homography_counts(Hc, x, 2.0)
    b = int(np.argmax(cnt))
    Hn = Hc[b] / Hc[b][2, 2]
    assert np.abs(Hn - sc["H"]).max() < 0.05 * np.abs(sc["H"]).max()
    assert cnt[b] > 0.85 * (~sc["is_outlier"]).sum()
    # fp32 emulation of the kernel's test agrees with the fp64 count up to borderline points
    m = O.homography_mask_f32(Hc[b], x, 2.0)
    c64, amb = O.homography_counts(Hc[b][None], x, 2.0)
    assert abs(int(m.sum()) - int(c64[0])) <= int(amb[0]) + 1


def test_adaptive_termination_rule(O):
    """Restatement of the RANSAC stopping bound used to check sfmb200_estimate_e_adaptive."""
    assert O.adaptive_rounds(1000, 64, 2) == [64, 128, 256, 512, 1000]
    assert O.adaptive_rounds(64, 64, 4) == [64]
    assert O.adaptive_rounds(65, 64, 4) == [64, 65]
    b = O.adaptive_rounds(4096, 256, 2)
    # w = 0.5: needed = ln(0.01) / ln(1 - 2^-8) = 1176.6 -> first boundary >= that is 2048
    assert O.adaptive_used([np.array([50])] * len(b), 100, 0.99, b) == 2048
    # perfect data stops after the first round; no inliers never stops early
    assert O.adaptive_used([np.array([100])] * len(b), 100, 0.99, b) == 256
    assert O.adaptive_used([np.array([0])] * len(b), 100, 0.99, b) == 4096
    # a batch waits for its worst pair; improving counts stop earlier
    assert O.adaptive_used([np.array([100, 50])] * len(b), 100, 0.99, b) == 2048
    rising = [np.array([30]), np.array([50]), np.array([80]), np.array([80]), np.array([80])]
    assert O.adaptive_used(rising, 100, 0.99, b) == 1024   # w=0.8: needed 25.1, reached when the count reaches 80


def test_chain_restatement_on_exact_reconstructions(O):
    """Chaining of per-pair reconstructions built from ground truth recovers baselines, cameras and points."""
    views, n = 4, 500
    sc = O.synthetic_sequence(views, n, outlier_frac=0.0, noise_px=0.0, seed=5)
    Ms, Xs, valids = [], [], []
    for b in range(views - 1):
        Grel = sc["G"][b + 1] @ np.linalg.inv(sc["G"][b])
        bl = np.linalg.norm(Grel[:3, 3])
        M = Grel.copy()
        M[:3, 3] /= bl                                              # unit baseline, like the two-view path
        Xb = (sc["X"] @ sc["G"][b][:3, :3].T + sc["G"][b][:3, 3]) / bl   # points in camera b's frame, pair units
        Ms.append(M)
        Xs.append(np.vstack([Xb.T, np.ones(n)]))
        valids.append(np.ones(n, bool))
    scales, used = O.chain_scales(Ms, Xs, valids)
    assert np.all(used[1:] == n)
    assert np.allclose(scales[1:], sc["baselines"][1:] / sc["baselines"][:-1], rtol=1e-9)
    G, S = O.chain_cameras(Ms, scales)
    Gt = sc["G"].copy()
    Gt[:, :3, 3] /= sc["baselines"][0]
    assert np.allclose(G, Gt, atol=1e-9)
    cloud, cnt = O.chain_merge(Xs, valids, G, S)
    assert np.all(cnt == views - 1) and np.allclose(cloud[:3].T, sc["X"] / sc["baselines"][0], atol=1e-9)


def test_bundle_adjust_restatement_converges_from_a_perturbed_pose(O):
    """The fp64 restatement of the two-view bundle adjustment (the checker of csrc/bundle.cu): from a perturbed
    camera and DLT points it lowers the cost monotonically, lands near the true pose, keeps |t| = 1 and R in SO(3)."""
    K, Kinv = O.reference_K()
    n = 400
    sc = O.synthetic_pair(n, outlier_frac=0.0, noise_px=0.3, seed=8)
    x = O.normalise_points(sc["px"], Kinv).astype(np.float64)
    R, t = sc["R"], sc["t"] / np.linalg.norm(sc["t"])
    w = np.array([0.004, -0.006, 0.003])
    M = np.eye(4)
    M[:3, :3] = O._rodrigues(w) @ R
    tp = t + np.array([0.0, 0.02, -0.03])
    M[:3, 3] = tp / np.linalg.norm(tp)
    X = O.triangulate(x, M)
    act = O.ba_active(x, M, X, np.ones(n, bool))
    assert act.sum() > 0.9 * n
    r10 = O.bundle_adjust(x, M, X, act, 10)
    r60 = O.bundle_adjust(x, M, X, act, 60)
    assert r60["cost"] <= r10["cost"] <= r10["cost0"] and r60["accepted"] >= r10["accepted"] >= 1
    Mo = r60["M"]
    assert abs(np.linalg.norm(Mo[:3, 3]) - 1) < 1e-12 and np.abs(Mo[:3, :3] @ Mo[:3, :3].T - np.eye(3)).max() < 1e-9
    e0 = np.linalg.norm(M[:3, :3] - R) + np.linalg.norm(M[:3, 3] - t)
    e1 = np.linalg.norm(Mo[:3, :3] - R) + np.linalg.norm(Mo[:3, 3] - t)
    assert e1 < 0.5 * e0
    # reprojection RMS at the optimum is of the order of the noise (0.3 px in both views)
    rms_px = np.sqrt(r60["cost"] / (4 * r60["n_active"])) * 2360
    assert 0.1 < rms_px < 0.4
    # the outer loop with E from the adjusted pose keeps every true inlier
    out = O.bundle_adjust_rounds(x, M, O.essential_from_pose(M), 1e-6, 2, 30)
    assert out["inliers"] >= 0.95 * n and np.allclose(np.linalg.svd(out["E"])[1], [1, 1, 0], atol=1e-9)


def test_global_bundle_adjustment_restatement_converges(O):
    """oracle.bundle_adjust_global (the fp64 restatement of chain.cu's global adjustment): from perturbed cameras and points of
    a 4-view synthetic sequence with noisy observations it must reduce the reprojection cost monotonically and end closer to
    the ground truth than it started."""
    K, Kinv = O.reference_K()
    views, n = 4, 150
    sc = O.synthetic_sequence(views, n, outlier_frac=0.0, noise_px=0.3, seed=12)
    xs = [O.normalise_points(sc["px_pairs"][b], Kinv).astype(np.float64) for b in range(views - 1)]
    valids = [np.ones(n, bool) for _ in range(views - 1)]
    uv, obs = O.gba_observations(xs, valids)
    assert obs.all() and uv.shape == (views, n, 2)
    Gt = sc["G"].copy()
    Gt[:, :3, 3] /= sc["baselines"][0]
    Xt = (sc["X"] / sc["baselines"][0]).T
    rng = np.random.default_rng(0)
    G0 = Gt.copy()
    for k in range(1, views):
        G0[k][:3, :3] = O._rodrigues(rng.normal(size=3) * 0.01) @ Gt[k][:3, :3]
        G0[k][:3, 3] += rng.normal(size=3) * 0.02
    X0 = Xt * (1 + rng.normal(size=Xt.shape) * 0.01)
    G, X, st = O.bundle_adjust_global(uv, obs, G0, X0, iterations=15)
    assert st["accepted"] >= 5 and st["cost"] < 0.05 * st["cost_entry"]
    # compare in the gauge the adjustment keeps (|t_1| of the START value)
    s_true = np.linalg.norm(G0[1][:3, 3]) / np.linalg.norm(Gt[1][:3, 3])
    for k in range(1, views):
        assert np.linalg.norm(G[k][:3, :3] - Gt[k][:3, :3]) < 0.5 * np.linalg.norm(G0[k][:3, :3] - Gt[k][:3, :3])
        assert np.linalg.norm(G[k][:3, 3] - s_true * Gt[k][:3, 3]) < 0.5 * np.linalg.norm(G0[k][:3, 3] - Gt[k][:3, 3]) + 1e-3
    err0 = np.median(np.linalg.norm(X0 - Xt, axis=0) / np.linalg.norm(Xt, axis=0))
    err1 = np.median(np.linalg.norm(X / s_true - Xt, axis=0) / np.linalg.norm(Xt, axis=0))
    assert err1 < err0
