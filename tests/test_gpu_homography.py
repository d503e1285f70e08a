"""GPU tier: homography RANSAC on the hypothesis / scoring skeleton (SURVEY.md 8f
rank 3; CudaSift FindHomography semantics) against the fp64 restatement."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
I3 = np.eye(3, dtype=np.float32)


def h_distance(Ha, Hb):
    a = Ha.reshape(len(Ha), -1).astype(np.float64)
    b = Hb.reshape(len(Hb), -1).astype(np.float64)
    a = a / np.linalg.norm(a, axis=1, keepdims=True)
    b = b / np.linalg.norm(b, axis=1, keepdims=True)
    return np.minimum(np.linalg.norm(a - b, axis=1), np.linalg.norm(a + b, axis=1))


@pytest.mark.parametrize("loops", [200, 1000, 4096])
def test_homography_against_oracle(pkg, O, loops):
    import torch

    sc = O.planar_pair(3000, outlier_frac=0.3, noise_px=0.5, seed=7)
    n, thresh, seed = len(sc["px"]), 3.0, 11
    h = pkg.BatchedPairs(I3, I3, 1, n, 4096)                 # K = I: pixel coordinates, like CudaSift
    d_px = torch.from_numpy(sc["px"]).cuda()
    h.set_points_xy(d_px)
    Hbest, matches = h.find_homography(loops, seed, thresh)
    Hc = h.get_E_candidates().cpu().numpy().reshape(loops, 3, 3)
    x = sc["px"]
    # (1) hypotheses: first 4 of the device-drawn rows of 8, fp64 4-point DLT
    idx = O.sample_indices(seed, loops, n)[:, :4]
    H64 = O.homography_hypotheses(x, idx)
    d = h_distance(Hc, H64)
    assert np.mean(d < 1e-4) >= 0.97 and np.median(d) < 1e-5, (np.mean(d < 1e-4), np.median(d))
    # (2) counts: exact vs the fp32 emulation of the same fma tree (rare double-rounding ties), banded vs fp64
    got = h.get_inlier_counts().cpu().numpy()
    sel = np.arange(0, loops, max(1, loops // 64))
    emu = np.array([O.homography_mask_f32(Hc[i], x, thresh).sum() for i in sel])
    assert np.abs(got[sel] - emu).max() <= 1
    c64, amb = O.homography_counts(Hc, x, thresh)
    assert np.all(np.abs(got - c64) <= amb + 1)
    # (3) selection: first maximum (matching.cu:1063-1068), h8 = 1 on return
    bi, bc = h.get_best()
    assert bc[0] == got.max() == matches[0] and bi[0] == int(np.argmax(got))
    assert abs(Hbest[0][2, 2] - 1) < 1e-6
    assert h_distance(Hbest[0][None], Hc[bi[0]][None])[0] < 1e-6
    # (4) ground truth: the winner maps points like the true homography; the mask separates outliers
    good = ~sc["is_outlier"]
    q = Hbest[0].astype(np.float64) @ np.stack([x[good, 0], x[good, 1], np.ones(good.sum())])
    q_true = sc["H"] @ np.stack([x[good, 0], x[good, 1], np.ones(good.sum())])
    transfer = np.hypot(q[0] / q[2] - q_true[0] / q_true[2], q[1] / q[2] - q_true[1] / q_true[2])
    if loops >= 1000:
        assert np.median(transfer) < 1.5          # pixels, from 4 points with 0.5 px noise
    mask = h.get_inlier_mask().cpu().numpy().astype(bool)
    assert mask.sum() == matches[0]
    if loops >= 1000:
        assert np.mean(mask[~sc["is_outlier"]]) > 0.9 and np.mean(mask[sc["is_outlier"]]) < 0.05
    # the pose stages refuse to run on a homography
    with pytest.raises(pkg.SfmError):
        h.pose_candidates()
    # and the handle goes back to essential-matrix work afterwards
    h.estimate_e(256, 1, 1e-6 * 2360 ** 2)
    h.pose_candidates()
    h.close()


def test_homography_batched_and_deterministic(pkg, O):
    import torch

    B, n, loops = 3, 1500, 1024
    px = np.stack([O.planar_pair(n, seed=20 + b)["px"] for b in range(B)])
    hb = pkg.BatchedPairs(I3, I3, B, n, loops)
    hb.set_points_xy(torch.from_numpy(px).cuda())
    H1, m1 = hb.find_homography(loops, 5, 2.5)
    H2, m2 = hb.find_homography(loops, 5, 2.5)
    assert np.array_equal(H1, H2) and np.array_equal(m1, m2)
    for b in range(B):
        hs = pkg.BatchedPairs(I3, I3, 1, n, loops)
        hs.set_points_xy(torch.from_numpy(px[b]).cuda())
        Hs, ms = hs.find_homography(loops, (5 + 0x632BE59BD9B4E019 * b) % (1 << 64), 2.5)
        assert np.array_equal(Hs[0], H1[b]) and ms[0] == m1[b]
        hs.close()
    with pytest.raises(pkg.SfmError):
        hb.find_homography(loops + 1, 5, 2.5)
    with pytest.raises(pkg.SfmError):
        hb.find_homography(loops, 5, 0.0)
    hb.close()


def test_homography_pixel_semantics_with_real_intrinsics(pkg, O):
    """CudaSift's FindHomography takes a PIXEL threshold and returns a pixel-to-pixel H.  A handle created with the
    reference's K (as SfM::Image_pair always is) holds K^-1-normalised coordinates: the call must still mean pixels
    (threshold thresh / f inside, K H K^-1 on return) and agree with a K = I handle fed the same pixel data."""
    import torch

    K, Kinv = O.reference_K()
    sc = O.planar_pair(3000, outlier_frac=0.3, noise_px=0.5, seed=7)
    n, loops, thresh, seed = len(sc["px"]), 2048, 3.0, 11
    d_px = torch.from_numpy(sc["px"]).cuda()
    hp = pkg.BatchedPairs(I3, I3, 1, n, loops)
    hp.set_points_xy(d_px)
    Hp, mp_ = hp.find_homography(loops, seed, thresh)
    hk = pkg.BatchedPairs(K, Kinv, 1, n, loops)
    hk.set_points_xy(d_px)
    Hk, mk = hk.find_homography(loops, seed, thresh)
    # same sample rows on both sides, same metric up to fp32 rounding of the normalisation: the consensus sets agree
    # to a handful of borderline points, nowhere near "every match is an inlier" (a 5.0 threshold in normalised units)
    assert abs(int(mk[0]) - int(mp_[0])) <= max(5, int(0.005 * n)), (mk, mp_)
    assert mk[0] < 0.8 * n
    mask_k = hk.get_inlier_mask().cpu().numpy().astype(bool)
    mask_p = hp.get_inlier_mask().cpu().numpy().astype(bool)
    assert mask_k.sum() == mk[0] and np.mean(mask_k != mask_p) < 0.01
    assert np.mean(mask_k[~sc["is_outlier"]]) > 0.9 and np.mean(mask_k[sc["is_outlier"]]) < 0.05
    # the returned H is in pixels: it transfers image-1 pixels like the true homography does
    good = ~sc["is_outlier"]
    x = sc["px"]
    hom = np.stack([x[good, 0], x[good, 1], np.ones(good.sum())])
    q, q_true = Hk[0].astype(np.float64) @ hom, sc["H"] @ hom
    transfer = np.hypot(q[0] / q[2] - q_true[0] / q_true[2], q[1] / q[2] - q_true[1] / q_true[2])
    assert np.median(transfer) < 1.5 and abs(Hk[0][2, 2] - 1) < 1e-6
    # anisotropic pixels: refused, not silently mis-scaled
    Kb = K.copy(); Kb[1, 1] *= 1.1
    hb = pkg.BatchedPairs(Kb, np.linalg.inv(Kb).astype(np.float32), 1, n, loops)
    hb.set_points_xy(d_px)
    with pytest.raises(pkg.SfmError):
        hb.find_homography(loops, seed, thresh)
    for h in (hp, hk, hb):
        h.close()


@pytest.mark.parametrize("n,H", [(1500, 300), (6000, 2048)])        # fused small-problem path / general five-launch path
def test_whole_path_after_homography_is_an_essential_estimate_again(pkg, O, n, H):
    """find_homography leaves a homography in the handle (the pose stages refuse it); a whole-path run afterwards must put the
    handle back into the essential-matrix state - inlier colouring, refit and the pose getters work, and the result equals a
    fresh handle's bit for bit.  (The general path once kept the stale model flag.)"""
    import torch

    K, Kinv = O.reference_K()
    px = O.synthetic_pair(n, seed=31)["px"]
    d_px = torch.from_numpy(px[None]).cuda()
    fresh = pkg.BatchedPairs(K, Kinv, 1, n, max(H, 1024))
    fresh.run_device(d_px, H, 9, 1e-6)
    h = pkg.BatchedPairs(K, Kinv, 1, n, max(H, 1024))
    h.set_points_xy(d_px)
    h.find_homography(1024, 3, 5.0)
    with pytest.raises(pkg.SfmError):
        h.pose_candidates()
    h.run_device(d_px, H, 9, 1e-6)
    assert np.array_equal(h.get_E(), fresh.get_E()) and np.array_equal(h.get_points_host(0), fresh.get_points_host(0))
    assert np.array_equal(h.get_best()[1], fresh.get_best()[1]) and np.array_equal(h.get_pose_index(), fresh.get_pose_index())
    pos = torch.empty((n, 4), device="cuda")
    col = torch.empty((n, 4), device="cuda")
    h.copy_to_vbo_coloured(pos, col, 0, 1.0, 1)            # needs an essential matrix: raised SFMB200_ERR_STATE before the fix
    green = int((col[:, 1] == 1.0).sum().item())
    assert green == int(h.get_best()[1][0])
    h.refine_e(2)
    h.close(); fresh.close()
