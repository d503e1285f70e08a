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
