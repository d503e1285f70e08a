#!/usr/bin/env python
"""The reference's own image pair (data/dino viff.000/001, BASELINE config 1) from the committed fixture through
the whole chain: RANSAC as the reference runs it (H = N/8) and with 65,536 hypotheses, refit, pose by vote,
bundle adjustment with inlier re-selection.  Real SIFT matches, unfiltered (the reference's policy)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
pkg = entry.load_package()
K, Kinv = pkg.synthetic.reference_K()
px = np.load(os.path.join(ROOT, "tests", "golden", "dino_000_001.npz"))["px"]
n = len(px)
for H in (n // 8, 65536):
    h = pkg.BatchedPairs(K, Kinv, 1, n, 65536)
    h.set_option(1, 0)
    h.set_points_xy(torch.from_numpy(px[None]).cuda())
    h.estimate_e(H, 2019, 1e-6)
    c0 = int(h.get_best()[1][0])
    acc = int(h.refine_e(6)[0])
    c1 = int(h.get_best()[1][0])
    h.pose_candidates(); h.choose_pose(); h.triangulate()
    st = h.bundle_adjust(4, 40)[0]
    c2 = int(h.get_best()[1][0])
    X = h.get_points_host(0)
    m = h.get_inlier_mask().cpu().numpy().astype(bool)
    rms_px = float(np.sqrt(st[2] / (4 * max(st[0], 1))) * 2360)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d_px = torch.from_numpy(px[None]).cuda()
    h.run_device(d_px, H, 2019, 1e-6); torch.cuda.synchronize()
    a.record()
    for _ in range(50):
        h.run_device(d_px, H, 2019, 1e-6)
    b.record(); torch.cuda.synchronize()
    print(json.dumps(dict(n=n, H=H, inliers_ransac=c0, refits_accepted=acc, inliers_refit=c1, inliers_bundle=c2, ba_active=int(st[0]),
                          reprojection_rms_px=rms_px, points_in_front=int(((X[2] > 0) & m).sum()),
                          whole_path_ms=a.elapsed_time(b) / 50)), flush=True)
    h.close()
