#!/usr/bin/env python
"""Round-2 timing record of the SURVEY 8f stages at config-2 size (a 3-view sequence = 2 pairs x 10,000 correspondences):
wall clock per call, device-synchronised on both sides (most of these calls return host results and synchronise themselves),
median of 15."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
pkg = entry.load_package()
S = pkg.synthetic
K, Kinv = S.reference_K()
n, H = 10000, 65536
seq = S.synthetic_sequence(3, n, seed=4321)
d_px = torch.from_numpy(seq["px_pairs"]).cuda()
h = pkg.BatchedPairs(K, Kinv, 2, n, H)

def timed(fn, reps=15, setup=None):
    ts = []
    for _ in range(reps):
        if setup:
            setup()
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return round(1e3 * sorted(ts)[len(ts) // 2], 4), r

out = {"pairs": 2, "n": n, "H_max": H, "unit": "ms per call, both pairs"}
h.set_points_xy(d_px)
out["estimate_e (65,536 hypotheses)"], _ = timed(lambda: h.estimate_e(H, 11, 1e-6))
out["estimate_e_adaptive (confidence 0.99; rounds 4096 x4)"], used = timed(lambda: h.estimate_e_adaptive(H, 11, 1e-6, 0.99))
out["hypotheses tried by the adaptive estimate"] = int(used)
out["refine_e (4 LO-RANSAC refits)"], acc = timed(lambda: h.refine_e(4), setup=lambda: h.estimate_e_adaptive(H, 11, 1e-6, 0.99))
def poses():
    h.pose_candidates(); h.choose_pose(); h.triangulate()
h.estimate_e_adaptive(H, 11, 1e-6, 0.99); h.refine_e(4)
out["pose_candidates + choose_pose + triangulate"], _ = timed(poses)
out["bundle_adjust (4 rounds x up to 40 LM iterations)"], st = timed(lambda: h.bundle_adjust(4, 40), reps=7, setup=poses)
out["chain_views (3 views, cloud merged)"], ch = timed(lambda: h.chain_views(), reps=7)
out["bundle_adjust_global (30 LM iterations, 3 cameras + 10,000 points)"], gst = timed(lambda: h.bundle_adjust_global(h.chain_views(), iterations=30), reps=7)
out["find_homography (10,000 loops, 5 px)"], (Hm, cnt) = timed(lambda: h.find_homography(10000, 3, 5.0))
pos = torch.empty((n, 4), device="cuda"); col = torch.empty((n, 4), device="cuda")
h.set_points_xy(d_px); h.run_device(d_px, 4096, 11, 1e-6)
out["copy_to_vbo_coloured (inlier colours)"], _ = timed(lambda: h.copy_to_vbo_coloured(pos, col, 0, 1.0, 1))
print(json.dumps(out, indent=1))
h.close()
