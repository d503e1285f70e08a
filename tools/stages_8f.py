#!/usr/bin/env python
"""One pass through the 8f stages at config-2 size (for ncu): adaptive estimate, refit, pose, bundle adjustment
(both paths), chaining on a 3-view sequence, homography."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
pkg = entry.load_package()
S = pkg.synthetic
K, Kinv = S.reference_K()
n, H = 10000, 65536
seq = S.synthetic_sequence(3, n, seed=4321)
h = pkg.BatchedPairs(K, Kinv, 2, n, H)
h.set_option(1, 0)
h.set_points_xy(torch.from_numpy(seq["px_pairs"]).cuda())
used = h.estimate_e_adaptive(H, 11, 1e-6, 0.99, 1024, 4)
acc = h.refine_e(4)
h.pose_candidates(); h.choose_pose(); h.triangulate()
st = h.bundle_adjust(1, 5)
h.set_option(6, 0)
st2 = h.bundle_adjust(1, 2)
ch = h.chain_views()
Hm, cnt = h.find_homography(4096, 3, 0.002)
h.synchronize()
print("used", used, "refits", acc.tolist(), "ba", st[:, [0, 2, 3, 6]].tolist(), "scales", ch["scales"].tolist(), "h matches", cnt.tolist())
h.close()
