#!/usr/bin/env python
"""Device time of sfmb200_bundle_adjust (CUDA events): one pair of 10,000 correspondences, one pair of
1M correspondences and a batch of 512 pairs x 4,096.  Per LM iteration the accumulate + update kernels read
2 x (16 B correspondence + 12 B point + 1 B flag) and write 12 B per correspondence."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry

pkg = entry.load_package()
S = pkg.synthetic
K, Kinv = S.reference_K()


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for pairs, n, H, reps in ((1, 10000, 65536, 20), (1, 1 << 20, 4096, 5), (512, 4096, 4096, 3)):
    base = [S.synthetic_pair(n, seed=900 + i)["px"] for i in range(min(pairs, 4))]
    px = np.stack([base[b % len(base)] for b in range(pairs)])
    h = pkg.BatchedPairs(K, Kinv, pairs, n, H)
    h.set_option(1, 0)
    d_px = torch.from_numpy(px).cuda()
    h.run_device(d_px, H, 1237, 1e-6)
    h.refine_e(4); h.pose_candidates(); h.choose_pose(); h.triangulate()
    C = __import__("ctypes")
    res = {}
    for outer, iters in ((1, 1), (1, 10), (3, 10)):
        fn = lambda: h.lib.call("sfmb200_bundle_adjust", h._h, outer, iters, None)
        res[f"ms_outer{outer}_iters{iters}"] = timed(fn, reps)
    per_iter = (res["ms_outer1_iters10"] - res["ms_outer1_iters1"]) / 9
    st = h.bundle_adjust(1, 1)
    bytes_iter = pairs * n * (2 * 29 + 12)
    print(json.dumps(dict(pairs=pairs, n=n, **res, ms_per_lm_iteration=per_iter, gbs_per_lm_iteration=bytes_iter / (per_iter * 1e-3) / 1e9,
                          points_per_s_per_iteration=pairs * n / (per_iter * 1e-3), active_mean=float(st[:, 0].mean()))), flush=True)
    h.close()
