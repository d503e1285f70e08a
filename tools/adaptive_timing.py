#!/usr/bin/env python
"""Device time of the adaptive RANSAC (sfmb200_estimate_e_adaptive) against the fixed-H estimate,
config-2 scene (10,000 correspondences, 30 % outliers, H_max = 65,536) and harder scenes."""
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
n, H = 10000, 65536
for outl in (0.3, 0.5, 0.6, 0.7):
    sc = S.synthetic_pair(n, outlier_frac=outl, seed=1234)
    h = pkg.BatchedPairs(K, Kinv, 1, n, H)
    h.set_points_xy(torch.from_numpy(sc["px"]).cuda())

    def timed(fn, reps=50):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    C = __import__("ctypes")
    t_fixed = timed(lambda: h.estimate_e(H, 1237, 1e-6))
    cnt_fixed = int(h.get_best()[1][0])
    # time without the D2H of `used` (h_used = NULL): pure enqueue
    for first, growth in ((1024, 4), (4096, 4), (4096, 16), (8192, 8)):
        call = lambda: h.lib.call("sfmb200_estimate_e_adaptive", h._h, None, H, first, growth, C.c_uint64(1237), C.c_float(1e-6), C.c_float(0.99), None)
        t_adapt = timed(call)
        used = h.estimate_e_adaptive(H, 1237, 1e-6, 0.99, first, growth)
        cnt = int(h.get_best()[1][0])
        print(json.dumps(dict(outlier_frac=outl, n=n, H_max=H, first_round=first, growth=growth, fixed_ms=t_fixed, fixed_inliers=cnt_fixed,
                              adaptive_ms=t_adapt, adaptive_used=used, adaptive_inliers=cnt)), flush=True)
    h.close()
