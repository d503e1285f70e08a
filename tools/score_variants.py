#!/usr/bin/env python
"""Scoring kernel variants at config 2 (and a config-4 slice): score-stage ms and whole-step ms per variant."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as entry
pkg = entry.load_package()
K, Kinv = pkg.synthetic.reference_K()
for (B, n, H) in ((1, 10000, 65536), (256, 4096, 4096)):
    px = np.stack([pkg.synthetic.synthetic_pair(n, 0.3, 1.0, seed=1234 + b % 8)["px"] for b in range(B)])
    d_px = torch.from_numpy(px).cuda()
    for variant in (4, 8, 11, 3):
        h = pkg.BatchedPairs(K, Kinv, B, n, H)
        h.set_option(2, variant)
        h.set_option(4, 1)
        for _ in range(5):
            h.run_device(d_px, H, 1237, 1e-6)
        h.set_option(4, 1)
        for _ in range(30):
            h.run_device(d_px, H, 1237, 1e-6)
        st = h.stage_times().mean(axis=0)
        idx, cnt = h.get_best()
        print(json.dumps(dict(B=B, n=n, H=H, variant=variant, plan=h.score_plan(), score_ms=float(st[2]), hypgen_ms=float(st[1]),
                              evals_per_s=B * n * H / (float(st[2]) * 1e-3), best=[int(idx[0]), int(cnt[0])])), flush=True)
        h.close()
