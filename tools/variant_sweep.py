#!/usr/bin/env python
"""Score-kernel time of every variant on config 2 (CUDA events around run_device's score stage, mean of 20)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
pkg = entry.load_package()
K, Kinv = pkg.synthetic.reference_K()
n, H = 10000, 65536
px = pkg.synthetic.synthetic_pair(n, 0.3, 1.0, seed=1234)["px"]
d_px = torch.from_numpy(px[None]).cuda()
h = pkg.BatchedPairs(K, Kinv, 1, n, H)
ref = None
for v in range(11):
    h.set_option(2, v)
    h.set_option(4, 1)
    for _ in range(3):
        h.run_device(d_px, H, 1237, 1e-6)
    h.set_option(4, 1)
    for _ in range(20):
        h.run_device(d_px, H, 1237, 1e-6)
    st = h.stage_times().mean(axis=0)
    idx, cnt = h.get_best()
    ref = ref or (int(idx[0]), int(cnt[0]))
    assert (int(idx[0]), int(cnt[0])) == ref
    print(json.dumps(dict(variant=v, plan=h.score_plan(), score_ms=float(st[2]), evals_per_s=n * H / (st[2] * 1e-3),
                          pct_nominal=100 * 34 * n * H / (st[2] * 1e-3) / 74.45e12), default=float), flush=True)
