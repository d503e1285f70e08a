#!/usr/bin/env python
"""Per-stage device time of the general (five-launch) path at a few shapes: ingest, hypgen, score, select, poses, choose, triangulate."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as entry
pkg = entry.load_package()
K, Kinv = pkg.synthetic.reference_K()
for (B, n, H) in ((1, 2153, 269), (64, 2153, 269), (1, 10000, 2048), (1, 10000, 65536), (256, 4096, 4096), (1, 1 << 20, 4096)):
    px = np.stack([pkg.synthetic.synthetic_pair(n, 0.3, 1.0, seed=1234 + b % 4)["px"] for b in range(B)])
    d_px = torch.from_numpy(px).cuda()
    h = pkg.BatchedPairs(K, Kinv, B, n, H)
    h.set_option(7, 0)
    h.set_option(4, 1)
    for _ in range(3):
        h.run_device(d_px, H, 1237, 1e-6)
    h.set_option(4, 1)
    for _ in range(12):
        h.run_device(d_px, H, 1237, 1e-6)
    st = h.stage_times().mean(axis=0)
    print(json.dumps(dict(B=B, n=n, H=H, plan=h.score_plan()["variant"], stage_us=[round(float(v) * 1e3, 2) for v in st], total_us=round(float(st.sum()) * 1e3, 2))), flush=True)
    h.close()
