#!/usr/bin/env python
"""Which scoring variant the plan should pick per shape: score-stage ms for every variant over small / medium / batched
shapes (the automatic choice of make_score_plan is marked)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as entry
pkg = entry.load_package()
K, Kinv = pkg.synthetic.reference_K()
SHAPES = ((1, 2153, 269), (64, 2153, 269), (256, 2153, 269), (1, 10000, 512), (1, 10000, 1024), (1, 10000, 2048), (1, 10000, 4096),
          (16, 4096, 512), (16, 4096, 1024), (64, 4096, 1024), (1, 100000, 8192))
if len(sys.argv) > 1 and sys.argv[1] == "more":
    SHAPES = ((1, 10000, 8192), (1, 10000, 16384), (1, 10000, 32768), (4, 10000, 4096), (1, 4096, 4096), (1, 50000, 2048), (1024, 2153, 269),
              (32, 4096, 4096), (8, 4096, 4096), (1, 1000, 125), (8, 1000, 125), (1, 3000, 700), (4, 3000, 700), (1, 30000, 300), (2, 10000, 65536))
for (B, n, H) in SHAPES:
    px = np.stack([pkg.synthetic.synthetic_pair(n, 0.3, 1.0, seed=1234 + b % 4)["px"] for b in range(B)])
    d_px = torch.from_numpy(px).cuda()
    row, auto = {}, None
    for variant in (-1, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 11):
        h = pkg.BatchedPairs(K, Kinv, B, n, H)
        h.set_option(7, 0)
        h.set_option(2, variant)
        h.set_option(4, 1)
        for _ in range(3):
            h.run_device(d_px, H, 1237, 1e-6)
        h.set_option(4, 1)
        for _ in range(12):
            h.run_device(d_px, H, 1237, 1e-6)
        st = h.stage_times().mean(axis=0)
        if variant < 0:
            auto = h.score_plan()["variant"]
        else:
            row[variant] = round(float(st[2]) * 1e3, 2)
        h.close()
    best = min(row, key=row.get)
    print(json.dumps(dict(B=B, n=n, H=H, auto=auto, auto_us=row[auto], best=best, best_us=row[best], all_us=row)), flush=True)
