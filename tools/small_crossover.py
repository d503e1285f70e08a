#!/usr/bin/env python
"""Where the fused small-problem kernel stops paying: run_device per pair with the fused path forced on / off
(SFMB200_OPT_SMALL_PATH = 1 / 0) over a sweep of sizes; device time with the stream kept busy while the host submits."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as entry
pkg = entry.load_package()
K, Kinv = pkg.synthetic.reference_K()
SWEEP = [(1, n, H) for n, H in ((1000, 125), (2153, 269), (2500, 500), (3000, 700), (4000, 800), (4000, 1000), (3000, 1500), (4000, 1500), (3000, 2048),
                                  (6000, 1000), (10000, 600))]
SWEEP += [(B, 2153, 269) for B in (2, 4, 8, 16, 64, 256)] + [(B, 1000, 125) for B in (4, 16, 64)]
for B, n, H in SWEEP:
    px = np.stack([pkg.synthetic.synthetic_pair(n, 0.3, 1.0, seed=77 + (b % 4))["px"] for b in range(B)])
    d_px = torch.from_numpy(px).cuda()
    h = pkg.BatchedPairs(K, Kinv, B, n, H)
    h.set_option(8, 2 ** 31 - 1)            # the size limit out of the way: option 7 alone decides
    row = {"pairs": B, "n": n, "H": H, "evals_per_pair": n * H}
    for name, small in (("fused", 1), ("general", 0)):
        h.set_option(7, small)
        for _ in range(5):
            h.run_device(d_px, H, 1237, 1e-6)
        ts = []
        for i in range(40):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(300000)
            a.record(); h.run_device(d_px, H, 1237, 1e-6); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        row[name + "_us"] = round(1e3 * sorted(ts)[len(ts) // 2], 2)
        row[name + "_launches"] = None
    print(json.dumps(row), flush=True)
    h.close()
