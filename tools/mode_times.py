#!/usr/bin/env python
"""Stage times of the whole path at config 2 in every mode (compat / vote, inliers-only triangulation, both solvers)."""
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
for compat, tri_only, solver in ((1, 0, 1), (0, 0, 1), (0, 1, 1), (1, 0, 0)):
    h = pkg.BatchedPairs(K, Kinv, 1, n, H)
    h.set_option(1, compat); h.set_option(3, tri_only); h.set_option(5, solver)
    for _ in range(3):
        h.run_device(d_px, H, 1237, 1e-6)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(50):
        h.run_device(d_px, H, 1237, 1e-6)
    b.record(); torch.cuda.synchronize()
    h.set_option(4, 1)
    for _ in range(12):
        h.run_device(d_px, H, 1237, 1e-6)
    st = h.stage_times()[2:].mean(axis=0)
    print(json.dumps(dict(compat=compat, tri_inliers_only=tri_only, solver=solver, step_ms=a.elapsed_time(b) / 50,
                          stage_ms=dict(zip(h.STAGES, [round(float(v), 4) for v in st])), pose=int(h.get_pose_index()[0]))), flush=True)
    h.close()
