#!/usr/bin/env python
"""How much do the per-stage CUDA events (SFMB200_OPT_PROFILE) cost inside a step?  Config 2, 300 steps each."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
pkg = entry.load_package()
K, Kinv = pkg.synthetic.reference_K()
px = pkg.synthetic.synthetic_pair(10000, 0.3, 1.0, seed=1234)["px"]
d_px = torch.from_numpy(px).cuda()
h = pkg.BatchedPairs(K, Kinv, 1, 10000, 65536)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for prof in (1, 0, 1, 0):
    h.set_option(4, prof)
    for _ in range(5):
        h.run_device(d_px, 65536, 1237, 1e-6); flush.fill_(1)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(300)]
    torch.cuda.synchronize()
    for i in range(300):
        ev[i][0].record(); h.run_device(d_px, 65536, 1237, 1e-6); ev[i][1].record(); flush.fill_(i & 0xFF)
    torch.cuda.synchronize()
    print("profile", prof, "ms/step", sum(a.elapsed_time(b) for a, b in ev) / 300)
