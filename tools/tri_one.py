#!/usr/bin/env python
"""One process that launches the triangulation kernel a few times at 1M points (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as entry
pkg = entry.load_package()
K, Kinv = pkg.synthetic.reference_K()
n, H = 1 << 20, 4096
px = pkg.synthetic.synthetic_pair(n, 0.3, 1.0, seed=77)["px"]
d_px = torch.from_numpy(px[None]).cuda()
lib = pkg.load_library(os.environ["SFMB200_LIB"]) if os.environ.get("SFMB200_LIB") else None      # an experimental build (tools/proto/explibs)
h = pkg.BatchedPairs(K, Kinv, 1, n, H, lib=lib) if lib else pkg.BatchedPairs(K, Kinv, 1, n, H)
h.run_device(d_px, H, 1237, 1e-6)
for _ in range(3):
    h.triangulate()
torch.cuda.synchronize()
