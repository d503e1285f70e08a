#!/usr/bin/env python
"""find_homography wall time per call (it returns host results, so the call is synchronous) for the product build and the
experimental builds in tools/proto/explibs."""
import glob, json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    path = sys.argv[2]
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    import __graft_entry__ as entry
    pkg = entry.load_package()
    lib = pkg.load_library(path)
    K, Kinv = pkg.synthetic.reference_K()
    out = {"lib": os.path.basename(path)}
    for n, loops in ((2153, 1000), (2153, 10000), (10000, 10000), (10000, 2000), (500, 10000)):
        px = pkg.synthetic.synthetic_pair(n, 0.3, 1.0, seed=5)["px"]
        h = pkg.BatchedPairs(K, Kinv, 1, n, max(loops, 1024), lib=lib)
        h.set_points_xy(torch.from_numpy(px[None]).cuda())
        for _ in range(5):
            Hm, cnt = h.find_homography(loops, 3, 5.0)
        ts = []
        for _ in range(40):
            t0 = time.perf_counter(); Hm, cnt = h.find_homography(loops, 3, 5.0); ts.append(time.perf_counter() - t0)
        out[f"n{n}_loops{loops}_us"] = round(1e6 * sorted(ts)[20], 1)
        out[f"n{n}_loops{loops}_matches"] = int(cnt[0])
        h.close()
    print(json.dumps(out), flush=True)
else:
    libs = [os.path.join(ROOT, "cuda-sfm_b200", "libsfmb200.so")] + sorted(glob.glob(os.path.join(ROOT, "tools", "proto", "explibs", "*.so")))
    for l in libs + libs[:1]:
        subprocess.run([sys.executable, __file__, "--one", l])
