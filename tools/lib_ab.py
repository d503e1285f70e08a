#!/usr/bin/env python
"""A/B of experimental builds of the library (tools/proto/libs/*.so) against the product build:
score-kernel and whole-step time of config 2, CUDA events, mean of 30; each build in its own process."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    path = sys.argv[2]
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    import __graft_entry__ as entry
    pkg = entry.load_package()
    pkg.load_library(path)
    K, Kinv = pkg.synthetic.reference_K()
    n, H = 10000, 65536
    px = pkg.synthetic.synthetic_pair(n, 0.3, 1.0, seed=1234)["px"]
    d_px = torch.from_numpy(px[None]).cuda()
    h = pkg.BatchedPairs(K, Kinv, 1, n, H, lib=pkg.load_library(path))
    if os.environ.get("SFMB200_AB_VARIANT"):
        h.set_option(2, int(os.environ["SFMB200_AB_VARIANT"]))
    h.set_option(4, 1)
    for _ in range(5):
        h.run_device(d_px, H, 1237, 1e-6)
    h.set_option(4, 1)
    for _ in range(30):
        h.run_device(d_px, H, 1237, 1e-6)
    st = h.stage_times().mean(axis=0)
    h.set_option(4, 0)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(50):
        h.run_device(d_px, H, 1237, 1e-6)
    b.record(); torch.cuda.synchronize()
    idx, cnt = h.get_best()
    out = dict(lib=os.path.basename(path), hypgen_ms=float(st[1]), score_ms=float(st[2]), step_ms=a.elapsed_time(b) / 50, best=[int(idx[0]), int(cnt[0])])
    h.close()
    # a config-4 slice: 256 pairs x 4,096 x 4,096
    B4, n4, H4 = 256, 4096, 4096
    px4 = np.stack([pkg.synthetic.synthetic_pair(n4, 0.3, 1.0, seed=1234 + (b % 8))["px"] for b in range(B4)])
    d_px4 = torch.from_numpy(px4).cuda()
    h = pkg.BatchedPairs(K, Kinv, B4, n4, H4, lib=pkg.load_library(path))
    if os.environ.get("SFMB200_AB_VARIANT"):
        h.set_option(2, int(os.environ["SFMB200_AB_VARIANT"]))
    h.set_option(4, 1)
    for _ in range(3):
        h.run_device(d_px4, H4, 1237, 1e-6)
    h.set_option(4, 1)
    for _ in range(10):
        h.run_device(d_px4, H4, 1237, 1e-6)
    st4 = h.stage_times().mean(axis=0)
    out.update(c4_slice_hypgen_ms=float(st4[1]), c4_slice_score_ms=float(st4[2]), c4_slice_total_ms=float(st4.sum()))
    print(json.dumps(out), flush=True)
else:
    libs = [os.path.join(ROOT, "cuda-sfm_b200", "libsfmb200.so")] + sorted(glob.glob(os.path.join(ROOT, "tools", "proto", "explibs", "*.so")))
    for l in libs + libs[:1]:
        subprocess.run([sys.executable, __file__, "--one", l])
