#!/usr/bin/env python
"""Is the whole path capturable in a CUDA graph, and what does replay buy at launch-bound sizes?"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
pkg = entry.load_package()
K, Kinv = pkg.synthetic.reference_K()
for n, H in ((1577, 197), (2000, 250), (10000, 65536)):
    px = pkg.synthetic.synthetic_pair(n, 0.3, 1.0, seed=5)["px"]
    d_px = torch.from_numpy(px[None]).cuda()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        h = pkg.BatchedPairs(K, Kinv, 1, n, H)          # the handle adopts torch's current stream
        for _ in range(3):
            h.run_device(d_px, H, 7, 1e-6)
        s.synchronize()
        ref = (h.get_best()[0].copy(), h.get_E().copy(), h.get_points_host(0).copy())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            h.run_device(d_px, H, 7, 1e-6)
        g.replay(); s.synchronize()
        got = (h.get_best()[0].copy(), h.get_E().copy(), h.get_points_host(0).copy())
        same = all(np.array_equal(a, b) for a, b in zip(ref, got))
        def timed(fn, reps=200):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.synchronize(); a.record(s)
            for _ in range(reps):
                fn()
            b.record(s); s.synchronize()
            return a.elapsed_time(b) / reps
        t_direct = timed(lambda: h.run_device(d_px, H, 7, 1e-6))
        t_graph = timed(g.replay)
        print(json.dumps(dict(n=n, H=H, same_bits=bool(same), ms_direct=t_direct, ms_graph_replay=t_graph)), flush=True)
        h.close()
