#!/usr/bin/env python
"""Phase breakdown of the fused small-problem kernel on the dino / CudaSift fixture (clock64 stamps of CTA 0)."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as entry
pkg = entry.load_package()
lib = pkg.load_library()
g = np.load(os.path.join(ROOT, "tests", "golden", "dino_cudasift_000_001.npz"))
px = np.ascontiguousarray(g["px"]); n, H = len(px), len(g["idx"])
K, Kinv = pkg.synthetic.reference_K(720, 576)
d_px = torch.from_numpy(px).cuda()
h = pkg.BatchedPairs(K, Kinv, 1, n, H)
stamps = torch.zeros(16, dtype=torch.int64, device="cuda")
lib.call("sfmb200_small_path_debug", C.c_void_p(stamps.data_ptr()))
for _ in range(5):
    h.run_device(d_px, H, 1237, 1e-6)
torch.cuda.synchronize()
s = stamps.cpu().numpy()
names = ["ingest", "hypgen", "score", "select+pose", "triangulate"]
d = np.diff(s[:6])
khz = torch.cuda.get_device_properties(0).clock_rate if hasattr(torch.cuda.get_device_properties(0), "clock_rate") else 1965000
print(json.dumps({"cycles": dict(zip(names, [int(v) for v in d])), "total_cycles": int(s[5] - s[0]), "us_at_1.9GHz": {k: round(float(v) / 1900.0, 2) for k, v in zip(names, d)}}))
sub = {"score: E staged (since hypgen end)": int(s[8] - s[2]), "score: loop done": int(s[9] - s[8]), "score: CTA barrier": int(s[10] - s[9]),
       "score: counts + arg-max atomics": int(s[11] - s[10]), "score: cluster barrier": int(s[3] - s[11]),
       "pose: select (best key + E read)": int(s[12] - s[3]), "pose: candidates (SVD + orientation replay)": int(s[13] - s[12]),
       "pose: cheirality + stores": int(s[14] - s[13]), "pose: cluster barrier": int(s[4] - s[14])}
print(json.dumps({"sub_phase_cycles": sub}))
lib.call("sfmb200_small_path_debug", C.c_void_p(0))
ts = []
for i in range(50):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); h.run_device(d_px, H, 1237, 1e-6); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
busy = []
for i in range(50):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(200000)          # the stream stays busy while the host submits: device time only between the events
    a.record(); h.run_device(d_px, H, 1237, 1e-6); b.record(); torch.cuda.synchronize(); busy.append(a.elapsed_time(b))
print(json.dumps({"run_device_ms_median": sorted(ts)[25], "min": min(ts), "stream_busy_ms_median": sorted(busy)[25], "stream_busy_min": min(busy),
                  "plan": h.score_plan()}))
