#!/usr/bin/env python
"""One-shot measurements on the GPU box (results go to stdout as JSON lines):
FP32 pipe probe, per-stage times of config 2 for both scoring variants,
hypothesis generation rate, triangulation at 1M points, batched config 4 slice."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry

pkg = entry.load_package()
O = pkg.synthetic          # scenes + intrinsics only: no checker code in these tools
lib = pkg.load_library()
K, Kinv = O.reference_K()
THR = 1e-6


def emit(**kw):
    print(json.dumps(kw, default=float), flush=True)


def probe():
    for mode, name in ((0, "ffma"), (1, "ffma2")):
        for it in (500, 4000):
            fmas, ms = C.c_double(), C.c_float()
            lib.call("sfmb200_fma_probe", mode, it, C.byref(fmas), C.byref(ms))
            emit(what="fma_probe", mode=name, iters=it, ms=ms.value, tflops=2 * fmas.value / (ms.value * 1e-3) / 1e12,
                 lane_fma_per_clk_per_sm_at_1965=fmas.value / (ms.value * 1e-3) / 148 / 1.965e9)


def stages(n, H, variant, pairs=1, reps=10, label="", solver=0):
    scenes = [O.synthetic_pair(n, seed=1234 + b) for b in range(min(pairs, 4))]
    px = np.stack([scenes[b % len(scenes)]["px"] for b in range(pairs)])
    d_px = torch.from_numpy(px).cuda()
    h = pkg.BatchedPairs(K, Kinv, pairs, n, H)
    h.set_option(2, variant)
    h.set_option(5, solver)
    h.set_option(4, 1)
    for _ in range(3):
        h.run_device(d_px, H, 1237, THR, n=n)
    h.set_option(4, 1)
    for _ in range(reps):
        h.run_device(d_px, H, 1237, THR, n=n)
    st = h.stage_times()
    m = st.mean(axis=0)
    evals = pairs * n * H
    emit(what="stages", label=label, n=n, H=H, pairs=pairs, variant=variant, solver=solver, plan=h.score_plan(),
         stage_ms=dict(zip(h.STAGES, [float(v) for v in m])), stage_min_ms=dict(zip(h.STAGES, [float(v) for v in st.min(axis=0)])),
         score_evals_per_s=evals / (m[2] * 1e-3), score_tflops_34=34 * evals / (m[2] * 1e-3) / 1e12,
         hyp_per_s=pairs * H / (m[1] * 1e-3), tri_points_per_s=pairs * n / (m[6] * 1e-3), tri_gbs=32 * pairs * n / (m[6] * 1e-3) / 1e9,
         best=[int(v) for v in h.get_best()[1][:4]])
    h.close()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    emit(what="device", name=torch.cuda.get_device_name(0), sms=torch.cuda.get_device_properties(0).multi_processor_count,
         hypgen_minb=os.environ.get("SFMB200_HYPGEN_MINB", "2"))
    if which == "hypgen":
        for solver in (0, 1):
            stages(10_000, 65_536, -1, label="config2", solver=solver)
            stages(4096, 4096, -1, pairs=256, reps=3, label="config4 slice: 256 pairs", solver=solver)
        sys.exit(0)
    if which == "score":
        for v in (4, 10, -1):
            stages(10_000, 65_536, v, label="config2", solver=1)
        for v in (4, 10, -1):
            stages(1 << 20, 16_384, v, reps=3, label="1M points x 16k hyp", solver=1)
        for v in (4, 10, -1):
            stages(4096, 4096, v, pairs=256, reps=3, label="config4 slice: 256 pairs", solver=1)
        for v in (9, 10, -1):
            stages(2000, 250, v, label="config1-like: 2k corr, 250 hyp", solver=1)
        sys.exit(0)
    probe()
    nv = 10
    for v in range(nv):
        stages(10_000, 65_536, v, label="config2", solver=1)
    for v in (3, 4, 8):
        stages(1 << 20, 16_384, v, reps=3, label="1M points x 16k hyp", solver=1)
    for v in (3, 4, 6, 8):
        stages(4096, 4096, v, pairs=256, reps=3, label="config4 slice: 256 pairs", solver=1)
    stages(2000, 250, -1, label="config1-like: 2k corr, 250 hyp")
    stages(10_000, 65_536, -1, label="config2 auto")
