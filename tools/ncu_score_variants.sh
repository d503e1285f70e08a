#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/run_v.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import __graft_entry__ as entry
pkg = entry.load_package(); S = pkg.synthetic
K, Kinv = S.reference_K()
px = torch.from_numpy(S.synthetic_pair(10000, seed=1234)["px"]).cuda()
for v in (3, 4, 0, 2):
    h = pkg.BatchedPairs(K, Kinv, 1, 10000, 65536)
    h.set_option(2, v)
    for _ in range(3):
        h.run_device(px, 65536, 1237, 1e-6)
    torch.cuda.synchronize()
    h.close()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_kernel -o gpurun_out/prof_variants python /tmp/run_v.py > gpurun_out/ncu_variants.log 2>&1
tail -2 gpurun_out/ncu_variants.log
