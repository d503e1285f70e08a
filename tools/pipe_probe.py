#!/usr/bin/env python
"""Issue rate of the packed FP32 instructions on this GPU (sfmb200_fma_probe): lane operations per second per mode."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as entry
lib = entry.load_package().load_library()
torch.cuda.init(); torch.zeros(1, device="cuda")
for mode, name in ((0, "FFMA"), (1, "FFMA2"), (2, "FMUL2"), (3, "FADD2"), (4, "FFMA2 with negated operand")):
    best = 0.0
    for _ in range(3):
        ops, ms = C.c_double(), C.c_float()
        lib.call("sfmb200_fma_probe", mode, 2000, C.byref(ops), C.byref(ms))
        best = max(best, ops.value / (ms.value * 1e-3))
    print(json.dumps({"mode": name, "lane_ops_per_s": best, "fraction_of_nominal_lane_rate": best / (148 * 128 * 1.965e9)}))
