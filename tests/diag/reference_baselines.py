#!/usr/bin/env python
"""Side-by-side timings of the reference's own CUDA path (oracle/_ref/libsfm_ref.so,
unmodified sources rebuilt for sm_100a) and sfmb200 on the BASELINE.json configs the
reference can run at all: config 1 (dino fixture), config 2, a
16-pair subset of config 4, and the pose + triangulation half of config 5 (1M points,
same injected E).  Host wall clock around synchronised calls for the reference (its
mallocs / syncs are part of its cost), CUDA events for ours.  JSON lines to stdout."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry

pkg, O = entry.load_package(), entry.load_oracle()
fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)
P = lambda a, t=fp: a.ctypes.data_as(t)
L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libsfm_ref.so"))
L.ref_create.restype = C.c_void_p
for name in ("ref_estimateE_injected", "ref_computePosecandidates", "ref_choosePose", "ref_linear_triangulation"):
    getattr(L, name).restype = C.c_float
THR = 1e-6


def emit(**kw):
    print(json.dumps(kw, default=float), flush=True)


def ours_ms(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        ev.append((a, b))
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev]))


def ref_pair(K, Kinv, px, idx, reps=3):
    n, H = len(px), len(idx)
    r = C.c_void_p(L.ref_create(P(K.reshape(9).copy()), P(Kinv.reshape(9).copy()), n))
    out = []
    for _ in range(reps + 1):
        t0 = time.perf_counter()
        L.ref_fillXU(r, P(px))
        t1 = time.perf_counter()
        te = L.ref_estimateE_injected(r, P(idx, ip), H, None)
        tp = L.ref_computePosecandidates(r)
        tc = L.ref_choosePose(r)
        tt = L.ref_linear_triangulation(r)
        out.append([1e3 * (t1 - t0), te, tp, tc, tt])
    L.ref_destroy(r)
    m = np.median(np.array(out[1:]), axis=0)
    return dict(zip(("fillXU", "estimateE", "computePosecandidates", "choosePose", "linear_triangulation"), m)), float(m.sum())


def ours_pair(K, Kinv, px, idx, pairs=1):
    n, H = px.shape[-2], idx.shape[-2]
    h = pkg.BatchedPairs(K, Kinv, pairs, n, H)
    h.set_option(4, 1)
    d_px, d_idx = torch.from_numpy(px).cuda(), torch.from_numpy(idx).cuda()

    def step():
        h.set_points_xy(d_px, n)
        h.estimate_e(H, 0, THR, d_idx=d_idx)
        h.pose_candidates(); h.choose_pose(); h.triangulate()
    ms = ours_ms(step)
    h.close()
    return ms


K, Kinv = O.reference_K()
# ---- config 1: dino fixture, H = N/8 disjoint rows (the reference's operating point)
g = np.load(os.path.join(ROOT, "tests", "golden", "dino_000_001.npz"))
px, idx = g["px"], np.ascontiguousarray(g["idx"])
st, tot = ref_pair(K, Kinv, px, idx)
mine = ours_pair(K, Kinv, px, idx)
emit(config="c1 dino 000/001", n=len(px), H=len(idx), reference_ms=tot, reference_stage_ms=st, ours_ms=mine, speedup=tot / mine,
     published_gtx1080ti_ms=38.92)
# ---- config 2
sc = O.synthetic_pair(10000, seed=1234)
idx = O.sample_indices(1237, 65536, 10000)
st, tot = ref_pair(K, Kinv, sc["px"], idx, reps=2)
mine = ours_pair(K, Kinv, sc["px"], idx)
emit(config="c2", n=10000, H=65536, reference_ms=tot, reference_stage_ms=st, ours_ms=mine, speedup=tot / mine)
# ---- config 4 subset: 16 pairs, sequential on the reference (it has no batching), one batch on ours
pairs = 16
scs = [O.synthetic_pair(4096, seed=500 + i)["px"] for i in range(pairs)]
idxs = np.stack([O.sample_indices(99 + i, 4096, 4096) for i in range(pairs)])
ref_tot = 0.0
for i in range(pairs):
    _, t = ref_pair(K, Kinv, scs[i], idxs[i], reps=1)
    ref_tot += t
mine = ours_pair(K, Kinv, np.stack(scs), idxs, pairs=pairs)
emit(config="c4 subset", pairs=pairs, n=4096, H=4096, reference_ms=ref_tot, ours_ms=mine, speedup=ref_tot / mine,
     reference_pairs_per_s=pairs / (ref_tot * 1e-3), ours_pairs_per_s=pairs / (mine * 1e-3),
     note="ours at 16 pairs underfills the GPU; see r01_scaling.md for 4,096 pairs")
# ---- config 5, pose + triangulation half: 1M points, injected E
n = 1 << 20
sc = O.synthetic_pair(n, seed=77)
h = pkg.BatchedPairs(K, Kinv, 1, n, 4096)
d_px = torch.from_numpy(sc["px"]).cuda()
h.set_points_xy(d_px)
h.estimate_e(4096, 1, THR)
E = h.get_E()[0].reshape(9).copy()
r = C.c_void_p(L.ref_create(P(K.reshape(9).copy()), P(Kinv.reshape(9).copy()), n))
L.ref_fillXU(r, P(sc["px"]))
rows = []
for _ in range(3):
    L.ref_set_E(r, P(E))
    rows.append([L.ref_computePosecandidates(r), L.ref_choosePose(r), L.ref_linear_triangulation(r)])
L.ref_destroy(r)
m = np.median(np.array(rows[1:]), axis=0)


def half():
    h.set_E(E)
    h.pose_candidates(); h.choose_pose(); h.triangulate()
mine = ours_ms(half)
emit(config="c5 pose+triangulation half", n=n, reference_ms=float(m.sum()),
     reference_stage_ms=dict(zip(("computePosecandidates", "choosePose", "linear_triangulation"), m)), ours_ms=mine,
     speedup=float(m.sum()) / mine, note="ours includes a blocking host->device copy of E in set_E")
h.close()
