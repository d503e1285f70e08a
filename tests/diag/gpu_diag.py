#!/usr/bin/env python
"""Diagnostics for parity questions that need a GPU (printed, not asserted)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry

pkg, O = entry.load_package(), entry.load_oracle()
K, Kinv = O.reference_K()
fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)
P = lambda a, t=fp: a.ctypes.data_as(t)
oc = C.CDLL(os.path.join(ROOT, "oracle/_ref/liboracle_c.so"))

sc = O.synthetic_pair(3000, seed=4321)
x = O.normalise_points(sc["px"], Kinv)
n, H, seed = 3000, 3000, 1237
h = pkg.BatchedPairs(K, Kinv, 1, n, H)
h.set_points_xy(torch.from_numpy(sc["px"]).cuda())
h.estimate_e(H, seed, 1e-6)
Eg = h.get_E_candidates().cpu().numpy()
idx = O.sample_indices(seed, H, n)
E64 = O.hypotheses(x, idx)
d = O.e_distance(Eg, E64)
print("ours vs fp64: within1e-4 %.4f median %.2e p99 %.2e max %.2e nan %d" % (np.mean(d < 1e-4), np.nanmedian(d), np.nanpercentile(d, 99), np.nanmax(d), np.isnan(d).sum()))
lib = pkg.load_library()
Eh = np.zeros((H, 9), np.float32)
for i in range(H):
    p = np.ascontiguousarray(x[idx[i]], np.float32)
    lib.raw("sfmb200_host_solve_hypothesis")(P(p), P(Eh[i]))
dh = O.e_distance(Eh, E64)
print("host solver vs fp64: within1e-4 %.4f ; gpu vs host max %.2e" % (np.mean(dh < 1e-4), O.e_distance(Eg, Eh).max()))
X0, X1 = h.get_X(0).cpu().numpy(), h.get_X(1).cpu().numpy()
xg = np.ascontiguousarray(np.stack([X0[0], X0[1], X1[0], X1[1]], 1))
print("x gpu vs oracle max diff", np.abs(xg - x).max())
got = h.get_inlier_counts().cpu().numpy()
c32 = np.zeros(H, np.int32)
oc.oracle_counts_f32(P(Eg), H, P(xg), n, C.c_float(1e-6), P(c32, ip))
print("counts equal fp32 port:", np.array_equal(got, c32), "mismatches", int((got != c32).sum()), "max abs diff", int(np.abs(got - c32).max()))
bi, bc = h.get_best()
print("best", bi, bc, "argmax", int(np.argmax(got)), int(got.max()))

# reference E candidates
rp = os.path.join(ROOT, "oracle/_ref/libsfm_ref.so")
if os.path.exists(rp):
    L = C.CDLL(rp)
    L.ref_create.restype = C.c_void_p
    r = C.c_void_p(L.ref_create(P(K.reshape(9).copy()), P(Kinv.reshape(9).copy()), n))
    L.ref_fillXU(r, P(sc["px"]))
    Er = np.zeros((H, 9), np.float32)
    Ar = np.zeros((H, 72), np.float32)
    Vr = np.zeros((H, 81), np.float32)
    L.ref_e_candidates(r, P(idx, ip), H, P(Er), P(Ar), P(Vr))
    dr = O.e_distance(Er, E64)
    dgr = O.e_distance(Eg, Er)
    print("ref vs fp64: within1e-4 %.4f median %.2e p99 %.2e max %.2e" % (np.mean(dr < 1e-4), np.median(dr), np.percentile(dr, 99), dr.max()))
    print("ours vs ref: within1e-4 %.4f median %.2e p99 %.2e" % (np.mean(dgr < 1e-4), np.median(dgr), np.percentile(dgr, 99)))
    bad = dgr >= 1e-4
    print("where ours-vs-ref >= 1e-4 (%d): ours-vs-fp64 median %.2e, ref-vs-fp64 median %.2e; ours worse in %d" % (bad.sum(), np.median(d[bad]), np.median(dr[bad]), int((d[bad] > dr[bad]).sum())))
    # is it the reference's null vector or its rank-2 projection?
    null = Vr[:, 72:81].astype(np.float64).reshape(H, 3, 3)
    A64 = O.design_matrix(x[idx])
    true_null = np.linalg.svd(A64)[2][:, -1, :].reshape(H, 3, 3)
    dn = O.e_distance(null, true_null)
    print("ref null vector vs fp64 null: within1e-4 %.4f median %.2e" % (np.mean(dn < 1e-4), np.median(dn)))
    proj_of_refnull = O.project_essential(null)
    print("fp64 projection of ref null vs ref E: within1e-4 %.4f (isolates normalizeE/svd.h error)" % np.mean(O.e_distance(proj_of_refnull, Er) < 1e-4))
