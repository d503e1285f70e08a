#!/usr/bin/env python
"""Generates tests/golden/dino_000_001.npz: BASELINE config 1 as a committed fixture.

Input: the reference's own image pair data/dino/viff.000.ppm / viff.001.ppm
(src/main.cpp:251-252).  The reference extracts and matches SIFT features with
the vendored CudaSift on a GPU; this container has no GPU, so the fixture uses
OpenCV's SIFT + brute-force L2 matching on the CPU, with the reference's policy
of keeping EVERY feature of image 1 with its best match, unfiltered
(main.cpp:298-299; the homography filter is commented out, main.cpp:283-290).
Sample rows follow sfm.cu:95-104: one permutation of the point indices cut into
H = N/8 disjoint groups of 8 (numpy PCG64 seed 2019 instead of random_device).

Golden outputs come from the fp64 oracle (oracle/oracle.py): per-hypothesis E,
inlier counts and the arg-max.  Run from the repo root, in the build container:
    python tests/golden/make_dino_fixture.py /root/reference
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
O = entry.load_oracle()
im1 = cv2.imread(os.path.join(ref, "data/dino/viff.000.ppm"), cv2.IMREAD_GRAYSCALE)
im2 = cv2.imread(os.path.join(ref, "data/dino/viff.001.ppm"), cv2.IMREAD_GRAYSCALE)
assert im1 is not None and im1.shape == (576, 720)
sift = cv2.SIFT_create(nfeatures=0, contrastThreshold=0.02)
k1, d1 = sift.detectAndCompute(im1, None)
k2, d2 = sift.detectAndCompute(im2, None)
matches = cv2.BFMatcher(cv2.NORM_L2).match(d1, d2)            # best match for every feature of image 1
px = np.array([[*k1[m.queryIdx].pt, *k2[m.trainIdx].pt] for m in matches], dtype=np.float32)
n = len(px)
H = n // 8
perm = np.random.Generator(np.random.PCG64(2019)).permutation(n).astype(np.int32)
idx = perm[: 8 * H].reshape(H, 8)
K, Kinv = O.reference_K(720, 576)
x = O.normalise_points(px, Kinv)
E = O.hypotheses(x, idx)
cnt, amb = O.inlier_counts(E.reshape(H, 9).astype(np.float32).astype(np.float64), x, 1e-6, band=1e-4)
out = os.path.join(ROOT, "tests", "golden", "dino_000_001.npz")
np.savez_compressed(out, px=px, idx=idx, E64=E, counts=cnt.astype(np.int32), borderline=amb.astype(np.int32),
                    best=np.int32(O.argmax_first(cnt)))
print(f"{out}: {n} correspondences, {H} hypotheses, best {int(np.argmax(cnt))} with {int(cnt.max())} inliers")
