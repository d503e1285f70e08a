#!/usr/bin/env python
"""Generates tests/golden/dino_cudasift_000_001.npz: BASELINE config 1 with the
REFERENCE'S OWN front-end and the REFERENCE'S OWN outputs.  Needs a GPU (run on the box):

    gpurun -- 'python tests/golden/make_dino_cudasift_fixture.py gpurun_out/dino_cudasift_000_001.npz'

Input: data/dino/viff.000.ppm / viff.001.ppm read with cv::imread(path, 0) as src/main.cpp:251-252
does (oracle/Makefile stores the grey images as oracle/_ref/dino_viff_00{0,1}.pgm).
Front-end: the unmodified CudaSift sources (oracle/_ref/libcudasift_ref.so, oracle/cudasift_harness.cu)
with main.cpp:260-282's parameters: ExtractSift(5 octaves, initBlur 1.5, thresh 1.0) on both images,
MatchSiftData(siftData1, siftData2).  Every feature of image 1 is a correspondence, unfiltered
(main.cpp:298-299).

Sample rows follow sfm.cu:95-104: one permutation of the point indices cut into H = N/8 disjoint
groups of 8 (numpy PCG64 seed 2019 instead of std::random_device, so the rows can be exported).

Reference outputs stored (oracle/_ref/libsfm_ref.so = SfM/sfm.cu rebuilt unmodified):
  X_ref [2][3][n]        fillXU
  E_ref [H][9]           kernels::kernels -> regular_svd -> row_extraction_kernel -> normalizeE on those rows
  P_ref [4][4][4]        computePosecandidates() on the winning E (winner picked by the fp64 oracle on E_ref,
                         because the reference's own inlier counting is undefined behaviour)
  P_ind_ref, Pinv_ref    choosePose()
  points_ref [4][n]      linear_triangulation()
fp64 oracle outputs stored: E64 [H][9], counts [H], borderline [H], best.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)


def read_pgm(path):
    with open(path, "rb") as f:
        assert f.readline().strip() == b"P5"
        line = f.readline()
        while line.startswith(b"#"):
            line = f.readline()
        w, h = map(int, line.split())
        assert int(f.readline()) == 255
        return np.frombuffer(f.read(w * h), np.uint8).reshape(h, w)


def cudasift_pair(a: int, b: int):
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libcudasift_ref.so"))
    im1 = read_pgm(os.path.join(ROOT, "oracle", "_ref", f"dino_viff_{a:03d}.pgm")).astype(np.float32)
    im2 = read_pgm(os.path.join(ROOT, "oracle", "_ref", f"dino_viff_{b:03d}.pgm")).astype(np.float32)
    h, w = im1.shape
    out = np.zeros((32768, 8), np.float32)
    sift = np.zeros((32768, 144), np.float32)
    n2 = C.c_int(0)
    n = L.cudasift_match_pair(im1.ctypes.data_as(fp), im2.ctypes.data_as(fp), w, h, out.ctypes.data_as(fp), 32768, C.byref(n2),
                              sift.ctypes.data_as(C.c_void_p))
    assert n > 8, n
    return out[:n].copy(), sift[:n].copy(), n2.value, (w, h)


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "dino_cudasift_000_001.npz")
    O = entry.load_oracle()
    feat, sift, n2, (w, h) = cudasift_pair(0, 1)
    n = len(feat)
    px = np.ascontiguousarray(feat[:, :4])
    print(f"CudaSift: {n} features in image 1, {n2} in image 2, image {w}x{h}")
    H = n // 8
    perm = np.random.Generator(np.random.PCG64(2019)).permutation(n).astype(np.int32)
    idx = np.ascontiguousarray(perm[: 8 * H].reshape(H, 8))
    K, Kinv = O.reference_K(w, h)
    x = O.normalise_points(px, Kinv)

    R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libsfm_ref.so"))
    R.ref_create.restype = C.c_void_p
    for name in ("ref_estimateE_injected", "ref_computePosecandidates", "ref_choosePose", "ref_linear_triangulation"):
        getattr(R, name).restype = C.c_float
    r = C.c_void_p(R.ref_create(K.reshape(9).copy().ctypes.data_as(fp), Kinv.reshape(9).copy().ctypes.data_as(fp), n))
    assert R.ref_fillXU(r, px.ctypes.data_as(fp)) == 0
    X_ref = np.zeros((2, 3, n), np.float32)
    for im in (0, 1):
        R.ref_get_X(r, im, X_ref[im].ctypes.data_as(fp))
    E_ref = np.zeros((H, 9), np.float32)
    assert R.ref_e_candidates(r, idx.ctypes.data_as(ip), H, E_ref.ctypes.data_as(fp), None, None) == 0
    # the winner among the REFERENCE's candidates, chosen by the fp64 restatement of the intended test
    cnt_ref, _ = O.inlier_counts(E_ref.astype(np.float64), x, 1e-6, band=1e-4)
    best_ref = int(O.argmax_first(cnt_ref))
    Ebest = np.ascontiguousarray(E_ref[best_ref])
    R.ref_set_E(r, Ebest.ctypes.data_as(fp))
    t_pc = R.ref_computePosecandidates(r)
    P_ref = np.zeros((4, 4, 4), np.float32)
    R.ref_get_P(r, P_ref.ctypes.data_as(fp))
    t_cp = R.ref_choosePose(r)
    P_ind = R.ref_get_P_ind(r)
    Pinv_ref = np.zeros((4, 4, 4), np.float32)
    R.ref_get_P(r, Pinv_ref.ctypes.data_as(fp))
    t_tr = R.ref_linear_triangulation(r)
    pts = np.zeros((4, n), np.float32)
    R.ref_get_points(r, pts.ctypes.data_as(fp))
    # reference timing of its estimateE body on this input (wall ms, its own mallocs and syncs included)
    t_e = [R.ref_estimateE_injected(r, idx.ctypes.data_as(ip), H, None) for _ in range(5)]
    R.ref_destroy(r)

    E64 = O.hypotheses(x, idx)
    cnt, amb = O.inlier_counts(E64.reshape(H, 9).astype(np.float32).astype(np.float64), x, 1e-6, band=1e-4)
    np.savez_compressed(
        out_path, px=px, score=feat[:, 4], ambiguity=feat[:, 5], match=feat[:, 6].astype(np.int32), match_error=feat[:, 7],
        n2=np.int32(n2), image_wh=np.array([w, h], np.int32), idx=idx, X_ref=X_ref, E_ref=E_ref, counts_ref_f64=cnt_ref.astype(np.int32),
        best_ref=np.int32(best_ref), P_ref=P_ref, P_ind_ref=np.int32(P_ind), Pinv_ref=Pinv_ref, points_ref=pts,
        E64=E64, counts=cnt.astype(np.int32), borderline=amb.astype(np.int32), best=np.int32(O.argmax_first(cnt)),
        ref_ms=np.array([min(t_e), t_pc, t_cp, t_tr], np.float32))
    print(f"{out_path}: {n} correspondences, {H} hypotheses; reference winner {best_ref} ({int(cnt_ref[best_ref])} inliers by fp64), "
          f"fp64 winner {int(np.argmax(cnt))} ({int(cnt.max())}); P_ind_ref {P_ind}; reference ms estimateE {min(t_e):.2f} "
          f"poses {t_pc:.2f} choose {t_cp:.2f} triangulate {t_tr:.2f}")


if __name__ == "__main__":
    main()
