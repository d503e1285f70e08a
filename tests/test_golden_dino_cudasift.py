"""BASELINE config 1 "as the reference runs it": tests/golden/dino_cudasift_000_001.npz holds the correspondences
the reference's OWN front-end produces on its data/dino pair (unmodified CudaSift ExtractSift + MatchSiftData with
src/main.cpp:260-282's parameters, run on a B200 by tests/golden/make_dino_cudasift_fixture.py) and the outputs of
the reference's OWN CUDA path (oracle/_ref/libsfm_ref.so) on them: X, per-hypothesis E, pose candidates, P_ind,
inverted poses, triangulated cloud.  The CUDA path here is compared with those REFERENCE outputs; the fp64 oracle
outputs in the same file pin what the reference cannot (inlier counts, arg-max: SURVEY Q9-Q13)."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
fp, ip, dp = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_double)
THR = 1e-6


def P(a, t=fp):
    return a.ctypes.data_as(t)


@pytest.fixture(scope="module")
def dino(O):
    g = np.load(os.path.join(ROOT, "tests", "golden", "dino_cudasift_000_001.npz"))
    d = {k: g[k] for k in g.files}
    w, h = (int(v) for v in d["image_wh"])
    d["K"], d["Kinv"] = O.reference_K(w, h)
    d["x"] = O.normalise_points(d["px"], d["Kinv"])
    return d


def test_fixture_is_the_reference_operating_point(dino):
    n, H = len(dino["px"]), len(dino["idx"])
    assert (n, H) == (2153, 269) and H == n // 8            # sfm.cu:95: ransac_count = num_points / 8
    assert tuple(dino["image_wh"]) == (720, 576)
    flat = dino["idx"].reshape(-1)
    assert len(set(flat.tolist())) == len(flat) and flat.min() >= 0 and flat.max() < n       # one shuffle, disjoint groups
    px = dino["px"]
    assert px.min() >= 0 and px[:, [0, 2]].max() < 720 and px[:, [1, 3]].max() < 576
    assert np.all((dino["match"] >= 0) & (dino["match"] < int(dino["n2"])))
    assert np.all((dino["ambiguity"] >= 0) & (dino["ambiguity"] <= 1.0 + 1e-6))


def test_oracle_against_reference_outputs(O, oracle_c, dino):
    """The fp64 restatement against what the reference itself computed on this input."""
    x, idx = dino["x"], dino["idx"]
    n, H = len(x), len(idx)
    # fillXU: the reference's cuBLAS product vs the restatement
    assert np.abs(dino["X_ref"][0][:2].T - x[:, :2]).max() < 3e-8 and np.abs(dino["X_ref"][1][:2].T - x[:, 2:]).max() < 3e-8
    assert np.all(dino["X_ref"][:, 2] == 1)
    # per-hypothesis E: the reference's own error (its null vector + 4-sweep normalizeE) bounds the agreement
    E = O.hypotheses(x, idx)
    assert O.e_distance(E, dino["E64"]).max() < 1e-9
    # (on real matches the 8x9 design matrices have sigma_8 / sigma_1 ~ 1e-5: the reference's un-normalised fp32 solve is
    # within 1e-4 of the fp64 truth for 77 % of the rows, within 1e-3 for 96 %, within 1e-2 for 99 %)
    d = O.e_distance(dino["E_ref"], E)
    assert np.mean(d < 1e-4) >= 0.75 and np.mean(d < 1e-3) >= 0.95 and np.mean(d < 1e-2) >= 0.99, (np.mean(d < 1e-4), np.mean(d < 1e-2))
    # counts / arg-max of the fp64 candidates
    cnt, _ = O.inlier_counts(dino["E64"].reshape(H, 9).astype(np.float32).astype(np.float64), x, THR)
    assert np.array_equal(cnt, dino["counts"]) and O.argmax_first(cnt) == int(dino["best"])
    # pose stage on the reference's winning E: candidates in the reference's order, its P_ind, its inverses, its cloud
    Eb = dino["E_ref"][int(dino["best_ref"])].reshape(3, 3).astype(np.float64)
    Pc = O.pose_candidates(Eb, compat=True)
    assert np.abs(Pc - dino["P_ref"]).max() < 5e-3
    ind, Pinv = O.choose_pose(x.astype(np.float64), Pc, compat=True)
    assert ind == int(dino["P_ind_ref"])
    assert np.abs(Pinv - dino["Pinv_ref"]).max() < 5e-3
    pts = O.triangulate(x.astype(np.float64), dino["Pinv_ref"][ind].astype(np.float64))
    rel = np.abs(pts[:3] - dino["points_ref"][:3]).max(axis=0) / np.maximum(np.abs(dino["points_ref"][:3]).max(axis=0), 1e-3)
    assert np.median(rel) < 1e-4 and np.mean(rel < 1e-2) > 0.95, (np.median(rel), np.mean(rel < 1e-2))


def test_host_solvers_on_cudasift_rows(lib, O, dino):
    x, idx = dino["x"], dino["idx"]
    H = len(idx)
    for name in ("sfmb200_host_solve_hypothesis", "sfmb200_host_solve_hypothesis_projector"):
        E32 = np.zeros((H, 9), np.float32)
        for h in range(H):
            p = np.ascontiguousarray(x[idx[h]], dtype=np.float32)
            lib.raw(name)(P(p), P(E32[h]))
        d = O.e_distance(E32, dino["E64"])
        assert np.mean(d < 1e-4) >= 0.97 and np.median(d) < 1e-5, (name, np.mean(d < 1e-4), np.median(d))


@pytest.mark.gpu
def test_cuda_path_against_reference_outputs(pkg, O, oracle_c, dino):
    """Whole path through the facade-shaped mirror (SiftPoint ingest like main.cpp:298-299) against the stored outputs of
    the reference's CUDA path, stage by stage, then end to end with NOTHING injected."""
    import torch

    n, H = len(dino["px"]), len(dino["idx"])
    x = dino["x"]
    ipair = pkg.ImagePair(dino["K"], dino["Kinv"], 2, n)          # default capacity H = N/8, like the reference
    sift = np.zeros((n, 144), np.float32)
    sift[:, 0], sift[:, 1], sift[:, 9], sift[:, 10] = dino["px"].T
    sift[:, 6], sift[:, 7] = dino["score"], dino["ambiguity"]
    d_sift = torch.from_numpy(sift).cuda()
    ipair.fillXU(d_sift)
    # fillXU vs the reference's X (cuBLAS Sgemm, k = 3): 1 ulp
    for image in (0, 1):
        Xg = ipair.get_X(image).cpu().numpy()
        assert np.abs(Xg - dino["X_ref"][image]).max() < 3e-8
    d_idx = torch.from_numpy(np.ascontiguousarray(dino["idx"])).cuda()
    stats = {}
    for solver in (1, 0):
        ipair.set_option(5, solver)
        ipair.estimateE(H, 0, THR, d_idx=d_idx)
        Eg = ipair.get_E_candidates().cpu().numpy()
        d_ref, d_64, dr64 = O.e_distance(Eg, dino["E_ref"]), O.e_distance(Eg, dino["E64"]), O.e_distance(dino["E_ref"], dino["E64"])
        stats[solver] = (np.mean(d_ref < 1e-4), np.mean(d_64 < 1e-4), np.mean(dr64 < 1e-4))
        # north_star: per-hypothesis E within 1e-4 of the reference.  Where that fails the reference is the one away from
        # the fp64 truth (its own error bounds the agreement): ours is within 1e-4 of fp64 at least as often, and
        # whenever the two disagree ours is the closer one.
        # On THIS input the reference is within 1e-4 of fp64 for only 77 % of the rows (ill-conditioned real-match samples,
        # no Hartley normalisation in the reference), so that is the ceiling of any agreement with it.
        assert np.mean(d_64 < 1e-4) >= 0.97 and np.mean(d_64 < 1e-4) >= np.mean(dr64 < 1e-4)
        bad = d_ref >= 1e-4
        assert np.mean(d_ref < 1e-4) >= np.mean(dr64 < 1e-4) - 0.03 and np.mean(d_ref < 1e-2) >= 0.98
        if bad.any():
            assert np.mean(d_64[bad] <= dr64[bad]) >= 0.9
        # counts: bit-exact vs the fp32 port on the device's own E; arg-max exact
        got = ipair.get_inlier_counts().cpu().numpy()
        X0, X1 = ipair.get_X(0).cpu().numpy(), ipair.get_X(1).cpu().numpy()
        xg = np.ascontiguousarray(np.stack([X0[0], X0[1], X1[0], X1[1]], 1))
        c32 = np.zeros(H, np.int32)
        oracle_c.oracle_counts_f32(P(Eg), H, P(xg), n, C.c_float(THR), P(c32, ip))
        assert np.array_equal(got, c32)
        ok = d_64 < 1e-5
        assert np.all(np.abs(got[ok] - dino["counts"][ok]) <= dino["borderline"][ok] + 3)
        bi, bc = ipair.get_best()
        assert bc[0] == got.max() and bi[0] == int(np.argmax(got))
        # the same winner as the fp64 oracle and as the reference's candidates scored by the oracle
        assert int(bi[0]) == int(dino["best"]) == int(dino["best_ref"])
        assert abs(int(bc[0]) - int(dino["counts"].max())) <= 3
    print(f"\ndino / CudaSift input, {H} hypotheses: within 1e-4 (ours vs reference, ours vs fp64, reference vs fp64): "
          f"projector {stats[1][0]:.4f} {stats[1][1]:.4f} {stats[1][2]:.4f}; Jacobi {stats[0][0]:.4f} {stats[0][1]:.4f} {stats[0][2]:.4f}")
    # ---- pose stage, nothing injected: our winner E (vs the reference's winner E: same hypothesis) ----
    ipair.set_option(5, 1)
    ipair.estimateE(H, 0, THR, d_idx=d_idx)
    E_g = ipair.get_E()[0].reshape(9)
    assert O.e_distance(E_g[None], dino["E_ref"][int(dino["best_ref"])][None])[0] < 1e-3
    ipair.computePosecandidates()
    Pg = ipair.get_poses()[0]
    # E is defined up to sign and the two sides' winners may differ by that sign: compare through the reference's E when they do
    if np.dot(E_g, dino["E_ref"][int(dino["best_ref"])]) < 0:
        ipair.set_E(-E_g)
        ipair.computePosecandidates()
        Pg = ipair.get_poses()[0]
    assert np.abs(Pg - dino["P_ref"]).max() < 5e-3, np.abs(Pg - dino["P_ref"]).reshape(4, -1).max(axis=1)
    ipair.choosePose()
    assert int(ipair.get_pose_index()[0]) == int(dino["P_ind_ref"])
    assert np.abs(ipair.get_poses()[0] - dino["Pinv_ref"]).max() < 5e-3
    ipair.linear_triangulation()
    pts = ipair.get_points_host()
    assert pts.shape == (4, n) and np.all(np.isfinite(pts)) and np.all(pts[3] == 1)
    rel = np.abs(pts[:3] - dino["points_ref"][:3]).max(axis=0) / np.maximum(np.abs(dino["points_ref"][:3]).max(axis=0), 1e-3)
    mask = ipair.get_inlier_mask().cpu().numpy().astype(bool)
    print(f"cloud vs the reference's (each side its own E, pose and solve): median rel {np.median(rel[mask]):.2e} over {mask.sum()} inliers, "
          f"within 1e-2: {np.mean(rel[mask] < 1e-2):.3f}")
    assert np.median(rel[mask]) < 2e-2
    # with the reference's selected pose injected the solve itself is compared: 1e-4 median
    ref_pose = dino["Pinv_ref"].copy()
    o = O.triangulate(x.astype(np.float64), ref_pose[int(dino["P_ind_ref"])].astype(np.float64))
    relo = np.abs(dino["points_ref"][:3] - o[:3]).max(axis=0) / np.maximum(np.abs(o[:3]).max(axis=0), 1e-3)
    assert np.median(relo) < 1e-4
    # the match filter of the facade on the real scores (CudaSift's own test, matching.cu:1035)
    kept = ipair.fillXU(d_sift, min_score=0.85, max_ambiguity=0.95)
    assert kept == int(((dino["score"] > np.float32(0.85)) & (dino["ambiguity"] < np.float32(0.95))).sum())
    ipair.close()
