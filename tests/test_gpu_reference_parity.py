"""GPU tier: our CUDA path against the REFERENCE'S OWN CUDA path, rebuilt
unmodified for sm_100a as oracle/_ref/libsfm_ref.so (oracle/ref_harness.cu),
fed identical inputs stage by stage.  Stages whose reference
behaviour is undefined (inlier scoring reads uninitialised memory, arg-max is
off by one: SURVEY Q9-Q13) are pinned by the fp64 oracle in test_gpu_parity.py
instead."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

fp = C.POINTER(C.c_float)
ip = C.POINTER(C.c_int32)


def P(a, t=fp):
    return a.ctypes.data_as(t)


@pytest.fixture(scope="module")
def ref_pair(ref_lib, scene_small):
    import torch

    assert torch.cuda.is_available()
    n = len(scene_small["px"])
    r = ref_lib.ref_create(P(scene_small["K"].reshape(9).copy()), P(scene_small["Kinv"].reshape(9).copy()), n)
    assert ref_lib.ref_fillXU(C.c_void_p(r), P(scene_small["px"])) == 0
    yield C.c_void_p(r)
    ref_lib.ref_destroy(C.c_void_p(r))


def test_fillXU_matches_reference(pkg, ref_lib, ref_pair, scene_small):
    import torch

    n = len(scene_small["px"])
    h = pkg.BatchedPairs(scene_small["K"], scene_small["Kinv"], 1, n, 16)
    h.set_points_xy(torch.from_numpy(scene_small["px"]).cuda())
    for image in (0, 1):
        Xr = np.zeros((3, n), np.float32)
        assert ref_lib.ref_get_X(ref_pair, image, P(Xr)) == 0
        Xg = h.get_X(image).cpu().numpy()
        # cuBLAS Sgemm (k = 3) vs our two fmas: same value to 1 ulp of 0.15
        assert np.abs(Xg - Xr).max() < 3e-8
        assert np.all(Xr[2] == 1)
    h.close()


def test_E_candidates_match_reference(pkg, O, ref_lib, ref_pair, scene_small):
    """north_star: per-hypothesis E matches up to sign and scale within 1e-4
    relative Frobenius norm, same sample rows on both sides."""
    import torch

    x, n = scene_small["x"], len(scene_small["x"])
    H = 2000
    idx = O.sample_indices(1237, H, n)
    Er = np.zeros((H, 9), np.float32)
    Ar = np.zeros((H, 72), np.float32)
    Vr = np.zeros((H, 81), np.float32)
    assert ref_lib.ref_e_candidates(ref_pair, P(idx, ip), H, P(Er), P(Ar), P(Vr)) == 0
    # does gesvdjBatched (m=8 < n=9) put the null vector in V's 9th column?
    null = Vr[:, 72:81].astype(np.float64)
    resid = np.linalg.norm(np.einsum("hij,hj->hi", Ar.reshape(H, 8, 9).astype(np.float64), null), axis=1)
    assert np.median(resid) < 1e-5, f"reference null-vector residual {np.median(resid)}"
    h = pkg.BatchedPairs(scene_small["K"], scene_small["Kinv"], 1, n, H)
    h.set_points_xy(torch.from_numpy(scene_small["px"]).cuda())
    h.estimate_e(H, 0, 1e-6, d_idx=torch.from_numpy(idx).cuda())
    Eg = h.get_E_candidates().cpu().numpy()
    d = O.e_distance(Eg, Er)
    E64 = O.hypotheses(x, idx)
    d_ours, d_ref = O.e_distance(Eg, E64), O.e_distance(Er, E64)
    print(f"\nE parity vs reference: median {np.median(d):.2e} p99 {np.percentile(d, 99):.2e} "
          f"within 1e-4: {np.mean(d < 1e-4):.4f}; vs fp64 ours {np.mean(d_ours < 1e-4):.4f} ref {np.mean(d_ref < 1e-4):.4f}")
    # Measured on B200: the reference itself is within 1e-4 of the fp64 truth for only
    # ~97 % of hypotheses (its null vector for 98.6 %, and normalizeE's 4-sweep
    # approximate svd.h loses more), ours for 99.97 %.  So: >= 95 % pairwise agreement,
    # ours within 1e-4 of fp64 wherever the reference is, and every disagreement is
    # the reference's error, not ours.
    assert np.mean(d < 1e-4) >= 0.95
    assert np.mean(d_ours < 1e-4) >= 0.995 and np.mean(d_ours < 1e-4) >= np.mean(d_ref < 1e-4)
    bad = d >= 1e-4
    if bad.any():
        assert np.mean(d_ours[bad] <= d_ref[bad]) > 0.95
    h.close()


def test_pose_and_triangulation_match_reference(pkg, O, ref_lib, ref_pair, scene_small):
    """Same E injected into both sides.  Candidates (incl. the det typo, Q15)
    match as a set up to the one discrete SVD freedom (oracle.match_candidates);
    with identical candidates injected, choosePose (in-place inversion,
    last-passing-index rule, Q17-Q18) gives the same index and inverses, and
    triangulation with that inverted pose (Q19) the same points."""
    import torch

    x, n = scene_small["x"], len(scene_small["x"])
    h = pkg.BatchedPairs(scene_small["K"], scene_small["Kinv"], 1, n, 4096)
    h.set_points_xy(torch.from_numpy(scene_small["px"]).cuda())
    h.estimate_e(4096, 5, 1e-6)
    Ebest = h.get_E()[0].reshape(9)
    cands = h.get_E_candidates().cpu().numpy()
    perms = []
    for E in [Ebest, -Ebest] + [cands[i] for i in (1, 10, 100, 1000, 2000, 3000)]:
        E = np.ascontiguousarray(E, np.float32)
        if not np.any(E):
            continue
        # 1. candidates from the same E
        assert ref_lib.ref_set_E(ref_pair, P(E)) == 0
        ref_lib.ref_computePosecandidates(ref_pair)
        Pr = np.zeros((4, 4, 4), np.float32)
        ref_lib.ref_get_P(ref_pair, P(Pr))
        h.set_E(E)
        h.pose_candidates()
        Pg = h.get_poses()[0].copy()
        # the reference's host svd() runs 4 approximate Jacobi sweeps: ~1e-3 agreement
        perm = O.match_candidates(Pg, Pr, tol=5e-3)
        assert perm is not None, (Pg, Pr)
        assert perm[0] == 0, "candidates must come in the reference's order (svd3_reference_orientation)"
        perms.append(perm[0])
        # 2. choosePose on IDENTICAL candidates (ours injected into the reference)
        assert ref_lib.ref_set_P(ref_pair, P(Pg)) == 0
        ref_lib.ref_choosePose(ref_pair)
        ind_r = ref_lib.ref_get_P_ind(ref_pair)
        ref_lib.ref_get_P(ref_pair, P(Pr))
        h.choose_pose()
        assert int(h.get_pose_index()[0]) == ind_r
        assert np.abs(h.get_poses()[0] - Pr).max() < 1e-4       # both hold the inverses now
        # 3. triangulation with the selected (inverted) pose
        ref_lib.ref_linear_triangulation(ref_pair)
        pr = np.zeros((4, n), np.float32)
        ref_lib.ref_get_points(ref_pair, P(pr))
        h.triangulate()
        pg = h.get_points_host()
        rel = np.abs(pg[:3] - pr[:3]).max(axis=0) / np.maximum(np.abs(pr[:3]).max(axis=0), 1e-3)
        inl = ~scene_small["is_outlier"]
        print(f"\ntriangulation vs reference: median rel {np.median(rel[inl]):.2e}, p99 {np.percentile(rel[inl], 99):.2e}, "
              f"all points within 1e-3: {np.mean(rel < 1e-3):.4f}")
        # stated tolerance: 1e-3 relative to the point's largest coordinate (fp32
        # Jacobi on both sides); the tail is ill-conditioned outlier geometry
        assert np.median(rel[inl]) < 1e-4 and np.mean(rel[inl] < 1e-3) > 0.98
        assert np.all(pg[3] == 1) and np.all(pr[3] == 1)
        # 4. egress
        pos_r = np.zeros((n, 4), np.float32)
        col_r = np.zeros((n, 4), np.float32)
        ref_lib.ref_vbo(ref_pair, P(pos_r), P(col_r))
        pos = torch.empty((n, 4), device="cuda")
        col = torch.empty((n, 4), device="cuda")
        h.copy_to_vbo(pos, col)
        assert np.array_equal(col.cpu().numpy(), col_r)
        assert np.array_equal(pos_r[:, :3], pr[:3].T) and np.array_equal(pos.cpu().numpy()[:, :3], pg[:3].T)
    print("candidate permutation per trial (0 = identity, 3 = swapped):", perms)
    h.close()


def test_end_to_end_pose_selection_own_candidates(pkg, O, ref_lib, ref_pair, scene_small):
    """every E goes through BOTH pipelines from E onwards - each side computes its OWN SVD, its own four
    candidates, its own cheirality test, picks its own index and triangulates with its own selected pose - and the
    physically selected pose and the cloud are compared.  This needs the candidate ORDER to agree, i.e. the discrete
    freedom of the SVD fixed the way the reference's svd() fixes it (smallmat.cuh: svd3_reference_orientation).
    1,024 essential matrices: the RANSAC candidates of the scene (good and bad models alike) and their negatives."""
    import torch

    x, n = scene_small["x"], len(scene_small["x"])
    hgen = pkg.BatchedPairs(scene_small["K"], scene_small["Kinv"], 1, n, 512)
    d_px = torch.from_numpy(scene_small["px"]).cuda()
    hgen.set_points_xy(d_px)
    hgen.estimate_e(512, 5, 1e-6)
    cands = hgen.get_E_candidates().cpu().numpy()
    hgen.close()
    Es = np.concatenate([cands, -cands]).astype(np.float32)
    Es = np.ascontiguousarray(Es[np.any(Es != 0, axis=1)])
    B = len(Es)
    assert B >= 1000
    # ours: one batched handle, pair b = the same correspondences with E_b injected
    h = pkg.BatchedPairs(scene_small["K"], scene_small["Kinv"], B, n, 8)
    h.set_points_xy(d_px.unsqueeze(0).expand(B, n, 4).contiguous())
    h.set_E(Es)
    h.pose_candidates()
    Pg = h.get_poses().copy()                    # [B][4][4][4] candidates
    h.choose_pose()
    ind_g = h.get_pose_index().copy()
    Pg_inv = h.get_poses().copy()                # inverses (Q18)
    h.triangulate()
    same_cand, same_ind, pose_err, cloud_med = 0, 0, [], []
    inl = ~scene_small["is_outlier"]
    mism = []
    for b in range(B):
        E = np.ascontiguousarray(Es[b])
        assert ref_lib.ref_set_E(ref_pair, P(E)) == 0
        ref_lib.ref_computePosecandidates(ref_pair)
        Pr = np.zeros((4, 4, 4), np.float32)
        ref_lib.ref_get_P(ref_pair, P(Pr))
        ref_lib.ref_choosePose(ref_pair)
        ind_r = ref_lib.ref_get_P_ind(ref_pair)
        Pr_inv = np.zeros((4, 4, 4), np.float32)
        ref_lib.ref_get_P(ref_pair, P(Pr_inv))
        # candidates index by index (the reference's host svd() runs 4 approximate sweeps: ~1e-3 agreement)
        ok_c = all(np.abs(Pg[b, i] - Pr[i]).max() < 5e-3 for i in range(4))
        same_cand += ok_c
        same_ind += int(ind_g[b]) == ind_r
        if int(ind_g[b]) != ind_r or not ok_c:
            mism.append((b, int(ind_g[b]), ind_r, bool(ok_c)))
            continue
        pose_err.append(np.abs(Pg_inv[b, ind_r] - Pr_inv[ind_r]).max())
        if b % 16 == 0:                             # the cloud under each side's own selected pose
            ref_lib.ref_linear_triangulation(ref_pair)
            pr = np.zeros((4, n), np.float32)
            ref_lib.ref_get_points(ref_pair, P(pr))
            pg = h.get_points_host(b)
            rel = np.abs(pg[:3] - pr[:3]).max(axis=0) / np.maximum(np.abs(pr[:3]).max(axis=0), 1e-3)
            cloud_med.append(np.median(rel[inl]))
    print(f"\nend-to-end pose selection, {B} essential matrices, each side on its own candidates: candidates equal index by index "
          f"{same_cand}/{B}, same P_ind {same_ind}/{B}, selected pose max abs diff median {np.median(pose_err):.2e} max {np.max(pose_err):.2e}, "
          f"cloud median rel diff (inliers) median {np.median(cloud_med):.2e} max {np.max(cloud_med):.2e}; mismatches {mism[:8]}")
    # stated tolerances: candidates / selected pose 5e-3 (the reference's own SVD is a 4-sweep approximation), P_ind exact;
    # the cloud inherits the pose difference amplified by the triangulation geometry: median relative 2e-2
    assert same_cand >= B - 2 and same_ind >= B - 2, mism
    assert np.max(pose_err) < 5e-3
    assert np.median(cloud_med) < 2e-2
    h.close()


def test_reference_injected_estimateE_runs(O, ref_lib, ref_pair, scene_small):
    """The timing baseline used by bench.py --impl reference: the reference's
    estimateE body with injected rows.  (Its as-built estimateE() cannot run on
    B200 at all: kernels::kernels' `index > ransac_iterations` guard, kernels.h:242,
    lets one extra thread gather through out-of-range sample indices and the
    launch dies with an illegal memory access - SURVEY Q5; the harness pads the
    index and design-matrix buffers by one row instead.)"""
    n = len(scene_small["x"])
    H = 512
    idx = O.sample_indices(3, H, n)
    best = C.c_int(-1)
    ms = ref_lib.ref_estimateE_injected(ref_pair, P(idx, ip), H, C.byref(best))
    assert ms > 0 and 0 <= best.value < H
    E = np.zeros(9, np.float32)
    ref_lib.ref_get_E(ref_pair, P(E))
    assert np.all(np.isfinite(E))
