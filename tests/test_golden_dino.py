"""BASELINE config 1 as a committed fixture (tests/golden/dino_000_001.npz, made by
tests/golden/make_dino_fixture.py from the reference's data/dino pair): real-image
correspondences, H = N/8 disjoint sample rows from one permutation like
sfm.cu:95-104, golden outputs from the fp64 oracle."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
fp, ip, dp = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_double)


def P(a, t=fp):
    return a.ctypes.data_as(t)


@pytest.fixture(scope="module")
def dino(O):
    g = np.load(os.path.join(ROOT, "tests", "golden", "dino_000_001.npz"))
    K, Kinv = O.reference_K(720, 576)
    d = {k: g[k] for k in g.files}
    d["K"], d["Kinv"] = K, Kinv
    d["x"] = O.normalise_points(d["px"], Kinv)
    return d


def test_fixture_shape(dino):
    n, H = len(dino["px"]), len(dino["idx"])
    assert n == 1577 and H == n // 8 == 197
    flat = dino["idx"].reshape(-1)
    assert len(set(flat.tolist())) == len(flat)          # disjoint samples, like one std::shuffle
    assert dino["px"].min() >= 0 and dino["px"][:, [0, 2]].max() < 720 and dino["px"][:, [1, 3]].max() < 576


def test_oracle_reproduces_golden(O, oracle_c, dino):
    x, idx = dino["x"], dino["idx"]
    H = len(idx)
    E = O.hypotheses(x, idx)
    assert O.e_distance(E, dino["E64"]).max() < 1e-9
    Ec = np.zeros((H, 9))
    oracle_c.oracle_hypotheses_f64(P(x), len(x), P(np.ascontiguousarray(idx), ip), H, P(Ec, dp))
    assert O.e_distance(Ec, dino["E64"]).max() < 1e-8
    cnt, _ = O.inlier_counts(dino["E64"].reshape(H, 9).astype(np.float32).astype(np.float64), x, 1e-6)
    assert np.array_equal(cnt, dino["counts"]) and O.argmax_first(cnt) == int(dino["best"])


def test_host_solver_on_dino(lib, O, dino):
    x, idx = dino["x"], dino["idx"]
    H = len(idx)
    E32 = np.zeros((H, 9), np.float32)
    for h in range(H):
        p = np.ascontiguousarray(x[idx[h]], dtype=np.float32)
        lib.raw("sfmb200_host_solve_hypothesis")(P(p), P(E32[h]))
    d = O.e_distance(E32, dino["E64"])
    assert np.mean(d < 1e-4) >= 0.98 and np.median(d) < 1e-5, (np.mean(d < 1e-4), np.median(d))


@pytest.mark.gpu
def test_cuda_path_on_dino(pkg, O, oracle_c, dino):
    import torch

    n, H = len(dino["px"]), len(dino["idx"])
    ipair = pkg.ImagePair(dino["K"], dino["Kinv"], 2, n)          # default capacity H = N/8, like the reference
    sift = np.zeros((n, 144), np.float32)
    sift[:, 0], sift[:, 1], sift[:, 9], sift[:, 10] = dino["px"].T
    ipair.fillXU(torch.from_numpy(sift).cuda())
    ipair.estimateE(H, 0, 1e-6, d_idx=torch.from_numpy(np.ascontiguousarray(dino["idx"])).cuda())
    Eg = ipair.get_E_candidates().cpu().numpy()
    d = O.e_distance(Eg, dino["E64"])
    assert np.mean(d < 1e-4) >= 0.98 and np.median(d) < 1e-5, (np.mean(d < 1e-4), np.median(d))
    got = ipair.get_inlier_counts().cpu().numpy()
    X0, X1 = ipair.get_X(0).cpu().numpy(), ipair.get_X(1).cpu().numpy()
    xg = np.ascontiguousarray(np.stack([X0[0], X0[1], X1[0], X1[1]], 1))
    c32 = np.zeros(H, np.int32)
    oracle_c.oracle_counts_f32(P(Eg), H, P(xg), n, C.c_float(1e-6), P(c32, ip))
    assert np.array_equal(got, c32)                                 # bit-exact vs the fp32 port
    # golden counts were computed from the fp64 E: allow the borderline band + the effect of E's fp32 rounding
    ok = d < 1e-5
    assert np.all(np.abs(got[ok] - dino["counts"][ok]) <= dino["borderline"][ok] + 3)
    bi, bc = ipair.get_best()
    assert bc[0] == got.max() and bi[0] == int(np.argmax(got))
    assert abs(int(bc[0]) - int(dino["counts"].max())) <= 3
    ipair.computePosecandidates()
    ipair.choosePose()
    ipair.linear_triangulation()
    pts = ipair.get_points_host()
    assert pts.shape == (4, n) and np.all(np.isfinite(pts)) and np.all(pts[3] == 1)


@pytest.mark.gpu
def test_reference_cuda_path_on_dino(O, ref_lib, dino):
    """The reference's own K3-K7 chain on the same rows (its real operating point)."""
    n, H = len(dino["px"]), len(dino["idx"])
    r = C.c_void_p(ref_lib.ref_create(P(dino["K"].reshape(9).copy()), P(dino["Kinv"].reshape(9).copy()), n))
    assert ref_lib.ref_fillXU(r, P(dino["px"])) == 0
    Er = np.zeros((H, 9), np.float32)
    assert ref_lib.ref_e_candidates(r, P(np.ascontiguousarray(dino["idx"]), ip), H, P(Er), None, None) == 0
    d = O.e_distance(Er, dino["E64"])
    print(f"\nreference on dino: within 1e-4 of fp64: {np.mean(d < 1e-4):.3f}, median {np.median(d):.2e}")
    assert np.mean(d < 1e-3) >= 0.9
    ref_lib.ref_destroy(r)
