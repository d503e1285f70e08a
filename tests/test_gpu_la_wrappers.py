"""GPU tier: the kernels.h-equivalent linear-algebra entry points
(include/sfmb200_la.h) against numpy, with the reference's own test literals
(SfM/sfm.cu:389-510) and its call-site shapes (3x3 * 3xN, 8x9 and 4x4 batched
SVD in cuSOLVER's column-major convention, 4x4 inverse)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def dp(t):
    return C.c_void_p(t.data_ptr())


@pytest.fixture(scope="module")
def T():
    import torch

    assert torch.cuda.is_available()
    return torch


def test_mmul_family(lib, T):
    rng = np.random.default_rng(0)
    A, B = rng.normal(size=(3, 3)).astype(np.float32), rng.normal(size=(3, 1000)).astype(np.float32)
    dA, dB = T.from_numpy(A).cuda(), T.from_numpy(B).cuda()
    dC = T.empty((3, 1000), device="cuda")
    lib.call("sfmb200_la_mmul", dp(dA), dp(dB), dp(dC), 3, 3, 1000, None)        # fillXU: K^-1 * U (sfm.cu:88)
    assert np.allclose(dC.cpu().numpy(), A @ B, atol=1e-6)
    # batched E_h * X with a shared right operand (calculateInliers, sfm.cu:174)
    E = rng.normal(size=(5, 3, 3)).astype(np.float32)
    dE = T.from_numpy(E).cuda()
    dO = T.empty((5, 3, 1000), device="cuda")
    lib.call("sfmb200_la_mmul_batched", dp(dE), dp(dB), dp(dO), 3, 3, 1000, 9, 0, 3000, 5, None)
    assert np.allclose(dO.cpu().numpy(), E @ B, atol=1e-5)
    # X^T * E_h with X stored 3 x N (sfm.cu:178)
    dO2 = T.empty((5, 1000, 3), device="cuda")
    lib.call("sfmb200_la_mmul_transpose_batched", dp(dB), dp(dE), dp(dO2), 1000, 3, 3, 0, 9, 3000, 5, None)
    assert np.allclose(dO2.cpu().numpy(), np.einsum("kn,bkj->bnj", B, E), atol=1e-5)
    # literal of testBatchedmultTranspose (sfm.cu:467-489)
    A2 = np.array([1, 2, 3, 1, 4, 5, 6, 1, 7, 8, 9, 1, 0, 1, 2, 1, 3, 4, 5, 1, 6, 7, 8, 1], np.float32)
    B2 = np.arange(1, 10, dtype=np.float32)
    dA2, dB2, dC2 = T.from_numpy(A2).cuda(), T.from_numpy(B2).cuda(), T.empty(24, device="cuda")
    lib.call("sfmb200_la_mmul_transpose_batched", dp(dA2), dp(dB2), dp(dC2), 4, 3, 3, 12, 0, 12, 2, None)
    want = np.stack([A2[:12].reshape(3, 4).T @ B2.reshape(3, 3), A2[12:].reshape(3, 4).T @ B2.reshape(3, 3)])
    assert np.array_equal(dC2.cpu().numpy().reshape(2, 4, 3), want)


def test_invert_and_singular(lib, T, pkg):
    rng = np.random.default_rng(1)
    M = rng.normal(size=(7, 4, 4)).astype(np.float32) + 3 * np.eye(4, dtype=np.float32)
    dM, dI = T.from_numpy(M).cuda(), T.empty((7, 4, 4), device="cuda")
    lib.call("sfmb200_la_invert", dp(dM), dp(dI), 4, 7, None)
    assert np.allclose(dI.cpu().numpy() @ M, np.eye(4), atol=1e-4)
    lib.call("sfmb200_la_invert", dp(dM), dp(dM), 4, 7, None)                    # in place, like choosePose (sfm.cu:286)
    assert np.allclose(dM.cpu().numpy(), dI.cpu().numpy())
    a = np.array([1, 2, 0, 0, 2, 0, 1, 2, 1], np.float32)                        # testInverse literal (sfm.cu:444-445)
    da, db = T.from_numpy(a).cuda(), T.empty(9, device="cuda")
    lib.call("sfmb200_la_invert", dp(da), dp(db), 3, 1, None)
    assert np.allclose(db.cpu().numpy().reshape(3, 3), np.linalg.inv(a.reshape(3, 3)), atol=1e-6)
    dz = T.zeros(16, device="cuda")
    with pytest.raises(pkg.SfmError) as e:
        lib.call("sfmb200_la_invert", dp(dz), dp(dz), 4, 1, None)
    assert e.value.code == -5                                                     # the reference exit()s here


@pytest.mark.parametrize("m,n", [(4, 4), (8, 9), (3, 3), (9, 8)])
def test_svd_batched_cusolver_convention(lib, T, m, n):
    rng = np.random.default_rng(2)
    batch = 50
    A = rng.normal(size=(batch, m, n)).astype(np.float32)
    Acm = np.ascontiguousarray(np.transpose(A, (0, 2, 1)))                       # column-major storage, lda = m
    dA = T.from_numpy(Acm).cuda()
    mn = min(m, n)
    dS, dU, dV = T.empty((batch, mn), device="cuda"), T.empty((batch, m, m), device="cuda"), T.empty((batch, n, n), device="cuda")
    lib.call("sfmb200_la_svd_batched", dp(dA), dp(dS), dp(dU), dp(dV), m, n, batch, None)
    S = dS.cpu().numpy()
    U = np.transpose(dU.cpu().numpy(), (0, 2, 1))                                # column-major -> math
    V = np.transpose(dV.cpu().numpy(), (0, 2, 1))
    sv = np.linalg.svd(A.astype(np.float64), compute_uv=False)
    assert np.allclose(S, sv, atol=2e-5 * sv.max())
    assert np.all(S[:, :-1] >= S[:, 1:] - 1e-6)                                   # sorted like gesvdj with sort_svd = 1
    for b in range(batch):
        Sm = np.zeros((m, n))
        Sm[:mn, :mn] = np.diag(S[b])
        assert np.abs(U[b] @ Sm @ V[b].T - A[b]).max() < 5e-5
        assert np.abs(U[b].T @ U[b] - np.eye(m)).max() < 1e-4 and np.abs(V[b].T @ V[b] - np.eye(n)).max() < 1e-4
    if (m, n) == (8, 9):
        # the reference's use: null vector = 9th column of V = floats 72..80 of the column-major block (kernels.h:457)
        dE = T.empty((batch, 9), device="cuda")
        lib.call("sfmb200_la_row_extraction", dp(dV), dp(dE), batch, None)
        e = dE.cpu().numpy()
        assert np.abs(np.einsum("bij,bj->bi", A, e)).max() < 1e-4
        assert np.array_equal(e, dV.cpu().numpy().reshape(batch, 81)[:, 72:81])


def test_transpose_batched(lib, T):
    rng = np.random.default_rng(5)
    A = rng.normal(size=(1000, 8, 9)).astype(np.float32)
    dA, dO = T.from_numpy(A).cuda(), T.empty((1000, 9, 8), device="cuda")
    lib.call("sfmb200_la_transpose_batched", dp(dA), dp(dO), 8, 9, 1000, None)
    assert np.array_equal(dO.cpu().numpy(), np.transpose(A, (0, 2, 1)))


def test_elementwise_vecnorm_threshold_argmax(lib, T):
    rng = np.random.default_rng(3)
    a, b = rng.normal(size=5000).astype(np.float32), rng.normal(size=5000).astype(np.float32)
    b[::7] = 0
    for op, f in ((0, lambda x, y: x * y), (2, lambda x, y: x + y)):
        da, db = T.from_numpy(a.copy()).cuda(), T.from_numpy(b).cuda()
        lib.call("sfmb200_la_elementwise", op, dp(da), dp(db), 5000, None)
        assert np.array_equal(da.cpu().numpy(), f(a, b))
    da, db = T.from_numpy(a.copy()).cuda(), T.from_numpy(b).cuda()
    lib.call("sfmb200_la_elementwise", 1, dp(da), dp(db), 5000, None)            # div with the b == 0 -> 0 guard (kernels.h:311)
    want = np.where(b == 0, 0, a / np.where(b == 0, 1, b)).astype(np.float32)
    assert np.allclose(da.cpu().numpy(), want, rtol=1e-6)
    # testVecnorm literal (sfm.cu:503-510): columns of [1 2 3; 4 5 6; 7 8 9], p = 2 -> 8.124, 9.644, 11.225
    t = T.arange(1, 10, dtype=T.float32, device="cuda")
    r = T.empty(3, device="cuda")
    lib.call("sfmb200_la_vecnorm", dp(t), dp(r), 3, 3, C.c_float(2), C.c_float(1), None)
    assert np.allclose(r.cpu().numpy(), [8.1240384, 9.6436508, 11.224972], rtol=1e-6)
    lib.call("sfmb200_la_vecnorm", dp(t), dp(r), 3, 3, C.c_float(2), C.c_float(2), None)   # exp == final_pow: sum of squares
    assert np.allclose(r.cpu().numpy(), [66, 93, 126])
    res = rng.uniform(0, 2e-6, size=(40, 333)).astype(np.float32)
    dres, dcnt = T.from_numpy(res).cuda(), T.empty(40, dtype=T.int32, device="cuda")
    lib.call("sfmb200_la_threshold_count", dp(dres), dp(dcnt), 333, 40, C.c_float(1e-6), None)
    assert np.array_equal(dcnt.cpu().numpy(), (res < np.float32(1e-6)).sum(1))
    v = T.tensor([1, 2, 3, 4, 5, 6, 4, 1, 3], dtype=T.int32, device="cuda")       # testThrust_max: 6 at position 5
    idx = C.c_int32(-1)
    lib.call("sfmb200_la_argmax_first", dp(v), 6, C.byref(idx), None)
    assert idx.value == 5
    v2 = T.tensor([-5, -2, -2, -9], dtype=T.int32, device="cuda")                 # negatives and ties: first maximum
    lib.call("sfmb200_la_argmax_first", dp(v2), 4, C.byref(idx), None)
    assert idx.value == 1
    big = T.from_numpy(rng.integers(0, 1000, 100000).astype(np.int32)).cuda()
    lib.call("sfmb200_la_argmax_first", dp(big), 100000, C.byref(idx), None)
    assert idx.value == int(np.argmax(big.cpu().numpy()))


def test_reference_stage_kernels_by_name(lib, T, O):
    """The reference's estimateE / pose / triangulation stages written with the kernels.h names
    (copy_point, gpu_blas_mmul, kernels, transpose, regular_svd, row_extraction_kernel, normalizeE,
    candidate_kernels, compute_linear_triangulation_A, svd_square, normalize_pt_kernal, kernCopy*ToVBO)
    against the fp64 restatement, stage by stage (sfm.cu:80-129, 238-252, 309-336, 374-383)."""
    n, H = 600, 40
    K, Kinv = O.reference_K()
    sc = O.synthetic_pair(n, seed=13)
    x = O.normalise_points(sc["px"], Kinv)
    sift = np.zeros((n, 144), np.float32)
    sift[:, 0], sift[:, 1], sift[:, 9], sift[:, 10] = sc["px"].T
    d_sift = T.from_numpy(sift).cuda()
    U1, U2 = T.empty((3, n), device="cuda"), T.empty((3, n), device="cuda")
    lib.call("sfmb200_la_copy_point", dp(d_sift), n, dp(U1), dp(U2), None)
    assert np.array_equal(U1.cpu().numpy(), np.vstack([sc["px"][:, :2].T, np.ones(n, np.float32)]))
    assert np.array_equal(U2.cpu().numpy(), np.vstack([sc["px"][:, 2:].T, np.ones(n, np.float32)]))
    dK = T.from_numpy(np.asarray(Kinv, np.float32).reshape(3, 3)).cuda()
    X1, X2 = T.empty((3, n), device="cuda"), T.empty((3, n), device="cuda")
    lib.call("sfmb200_la_mmul", dp(dK), dp(U1), dp(X1), 3, 3, n, None)
    lib.call("sfmb200_la_mmul", dp(dK), dp(U2), dp(X2), 3, 3, n, None)
    assert np.allclose(X1.cpu().numpy()[:2].T, x[:, :2], atol=1e-6) and np.allclose(X2.cpu().numpy()[:2].T, x[:, 2:], atol=1e-6)
    # design matrices -> null vectors -> E candidates
    idx = O.sample_indices(3, H, n).astype(np.int32)
    d_idx = T.from_numpy(idx).cuda()
    A = T.empty((H, 8, 9), device="cuda")
    lib.call("sfmb200_la_design_matrix", dp(X1), dp(X2), dp(A), dp(d_idx), H, n, None)
    want_A = np.stack([O.design_matrix(x[idx[h]]) for h in range(H)])
    assert np.allclose(A.cpu().numpy(), want_A, atol=1e-6)
    At = T.empty((H, 9, 8), device="cuda")
    lib.call("sfmb200_la_transpose_batched", dp(A), dp(At), 8, 9, H, None)
    S, Uo, Vo = T.empty((H, 8), device="cuda"), T.empty((H, 64), device="cuda"), T.empty((H, 81), device="cuda")
    lib.call("sfmb200_la_svd_batched", dp(At), dp(S), dp(Uo), dp(Vo), 8, 9, H, None)
    E = T.empty((H, 9), device="cuda")
    lib.call("sfmb200_la_row_extraction", dp(Vo), dp(E), H, None)
    lib.call("sfmb200_la_normalize_E", dp(E), H, None)
    Eo = O.hypotheses(x, idx)
    dist = O.e_distance(E.cpu().numpy().reshape(H, 3, 3), Eo)
    assert (dist < 1e-3).mean() > 0.9, dist
    sv = np.linalg.svd(E.cpu().numpy().reshape(H, 3, 3).astype(np.float64), compute_uv=False)
    assert np.allclose(sv, np.array([1, 1, 0])[None], atol=1e-5)
    # pose candidates from a host SVD of the first candidate, like computePosecandidates (sfm.cu:239-250)
    u, _, v = O.svd_rot(Eo[0])
    if O.det_reference_typo(u @ v.T) < 0:
        v = -v
    du = T.from_numpy(np.ascontiguousarray(u, dtype=np.float32)).cuda()
    dv = T.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).cuda()       # V = Vt.T is a strided view
    dP = T.empty((4, 4, 4), device="cuda")
    lib.call("sfmb200_la_candidate_poses", dp(dP), dp(du), dp(dv), None)
    P = dP.cpu().numpy().astype(np.float64)
    Wm = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]])
    for i in range(4):
        R = (u @ (Wm if i < 2 else Wm.T) @ v.T).T
        t = (-1 if i in (0, 2) else 1) * u[:, 2]
        assert np.allclose(P[i, :3, :3], R, atol=1e-6) and np.allclose(P[i, :3, 3], t, atol=1e-6) and np.allclose(P[i, 3], [0, 0, 0, 1])
    # triangulation: A per point with camera 2 = P[1], 4x4 SVD, de-homogenise, VBO
    I4 = T.eye(4, device="cuda")
    A4 = T.empty((n, 4, 4), device="cuda")
    lib.call("sfmb200_la_triangulation_A", dp(A4), dp(X1), dp(X2), n, n, dp(I4), dp(dP), 1, 0, None)
    assert np.allclose(A4.cpu().numpy(), O.dlt_rows(x, P[1]), atol=1e-6)
    A4c = T.empty((4, 4, 4), device="cuda")
    lib.call("sfmb200_la_triangulation_A", dp(A4c), dp(X1), dp(X2), 4, n, dp(I4), dp(dP), 0, 1, None)
    assert np.allclose(A4c.cpu().numpy(), np.stack([O.dlt_rows(x[:1], P[i])[0] for i in range(4)]), atol=1e-6)
    A4t = T.empty((n, 4, 4), device="cuda")
    lib.call("sfmb200_la_transpose_batched", dp(A4), dp(A4t), 4, 4, n, None)
    S4, U4, V4 = T.empty((n, 4), device="cuda"), T.empty((n, 16), device="cuda"), T.empty((n, 16), device="cuda")
    lib.call("sfmb200_la_svd_batched", dp(A4t), dp(S4), dp(U4), dp(V4), 4, 4, n, None)
    pts = T.empty((4, n), device="cuda")
    lib.call("sfmb200_la_normalize_pt", dp(V4), dp(pts), n, None)
    want = O.triangulate(x, P[1])
    got = pts.cpu().numpy()
    ok = (np.abs(want[2]) < 50) & (np.abs(want[2]) > 1e-3)
    assert ok.mean() > 0.6 and np.all(got[3] == 1)
    assert np.median(np.abs(got[:3, ok] - want[:3, ok]) / (np.abs(want[:3, ok]) + 1e-3)) < 1e-3
    v0 = T.tensor([[0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 4, 6, 2], [0] * 12 + [1, 1, 1, 0], [0] * 12 + [6, 6, 6, 6]],
                  dtype=T.float32, device="cuda")
    o3 = T.empty((4, 3), device="cuda")
    lib.call("sfmb200_la_normalize_pt", dp(v0), dp(o3), 3, None)
    assert np.array_equal(o3.cpu().numpy(), np.array([[1, 0, 0], [2, 0, 0], [3, 0, 0], [1, 1, 1]], np.float32))   # w = 0 and |w| > 5 -> origin
    vbo, col = T.empty((n, 4), device="cuda"), T.zeros((n, 4), device="cuda")
    lib.call("sfmb200_la_copy_to_vbo", n, dp(pts), dp(vbo), C.c_float(2.0), None)
    lib.call("sfmb200_la_copy_to_vbo", n, None, dp(col), C.c_float(1.0), None)
    assert np.array_equal(vbo.cpu().numpy()[:, :3], 2 * got[:3].T) and np.all(vbo.cpu().numpy()[:, 3] == 1) and np.all(col.cpu().numpy() == 1)
