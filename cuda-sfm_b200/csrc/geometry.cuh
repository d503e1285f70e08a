// Device functions of the geometry stages shared by the stage kernels (geometry.cu) and the fused small-problem
// kernel (small.cu): K^-1 normalisation + scaled copies, pose candidates, DLT rows, de-homogenisation, cheirality.
// Reference: SfM/sfm.cu:80-92, 238-307; SfM/kernels.h:261-279, 357-450.
#pragma once
#include "internal.cuh"
#include "sampson.cuh"
#include "smallmat.cuh"

namespace sfmb200 {

struct Mat9 { float v[9]; };

// ---------------------------------------------------------------------------
// Ingest.  X = K^-1 [u v 1]^T for both images (copy_point + 2 cublasSgemm in the
// reference).  Writes the float4 correspondence array every later stage reads
// and the threshold-scaled copy the scoring kernels stage through TMA.
// ---------------------------------------------------------------------------
// Threshold-scaled copy for the scoring kernels (sampson.cuh): coordinates * pt_scale.
__device__ __forceinline__ void store_scaled(const DeviceState& s, size_t o, float x1, float y1, float x2, float y2) {
    const float k = s.pt_scale;
    x1 *= k; y1 *= k; x2 *= k; y2 *= k;
    s.corr_s[o] = make_float4(x1, y1, x2, y2);
}
__device__ __forceinline__ float4 normalise_point(float u1, float v1, float u2, float v2, const Mat9& k) {
    float x1 = fmaf(k.v[1], v1, fmaf(k.v[0], u1, k.v[2]));
    float y1 = fmaf(k.v[4], v1, fmaf(k.v[3], u1, k.v[5]));
    float z1 = fmaf(k.v[7], v1, fmaf(k.v[6], u1, k.v[8]));
    float x2 = fmaf(k.v[1], v2, fmaf(k.v[0], u2, k.v[2]));
    float y2 = fmaf(k.v[4], v2, fmaf(k.v[3], u2, k.v[5]));
    float z2 = fmaf(k.v[7], v2, fmaf(k.v[6], u2, k.v[8]));
    // The reference keeps a z row (== 1 for any K with last row 0 0 1); the
    // float4 layout fixes z = 1, which is the same projective point.
    if (z1 != 1.0f) { x1 /= z1; y1 /= z1; }
    if (z2 != 1.0f) { x2 /= z2; y2 /= z2; }
    return make_float4(x1, y1, x2, y2);
}
__device__ __forceinline__ void normalise_store(const DeviceState& s, int b, int i, float u1, float v1, float u2, float v2,
                                                const Mat9& k) {
    const float4 c = normalise_point(u1, v1, u2, v2, k);
    size_t o = (size_t)b * s.n_stride + i;
    s.corr[o] = c;
    store_scaled(s, o, c.x, c.y, c.z, c.w);
}

// ---------------------------------------------------------------------------
// Pose candidates (computePosecandidates sfm.cu:238-252 + candidate_kernels
// kernels.h:357-385).  One thread per pair; the reference does the SVD on the
// host between two synchronous copies.
// compat = 1: P_i = [ (U W(^T) V^T)^T | +-u3 ], det-typo sign fix (Q15, Q16).
// compat = 0: textbook pose for x1^T E x2 = 0: X2 = R X1 + t, R = V W(^T) U^T, t = +-v3.
// ---------------------------------------------------------------------------
// Candidate c (0..3) of E into P[16].
// Candidate c from the (oriented) SVD factors of E; v is modified.
__device__ __forceinline__ void pose_from_svd(const float* u, float* v, int c, int compat, float* P);
__device__ __forceinline__ void pose_candidate(const float* E, int c, int compat, float* P) {
    float u[9], sg[9], v[9];
    if (compat) svd3_reference_orientation(E, u, sg, v);     // candidate ORDER as the reference's svd() yields it
    else svd3<5>(E, u, sg, v);
    pose_from_svd(u, v, c, compat, P);
}
__device__ __forceinline__ void pose_from_svd(const float* u, float* v, int c, int compat, float* P) {
    const float W[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1};
    const float Wt[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1};
    if (compat) {
        float tmp[9];
        mul33_ABt(u, v, tmp);
        if (det33_reference_typo(tmp) < 0.0f) {
#pragma unroll
            for (int i = 0; i < 9; i++) v[i] = -v[i];
        }
    }
    float t1[9], R[9];
    const bool first = c < 2;
    float Wc[9];
#pragma unroll
    for (int i = 0; i < 9; i++) Wc[i] = first ? W[i] : Wt[i];
    float sign = (c == 0 || c == 2) ? -1.0f : 1.0f;
    float tx, ty, tz;
    if (compat) {
        mul33_ABt(Wc, v, t1);     // W V^T
        mul33(u, t1, R);          // U W V^T ; stored transposed
        float Rt[9];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) Rt[3 * i + j] = R[3 * j + i];
#pragma unroll
        for (int i = 0; i < 9; i++) R[i] = Rt[i];
        tx = sign * u[2]; ty = sign * u[5]; tz = sign * u[8];
    } else {
        mul33_ABt(Wc, u, t1);     // W U^T
        mul33(v, t1, R);          // V W U^T
        if (det33(R) < 0.0f) {
#pragma unroll
            for (int i = 0; i < 9; i++) R[i] = -R[i];
        }
        tx = sign * v[2]; ty = sign * v[5]; tz = sign * v[8];
    }
    P[0] = R[0]; P[1] = R[1]; P[2] = R[2];  P[3] = tx;
    P[4] = R[3]; P[5] = R[4]; P[6] = R[5];  P[7] = ty;
    P[8] = R[6]; P[9] = R[7]; P[10] = R[8]; P[11] = tz;
    P[12] = 0.0f; P[13] = 0.0f; P[14] = 0.0f; P[15] = 1.0f;
}

// DLT rows for one correspondence, camera 1 = I4, camera 2 = M
// (compute_linear_triangulation_A, kernels.h:387-431).
__device__ __forceinline__ void dlt_matrix(float x1, float y1, float x2, float y2, const float* M, float* A) {
    A[0] = -1.0f; A[1] = 0.0f;  A[2] = x1; A[3] = 0.0f;
    A[4] = 0.0f;  A[5] = -1.0f; A[6] = y1; A[7] = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        A[8 + i] = fmaf(x2, M[8 + i], -M[i]);
        A[12 + i] = fmaf(y2, M[8 + i], -M[4 + i]);
    }
}
// De-homogenise like normalize_pt_kernal (kernels.h:433-450): w == 0 -> origin (as v * 0: the packed form of
// triangulate_pairs_kernel does the same arithmetic, so both give the same bits).
__device__ __forceinline__ void dehomogenise(const float* v, float& X, float& Y, float& Z) {
#if defined(__CUDA_ARCH__)
    const float iw = v[3] == 0.0f ? 0.0f : __fdividef(1.0f, v[3]);      // MUFU.RCP: 1 ulp, |v| = 1
#else
    const float iw = v[3] == 0.0f ? 0.0f : 1.0f / v[3];
#endif
    X = v[0] * iw; Y = v[1] * iw; Z = v[2] * iw;
}
__device__ __forceinline__ void dehomogenise2(const float2* v, float2& X, float2& Y, float2& Z) {
#if defined(__CUDA_ARCH__)
    const float2 iw = make_float2(v[3].x == 0.0f ? 0.0f : __fdividef(1.0f, v[3].x), v[3].y == 0.0f ? 0.0f : __fdividef(1.0f, v[3].y));
    X = __fmul2_rn(v[0], iw); Y = __fmul2_rn(v[1], iw); Z = __fmul2_rn(v[2], iw);
#endif
}

// Cheirality of candidate M on correspondence c0 (reference semantics): returns
// whether the triangulated point is in front of both cameras; Minv = M^-1.
__device__ __forceinline__ bool cheirality_compat(const float4& c0, const float* M, float* Minv) {
    float A[16], v[4];
    dlt_matrix(c0.x, c0.y, c0.z, c0.w, M, A);
    if (!dlt_null_adjugate1(A, v)) null4<5>(A, v);     // same solve as triangulate_kernel
    float X, Y, Z;
    dehomogenise(v, X, Y, Z);
    inv4(M, Minv);
    float z2 = fmaf(Minv[8], X, fmaf(Minv[9], Y, fmaf(Minv[10], Z, Minv[11])));
    return Z > 0.0f && z2 > 0.0f;
}

}  // namespace sfmb200
