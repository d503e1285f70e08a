// The inlier test shared by every kernel that classifies a correspondence:
//   Sampson error  num^2 / den < thr,  num = x1^T E x2,  den = (E x2)_0^2 + (E x2)_1^2 + (E^T x1)_0^2 + (E^T x1)_1^2
// (z = 1 in both views; reference convention x1^T E x2 = 0; threshold literal 1e-6 from
// SfM/sfm.cu:220).  ONE fma tree, so scoring, cheirality vote, inlier mask, triangulation mask and
// refit agree bit for bit; oracle/oracle_c.c:sampson_d_f32 mirrors it on the CPU.
#pragma once
#include <cuda_runtime.h>

namespace sfmb200 {

// The threshold is folded into the coordinates: with k = sqrt(thr), points scaled by 1/k and
// E~ = D E D, D = diag(k, k, 1), the Sampson error in the scaled coordinates is error / thr, so
//   inlier  <=>  num~^2 - den~ < 0,   num~ = x1~^T E~ x2~,  den~ = (E~ x2~)_0^2 + (E~ x2~)_1^2 + (E~^T x1~)_0^2 + (E~^T x1~)_1^2:
// 17 FP32-pipe instructions per evaluation instead of 18 (no thr * den product).  The scoring kernels read
// pre-scaled correspondences (DeviceState::corr_s) and scale each hypothesis once when it is
// loaded; every other classifier calls sampson_d, which applies the same scaling per call - same operations in
// the same order, so all of them agree bit for bit.
struct ThrScale { float k, ik, k2; };
__host__ __device__ __forceinline__ ThrScale make_thr_scale(float thr) {
    ThrScale t;
    t.k = sqrtf(thr);          // correctly rounded on host and device (no fast-math)
    t.ik = 1.0f / t.k;
    t.k2 = t.k * t.k;
    return t;
}
// factor of entry q of E~ = D E D: k^2 for the upper-left 2x2, k for the rest of row / column 2, 1 for e8
__host__ __device__ __forceinline__ float thr_scale_factor(const ThrScale& t, int q) {
    return (q == 8) ? 1.0f : ((q == 2 || q == 5 || q == 6 || q == 7) ? t.k : t.k2);
}

// scaled E, scaled coordinates -> d (negative = inlier)
__device__ __forceinline__ float sampson_unit_d(const float* e, float x1, float y1, float x2, float y2) {
    float l0 = fmaf(e[0], x2, fmaf(e[1], y2, e[2]));
    float l1 = fmaf(e[3], x2, fmaf(e[4], y2, e[5]));
    float l2 = fmaf(e[6], x2, fmaf(e[7], y2, e[8]));
    float num = fmaf(x1, l0, fmaf(y1, l1, l2));
    float m0 = fmaf(e[0], x1, fmaf(e[3], y1, e[6]));
    float m1 = fmaf(e[1], x1, fmaf(e[4], y1, e[7]));
    float den = fmaf(l0, l0, fmaf(l1, l1, fmaf(m0, m0, m1 * m1)));
    return fmaf(num, num, -den);
}
// Two hypotheses at once with packed FFMA2 / FMUL2 (fma.rn.f32x2, sm_100+): lane-wise identical to sampson_unit_d.
__device__ __forceinline__ float2 sampson_unit_d2(const float2* e, float2 x1, float2 y1, float2 x2, float2 y2) {
    float2 l0 = __ffma2_rn(e[0], x2, __ffma2_rn(e[1], y2, e[2]));
    float2 l1 = __ffma2_rn(e[3], x2, __ffma2_rn(e[4], y2, e[5]));
    float2 l2 = __ffma2_rn(e[6], x2, __ffma2_rn(e[7], y2, e[8]));
    float2 num = __ffma2_rn(x1, l0, __ffma2_rn(y1, l1, l2));
    float2 m0 = __ffma2_rn(e[0], x1, __ffma2_rn(e[3], y1, e[6]));
    float2 m1 = __ffma2_rn(e[1], x1, __ffma2_rn(e[4], y1, e[7]));
    float2 den = __ffma2_rn(l0, l0, __ffma2_rn(l1, l1, __ffma2_rn(m0, m0, __fmul2_rn(m1, m1))));
    return __ffma2_rn(num, num, make_float2(-den.x, -den.y));
}
// unscaled E, unscaled coordinates, nthr = -thr: the form every classifier outside the scoring kernels uses
__device__ __forceinline__ float sampson_d(const float* e, float x1, float y1, float x2, float y2, float nthr) {
    const ThrScale t = make_thr_scale(-nthr);
    float es[9];
#pragma unroll
    for (int q = 0; q < 9; q++) es[q] = (q == 8) ? e[q] : e[q] * thr_scale_factor(t, q);
    return sampson_unit_d(es, x1 * t.ik, y1 * t.ik, x2 * t.ik, y2 * t.ik);
}

// Symmetric epipolar distance (SFMB200_OPT_SCORE_METRIC = 1): what the reference's calculateInliers was written to
// compute - n^2 / |(E x2)_{0,1}|^2 summed with n^2 / |(E^T x1)_{0,1}|^2 (sfm.cu:155-221) - before its indexing
// bugs (Q8-Q11).  Same scaling as above, division-free:
//   inlier  <=>  num~^2 (A + B) - A B < 0,   A = l0^2 + l1^2,  B = m0^2 + m1^2        (20 FP32-pipe instructions)
__device__ __forceinline__ float symmetric_unit_d(const float* e, float x1, float y1, float x2, float y2) {
    float l0 = fmaf(e[0], x2, fmaf(e[1], y2, e[2]));
    float l1 = fmaf(e[3], x2, fmaf(e[4], y2, e[5]));
    float l2 = fmaf(e[6], x2, fmaf(e[7], y2, e[8]));
    float num = fmaf(x1, l0, fmaf(y1, l1, l2));
    float m0 = fmaf(e[0], x1, fmaf(e[3], y1, e[6]));
    float m1 = fmaf(e[1], x1, fmaf(e[4], y1, e[7]));
    float A = fmaf(l0, l0, l1 * l1);
    float B = fmaf(m0, m0, m1 * m1);
    return fmaf(num * num, A + B, -(A * B));
}
__device__ __forceinline__ float2 symmetric_unit_d2(const float2* e, float2 x1, float2 y1, float2 x2, float2 y2) {
    float2 l0 = __ffma2_rn(e[0], x2, __ffma2_rn(e[1], y2, e[2]));
    float2 l1 = __ffma2_rn(e[3], x2, __ffma2_rn(e[4], y2, e[5]));
    float2 l2 = __ffma2_rn(e[6], x2, __ffma2_rn(e[7], y2, e[8]));
    float2 num = __ffma2_rn(x1, l0, __ffma2_rn(y1, l1, l2));
    float2 m0 = __ffma2_rn(e[0], x1, __ffma2_rn(e[3], y1, e[6]));
    float2 m1 = __ffma2_rn(e[1], x1, __ffma2_rn(e[4], y1, e[7]));
    float2 A = __ffma2_rn(l0, l0, __fmul2_rn(l1, l1));
    float2 B = __ffma2_rn(m0, m0, __fmul2_rn(m1, m1));
    float2 AB = __fmul2_rn(A, B);
    return __ffma2_rn(__fmul2_rn(num, num), __fadd2_rn(A, B), make_float2(-AB.x, -AB.y));
}
// unscaled E and coordinates, the metric chosen at run time: what the classifiers outside the scoring kernels use
// (inlier mask, triangulation mask, egress colours, cheirality vote)
__device__ __forceinline__ float epipolar_d(int metric, const float* e, float x1, float y1, float x2, float y2, float nthr) {
    const ThrScale t = make_thr_scale(-nthr);
    float es[9];
#pragma unroll
    for (int q = 0; q < 9; q++) es[q] = (q == 8) ? e[q] : e[q] * thr_scale_factor(t, q);
    return metric == 0 ? sampson_unit_d(es, x1 * t.ik, y1 * t.ik, x2 * t.ik, y2 * t.ik)
                       : symmetric_unit_d(es, x1 * t.ik, y1 * t.ik, x2 * t.ik, y2 * t.ik);
}

// Homography model (CudaSift's TestHomographies, matching.cu:953-996):
// one-sided transfer error of x1 -> x2 under H, division-free exactly like the original:
//   d = (x2*W - X)^2 + (y2*W - Y)^2 - thr^2 * W^2,  (X, Y, W) = H (x1, y1, 1);  inlier <=> d < 0.
// nthr is -(thr^2).
__device__ __forceinline__ float homography_d(const float* e, float x1, float y1, float x2, float y2, float nthr) {
    float X = fmaf(e[0], x1, fmaf(e[1], y1, e[2]));
    float Y = fmaf(e[3], x1, fmaf(e[4], y1, e[5]));
    float W = fmaf(e[6], x1, fmaf(e[7], y1, e[8]));
    float ex = fmaf(x2, W, -X);
    float ey = fmaf(y2, W, -Y);
    float err2 = fmaf(ex, ex, ey * ey);
    return fmaf(W * W, nthr, err2);
}
__device__ __forceinline__ float2 homography_d2(const float2* e, float2 x1, float2 y1, float2 x2, float2 y2, float2 nthr) {
    float2 X = __ffma2_rn(e[0], x1, __ffma2_rn(e[1], y1, e[2]));
    float2 Y = __ffma2_rn(e[3], x1, __ffma2_rn(e[4], y1, e[5]));
    float2 W = __ffma2_rn(e[6], x1, __ffma2_rn(e[7], y1, e[8]));
    float2 nX = make_float2(-X.x, -X.y), nY = make_float2(-Y.x, -Y.y);
    float2 ex = __ffma2_rn(x2, W, nX);
    float2 ey = __ffma2_rn(y2, W, nY);
    float2 err2 = __ffma2_rn(ex, ex, __fmul2_rn(ey, ey));
    return __ffma2_rn(__fmul2_rn(W, W), nthr, err2);
}

// MODEL 0: essential matrix / Sampson; MODEL 1: homography / transfer error; MODEL 2: essential matrix / symmetric
// epipolar distance.
template <int MODEL>
__device__ __forceinline__ float model_d(const float* e, float x1, float y1, float x2, float y2, float nthr) {
    // scoring kernels: MODEL 0 / 2 get pre-scaled E and coordinates (see above); MODEL 1 is unscaled (pt_scale = 1)
    if (MODEL == 0) return sampson_unit_d(e, x1, y1, x2, y2);
    if (MODEL == 2) return symmetric_unit_d(e, x1, y1, x2, y2);
    return homography_d(e, x1, y1, x2, y2, nthr);
}
template <int MODEL>
__device__ __forceinline__ float2 model_d2(const float2* e, float2 x1, float2 y1, float2 x2, float2 y2, float2 nthr) {
    if (MODEL == 0) return sampson_unit_d2(e, x1, y1, x2, y2);
    if (MODEL == 2) return symmetric_unit_d2(e, x1, y1, x2, y2);
    return homography_d2(e, x1, y1, x2, y2, nthr);
}

}  // namespace sfmb200
