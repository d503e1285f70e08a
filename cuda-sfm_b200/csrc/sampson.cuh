// The inlier test shared by every kernel that classifies a correspondence:
//   d = num^2 - thr * den,  num = x1^T E x2,  den = (E x2)_0^2 + (E x2)_1^2 + (E^T x1)_0^2 + (E^T x1)_1^2
// (Sampson error < thr  <=>  d < 0; z = 1 in both views; reference convention
// x1^T E x2 = 0; threshold literal 1e-6 from SfM/sfm.cu:220).
// ONE fma tree, so scoring, cheirality vote, inlier mask, triangulation mask and
// refit agree bit for bit; oracle/oracle_c.c:sampson_d_f32 mirrors it on the CPU.
#pragma once
#include <cuda_runtime.h>

namespace sfmb200 {

__device__ __forceinline__ float sampson_d(const float* e, float x1, float y1, float x2, float y2, float nthr) {
    float l0 = fmaf(e[0], x2, fmaf(e[1], y2, e[2]));
    float l1 = fmaf(e[3], x2, fmaf(e[4], y2, e[5]));
    float l2 = fmaf(e[6], x2, fmaf(e[7], y2, e[8]));
    float num = fmaf(x1, l0, fmaf(y1, l1, l2));
    float m0 = fmaf(e[0], x1, fmaf(e[3], y1, e[6]));
    float m1 = fmaf(e[1], x1, fmaf(e[4], y1, e[7]));
    float den = fmaf(l0, l0, fmaf(l1, l1, fmaf(m0, m0, m1 * m1)));
    return fmaf(den, nthr, num * num);
}

// Two hypotheses at once with packed FFMA2 / FMUL2 (fma.rn.f32x2, sm_100+):
// lane-wise identical to sampson_d.
__device__ __forceinline__ float2 sampson_d2(const float2* e, float2 x1, float2 y1, float2 x2, float2 y2, float2 nthr) {
    float2 l0 = __ffma2_rn(e[0], x2, __ffma2_rn(e[1], y2, e[2]));
    float2 l1 = __ffma2_rn(e[3], x2, __ffma2_rn(e[4], y2, e[5]));
    float2 l2 = __ffma2_rn(e[6], x2, __ffma2_rn(e[7], y2, e[8]));
    float2 num = __ffma2_rn(x1, l0, __ffma2_rn(y1, l1, l2));
    float2 m0 = __ffma2_rn(e[0], x1, __ffma2_rn(e[3], y1, e[6]));
    float2 m1 = __ffma2_rn(e[1], x1, __ffma2_rn(e[4], y1, e[7]));
    float2 den = __ffma2_rn(l0, l0, __ffma2_rn(l1, l1, __ffma2_rn(m0, m0, __fmul2_rn(m1, m1))));
    return __ffma2_rn(den, nthr, __fmul2_rn(num, num));
}

}  // namespace sfmb200
