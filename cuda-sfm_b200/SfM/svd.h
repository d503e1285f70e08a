// Source-compatible replacement of the reference's SfM/svd.h call surface:
// 3x3 row-major float[9] small-matrix helpers, usable from host code.  The
// arithmetic lives in libsfmb200 (cuda-sfm_b200/csrc/smallmat.cuh: one-sided
// Jacobi + Givens QR); these inline functions only forward.
//
// Contract of svd(a, u, s, v) as in SfM/svd.h:311-335: a = u * s * v^T, v NOT
// transposed, s 3x3 (diagonal up to rounding) with s00 >= s11 >= |s22|, u and
// v proper rotations (the sign lives in s22).
//
// Not provided: the internal steps of the reference's own SVD algorithm
// (approximateGivensQuaternion, jacobiConjugation, jacobiEigenanlysis,
// sortSingularValues, QRGivensQuaternion, QRDecomposition, svd.h:120-309): no
// code outside svd.h calls them.
#ifndef SFMB200_FACADE_SVD_H
#define SFMB200_FACADE_SVD_H

#include <cmath>

#include "../../include/sfmb200.h"
#include "common.h"

inline float accurateSqrt(float x) { return std::sqrt(x); }
inline void condSwap(bool c, float& X, float& Y) { float Z = X; X = c ? Y : X; Y = c ? Z : Y; }
inline void condNegSwap(bool c, float& X, float& Y) { float Z = -X; X = c ? Y : X; Y = c ? Z : Y; }
inline float dist2(float x, float y, float z) { return x * x + y * y + z * z; }

// M = A B, M = A^T B, M = A B^T (svd.h:58-83)
inline void multAB(const float* a, const float* b, float* m) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) m[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
inline void multAtB(const float* a, const float* b, float* m) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) m[3 * i + j] = a[i] * b[j] + a[3 + i] * b[3 + j] + a[6 + i] * b[6 + j];
}
inline void multABt(const float* a, const float* b, float* m) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) m[3 * i + j] = a[3 * i] * b[3 * j] + a[3 * i + 1] * b[3 * j + 1] + a[3 * i + 2] * b[3 * j + 2];
}
inline void neg(float* a) {
    for (int i = 0; i < 9; i++) a[i] = -a[i];
}
// quaternion (x, y, z, w) -> rotation matrix (svd.h:97-118)
inline void quatToMat3(const float* q, float* m) {
    float x = q[0], y = q[1], z = q[2], w = q[3];
    m[0] = 1 - 2 * (y * y + z * z); m[1] = 2 * (x * y - w * z);     m[2] = 2 * (x * z + w * y);
    m[3] = 2 * (x * y + w * z);     m[4] = 1 - 2 * (x * x + z * z); m[5] = 2 * (y * z - w * x);
    m[6] = 2 * (x * z - w * y);     m[7] = 2 * (y * z + w * x);     m[8] = 1 - 2 * (x * x + y * y);
}
inline void svd(const float* a, float* u, float* s, float* v) { sfmb200_host_svd3(a, u, s, v); }

// Determinant.  The reference's det() (svd.h:337-341) has a typo in its third
// term and is wrong for general matrices; det() here is the true determinant,
// det_reference() reproduces the reference's expression for bug-compatible
// callers (the compat pose candidates use it internally).
inline float det(const float* a) {
    return a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
}
inline float det_reference(const float* a) {
    return a[0] * a[4] * a[8] - a[0] * a[5] * a[7] - a[0] * a[3] * a[8] + a[1] * a[5] * a[6] + a[2] * a[3] * a[7] - a[2] * a[4] * a[6];
}
// transpose the leading 3x3 of a (row stride a_size) into b (row stride b_size), svd.h:343-349
inline void transpose_copy3x3(const float* a, float* b, int a_size, int b_size) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) b[access2(j, i, b_size)] = a[access2(i, j, a_size)];
}
inline bool InvertMatrix4x4(const float m[16], float invOut[16]) { return sfmb200_host_inv4(m, invOut) == 0; }
// polar decomposition a = u p (svd.h:483-501)
inline void pd(const float* a, float* u, float* p) {
    float w[9], s[9], v[9], t[9], vt[9];
    svd(a, w, s, v);
    multAB(v, s, t);
    transpose_copy3x3(v, vt, 3, 3);
    multAB(t, vt, p);      // P = V S V^T
    multAB(w, vt, u);      // U = W V^T
}
#endif
