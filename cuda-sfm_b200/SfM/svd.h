// Source-compatible replacement of the reference's SfM/svd.h call surface (SfM/svd.h:33-501): 3x3 row-major
// float[9] small-matrix helpers, every one `__host__ __device__ __forceinline__` like the reference's, so the
// header works from host code (g++ or nvcc) and inside kernels.  Header-only: the arithmetic is the library's own
// small-matrix code (cuda-sfm_b200/csrc/smallmat.cuh: one-sided Jacobi + Givens QR), not the reference's.
//
// Contract of svd(a, u, s, v) as in SfM/svd.h:311-335: a = u * s * v^T, v NOT transposed, s 3x3 (upper triangular,
// diagonal up to rounding) with s00 >= s11 >= |s22|, u and v proper rotations (the sign lives in s22).
//
// The internal steps of the reference's SVD (svd.h:120-309: approximateGivensQuaternion, jacobiConjugation,
// jacobiEigenanlysis, sortSingularValues, QRGivensQuaternion, QRDecomposition) are kept by name, signature and
// meaning as thin, exact equivalents (the reference's are 4-sweep approximations): nothing outside svd.h calls them,
// svd() above does not go through them.
#ifndef SFMB200_FACADE_SVD_H
#define SFMB200_FACADE_SVD_H

#include <cmath>

#include "../csrc/smallmat.cuh"
#include "common.h"

#define SFM_SVD_HD __host__ __device__ __forceinline__

SFM_SVD_HD float accurateSqrt(float x) { return sqrtf(x); }
SFM_SVD_HD void condSwap(bool c, float& X, float& Y) { float Z = X; X = c ? Y : X; Y = c ? Z : Y; }
SFM_SVD_HD void condNegSwap(bool c, float& X, float& Y) { float Z = -X; X = c ? Y : X; Y = c ? Z : Y; }
SFM_SVD_HD float dist2(float x, float y, float z) { return x * x + y * y + z * z; }

// M = A B, M = A^T B, M = A B^T (svd.h:58-83)
SFM_SVD_HD void multAB(const float* a, const float* b, float* m) { sfmb200::mul33(a, b, m); }
SFM_SVD_HD void multAtB(const float* a, const float* b, float* m) { sfmb200::mul33_AtB(a, b, m); }
SFM_SVD_HD void multABt(const float* a, const float* b, float* m) { sfmb200::mul33_ABt(a, b, m); }
SFM_SVD_HD void neg(float* a) {
    for (int i = 0; i < 9; i++) a[i] = -a[i];
}
// quaternion (x, y, z, w) -> rotation matrix (svd.h:97-118)
SFM_SVD_HD void quatToMat3(const float* q, float* m) {
    float x = q[0], y = q[1], z = q[2], w = q[3];
    m[0] = 1 - 2 * (y * y + z * z); m[1] = 2 * (x * y - w * z);     m[2] = 2 * (x * z + w * y);
    m[3] = 2 * (x * y + w * z);     m[4] = 1 - 2 * (x * x + z * z); m[5] = 2 * (y * z - w * x);
    m[6] = 2 * (x * z - w * y);     m[7] = 2 * (y * z + w * x);     m[8] = 1 - 2 * (x * x + y * y);
}

// ---- the reference's internal steps, by name (svd.h:120-309) ----
// (ch, sh) = (cos, sin) of HALF the Jacobi angle that diagonalises [a11 a12; a12 a22] (svd.h:120-133 approximates it
// from ch ~ 2 (a11 - a22), sh ~ a12; this is the exact angle): the rotation [c -s; s c], c = ch^2 - sh^2, s = 2 ch sh.
SFM_SVD_HD void approximateGivensQuaternion(float a11, float a12, float a22, float& ch, float& sh) {
    // Jacobi angle, inner rotation (|theta| <= pi/4: the choice cyclic Jacobi needs to converge)
    const float d = a11 - a22;
    const float theta = 0.5f * atan2f(d < 0.0f ? -2.0f * a12 : 2.0f * a12, fabsf(d));
    ch = cosf(0.5f * theta);
    sh = sinf(0.5f * theta);
}
// One Jacobi conjugation of the symmetric matrix kept in the lower triangle of s (s[0] s[3] s[4] s[6] s[7] s[8])
// on its (0, 1) pair, S <- Q^T S Q, accumulated into the quaternion qV (x, y, z, w) as qV <- qV * Q with Q the
// rotation about axis z; then the matrix is cyclically re-indexed (i -> i + 1 mod 3) so that three calls with
// (x, y, z) = (0,1,2), (1,2,0), (2,0,1) visit the pairs (0,1), (1,2), (0,2) (svd.h:135-190).
SFM_SVD_HD void jacobiConjugation(const int x, const int y, const int z, float* s, float* qV) {
    float ch, sh;
    approximateGivensQuaternion(s[0], s[3], s[4], ch, sh);
    const float c = ch * ch - sh * sh, sn = 2.0f * sh * ch;
    const float s00 = s[0], s10 = s[3], s11 = s[4], s20 = s[6], s21 = s[7], s22 = s[8];
    // Q = [c -sn 0; sn c 0; 0 0 1]
    const float t00 = c * s00 + sn * s10, t01 = -sn * s00 + c * s10;     // row 0 of S Q
    const float t10 = c * s10 + sn * s11, t11 = -sn * s10 + c * s11;     // row 1 of S Q
    const float n00 = c * t00 + sn * t10;
    const float n10 = -sn * t00 + c * t10;
    const float n11 = -sn * t01 + c * t11;
    const float n20 = c * s20 + sn * s21, n21 = -sn * s20 + c * s21;
    // quaternion product qV * (sh e_z + ch)
    const float qx = qV[x], qy = qV[y], qz = qV[z], qw = qV[3];
    qV[x] = ch * qx + sh * qy;
    qV[y] = ch * qy - sh * qx;
    qV[z] = ch * qz + sh * qw;
    qV[3] = ch * qw - sh * qz;
    // cyclic re-indexing: new(i, j) = old(i + 1, j + 1)
    s[0] = n11;
    s[3] = n21; s[4] = s22;
    s[6] = n10; s[7] = n20; s[8] = n00;
}
// V (as a quaternion) that diagonalises the symmetric s (svd.h:198-215; 4 cyclic sweeps there, 6 exact ones here)
SFM_SVD_HD void jacobiEigenanlysis(float* s, float* qV) {
    qV[0] = 0; qV[1] = 0; qV[2] = 0; qV[3] = 1;
    for (int sweep = 0; sweep < 6; sweep++) {
        jacobiConjugation(0, 1, 2, s, qV);
        jacobiConjugation(1, 2, 0, s, qV);
        jacobiConjugation(2, 0, 1, s, qV);
    }
}
// columns of b (and v) ordered by decreasing column norm of b; a swap negates one column so det v is kept (svd.h:217-241)
SFM_SVD_HD void sortSingularValues(float* b, float* v) {
    float r0 = dist2(b[0], b[3], b[6]), r1 = dist2(b[1], b[4], b[7]), r2 = dist2(b[2], b[5], b[8]);
    const int pairs[3][2] = {{0, 1}, {0, 2}, {1, 2}};
    for (int k = 0; k < 3; k++) {
        const int i = pairs[k][0], j = pairs[k][1];
        float& ri = i == 0 ? r0 : r1;
        float& rj = j == 1 ? r1 : r2;
        const bool c = ri < rj;
        for (int row = 0; row < 3; row++) {
            condNegSwap(c, b[3 * row + i], b[3 * row + j]);
            condNegSwap(c, v[3 * row + i], v[3 * row + j]);
        }
        condSwap(c, ri, rj);
    }
}
// (ch, sh) = (cos, sin) of HALF the Givens angle that annihilates a2 against the pivot a1 (svd.h:243-258)
SFM_SVD_HD void QRGivensQuaternion(float a1, float a2, float& ch, float& sh) {
    const float rho = accurateSqrt(a1 * a1 + a2 * a2);
    if (!(rho > 1e-6f)) { ch = 1.0f; sh = 0.0f; return; }
    // half-angle from (cos, sin) = (a1, a2) / rho without cancellation: (|a1| + rho, a2), swapped for a1 < 0
    ch = fabsf(a1) + rho;
    sh = a2;
    condSwap(a1 < 0.0f, sh, ch);
    const float w = 1.0f / accurateSqrt(ch * ch + sh * sh);
    ch *= w;
    sh *= w;
}
// b = q r, q a rotation, r upper triangular, by three Givens rotations (svd.h:260-309)
SFM_SVD_HD void QRDecomposition(const float* b, float* q, float* r) {
    float R[9], Qt[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int i = 0; i < 9; i++) R[i] = b[i];
    const int steps[3][3] = {{0, 1, 0}, {0, 2, 0}, {1, 2, 1}};          // rotate rows (p, q) to zero R[q][col]
    for (int k = 0; k < 3; k++) {
        const int p = steps[k][0], qq = steps[k][1], col = steps[k][2];
        float c, sn;
        sfmb200::givens(R[3 * p + col], R[3 * qq + col], c, sn);
        for (int j = 0; j < 3; j++) {
            const float rp = R[3 * p + j], rq = R[3 * qq + j];
            R[3 * p + j] = c * rp + sn * rq; R[3 * qq + j] = -sn * rp + c * rq;
            const float tp = Qt[3 * p + j], tq = Qt[3 * qq + j];
            Qt[3 * p + j] = c * tp + sn * tq; Qt[3 * qq + j] = -sn * tp + c * tq;
        }
    }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) { q[3 * i + j] = Qt[3 * j + i]; r[3 * i + j] = R[3 * i + j]; }
}

SFM_SVD_HD void svd(const float* a, float* u, float* s, float* v) { sfmb200::svd3<5>(a, u, s, v); }

// Determinant.  The reference's det() (svd.h:337-341) has a typo in its third term and is wrong for general
// matrices; det() here is the true determinant, det_reference() reproduces the reference's expression for
// bug-compatible callers (the compat pose candidates use it internally).
SFM_SVD_HD float det(const float* a) { return sfmb200::det33(a); }
SFM_SVD_HD float det_reference(const float* a) { return sfmb200::det33_reference_typo(a); }
// transpose the leading 3x3 of a (row stride a_size) into b (row stride b_size), svd.h:343-349
SFM_SVD_HD void transpose_copy3x3(const float* a, float* b, int a_size, int b_size) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) b[access2(j, i, b_size)] = a[access2(i, j, a_size)];
}
SFM_SVD_HD bool InvertMatrix4x4(const float m[16], float invOut[16]) { return sfmb200::inv4(m, invOut); }
// polar decomposition a = u p (svd.h:483-501)
SFM_SVD_HD void pd(const float* a, float* u, float* p) {
    float w[9], s[9], v[9], t[9], vt[9];
    svd(a, w, s, v);
    multAB(v, s, t);
    transpose_copy3x3(v, vt, 3, 3);
    multAB(t, vt, p);      // P = V S V^T
    multAB(w, vt, u);      // U = W V^T
}
#endif
