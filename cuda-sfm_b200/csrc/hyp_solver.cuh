// One RANSAC hypothesis: 8 correspondences -> essential matrix candidate.
//
// Replaces the reference chain K3 kernels::kernels (SfM/kernels.h:236-259),
// K4+K5 regular_svd / cusolverDnSgesvdjBatched 8x9 (kernels.h:196-234),
// K6 row_extraction_kernel (kernels.h:452-458) and K7 normalizeE
// (kernels.h:281-295) with one register-resident solve per thread:
//
//   Hartley-normalise both point sets -> 8x9 design rows kron(x1, x2)
//   (reference convention: x1^T E x2 = 0) -> 9x9 Gram matrix ->
//   cyclic Jacobi eigensolve with accumulated V -> eigenvector of the smallest
//   eigenvalue -> refinement steps that use the design rows themselves (the
//   Gram matrix squares the condition number; the refinement brings the error
//   back to eps*cond, which the 1e-4 parity bar needs) -> de-normalise ->
//   rank-2 projection U diag(1,1,0) V^T.
#pragma once
#include "smallmat.cuh"

namespace sfmb200 {

struct Corr { float x1, y1, x2, y2; };

#ifndef SFM_HYP_SWEEPS
#define SFM_HYP_SWEEPS 5
#endif
#ifndef SFM_HYP_REFINE
#define SFM_HYP_REFINE 3
#endif
#ifndef SFM_HYP_FAST_ANGLE
#define SFM_HYP_FAST_ANGLE 1     // device: approximate rotation angles in the 9x9 eigensolve
#endif

SFM_HD void design_row(float x1, float y1, float x2, float y2, float* a) {
    a[0] = x1 * x2; a[1] = x1 * y2; a[2] = x1;
    a[3] = y1 * x2; a[4] = y1 * y2; a[5] = y1;
    a[6] = x2;      a[7] = y2;      a[8] = 1.0f;
}

// Hartley similarity for one image: returns scale s and centroid (cx, cy) so
// that xh = s * (x - cx) has zero mean and mean distance sqrt(2).
SFM_HD void hartley(const float* x, const float* y, float& s, float& cx, float& cy) {
    float sx = 0.0f, sy = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; i++) { sx += x[i]; sy += y[i]; }
    cx = sx * 0.125f;
    cy = sy * 0.125f;
    float d = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        float dx = x[i] - cx, dy = y[i] - cy;
        d += sqrtf(fmaf(dx, dx, dy * dy));
    }
    s = 11.3137085f / d;   // sqrt(2) / (d / 8)
}

// pts: the 8 sampled correspondences in normalised camera coordinates.
// E (row-major 3x3): projected essential matrix, Frobenius norm sqrt(2);
// all zeros when the sample is degenerate (non-finite anywhere).
// SYNC (device only): 0 = none; 1 = __syncthreads() before every sweep; 2 = before
// every row of rotations.  A sweep is ~3,000 unrolled instructions (~46 KB), more
// than the instruction cache: keeping the CTA's warps in step lets them share
// each fetched line instead of thrashing it.  Every thread of the CTA must call.
template <int SYNC>
SFM_HD void sweep_sync() {
#if defined(__CUDA_ARCH__)
    if (SYNC > 0) __syncthreads();
#endif
}

template <int SYNC = 0>
SFM_HD void solve_hypothesis(const Corr* pts, float* E) {
    float x1[8], y1[8], x2[8], y2[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x1[i] = pts[i].x1; y1[i] = pts[i].y1; x2[i] = pts[i].x2; y2[i] = pts[i].y2; }
    float s1, c1x, c1y, s2, c2x, c2y;
    hartley(x1, y1, s1, c1x, c1y);
    hartley(x2, y2, s2, c2x, c2y);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        x1[i] = s1 * (x1[i] - c1x); y1[i] = s1 * (y1[i] - c1y);
        x2[i] = s2 * (x2[i] - c2x); y2[i] = s2 * (y2[i] - c2y);
    }
    // Gram matrix, upper triangle
    float g[9][9];
#pragma unroll
    for (int i = 0; i < 9; i++)
#pragma unroll
        for (int j = 0; j < 9; j++) g[i][j] = 0.0f;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        float a[9];
        design_row(x1[r], y1[r], x2[r], y2[r], a);
#pragma unroll
        for (int i = 0; i < 9; i++)
#pragma unroll
            for (int j = i; j < 9; j++) g[i][j] = fmaf(a[i], a[j], g[i][j]);
    }
    float V[9][9];
#pragma unroll
    for (int i = 0; i < 9; i++)
#pragma unroll
        for (int j = 0; j < 9; j++) V[i][j] = (i == j) ? 1.0f : 0.0f;

#pragma unroll 1
    for (int sw = 0; sw < SFM_HYP_SWEEPS; sw++) {
        sweep_sync<(SYNC >= 1) ? 1 : 0>();
#pragma unroll
        for (int p = 0; p < 8; p++) {
            if (p > 0) sweep_sync<(SYNC >= 2) ? 1 : 0>();
#pragma unroll
            for (int q = p + 1; q < 9; q++) {
                float c, s, t;
#if SFM_HYP_FAST_ANGLE
                jacobi_angle_fast(g[p][p], g[q][q], g[p][q], c, s, t);
#else
                jacobi_angle(g[p][p], g[q][q], g[p][q], c, s, t);
#endif
                g[p][p] = fmaf(-t, g[p][q], g[p][p]);
                g[q][q] = fmaf(t, g[p][q], g[q][q]);
                g[p][q] = 0.0f;
#pragma unroll
                for (int k = 0; k < 9; k++) {
                    if (k == p || k == q) continue;
                    float& akp = (k < p) ? g[k][p] : g[p][k];
                    float& akq = (k < q) ? g[k][q] : g[q][k];
                    float a = akp, b = akq;
                    akp = fmaf(c, a, -s * b);
                    akq = fmaf(s, a, c * b);
                }
#pragma unroll
                for (int k = 0; k < 9; k++) {
                    float a = V[k][p], b = V[k][q];
                    V[k][p] = fmaf(c, a, -s * b);
                    V[k][q] = fmaf(s, a, c * b);
                }
            }
        }
    }
    float lam[9];
#pragma unroll
    for (int i = 0; i < 9; i++) lam[i] = g[i][i];
    int m = 0;
    float lmin = lam[0], lmax = lam[0];
#pragma unroll
    for (int i = 1; i < 9; i++) {
        if (lam[i] < lmin) { lmin = lam[i]; m = i; }
        lmax = fmaxf(lmax, lam[i]);
    }
    float e[9];
#pragma unroll
    for (int k = 0; k < 9; k++) {
        float val = V[k][0];
#pragma unroll
        for (int i = 1; i < 9; i++) val = (m == i) ? V[k][i] : val;
        e[k] = val;
    }
    // refinement with the design rows: e <- e - G~^+ A^T (A e)
#pragma unroll 1
    for (int it = 0; it < SFM_HYP_REFINE; it++) {
        float gg[9];
#pragma unroll
        for (int k = 0; k < 9; k++) gg[k] = 0.0f;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            float a[9];
            design_row(x1[r], y1[r], x2[r], y2[r], a);
            float res = a[0] * e[0];
#pragma unroll
            for (int k = 1; k < 9; k++) res = fmaf(a[k], e[k], res);
#pragma unroll
            for (int k = 0; k < 9; k++) gg[k] = fmaf(a[k], res, gg[k]);
        }
#pragma unroll
        for (int i = 0; i < 9; i++) {
            float d = V[0][i] * gg[0];
#pragma unroll
            for (int k = 1; k < 9; k++) d = fmaf(V[k][i], gg[k], d);
            bool use = (i != m) && (lam[i] > 1e-10f * lmax);
            float coef = use ? d / lam[i] : 0.0f;
#pragma unroll
            for (int k = 0; k < 9; k++) e[k] = fmaf(-coef, V[k][i], e[k]);
        }
        float n2 = e[0] * e[0];
#pragma unroll
        for (int k = 1; k < 9; k++) n2 = fmaf(e[k], e[k], n2);
        float inv = 1.0f / sqrtf(n2);
#pragma unroll
        for (int k = 0; k < 9; k++) e[k] *= inv;
    }
    // de-normalise: E = T1^T Eh T2
    float M[9];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        M[3 * i + 0] = e[3 * i + 0] * s2;
        M[3 * i + 1] = e[3 * i + 1] * s2;
        M[3 * i + 2] = fmaf(-s2 * c2x, e[3 * i + 0], fmaf(-s2 * c2y, e[3 * i + 1], e[3 * i + 2]));
    }
#pragma unroll
    for (int j = 0; j < 3; j++) {
        E[0 + j] = s1 * M[0 + j];
        E[3 + j] = s1 * M[3 + j];
        E[6 + j] = fmaf(-s1 * c1x, M[0 + j], fmaf(-s1 * c1y, M[3 + j], M[6 + j]));
    }
    project_essential(E);
    bool finite = true;
#pragma unroll
    for (int i = 0; i < 9; i++) finite = finite && (fabsf(E[i]) <= 3.0e38f);   // false for NaN/Inf
    if (!finite) {
#pragma unroll
        for (int i = 0; i < 9; i++) E[i] = 0.0f;
    }
}

// Counter-based sample-index generator shared (bit-exactly) with the oracle
// (oracle/oracle.py: sample_indices).  Hypothesis h of a pair draws 8 distinct
// indices in [0, n): splitmix64 of (seed, h, draw counter), mapped with a
// multiply-shift, redrawn while it repeats an earlier index.  n >= 8.
SFM_HD unsigned long long splitmix64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
SFM_HD void sample_indices(unsigned long long seed, unsigned long long h, int n, int* idx) {
    unsigned long long key = splitmix64(seed ^ (h * 0xD1342543DE82EF95ull));
    unsigned int ctr = 0;
#pragma unroll 1
    for (int j = 0; j < 8; j++) {
        int cand;
        bool dup;
        do {
            unsigned long long r = splitmix64(key + ctr);
            ctr++;
            cand = (int)(((r >> 32) * (unsigned long long)n) >> 32);
            dup = false;
            for (int k = 0; k < j; k++) dup = dup || (idx[k] == cand);
        } while (dup);
        idx[j] = cand;
    }
}

}  // namespace sfmb200
