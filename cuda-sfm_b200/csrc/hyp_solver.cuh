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
#ifndef SFM_PROJ_ITERS
#define SFM_PROJ_ITERS 3        // projector solver: projection + (SFM_PROJ_ITERS - 1) re-projections against the design rows
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
        d += sfm_sqrt_approx(fmaf(dx, dx, dy * dy));
    }
    // The similarity only conditions the solve: ANY scale is exact as long as normalisation and de-normalisation use the
    // same one (they do), so hardware approximations (MUFU.RSQ / RCP, ~2 ulp) are free accuracy-wise and take ~150
    // instructions and 18 slow-path branches out of every hypothesis.
    s = 11.3137085f * sfm_rcp_approx(d);   // sqrt(2) / (d / 8)
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

// ---------------------------------------------------------------------------
// Alternative null-vector solve (SFMB200_OPT_HYP_SOLVER = 1): orthogonal
// projector onto null(A) through ONE 8x8 Cholesky factorisation,
//     e = (I - A^T (A A^T)^-1 A) r,
// ~8x fewer instructions than the 9x9 Jacobi eigensolve.  K = A A^T needs no
// design rows at all: <kron(p,q), kron(p',q')> = <p,p'><q,q'>.  The start
// vector r = e_i is the basis vector with the largest null-space component
// (diag of the projector, so |<r, n>| >= 1/3), and two re-projections against
// the design rows bring the error down to eps*cond like the Jacobi path's
// refinement.  Same pre/post-processing as solve_hypothesis().
// ---------------------------------------------------------------------------
struct Chol8 {
    float l[8][8];      // lower triangle; the diagonal holds 1 / L_ii
    bool ok;
};
SFM_HD void chol8_factor(float k[8][8], Chol8& c) {
    c.ok = true;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        float d = k[j][j];
#pragma unroll
        for (int m = 0; m < j; m++) d = fmaf(-c.l[j][m], c.l[j][m], d);
        c.ok = c.ok && (d > 0.0f);
        // 2-ulp reciprocal square root: the factor only has to be a good preconditioner, the re-projections against the
        // design rows in the callers remove what it leaves (same argument as for fp32 rounding in the factor itself)
        float inv = sfm_rsqrt(fmaxf(d, 1e-30f));
        c.l[j][j] = inv;
#pragma unroll
        for (int i = j + 1; i < 8; i++) {
            float v = k[i][j];
#pragma unroll
            for (int m = 0; m < j; m++) v = fmaf(-c.l[i][m], c.l[j][m], v);
            c.l[i][j] = v * inv;
        }
    }
}
// y <- K^-1 y  (forward + backward substitution)
SFM_HD void chol8_solve(const Chol8& c, float* y) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        float v = y[i];
#pragma unroll
        for (int m = 0; m < i; m++) v = fmaf(-c.l[i][m], y[m], v);
        y[i] = v * c.l[i][i];
    }
#pragma unroll
    for (int i = 7; i >= 0; i--) {
        float v = y[i];
#pragma unroll
        for (int m = i + 1; m < 8; m++) v = fmaf(-c.l[m][i], y[m], v);
        y[i] = v * c.l[i][i];
    }
}
// |L^-1 y|^2 = y^T K^-1 y (forward substitution only)
SFM_HD float chol8_quad(const Chol8& c, const float* y0) {
    float y[8], q = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        float v = y0[i];
#pragma unroll
        for (int m = 0; m < i; m++) v = fmaf(-c.l[i][m], y[m], v);
        y[i] = v * c.l[i][i];
        q = fmaf(y[i], y[i], q);
    }
    return q;
}

SFM_HD void solve_hypothesis_projector(const Corr* pts, float* E) {
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
    // K = A A^T through the Kronecker identity
    float k[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) {
            float a = fmaf(x1[i], x1[j], fmaf(y1[i], y1[j], 1.0f));
            float b = fmaf(x2[i], x2[j], fmaf(y2[i], y2[j], 1.0f));
            k[i][j] = a * b;
        }
    Chol8 ch;
    chol8_factor(k, ch);
    // diag of the projector: 1 - a_c^T K^-1 a_c for every column c of A; pick the largest
    int best = 0;
    float bestv = -1.0f;
#pragma unroll
    for (int c = 0; c < 9; c++) {
        float col[8];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            float a[9];
            design_row(x1[r], y1[r], x2[r], y2[r], a);
            col[r] = a[c];
        }
        float v = 1.0f - chol8_quad(ch, col);
        if (v > bestv) { bestv = v; best = c; }
    }
    float e[9];
#pragma unroll
    for (int c = 0; c < 9; c++) e[c] = (c == best) ? 1.0f : 0.0f;
    // e <- e - A^T K^-1 (A e), three times (projection + two re-projections)
#pragma unroll 1
    for (int it = 0; it < SFM_PROJ_ITERS; it++) {
        float y[8];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            float a[9];
            design_row(x1[r], y1[r], x2[r], y2[r], a);
            float res = a[0] * e[0];
#pragma unroll
            for (int c = 1; c < 9; c++) res = fmaf(a[c], e[c], res);
            y[r] = res;
        }
        chol8_solve(ch, y);
#pragma unroll
        for (int r = 0; r < 8; r++) {
            float a[9];
            design_row(x1[r], y1[r], x2[r], y2[r], a);
#pragma unroll
            for (int c = 0; c < 9; c++) e[c] = fmaf(-a[c], y[r], e[c]);
        }
        float n2 = e[0] * e[0];
#pragma unroll
        for (int c = 1; c < 9; c++) n2 = fmaf(e[c], e[c], n2);
        float inv = sfm_rsqrt(n2);          // keeps |e| ~ 1 between steps; the exact normalisation is project_essential's
#pragma unroll
        for (int c = 0; c < 9; c++) e[c] *= inv;
    }
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
    bool finite = ch.ok;
#pragma unroll
    for (int i = 0; i < 9; i++) finite = finite && (fabsf(E[i]) <= 3.0e38f);
    if (!finite) {
#pragma unroll
        for (int i = 0; i < 9; i++) E[i] = 0.0f;
    }
}

// ---------------------------------------------------------------------------
// 4-point homography hypothesis (replaces CudaSift's
// ComputeHomographies, matching.cu:907-948, an 8x8 Gaussian elimination with
// h8 = 1): Hartley-normalise, 8x9 DLT rows, null vector through the same 8x8
// Cholesky projector as the essential-matrix solver, de-normalise
// H = T2^-1 Hh T1, unit Frobenius norm.  Maps image-1 points to image-2 points.
// ---------------------------------------------------------------------------
SFM_HD void hartley4(const float* x, const float* y, float& s, float& cx, float& cy) {
    cx = 0.25f * (x[0] + x[1] + x[2] + x[3]);
    cy = 0.25f * (y[0] + y[1] + y[2] + y[3]);
    float d = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float dx = x[i] - cx, dy = y[i] - cy;
        d += sqrtf(fmaf(dx, dx, dy * dy));
    }
    s = 5.65685425f / d;   // sqrt(2) / (d / 4)
}
SFM_HD void solve_homography(const Corr* pts, float* Hm) {
    float x1[4], y1[4], x2[4], y2[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { x1[i] = pts[i].x1; y1[i] = pts[i].y1; x2[i] = pts[i].x2; y2[i] = pts[i].y2; }
    float s1, c1x, c1y, s2, c2x, c2y;
    hartley4(x1, y1, s1, c1x, c1y);
    hartley4(x2, y2, s2, c2x, c2y);
    float A[8][9];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float x = s1 * (x1[i] - c1x), y = s1 * (y1[i] - c1y);
        float u = s2 * (x2[i] - c2x), v = s2 * (y2[i] - c2y);
        float* r0 = A[2 * i];
        float* r1 = A[2 * i + 1];
        r0[0] = -x; r0[1] = -y; r0[2] = -1.0f; r0[3] = 0.0f; r0[4] = 0.0f; r0[5] = 0.0f; r0[6] = u * x; r0[7] = u * y; r0[8] = u;
        r1[0] = 0.0f; r1[1] = 0.0f; r1[2] = 0.0f; r1[3] = -x; r1[4] = -y; r1[5] = -1.0f; r1[6] = v * x; r1[7] = v * y; r1[8] = v;
    }
    float k[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) {
            float acc = A[i][0] * A[j][0];
#pragma unroll
            for (int c = 1; c < 9; c++) acc = fmaf(A[i][c], A[j][c], acc);
            k[i][j] = acc;
        }
    Chol8 ch;
    chol8_factor(k, ch);
    int best = 0;
    float bestv = -1.0f;
#pragma unroll
    for (int c = 0; c < 9; c++) {
        float col[8];
#pragma unroll
        for (int r = 0; r < 8; r++) col[r] = A[r][c];
        float v = 1.0f - chol8_quad(ch, col);
        if (v > bestv) { bestv = v; best = c; }
    }
    float e[9];
#pragma unroll
    for (int c = 0; c < 9; c++) e[c] = (c == best) ? 1.0f : 0.0f;
#pragma unroll 1
    for (int it = 0; it < 3; it++) {
        float y[8];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            float res = A[r][0] * e[0];
#pragma unroll
            for (int c = 1; c < 9; c++) res = fmaf(A[r][c], e[c], res);
            y[r] = res;
        }
        chol8_solve(ch, y);
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int c = 0; c < 9; c++) e[c] = fmaf(-A[r][c], y[r], e[c]);
        float n2 = e[0] * e[0];
#pragma unroll
        for (int c = 1; c < 9; c++) n2 = fmaf(e[c], e[c], n2);
        float inv = 1.0f / sqrtf(n2);
#pragma unroll
        for (int c = 0; c < 9; c++) e[c] *= inv;
    }
    // H = T2^-1 Hh T1 with T = [s 0 -s cx; 0 s -s cy; 0 0 1], T^-1 = [1/s 0 cx; 0 1/s cy; 0 0 1]
    float M[9];                       // M = Hh T1
#pragma unroll
    for (int i = 0; i < 3; i++) {
        M[3 * i + 0] = e[3 * i + 0] * s1;
        M[3 * i + 1] = e[3 * i + 1] * s1;
        M[3 * i + 2] = fmaf(-s1 * c1x, e[3 * i + 0], fmaf(-s1 * c1y, e[3 * i + 1], e[3 * i + 2]));
    }
    const float is2 = 1.0f / s2;
    float n2 = 0.0f;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        Hm[0 + j] = fmaf(c2x, M[6 + j], is2 * M[0 + j]);
        Hm[3 + j] = fmaf(c2y, M[6 + j], is2 * M[3 + j]);
        Hm[6 + j] = M[6 + j];
    }
#pragma unroll
    for (int i = 0; i < 9; i++) n2 = fmaf(Hm[i], Hm[i], n2);
    float inv = 1.0f / sqrtf(n2);
    bool finite = ch.ok;
#pragma unroll
    for (int i = 0; i < 9; i++) { Hm[i] *= inv; finite = finite && (fabsf(Hm[i]) <= 3.0e38f); }
    if (!finite) {
#pragma unroll
        for (int i = 0; i < 9; i++) Hm[i] = 0.0f;
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

// The reference's sampling scheme (sfm.cu:95-104): ONE random permutation of the point indices cut into H = N / 8
// DISJOINT groups of 8, so no correspondence is used by two hypotheses (SFMB200_OPT_SAMPLER = 1).  The reference
// shuffles on the host (std::shuffle + std::random_device) and uploads the rows; here row h is computed where it is
// needed from a keyed permutation of [0, n) - a 4-round Feistel network over 2 * half bits (2^(2 half) >= n) with cycle
// walking - so any hypothesis slice regenerates on any GPU, and the oracle mirrors it bit for bit
// (oracle.py: sample_indices_disjoint).  Needs 8 * H_total <= n.
SFM_HD unsigned int perm_mix(unsigned int v, unsigned int key) {
    v ^= key;
    v *= 0x85EBCA6Bu;
    v ^= v >> 13;
    v *= 0xC2B2AE35u;
    v ^= v >> 16;
    return v;
}
SFM_HD void sample_indices_disjoint(unsigned long long seed, unsigned long long h, int n, int* idx) {
    const unsigned long long k0 = splitmix64(seed ^ 0xA5A5A5A5A5A5A5A5ull), k1 = splitmix64(k0);
    const unsigned int key[4] = {(unsigned int)k0, (unsigned int)(k0 >> 32), (unsigned int)k1, (unsigned int)(k1 >> 32)};
    int bits = 1;
    while ((1ll << bits) < (long long)n) bits++;
    const int half = (bits + 1) / 2;
    const unsigned int mask = (1u << half) - 1u;
#pragma unroll 1
    for (int j = 0; j < 8; j++) {
        unsigned int x = (unsigned int)(8ull * h + (unsigned long long)j);
        if (x >= (unsigned int)n) { idx[j] = -1; continue; }          // beyond the permutation: a degenerate row
        do {
            unsigned int L = x >> half, R = x & mask;
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const unsigned int t = L ^ (perm_mix(R, key[r]) & mask);
                L = R;
                R = t;
            }
            x = (L << half) | R;
        } while (x >= (unsigned int)n);
        idx[j] = (int)x;
    }
}

#if defined(__CUDACC__)
// The 8 correspondence indices of hypothesis hg (caller's rows, or one of the two samplers); false = not 8 distinct
// indices inside [0, n) - the hypothesis is void (E = 0) and the indices are replaced by 0.
__device__ __forceinline__ bool sample_ids(int n, const int32_t* __restrict__ idx_rows, unsigned long long seed, long long hg,
                                           int sampler, int* id) {
    if (idx_rows != nullptr) {
        const int4* row = reinterpret_cast<const int4*>(idx_rows + 8 * hg);
        int4 a = __ldg(row), b = __ldg(row + 1);
        id[0] = a.x; id[1] = a.y; id[2] = a.z; id[3] = a.w;
        id[4] = b.x; id[5] = b.y; id[6] = b.z; id[7] = b.w;
    } else if (sampler == 1) {
        sample_indices_disjoint(seed, (unsigned long long)hg, n, id);
    } else {
        sample_indices(seed, (unsigned long long)hg, n, id);
    }
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        ok = ok && (id[i] >= 0) && (id[i] < n);
#pragma unroll
        for (int j = 0; j < i; j++) ok = ok && (id[i] != id[j]);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) id[i] = ok ? id[i] : 0;
    return ok;
}
// Loads the 8 sampled correspondences of hypothesis (pair b, global index hg).
// A sample with an out-of-range or repeated index is degenerate: returns false.
// COHERENT: the correspondences were written earlier in the SAME kernel (small.cu): read them through L2 (ld.global.cg)
// instead of the non-coherent read-only path.
template <bool COHERENT = false>
__device__ __forceinline__ bool load_sample(const float4* __restrict__ corr, int n, const int32_t* __restrict__ idx_rows,
                                            unsigned long long seed, long long hg, Corr* pts, int sampler = 0) {
    int id[8];
    const bool ok = sample_ids(n, idx_rows, seed, hg, sampler, id);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        float4 c = COHERENT ? __ldcg(corr + id[i]) : __ldg(corr + id[i]);
        pts[i] = Corr{c.x, c.y, c.z, c.w};
    }
    return ok;
}
#endif

}  // namespace sfmb200
