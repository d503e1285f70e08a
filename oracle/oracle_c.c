/* C restatement of the reference hot path (TEST INFRASTRUCTURE ONLY).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product never does.
 *
 * Reference: Black-Phoenix/CUDA-SfM.  What each function restates:
 *   oracle_hypotheses_f64   kernels::kernels (SfM/kernels.h:236-259) +
 *                           regular_svd / cusolverDnSgesvdjBatched 8x9
 *                           (kernels.h:211-234; gesvdj = one-sided Jacobi,
 *                           restated here in fp64) + row_extraction_kernel
 *                           (kernels.h:452-458) + normalizeE (kernels.h:281-295)
 *   oracle_counts_f64/_f32  calculateInliers' stated intent (SfM/sfm.cu:155-221)
 *                           with the Sampson error BASELINE.json mandates and
 *                           the 1e-6 literal of sfm.cu:220
 *   oracle_argmax_first     thrust::max_element (sfm.cu:135-136), without the
 *                           reference's off-by-one (sfm.cu:137)
 *   oracle_triangulate_f64  compute_linear_triangulation_A (kernels.h:387-431) +
 *                           svd_square 4x4 (kernels.h:175-194) +
 *                           normalize_pt_kernal (kernels.h:433-450)
 * Parity pinning: see the header of oracle/oracle.py.
 *
 * oracle_counts_f32 evaluates the SAME fp32 fma tree as the CUDA scoring
 * kernel (cuda-sfm_b200/csrc/score.cu: sampson_d), so its counts must equal
 * the GPU's bit for bit; oracle_counts_f64 is the mathematical truth with a
 * borderline band.  Build with -ffp-contract=off (see oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---------- one-sided Jacobi: right singular vectors of an m x n matrix ---------- */
/* A is m x n row-major (overwritten by A V); V n x n row-major.  Returns sweeps used. */
static int jacobi_right(double* A, double* V, int m, int n, int max_sweeps, double tol) {
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) V[i * n + j] = (i == j) ? 1.0 : 0.0;
    int sw;
    for (sw = 0; sw < max_sweeps; sw++) {
        int rotated = 0;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                double a = 0, b = 0, c = 0;
                for (int k = 0; k < m; k++) {
                    a += A[k * n + p] * A[k * n + p];
                    b += A[k * n + q] * A[k * n + q];
                    c += A[k * n + p] * A[k * n + q];
                }
                if (fabs(c) <= tol * sqrt(a * b) || c == 0.0) continue;
                rotated = 1;
                double theta = 0.5 * (b - a) / c;
                double t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));
                if (theta < 0) t = -t;
                double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
                for (int k = 0; k < m; k++) {
                    double x = A[k * n + p], y = A[k * n + q];
                    A[k * n + p] = cs * x - sn * y;
                    A[k * n + q] = sn * x + cs * y;
                }
                for (int k = 0; k < n; k++) {
                    double x = V[k * n + p], y = V[k * n + q];
                    V[k * n + p] = cs * x - sn * y;
                    V[k * n + q] = sn * x + cs * y;
                }
            }
        if (!rotated) break;
    }
    return sw;
}

/* column of V belonging to the smallest column norm of A V */
static void smallest_right_vector(const double* AV, const double* V, int m, int n, double* out) {
    int best = 0;
    double bn = INFINITY;
    for (int j = 0; j < n; j++) {
        double s = 0;
        for (int k = 0; k < m; k++) s += AV[k * n + j] * AV[k * n + j];
        if (s < bn) { bn = s; best = j; }
    }
    for (int k = 0; k < n; k++) out[k] = V[k * n + best];
}

/* E <- U diag(1,1,0) V^T through a 3x3 one-sided Jacobi SVD */
static void project_essential(double* E) {
    double B[9], V[9];
    memcpy(B, E, sizeof(B));
    jacobi_right(B, V, 3, 3, 60, 1e-15);
    double nrm[3];
    int ord[3] = {0, 1, 2};
    for (int j = 0; j < 3; j++) nrm[j] = sqrt(B[j] * B[j] + B[3 + j] * B[3 + j] + B[6 + j] * B[6 + j]);
    for (int i = 0; i < 2; i++)
        for (int j = i + 1; j < 3; j++)
            if (nrm[ord[j]] > nrm[ord[i]]) { int t = ord[i]; ord[i] = ord[j]; ord[j] = t; }
    memset(E, 0, 9 * sizeof(double));
    for (int r = 0; r < 2; r++) {
        int j = ord[r];
        if (nrm[j] == 0.0) continue;
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) E[3 * a + b] += (B[3 * a + j] / nrm[j]) * V[3 * b + j];
    }
}

/* x: n x 4 (x1,y1,x2,y2) fp32 normalised coords; idx: H x 8; E out: H x 9 fp64 */
void oracle_hypotheses_f64(const float* x, int n, const int32_t* idx, int H, double* E) {
    (void)n;
#pragma omp parallel for schedule(static)
    for (int h = 0; h < H; h++) {
        double A[72], V[81], e[9];
        for (int r = 0; r < 8; r++) {
            const float* p = x + 4 * (size_t)idx[8 * (size_t)h + r];
            double x1 = p[0], y1 = p[1], x2 = p[2], y2 = p[3];
            double* a = A + 9 * r;
            a[0] = x1 * x2; a[1] = x1 * y2; a[2] = x1;
            a[3] = y1 * x2; a[4] = y1 * y2; a[5] = y1;
            a[6] = x2;      a[7] = y2;      a[8] = 1.0;
        }
        jacobi_right(A, V, 8, 9, 60, 1e-15);
        smallest_right_vector(A, V, 8, 9, e);
        project_essential(e);
        memcpy(E + 9 * (size_t)h, e, sizeof(e));
    }
}

/* ---------- scoring ---------- */
/* Mirrors cuda-sfm_b200/csrc/sampson.cuh: the threshold is folded into the coordinates (k = sqrt(thr),
 * points * 1/k, E~ = D E D with D = diag(k,k,1)), then d = num~^2 - den~ with one fma tree. */
static inline void thr_scale_E(const float* e, float thr, float* s) {
    const float k = sqrtf(thr), k2 = k * k;
    for (int q = 0; q < 9; q++) {
        float f = (q == 2 || q == 5 || q == 6 || q == 7) ? k : k2;
        s[q] = (q == 8) ? e[q] : e[q] * f;
    }
}
static inline float sampson_unit_d_f32(const float* s, float x1, float y1, float x2, float y2) {
    float l0 = fmaf(s[0], x2, fmaf(s[1], y2, s[2]));
    float l1 = fmaf(s[3], x2, fmaf(s[4], y2, s[5]));
    float l2 = fmaf(s[6], x2, fmaf(s[7], y2, s[8]));
    float num = fmaf(x1, l0, fmaf(y1, l1, l2));
    float m0 = fmaf(s[0], x1, fmaf(s[3], y1, s[6]));
    float m1 = fmaf(s[1], x1, fmaf(s[4], y1, s[7]));
    float den = fmaf(l0, l0, fmaf(l1, l1, fmaf(m0, m0, m1 * m1)));
    return fmaf(num, num, -den);
}
static inline float sampson_d_f32(const float* e, float x1, float y1, float x2, float y2, float nthr) {
    const float ik = 1.0f / sqrtf(-nthr);
    float s[9];
    thr_scale_E(e, -nthr, s);
    return sampson_unit_d_f32(s, x1 * ik, y1 * ik, x2 * ik, y2 * ik);
}

/* E: H x 9 fp32 (exactly what the GPU scored); counts[h] = #{i : d < 0}.  Like the GPU path the coordinates are
 * scaled once and each hypothesis once (same operations as sampson_d_f32, hoisted). */
void oracle_counts_f32(const float* E, int H, const float* x, int n, float thr, int32_t* counts) {
    const float ik = 1.0f / sqrtf(thr);
    float* xs = (float*)malloc(sizeof(float) * 4 * (size_t)n);
    for (size_t i = 0; i < 4 * (size_t)n; i++) xs[i] = x[i] * ik;
#pragma omp parallel for schedule(static)
    for (int h = 0; h < H; h++) {
        float s[9];
        thr_scale_E(E + 9 * (size_t)h, thr, s);
        int c = 0;
        for (int i = 0; i < n; i++) {
            const float* p = xs + 4 * (size_t)i;
            float d = sampson_unit_d_f32(s, p[0], p[1], p[2], p[3]);
            uint32_t bits;
            memcpy(&bits, &d, 4);
            c += (int)(bits >> 31);
        }
        counts[h] = c;
    }
    free(xs);
}

/* Symmetric epipolar distance (SFMB200_OPT_SCORE_METRIC = 1; the intent of sfm.cu:155-221): mirrors
 * sampson.cuh:symmetric_unit_d - same scaling, d = num~^2 (A + B) - A B, A = l0^2 + l1^2, B = m0^2 + m1^2. */
static inline float symmetric_unit_d_f32(const float* s, float x1, float y1, float x2, float y2) {
    float l0 = fmaf(s[0], x2, fmaf(s[1], y2, s[2]));
    float l1 = fmaf(s[3], x2, fmaf(s[4], y2, s[5]));
    float l2 = fmaf(s[6], x2, fmaf(s[7], y2, s[8]));
    float num = fmaf(x1, l0, fmaf(y1, l1, l2));
    float m0 = fmaf(s[0], x1, fmaf(s[3], y1, s[6]));
    float m1 = fmaf(s[1], x1, fmaf(s[4], y1, s[7]));
    float A = fmaf(l0, l0, l1 * l1);
    float B = fmaf(m0, m0, m1 * m1);
    return fmaf(num * num, A + B, -(A * B));
}
void oracle_counts_sym_f32(const float* E, int H, const float* x, int n, float thr, int32_t* counts) {
    const float ik = 1.0f / sqrtf(thr);
    float* xs = (float*)malloc(sizeof(float) * 4 * (size_t)n);
    for (size_t i = 0; i < 4 * (size_t)n; i++) xs[i] = x[i] * ik;
#pragma omp parallel for schedule(static)
    for (int h = 0; h < H; h++) {
        float s[9];
        thr_scale_E(E + 9 * (size_t)h, thr, s);
        int c = 0;
        for (int i = 0; i < n; i++) {
            const float* p = xs + 4 * (size_t)i;
            float d = symmetric_unit_d_f32(s, p[0], p[1], p[2], p[3]);
            uint32_t bits;
            memcpy(&bits, &d, 4);
            c += (int)(bits >> 31);
        }
        counts[h] = c;
    }
    free(xs);
}
/* fp64 truth: inlier <=> num^2 (A + B) < thr A B; borderline within band (relative to thr A B) */
void oracle_counts_sym_f64(const float* E, int H, const float* x, int n, double thr, double band, int32_t* counts,
                           int32_t* borderline) {
#pragma omp parallel for schedule(static)
    for (int h = 0; h < H; h++) {
        double e[9];
        for (int k = 0; k < 9; k++) e[k] = E[9 * (size_t)h + k];
        int c = 0, bl = 0;
        for (int i = 0; i < n; i++) {
            const float* p = x + 4 * (size_t)i;
            double x1 = p[0], y1 = p[1], x2 = p[2], y2 = p[3];
            double l0 = e[0] * x2 + e[1] * y2 + e[2], l1 = e[3] * x2 + e[4] * y2 + e[5], l2 = e[6] * x2 + e[7] * y2 + e[8];
            double num = x1 * l0 + y1 * l1 + l2;
            double m0 = e[0] * x1 + e[3] * y1 + e[6], m1 = e[1] * x1 + e[4] * y1 + e[7];
            double A = l0 * l0 + l1 * l1, B = m0 * m0 + m1 * m1;
            double d = num * num * (A + B) - thr * A * B;
            c += d < 0;
            bl += fabs(d) <= band * thr * A * B + 1e-300;
        }
        counts[h] = c;
        if (borderline) borderline[h] = bl;
    }
}

/* fp64 truth on the same fp32 inputs; borderline[h] = #{i : |num^2 - thr*den| <= band*thr*den} */
void oracle_counts_f64(const float* E, int H, const float* x, int n, double thr, double band, int32_t* counts,
                       int32_t* borderline) {
#pragma omp parallel for schedule(static)
    for (int h = 0; h < H; h++) {
        double e[9];
        for (int k = 0; k < 9; k++) e[k] = E[9 * (size_t)h + k];
        int c = 0, bl = 0;
        for (int i = 0; i < n; i++) {
            const float* p = x + 4 * (size_t)i;
            double x1 = p[0], y1 = p[1], x2 = p[2], y2 = p[3];
            double l0 = e[0] * x2 + e[1] * y2 + e[2], l1 = e[3] * x2 + e[4] * y2 + e[5], l2 = e[6] * x2 + e[7] * y2 + e[8];
            double num = x1 * l0 + y1 * l1 + l2;
            double m0 = e[0] * x1 + e[3] * y1 + e[6], m1 = e[1] * x1 + e[4] * y1 + e[7];
            double den = l0 * l0 + l1 * l1 + m0 * m0 + m1 * m1;
            double d = num * num - thr * den;
            c += d < 0;
            bl += fabs(d) <= band * thr * den + 1e-300;
        }
        counts[h] = c;
        if (borderline) borderline[h] = bl;
    }
}

int oracle_argmax_first(const int32_t* counts, int H) {
    int best = 0;
    for (int h = 1; h < H; h++)
        if (counts[h] > counts[best]) best = h;
    return best;
}

/* ---------- triangulation ---------- */
/* x: n x 4; M: 4x4 row-major camera 2 (camera 1 = I4); out: 4 x n SoA fp64 */
void oracle_triangulate_f64(const float* x, int n, const double* M, double* out) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        const float* p = x + 4 * (size_t)i;
        double A[16] = {-1, 0, p[0], 0, 0, -1, p[1], 0};
        for (int k = 0; k < 4; k++) {
            A[8 + k] = (double)p[2] * M[8 + k] - M[k];
            A[12 + k] = (double)p[3] * M[8 + k] - M[4 + k];
        }
        double V[16], v[4];
        jacobi_right(A, V, 4, 4, 60, 1e-15);
        smallest_right_vector(A, V, 4, 4, v);
        if (v[3] == 0.0) {
            out[i] = out[n + i] = out[2 * (size_t)n + i] = 0.0;
        } else {
            out[i] = v[0] / v[3];
            out[n + i] = v[1] / v[3];
            out[2 * (size_t)n + i] = v[2] / v[3];
        }
        out[3 * (size_t)n + i] = 1.0;
    }
}

#ifdef _OPENMP
#include <omp.h>
int oracle_threads(void) { return omp_get_max_threads(); }
void oracle_set_threads(int t) { omp_set_num_threads(t); }
#else
int oracle_threads(void) { return 1; }
void oracle_set_threads(int t) { (void)t; }
#endif
