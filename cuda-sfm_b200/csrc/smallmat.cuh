// Small fixed-size matrix kernels shared by every stage of the hot path.
//
// Everything here is __host__ __device__ so the same code backs the device
// kernels (hypgen / pose / triangulate) and the host-callable svd.h facade
// (reference surface: SfM/svd.h:33-501).  The algorithms are our own (one-sided
// Jacobi + Givens QR for the 3x3 SVD; Jacobi / inverse iteration for the 4x4
// null vector), with the reference's output contract for svd() (SfM/svd.h:311-335):
//   a = u * s * v^T, v NOT transposed, s (nearly) diagonal with
//   |s00| >= |s11| >= |s22|, u and v proper rotations (s22 may be negative).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#ifndef SFM_HD
#define SFM_HD __host__ __device__ __forceinline__
#endif

namespace sfmb200 {

// Raw MUFU.RSQ (rsqrt.approx.ftz.f32): rsqrtf() without -use_fast_math wraps the same instruction in denormal handling
// (two FSETP, an FSEL and two FMUL by 2^24 / 2^12 per call); every caller here feeds it a normal number (sums of squares
// behind explicit floors), for which the two give the same bits.
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ float sfm_mufu_rsq(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
#endif

// Symmetric Jacobi rotation that annihilates a_pq.  Returns (c, s, t) of the
// rotation J = [c s; -s c] applied as A <- J^T A J, i.e. the classic
//   t = sgn(theta) / (|theta| + sqrt(theta^2 + 1)), theta = (a_qq - a_pp) / (2 a_pq).
// Tiny off-diagonals give the identity (no NaN from 0/0).
SFM_HD void jacobi_angle(float app, float aqq, float apq, float& c, float& s, float& t) {
    const float tiny = 1e-37f;
    bool skip = fabsf(apq) < tiny;
    float theta = 0.5f * (aqq - app) / (skip ? 1.0f : apq);
    float at = fabsf(theta);
    // for |theta| large, sqrt(theta^2+1) overflows to inf and t -> 0: fine.
    float tt = 1.0f / (at + sqrtf(fmaf(theta, theta, 1.0f)));
    tt = theta < 0.0f ? -tt : tt;
    tt = skip ? 0.0f : tt;
    float cc = 1.0f / sqrtf(fmaf(tt, tt, 1.0f));
    c = cc;
    s = tt * cc;
    t = tt;
}

// Same rotation from hardware approximations (MUFU.RSQ / MUFU.RCP): a Jacobi
// sweep converges for any angle close to the annihilating one, and (c, s) stays
// orthonormal to ~2 ulp because s = t*c with c = rsqrt(1 + t^2).  Used where
// the rotation sequence is followed by a refinement against the original
// matrix (hyp_solver.cuh); it shortens the serial angle chain ~4x.
//   t = beta / (alpha + sgn(alpha) * sqrt(alpha^2 + beta^2)), alpha = (a_qq - a_pp)/2, beta = a_pq
SFM_HD void jacobi_angle_fast(float app, float aqq, float apq, float& c, float& s, float& t) {
#if defined(__CUDA_ARCH__)
    float alpha = 0.5f * (aqq - app);
    float rho2 = fmaf(alpha, alpha, apq * apq);
    bool skip = !(rho2 > 1e-36f) || apq == 0.0f;
    float rho = rho2 * sfm_mufu_rsq(rho2);
    float denom = alpha + copysignf(rho, alpha);
    float tt = skip ? 0.0f : __fdividef(apq, denom);
    float cc = sfm_mufu_rsq(fmaf(tt, tt, 1.0f));
    c = cc;
    s = tt * cc;
    t = tt;
#else
    jacobi_angle(app, aqq, apq, c, s, t);
#endif
}

// 1/sqrt(x): MUFU.RSQ (2 ulp) on the device, exact on the host.  Only for places
// where the result feeds a self-correcting iteration or a normalisation.
SFM_HD float sfm_rsqrt(float x) {
#if defined(__CUDA_ARCH__)
    return sfm_mufu_rsq(x);
#else
    return 1.0f / sqrtf(x);
#endif
}

// sqrt(x) and 1/x from MUFU.RSQ / MUFU.RCP on the device (x >= 0; 0 -> 0), exact on the host.  For quantities whose
// exact value does not matter (conditioning scales).
SFM_HD float sfm_sqrt_approx(float x) {
#if defined(__CUDA_ARCH__)
    return x > 1e-30f ? x * sfm_mufu_rsq(x) : 0.0f;
#else
    return sqrtf(x);
#endif
}
SFM_HD float sfm_rcp_approx(float x) {
#if defined(__CUDA_ARCH__)
    return __fdividef(1.0f, x);
#else
    return 1.0f / x;
#endif
}

// Givens pair (c, s) with c*a + s*b = r >= 0 and -s*a + c*b = 0.
SFM_HD void givens(float a, float b, float& c, float& s) {
    float r2 = fmaf(a, a, b * b);
    if (r2 < 1e-37f) { c = 1.0f; s = 0.0f; return; }
#if defined(__CUDA_ARCH__)
    float ir = sfm_mufu_rsq(r2);
#else
    float ir = 1.0f / sqrtf(r2);
#endif
    c = a * ir;
    s = b * ir;
}

SFM_HD void mul33(const float* a, const float* b, float* m) {      // m = a b
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            m[3 * i + j] = fmaf(a[3 * i + 2], b[6 + j], fmaf(a[3 * i + 1], b[3 + j], a[3 * i] * b[j]));
}
SFM_HD void mul33_AtB(const float* a, const float* b, float* m) {  // m = a^T b
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            m[3 * i + j] = fmaf(a[6 + i], b[6 + j], fmaf(a[3 + i], b[3 + j], a[i] * b[j]));
}
SFM_HD void mul33_ABt(const float* a, const float* b, float* m) {  // m = a b^T
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            m[3 * i + j] = fmaf(a[3 * i + 2], b[3 * j + 2], fmaf(a[3 * i + 1], b[3 * j + 1], a[3 * i] * b[3 * j]));
}

SFM_HD float det33(const float* a) {
    return a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) +
           a[2] * (a[3] * a[7] - a[4] * a[6]);
}
// The reference's det() as written (SfM/svd.h:337-341): third term uses a0
// where the cofactor expansion needs a1.  Needed for compat pose candidates
// (SURVEY Q15); never use it for anything else.
SFM_HD float det33_reference_typo(const float* a) {
    return a[0] * a[4] * a[8] - a[0] * a[5] * a[7] - a[0] * a[3] * a[8] + a[1] * a[5] * a[6] +
           a[2] * a[3] * a[7] - a[2] * a[4] * a[6];
}

// 3x3 SVD, contract in the file header.  One-sided (Hestenes) Jacobi: the
// columns of B = a V are rotated pairwise until orthogonal, which keeps high
// relative accuracy for small singular values (a Gram-matrix eigensolve would
// lose sigma_2 of a nearly rank-1 matrix to fp32 rounding, and the rank-2
// projection of E needs it).
template <int SWEEPS = 5>
SFM_HD void svd3(const float* a, float* u, float* s, float* v) {
    float B[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
#pragma unroll
    for (int i = 0; i < 9; i++) B[i] = a[i];
    float c, sn, t;
#pragma unroll 1
    for (int sw = 0; sw < SWEEPS; sw++) {
#pragma unroll
        for (int p = 0; p < 2; p++)
#pragma unroll
            for (int q = p + 1; q < 3; q++) {
                float app = fmaf(B[6 + p], B[6 + p], fmaf(B[3 + p], B[3 + p], B[p] * B[p]));
                float aqq = fmaf(B[6 + q], B[6 + q], fmaf(B[3 + q], B[3 + q], B[q] * B[q]));
                float apq = fmaf(B[6 + p], B[6 + q], fmaf(B[3 + p], B[3 + q], B[p] * B[q]));
                // skip when already orthogonal to working precision
                bool small = apq * apq <= 1e-18f * (app * aqq);
                jacobi_angle_fast(app, aqq, small ? 0.0f : apq, c, sn, t);   // MUFU angles on the device, exact on the host
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    float bp = B[3 * k + p], bq = B[3 * k + q];
                    B[3 * k + p] = fmaf(c, bp, -sn * bq); B[3 * k + q] = fmaf(sn, bp, c * bq);
                    float vp = V[3 * k + p], vq = V[3 * k + q];
                    V[3 * k + p] = fmaf(c, vp, -sn * vq); V[3 * k + q] = fmaf(sn, vp, c * vq);
                }
            }
    }
    // B = a V ; columns are (numerically) orthogonal with norms = singular values
    float n0 = fmaf(B[6], B[6], fmaf(B[3], B[3], B[0] * B[0]));
    float n1 = fmaf(B[7], B[7], fmaf(B[4], B[4], B[1] * B[1]));
    float n2 = fmaf(B[8], B[8], fmaf(B[5], B[5], B[2] * B[2]));
    // sort columns by norm, descending.  (col_i, col_j) <- (col_j, -col_i)
    // is a rotation, so det V stays +1.
#define SFM_SWAPNEG(i, j, ni, nj)                                   \
    if (ni < nj) {                                                  \
        float tn = ni; ni = nj; nj = tn;                            \
        _Pragma("unroll") for (int k = 0; k < 3; k++) {             \
            float bi = B[3 * k + i], vi = V[3 * k + i];             \
            B[3 * k + i] = B[3 * k + j]; B[3 * k + j] = -bi;        \
            V[3 * k + i] = V[3 * k + j]; V[3 * k + j] = -vi;        \
        }                                                           \
    }
    SFM_SWAPNEG(0, 1, n0, n1)
    SFM_SWAPNEG(0, 2, n0, n2)
    SFM_SWAPNEG(1, 2, n1, n2)
#undef SFM_SWAPNEG
    // Givens QR of B: Qt accumulates the row rotations, u = Qt^T.
    float Qt[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
#define SFM_ROWROT(p, q, col)                                                           \
    {                                                                                   \
        givens(B[3 * p + col], B[3 * q + col], c, sn);                                  \
        _Pragma("unroll") for (int k = 0; k < 3; k++) {                                 \
            float bp = B[3 * p + k], bq = B[3 * q + k];                                 \
            B[3 * p + k] = fmaf(c, bp, sn * bq); B[3 * q + k] = fmaf(-sn, bp, c * bq);  \
            float qp = Qt[3 * p + k], qq = Qt[3 * q + k];                               \
            Qt[3 * p + k] = fmaf(c, qp, sn * qq); Qt[3 * q + k] = fmaf(-sn, qp, c * qq);\
        }                                                                               \
    }
    SFM_ROWROT(0, 1, 0)
    SFM_ROWROT(0, 2, 0)
    SFM_ROWROT(1, 2, 1)
#undef SFM_ROWROT
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            u[3 * i + j] = Qt[3 * j + i];
            s[3 * i + j] = B[3 * i + j];
            v[3 * i + j] = V[3 * i + j];
        }
}

// Orientation of the null direction as the REFERENCE's svd() returns it (SfM/svd.h:311-335).  Under the contract
// a = u s v^T with u, v proper rotations an SVD keeps one discrete freedom: negating (u1, v1, u3, v3) - equivalently
// the sign of v3 - changes nothing in a, u v^T or det, but it swaps W <-> W^T and the sign of the translation in
// candidate_kernels (kernels.h:357-385), i.e. it permutes the four pose candidates 0 <-> 3, 1 <-> 2, and with them
// the index choosePose reports (sfm.cu:284-297).  Which sign the reference gets is decided by the rotation path of
// its SVD algorithm (McAdams, Selle, Tamstorf, Teran, Sifakis: "Computing the singular value decomposition of 3x3
// matrices with minimal branching and elementary floating point operations", 2011): cyclic Jacobi on a^T a from the
// identity over the pairs (0,1), (1,2), (2,0), 4 sweeps, each rotation taken from the half-angle estimate
// (ch, sh) ~ (2 (s_pp - s_qq), s_pq), replaced by the fixed angle pi/8 when gamma sh^2 >= ch^2 (svd.h:120-215), then the
// columns ordered by decreasing |a v_i| with negating swaps (svd.h:217-241).  This function replays that path on plain
// 3x3 matrices for the one bit it decides and returns the third column of the resulting V; the decomposition itself
// comes from svd3() (the replayed path is a 4-sweep approximation and is not used for values).  Checked against the
// reference's own host svd() on 4,000 essential matrices: same orientation every time (tests/test_cpu_oracle.py).
SFM_HD void reference_null_direction(const float* a, float* v3) {
    const float gamma = 5.828427124746190f, cstar = 0.923879532511287f, sstar = 0.382683432365090f;
    float S[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    mul33_AtB(a, a, S);
#pragma unroll 1
    for (int sweep = 0; sweep < 4; sweep++) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int p = k, q = (k + 1) % 3, r = (k + 2) % 3;
            float ch = 2.0f * (S[4 * p] - S[4 * q]), sh = S[3 * p + q];
            const bool inner = gamma * sh * sh < ch * ch;
            const float w = 1.0f / sqrtf(ch * ch + sh * sh);
            ch = inner ? w * ch : cstar;
            sh = inner ? w * sh : sstar;
            const float c = ch * ch - sh * sh, sn = 2.0f * sh * ch;       // rotation [c -sn; sn c] in the (p, q) plane
            const float spp = S[4 * p], sqq = S[4 * q], spq = S[3 * p + q], spr = S[3 * p + r], sqr = S[3 * q + r];
            const float npp = c * (c * spp + sn * spq) + sn * (c * spq + sn * sqq);
            const float npq = c * (-sn * spp + c * spq) + sn * (-sn * spq + c * sqq);
            const float nqq = -sn * (-sn * spp + c * spq) + c * (-sn * spq + c * sqq);
            const float npr = c * spr + sn * sqr, nqr = -sn * spr + c * sqr;
            S[4 * p] = npp; S[4 * q] = nqq;
            S[3 * p + q] = npq; S[3 * q + p] = npq;
            S[3 * p + r] = npr; S[3 * r + p] = npr;
            S[3 * q + r] = nqr; S[3 * r + q] = nqr;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const float vp = V[3 * i + p], vq = V[3 * i + q];
                V[3 * i + p] = c * vp + sn * vq;
                V[3 * i + q] = -sn * vp + c * vq;
            }
        }
    }
    float B[9], rho[3];
    mul33(a, V, B);
#pragma unroll
    for (int j = 0; j < 3; j++) rho[j] = B[j] * B[j] + B[3 + j] * B[3 + j] + B[6 + j] * B[6 + j];
#define SFM_REFSORT(i, j)                                                              \
    if (rho[i] < rho[j]) {                                                             \
        const float t = rho[i]; rho[i] = rho[j]; rho[j] = t;                           \
        _Pragma("unroll") for (int row = 0; row < 3; row++) {                          \
            const float vi = V[3 * row + i];                                           \
            V[3 * row + i] = V[3 * row + j]; V[3 * row + j] = -vi;                     \
        }                                                                              \
    }
    SFM_REFSORT(0, 1)
    SFM_REFSORT(0, 2)
    SFM_REFSORT(1, 2)
#undef SFM_REFSORT
    v3[0] = V[2]; v3[1] = V[5]; v3[2] = V[8];
}
// svd3() with the discrete freedom fixed the way the reference's svd() fixes it (see reference_null_direction).
// (the two halves are independent until the last step: small.cu runs them on two warps)
SFM_HD void svd3_orient(const float* r, float* u, float* s, float* v) {
    if (r[0] * v[2] + r[1] * v[5] + r[2] * v[8] < 0.0f) {
#pragma unroll
        for (int row = 0; row < 3; row++) {
            u[3 * row] = -u[3 * row]; u[3 * row + 2] = -u[3 * row + 2];
            v[3 * row] = -v[3 * row]; v[3 * row + 2] = -v[3 * row + 2];
        }
        // s = u^T a v: negating columns 0 and 2 of both sides leaves the diagonal and flips s01, s10, s12, s21 (all ~0)
        s[1] = -s[1]; s[3] = -s[3]; s[5] = -s[5]; s[7] = -s[7];
    }
}
SFM_HD void svd3_reference_orientation(const float* a, float* u, float* s, float* v) {
    svd3<5>(a, u, s, v);
    float r[3];
    reference_null_direction(a, r);
    svd3_orient(r, u, s, v);
}

#ifndef SFM_PROJECT_SWEEPS
#define SFM_PROJECT_SWEEPS 4    // one-sided Jacobi sweeps of the rank-2 projection's 3x3 SVD (3 already reach fp32 accuracy: tools record in DESIGN.md)
#endif
// Closest essential matrix in the reference's sense (SfM/kernels.h:281-295):
// E <- U diag(1,1,0) V^T.  The reference leaves the QR residue of S
// off-diagonals in (<= 1e-6 typical); we use the exact diag.
SFM_HD void project_essential(float* E) {
    float n = 0.0f;
#pragma unroll
    for (int i = 0; i < 9; i++) n = fmaf(E[i], E[i], n);
    float inv = n > 0.0f ? sfm_rsqrt(n) : 0.0f;      // conditioning only: the result U diag(1,1,0) V^T does not depend on the scale
    float En[9];
#pragma unroll
    for (int i = 0; i < 9; i++) En[i] = E[i] * inv;
    float u[9], s[9], v[9];
    svd3<SFM_PROJECT_SWEEPS>(En, u, s, v);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            E[3 * i + j] = fmaf(u[3 * i + 1], v[3 * j + 1], u[3 * i] * v[3 * j]);
}

// Null vector (right singular vector of the smallest singular value) of a
// row-major 4x4 A: cyclic Jacobi on A^T A with accumulated V, then one
// refinement step that uses A itself (restores eps*cond instead of
// eps*cond^2).  Replaces cusolverDnSgesvdjBatched 4x4 at
// SfM/kernels.h:175-194 as used by sfm.cu:278,329.
template <int SWEEPS = 5>
SFM_HD void null4(const float* A, float* x) {
    float g[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = i; j < 4; j++) {
            float acc = A[i] * A[j];
#pragma unroll
            for (int r = 1; r < 4; r++) acc = fmaf(A[4 * r + i], A[4 * r + j], acc);
            g[i][j] = acc;
        }
    float V[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) V[i][j] = (i == j) ? 1.0f : 0.0f;
#pragma unroll 1
    for (int sw = 0; sw < SWEEPS; sw++) {
#pragma unroll
        for (int p = 0; p < 3; p++)
#pragma unroll
            for (int q = p + 1; q < 4; q++) {
                float c, s, t;
                jacobi_angle(g[p][p], g[q][q], g[p][q], c, s, t);
                g[p][p] = fmaf(-t, g[p][q], g[p][p]);
                g[q][q] = fmaf(t, g[p][q], g[q][q]);
                g[p][q] = 0.0f;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (k == p || k == q) continue;
                    // upper-triangular storage: element (min,max)
                    float& akp = (k < p) ? g[k][p] : g[p][k];
                    float& akq = (k < q) ? g[k][q] : g[q][k];
                    float a = akp, b = akq;
                    akp = fmaf(c, a, -s * b);
                    akq = fmaf(s, a, c * b);
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    float a = V[k][p], b = V[k][q];
                    V[k][p] = fmaf(c, a, -s * b);
                    V[k][q] = fmaf(s, a, c * b);
                }
            }
    }
    // smallest eigenvalue (first minimum)
    float lam[4] = {g[0][0], g[1][1], g[2][2], g[3][3]};
    int m = 0;
    float lm = lam[0];
#pragma unroll
    for (int i = 1; i < 4; i++)
        if (lam[i] < lm) { lm = lam[i]; m = i; }
    float v0[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        float val = V[k][0];
#pragma unroll
        for (int i = 1; i < 4; i++) val = (m == i) ? V[k][i] : val;
        v0[k] = val;
    }
    // refinement: v <- v - sum_{i != m} V_i (V_i . A^T A v) / lam_i
    float r[4], gg[4];
#pragma unroll
    for (int i = 0; i < 4; i++)
        r[i] = fmaf(A[4 * i + 3], v0[3], fmaf(A[4 * i + 2], v0[2], fmaf(A[4 * i + 1], v0[1], A[4 * i] * v0[0])));
#pragma unroll
    for (int j = 0; j < 4; j++)
        gg[j] = fmaf(A[12 + j], r[3], fmaf(A[8 + j], r[2], fmaf(A[4 + j], r[1], A[j] * r[0])));
    float lmax = fmaxf(fmaxf(lam[0], lam[1]), fmaxf(lam[2], lam[3]));
    float v1[4] = {v0[0], v0[1], v0[2], v0[3]};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float d = fmaf(V[3][i], gg[3], fmaf(V[2][i], gg[2], fmaf(V[1][i], gg[1], V[0][i] * gg[0])));
        bool use = (i != m) && (lam[i] > 1e-12f * lmax);
        float coef = use ? d / lam[i] : 0.0f;
#pragma unroll
        for (int k = 0; k < 4; k++) v1[k] = fmaf(-coef, V[k][i], v1[k]);
    }
    float n2 = fmaf(v1[3], v1[3], fmaf(v1[2], v1[2], fmaf(v1[1], v1[1], v1[0] * v1[0])));
    float inv = 1.0f / sqrtf(n2);
#pragma unroll
    for (int k = 0; k < 4; k++) x[k] = v1[k] * inv;
}

// Fast path for the same null vector: inverse iteration on G = A^T A through one
// 4x4 Cholesky factorisation (G + eps*tr(G) I = L L^T) and ITERS solves.  The
// iteration contracts by lambda_4 / lambda_3 per step (<= 0.08 for two-view DLT
// matrices, ~1e-6 for inliers), so a handful of steps reach fp32 precision at
// ~1/7 of the Jacobi solve's instructions.  Returns false when the last step
// still moved the vector by more than 1e-5 (or anything is non-finite): the
// caller then falls back to null4().
// ADAPTIVE (device only): a thread freezes its vector at the first step that moved it by less than
// 1e-5 - the same acceptance test - and the warp leaves the loop once every lane has frozen, so the
// result of a point does not depend on its neighbours; inliers need 2 of the ITERS solves.
// With ADAPTIVE every lane of the warp must make the call; lanes without a point pass live = false.
template <int ITERS = 4, bool ADAPTIVE = false>
SFM_HD bool null4_inverse_iteration(const float* A, float* x, bool live = true) {
    float g[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = i; j < 4; j++) {
            float acc = A[i] * A[j];
#pragma unroll
            for (int r = 1; r < 4; r++) acc = fmaf(A[4 * r + i], A[4 * r + j], acc);
            g[i][j] = acc;
        }
    const float eps = 1e-7f * (g[0][0] + g[1][1] + g[2][2] + g[3][3]);
    // Cholesky, inverse diagonal kept
    float d0 = g[0][0] + eps;
    float i0 = sfm_rsqrt(d0);
    float l10 = g[0][1] * i0, l20 = g[0][2] * i0, l30 = g[0][3] * i0;
    float d1 = fmaf(-l10, l10, g[1][1] + eps);
    d1 = fmaxf(d1, eps);
    float i1 = sfm_rsqrt(d1);
    float l21 = fmaf(-l20, l10, g[1][2]) * i1, l31 = fmaf(-l30, l10, g[1][3]) * i1;
    float d2 = fmaf(-l21, l21, fmaf(-l20, l20, g[2][2] + eps));
    d2 = fmaxf(d2, eps);
    float i2 = sfm_rsqrt(d2);
    float l32 = fmaf(-l31, l21, fmaf(-l30, l20, g[2][3])) * i2;
    float d3 = fmaf(-l32, l32, fmaf(-l31, l31, fmaf(-l30, l30, g[3][3] + eps)));
    d3 = fmaxf(d3, eps);
    float i3 = sfm_rsqrt(d3);
    // Start from the exact null vector of the first three rows (A has rows
    // (-1,0,x1,0), (0,-1,y1,0) in DLT use, but any A works): generalised cross
    // product of rows 0..2.  For an inlier this is already the answer to within
    // the noise, so the iteration only has to remove an O(1e-2) component.
    float v0, v1, v2, v3;
    {
        const float *r0 = A, *r1 = A + 4, *r2 = A + 8;
        float m01 = r0[0] * r1[1] - r0[1] * r1[0], m02 = r0[0] * r1[2] - r0[2] * r1[0], m03 = r0[0] * r1[3] - r0[3] * r1[0];
        float m12 = r0[1] * r1[2] - r0[2] * r1[1], m13 = r0[1] * r1[3] - r0[3] * r1[1], m23 = r0[2] * r1[3] - r0[3] * r1[2];
        v0 = -(r2[1] * m23 - r2[2] * m13 + r2[3] * m12);
        v1 = (r2[0] * m23 - r2[2] * m03 + r2[3] * m02);
        v2 = -(r2[0] * m13 - r2[1] * m03 + r2[3] * m01);
        v3 = (r2[0] * m12 - r2[1] * m02 + r2[2] * m01);
        float n2 = fmaf(v3, v3, fmaf(v2, v2, fmaf(v1, v1, v0 * v0)));
        bool ok = n2 > 1e-30f;
        float n = ok ? sfm_rsqrt(n2) : 0.0f;
        v0 = ok ? v0 * n : 0.5f; v1 = ok ? v1 * n : 0.5f; v2 = ok ? v2 * n : 0.5f; v3 = ok ? v3 * n : 0.5f;
    }
    float diff2 = 1.0f;
    bool frozen = !live;
    (void)frozen;
#pragma unroll
    for (int it = 0; it < ITERS; it++) {
        float y0 = v0 * i0;
        float y1 = fmaf(-l10, y0, v1) * i1;
        float y2 = fmaf(-l21, y1, fmaf(-l20, y0, v2)) * i2;
        float y3 = fmaf(-l32, y2, fmaf(-l31, y1, fmaf(-l30, y0, v3))) * i3;
        float z3 = y3 * i3;
        float z2 = fmaf(-l32, z3, y2) * i2;
        float z1 = fmaf(-l31, z3, fmaf(-l21, z2, y1)) * i1;
        float z0 = fmaf(-l30, z3, fmaf(-l20, z2, fmaf(-l10, z1, y0))) * i0;
        float n = sfm_rsqrt(fmaf(z3, z3, fmaf(z2, z2, fmaf(z1, z1, z0 * z0))));
        z0 *= n; z1 *= n; z2 *= n; z3 *= n;
        float e0 = z0 - v0, e1 = z1 - v1, e2 = z2 - v2, e3 = z3 - v3;
#ifdef __CUDA_ARCH__
        if (ADAPTIVE) {
            if (!frozen) {
                diff2 = fmaf(e3, e3, fmaf(e2, e2, fmaf(e1, e1, e0 * e0)));
                v0 = z0; v1 = z1; v2 = z2; v3 = z3;
                frozen = diff2 < 1e-10f;
            }
            if (__all_sync(0xFFFFFFFFu, frozen)) break;
            continue;
        }
#endif
        diff2 = fmaf(e3, e3, fmaf(e2, e2, fmaf(e1, e1, e0 * e0)));
        v0 = z0; v1 = z1; v2 = z2; v3 = z3;
    }
    x[0] = v0; x[1] = v1; x[2] = v2; x[3] = v3;
    return diff2 < 1e-10f;     // false also for NaN
}

// Null vector of the two-view DLT matrix (compute_linear_triangulation_A, SfM/kernels.h:387-431, with camera 1 = I4)
//     A = [ -1  0 x1 0 ;  0 -1 y1 0 ;  a ;  b ],   a = x2 M[2,:] - M[0,:],  b = y2 M[2,:] - M[1,:]
// i.e. the right singular vector of the smallest singular value, which is what the reference takes from
// cusolverDnSgesvdjBatched (sfm.cu:329-333).  Power iteration on adj(A) adj(A)^T = det(A)^2 (A^T A)^-1: the
// columns k_j of adj(A) are generalised cross products of three rows of A, and with two rows this sparse they
// cost a handful of instructions - no Gram matrix, no factorisation, no division:
//     k3 = r0 x r1 x a = (-a3 x1, -a3 y1, -a3, a0 x1 + a1 y1 + a2)        (k2: the same with b)
//     k0 = r1 x a x b,  k1 = r0 x a x b   from the six 2x2 minors m_ij = a_i b_j - a_j b_i
// One step v <- sum_j k_j (k_j . v) contracts the error by (sigma_4 / sigma_3)^2, exactly like inverse iteration
// on A^T A; the start vector k3 is the exact null vector of the first three rows, which for an inlier is the
// answer to within the noise.  MIN_ITERS steps are always taken; after that the vector is frozen at the first
// step that moved it (normalised) by less than 1e-5, up to MAX_ITERS steps; returns false if it never did (the
// caller falls back to the Jacobi solve null4()).  A point's result does not depend on its warp neighbours.
// ~40 instructions of set-up + ~41 per step (the inverse iteration it replaces: ~110 + 45 per step).
// PTS points are solved side by side (independent instruction streams for the scheduler to interleave); ok[p] / v[p].
// WARP_VOTE (device): the warp leaves the loop together once every lane has frozen - every lane of the warp must
// then make the call (live = false for idle points); without it each lane leaves on its own.
template <int PTS = 1, bool WARP_VOTE = true, int MIN_ITERS = 3, int MAX_ITERS = 9>
SFM_HD void dlt_null_adjugate(const float* x1, const float* y1, const float (*a)[4], const float (*b)[4], float (*v)[4],
                              const bool* live, bool* ok) {
    float k[PTS][4][4];
    float u[PTS][4];
    bool frozen[PTS];
    float diff2[PTS];
#pragma unroll
    for (int p = 0; p < PTS; p++) {
        const float sa = fmaf(a[p][0], x1[p], fmaf(a[p][1], y1[p], a[p][2])), sb = fmaf(b[p][0], x1[p], fmaf(b[p][1], y1[p], b[p][2]));
        k[p][3][0] = -a[p][3] * x1[p]; k[p][3][1] = -a[p][3] * y1[p]; k[p][3][2] = -a[p][3]; k[p][3][3] = sa;
        k[p][2][0] = -b[p][3] * x1[p]; k[p][2][1] = -b[p][3] * y1[p]; k[p][2][2] = -b[p][3]; k[p][2][3] = sb;
        const float m01 = fmaf(a[p][0], b[p][1], -a[p][1] * b[p][0]), m02 = fmaf(a[p][0], b[p][2], -a[p][2] * b[p][0]);
        const float m03 = fmaf(a[p][0], b[p][3], -a[p][3] * b[p][0]), m12 = fmaf(a[p][1], b[p][2], -a[p][2] * b[p][1]);
        const float m13 = fmaf(a[p][1], b[p][3], -a[p][3] * b[p][1]), m23 = fmaf(a[p][2], b[p][3], -a[p][3] * b[p][2]);
        k[p][0][0] = -fmaf(y1[p], m13, m23); k[p][0][1] = y1[p] * m03; k[p][0][2] = m03; k[p][0][3] = -fmaf(y1[p], m01, m02);
        k[p][1][0] = -x1[p] * m13; k[p][1][1] = fmaf(x1[p], m03, m23); k[p][1][2] = -m13; k[p][1][3] = fmaf(-x1[p], m01, m12);
        const float n2 = fmaf(k[p][3][3], k[p][3][3], fmaf(k[p][3][2], k[p][3][2], fmaf(k[p][3][1], k[p][3][1], k[p][3][0] * k[p][3][0])));
        const bool nz = n2 > 1e-30f;
        const float n = nz ? sfm_rsqrt(n2) : 0.0f;
#pragma unroll
        for (int c = 0; c < 4; c++) u[p][c] = nz ? k[p][3][c] * n : 0.5f;
        frozen[p] = !live[p];
        diff2[p] = 1.0f;
    }
#pragma unroll
    for (int it = 0; it < MAX_ITERS; it++) {
        bool all_frozen = true;
#pragma unroll
        for (int p = 0; p < PTS; p++) {
            float w[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float d = fmaf(k[p][j][3], u[p][3], fmaf(k[p][j][2], u[p][2], fmaf(k[p][j][1], u[p][1], k[p][j][0] * u[p][0])));
#pragma unroll
                for (int c = 0; c < 4; c++) w[c] = fmaf(k[p][j][c], d, w[c]);
            }
            const float n = sfm_rsqrt(fmaf(w[3], w[3], fmaf(w[2], w[2], fmaf(w[1], w[1], w[0] * w[0]))));
#pragma unroll
            for (int c = 0; c < 4; c++) w[c] *= n;        // w . u = sum_j d_j^2 >= 0: no sign flip between steps
            if (it < MIN_ITERS - 1) {                      // unconditional steps: no convergence bookkeeping
#pragma unroll
                for (int c = 0; c < 4; c++) u[p][c] = w[c];
                all_frozen = false;
            } else {
                if (!frozen[p]) {
                    const float e0 = w[0] - u[p][0], e1 = w[1] - u[p][1], e2 = w[2] - u[p][2], e3 = w[3] - u[p][3];
                    diff2[p] = fmaf(e3, e3, fmaf(e2, e2, fmaf(e1, e1, e0 * e0)));
#pragma unroll
                    for (int c = 0; c < 4; c++) u[p][c] = w[c];
                    frozen[p] = diff2[p] < 1e-10f;
                }
                all_frozen = all_frozen && frozen[p];
            }
        }
        if (it >= MIN_ITERS - 1) {
#if defined(__CUDA_ARCH__)
            if (WARP_VOTE ? __all_sync(0xFFFFFFFFu, all_frozen) : all_frozen) break;
#else
            if (all_frozen) break;
#endif
        }
    }
#pragma unroll
    for (int p = 0; p < PTS; p++) {
#pragma unroll
        for (int c = 0; c < 4; c++) v[p][c] = u[p][c];
        ok[p] = diff2[p] < 1e-10f;      // false also for NaN
    }
}
// The same null vector for the triangulation kernel, where instruction count is everything (1M points, 32 B of HBM traffic
// each: ~5 us of memory time): no convergence bookkeeping, no normalisation inside, no warp vote.  With the adjugate
// columns k_j as above, S = sum_j k_j k_j^T / trace is formed once (10 entries, largest eigenvalue in [1/4, 1]) and
//     v = S (S (S (S (k3 / |k|))))        = FOUR power-iteration steps from the start vector k3
// costs 40 + 64 FMAs; only the direction of v matters (the caller de-homogenises), so nothing is normalised on the way.
// Error after four steps: (sigma_4 / sigma_3)^8 of the start error - inliers are exact to fp32 rounding after the first,
// 99.99 % of gross outliers are within 1e-3 (oracle comparison in tests/test_cpu_hostlogic.py); a point whose two smallest
// singular values nearly coincide has no well-defined null vector in the reference's SVD either.  Degenerate input
// (all cofactors zero or non-finite) gives v = 0, which de-homogenises to the reference's (0,0,0,1) (kernels.h:439).
// Written once over a value type: float (one point), or float2 on the device = TWO points per thread in packed
// FFMA2 / FMUL2 / FADD2 (sm_100): the kernel is bound by issue slots, not by the FMA pipe, and a packed instruction
// does two points' worth of work in one slot.  Lane-wise the packed form is the scalar fma tree, bit for bit.
struct LaneF1 {
    typedef float T;
    static SFM_HD T fma(T a, T b, T c) { return fmaf(a, b, c); }
    static SFM_HD T mul(T a, T b) { return a * b; }
    static SFM_HD T add(T a, T b) { return a + b; }
    static SFM_HD T neg(T a) { return -a; }
    static SFM_HD T rsqrt_scale(T tr) { return (tr > 0.0f && tr < 3.0e38f) ? sfm_rsqrt(tr) : 0.0f; }
    typedef bool M;                                                          // per-lane flags
    static SFM_HD T splat(float c) { return c; }
    static SFM_HD M none() { return false; }
    static SFM_HD M below(T num, T lim) { return num <= lim; }              // also true for 0 <= 0 (degenerate: nothing to refine)
    static SFM_HD M join(M a, M b) { return a || b; }
    static SFM_HD bool all(M m) { return m; }
    static SFM_HD T keep(M m, T old_v, T new_v) { return m ? old_v : new_v; }
};
#if defined(__CUDACC__)
struct LaneF2 {
    typedef float2 T;
    static __device__ __forceinline__ T fma(T a, T b, T c) { return __ffma2_rn(a, b, c); }
#ifdef SFM_LANEF2_MUL_AS_FMA      // experiment: FMUL2 / FADD2 through FFMA2 (a * b + 0, a * 1 + b)
    static __device__ __forceinline__ T mul(T a, T b) { return __ffma2_rn(a, b, make_float2(0.0f, 0.0f)); }
    static __device__ __forceinline__ T add(T a, T b) { return __ffma2_rn(a, make_float2(1.0f, 1.0f), b); }
#else
    static __device__ __forceinline__ T mul(T a, T b) { return __fmul2_rn(a, b); }
    static __device__ __forceinline__ T add(T a, T b) { return __fadd2_rn(a, b); }
#endif
    static __device__ __forceinline__ T neg(T a) { return make_float2(-a.x, -a.y); }
    static __device__ __forceinline__ T rsqrt_scale(T tr) { return make_float2(LaneF1::rsqrt_scale(tr.x), LaneF1::rsqrt_scale(tr.y)); }
    struct M { bool x, y; };
    static __device__ __forceinline__ T splat(float c) { return make_float2(c, c); }
    static __device__ __forceinline__ M none() { return M{false, false}; }
    static __device__ __forceinline__ M below(T num, T lim) { return M{num.x <= lim.x, num.y <= lim.y}; }
    static __device__ __forceinline__ M join(M a, M b) { return M{a.x || b.x, a.y || b.y}; }
    static __device__ __forceinline__ bool all(M m) { return m.x && m.y; }
    static __device__ __forceinline__ T keep(M m, T old_v, T new_v) { return make_float2(m.x ? old_v.x : new_v.x, m.y ? old_v.y : new_v.y); }
};
#endif
// Convergence: sin^2 of the angle between the last two iterates, from (w.w)(u.u) - (w.u)^2 - resolvable in fp32 down to
// ~2e-7.  Every point of a sane two-view geometry passes after the four steps.  Points with sigma_4 / sigma_3 close to 1
// (gross outliers, a grossly wrong pose) do not: for them S is squared and re-normalised, S <- S S / trace, which doubles
// the number of power steps per product, and two more products are taken - up to DLT_MAX_SQUARINGS rounds, i.e.
// 2^6 = 64 steps per product in the last one; whatever has not converged by then has no well-defined null vector in the
// reference's SVD either.  ~105 instructions per round, in registers, no call; a converged lane that shares a packed
// pair with an unconverged one just gets more accurate.
#ifndef DLT_POWER_STEPS
#define DLT_POWER_STEPS 3      // fixed products with S before the first convergence check
#endif
#ifndef DLT_MAX_SQUARINGS
#define DLT_MAX_SQUARINGS 6
#endif
constexpr float DLT_SIN2_CONVERGED = 5e-7f;      // angle between the last two iterates below 7e-4
// The refinement rounds, out of line: the common path (every point converged after the fixed steps) keeps nothing alive
// for them - no spills at 64 registers in triangulate_kernel, 5 % off its time per point at large n - and only a thread
// whose pair needs them pays the call and the copies to the stack.
#if defined(__CUDACC__)
#define SFM_HD_NOINLINE __host__ __device__ __noinline__
#else
#define SFM_HD_NOINLINE
#endif
template <class L>
SFM_HD_NOINLINE void dlt_refine_rounds(typename L::T (&S)[4][4], typename L::T* u, typename L::T* w) {
    typedef typename L::T T;
    T r, it;
    {
        typename L::M done = L::none();
#pragma unroll 1
        for (int round = 0; round < DLT_MAX_SQUARINGS; round++) {
            const T ww = L::fma(w[3], w[3], L::fma(w[2], w[2], L::fma(w[1], w[1], L::mul(w[0], w[0]))));
            const T uu = L::fma(u[3], u[3], L::fma(u[2], u[2], L::fma(u[1], u[1], L::mul(u[0], u[0]))));
            const T wu = L::fma(w[3], u[3], L::fma(w[2], u[2], L::fma(w[1], u[1], L::mul(w[0], u[0]))));
            const T den = L::mul(ww, uu);
            // a lane that has converged is frozen for good: the result of a point never depends on the point it shares
            // a packed pair with
            done = L::join(done, L::below(L::fma(L::neg(wu), wu, den), L::mul(den, L::splat(DLT_SIN2_CONVERGED))));
            if (L::all(done)) break;
            T Q[4][4];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = i; j < 4; j++)
                    Q[i][j] = L::fma(S[i][3], S[3][j], L::fma(S[i][2], S[2][j], L::fma(S[i][1], S[1][j], L::mul(S[i][0], S[0][j]))));
            r = L::rsqrt_scale(L::add(L::add(Q[0][0], Q[1][1]), L::add(Q[2][2], Q[3][3])));
            it = L::mul(r, r);
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = i; j < 4; j++) {
                    S[i][j] = L::mul(Q[i][j], it);
                    S[j][i] = S[i][j];
                }
            // keep |u| near 1 (two products shrink it by up to 16x per round)
            const T ru = L::rsqrt_scale(uu);
            T un[4], wn[4], u2[4];
#pragma unroll
            for (int i = 0; i < 4; i++) un[i] = L::mul(u[i], ru);
#pragma unroll
            for (int i = 0; i < 4; i++) wn[i] = L::fma(S[i][3], un[3], L::fma(S[i][2], un[2], L::fma(S[i][1], un[1], L::mul(S[i][0], un[0]))));
#pragma unroll
            for (int i = 0; i < 4; i++) u2[i] = L::fma(S[i][3], wn[3], L::fma(S[i][2], wn[2], L::fma(S[i][1], wn[1], L::mul(S[i][0], wn[0]))));
#pragma unroll
            for (int i = 0; i < 4; i++) {
                w[i] = L::keep(done, w[i], wn[i]);
                u[i] = L::keep(done, u[i], u2[i]);
            }
        }
    }
}
template <class L>
SFM_HD void dlt_null_power4_lanes(typename L::T x1, typename L::T y1, const typename L::T* a, const typename L::T* b, typename L::T* v,
                                  bool refine = true) {
    typedef typename L::T T;
    T k[4][4];
    const T na3 = L::neg(a[3]), nb3 = L::neg(b[3]);
    k[3][0] = L::mul(na3, x1); k[3][1] = L::mul(na3, y1); k[3][2] = na3; k[3][3] = L::fma(a[0], x1, L::fma(a[1], y1, a[2]));
    k[2][0] = L::mul(nb3, x1); k[2][1] = L::mul(nb3, y1); k[2][2] = nb3; k[2][3] = L::fma(b[0], x1, L::fma(b[1], y1, b[2]));
    const T m01 = L::fma(a[0], b[1], L::neg(L::mul(a[1], b[0]))), m02 = L::fma(a[0], b[2], L::neg(L::mul(a[2], b[0])));
    const T m03 = L::fma(a[0], b[3], L::neg(L::mul(a[3], b[0]))), m12 = L::fma(a[1], b[2], L::neg(L::mul(a[2], b[1])));
    const T m13 = L::fma(a[1], b[3], L::neg(L::mul(a[3], b[1]))), m23 = L::fma(a[2], b[3], L::neg(L::mul(a[3], b[2])));
    k[0][0] = L::neg(L::fma(y1, m13, m23)); k[0][1] = L::mul(y1, m03); k[0][2] = m03; k[0][3] = L::neg(L::fma(y1, m01, m02));
    k[1][0] = L::neg(L::mul(x1, m13)); k[1][1] = L::fma(x1, m03, m23); k[1][2] = L::neg(m13); k[1][3] = L::fma(L::neg(x1), m01, m12);
    // S = sum_j k_j k_j^T, upper triangle, then / trace
    T S[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = i; j < 4; j++)
            S[i][j] = L::fma(k[3][i], k[3][j], L::fma(k[2][i], k[2][j], L::fma(k[1][i], k[1][j], L::mul(k[0][i], k[0][j]))));
    T r = L::rsqrt_scale(L::add(L::add(S[0][0], S[1][1]), L::add(S[2][2], S[3][3])));
    T it = L::mul(r, r);
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = i; j < 4; j++) {
            S[i][j] = L::mul(S[i][j], it);
            S[j][i] = S[i][j];
        }
    // DLT_POWER_STEPS products with S from the start vector k3 / |.|; (w, u) = the last two iterates
    T u[4], w[4];
#pragma unroll
    for (int i = 0; i < 4; i++) u[i] = L::mul(k[3][i], r);
#pragma unroll
    for (int step = 0; step < DLT_POWER_STEPS; step++) {
#pragma unroll
        for (int i = 0; i < 4; i++) w[i] = u[i];
#pragma unroll
        for (int i = 0; i < 4; i++) u[i] = L::fma(S[i][3], w[3], L::fma(S[i][2], w[2], L::fma(S[i][1], w[1], L::mul(S[i][0], w[0]))));
    }
    if (refine) {
        // the first convergence check inline; the rounds themselves out of line and only for a pair that needs them
        const T ww = L::fma(w[3], w[3], L::fma(w[2], w[2], L::fma(w[1], w[1], L::mul(w[0], w[0]))));
        const T uu = L::fma(u[3], u[3], L::fma(u[2], u[2], L::fma(u[1], u[1], L::mul(u[0], u[0]))));
        const T wu = L::fma(w[3], u[3], L::fma(w[2], u[2], L::fma(w[1], u[1], L::mul(w[0], u[0]))));
        const T den = L::mul(ww, uu);
        if (!L::all(L::below(L::fma(L::neg(wu), wu, den), L::mul(den, L::splat(DLT_SIN2_CONVERGED))))) {
            // copies whose address escapes: S, u, w themselves stay in registers on the common path
            T Sc[4][4], uc[4], wc[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uc[i] = u[i]; wc[i] = w[i];
#pragma unroll
                for (int j = 0; j < 4; j++) Sc[i][j] = S[i][j];
            }
            dlt_refine_rounds<L>(Sc, uc, wc);
#pragma unroll
            for (int i = 0; i < 4; i++) u[i] = uc[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] = u[i];
}
template <int PTS = 1>
SFM_HD void dlt_null_power4(const float* x1, const float* y1, const float (*a)[4], const float (*b)[4], float (*v)[4]) {
#pragma unroll
    for (int p = 0; p < PTS; p++) dlt_null_power4_lanes<LaneF1>(x1[p], y1[p], a[p], b[p], v[p]);
}

// one point; the matrix given row-major (rows 0 and 1 must be the camera-1 = I4 rows above)
SFM_HD bool dlt_null_adjugate1(const float* A, float* v) {
    const float x1[1] = {A[2]}, y1[1] = {A[6]};
    float a[1][4], b[1][4], out[1][4];
    const bool live[1] = {true};
    bool ok[1];
#pragma unroll
    for (int c = 0; c < 4; c++) { a[0][c] = A[8 + c]; b[0][c] = A[12 + c]; }
    dlt_null_adjugate<1, false>(x1, y1, a, b, out, live, ok);
#pragma unroll
    for (int c = 0; c < 4; c++) v[c] = out[0][c];
    return ok[0];
}

// Inverse of a rigid transform-shaped 4x4 done generally (Gauss-Jordan with
// partial pivoting), like the LU the reference calls
// (cublasSgetrfBatched/SgetriBatched, SfM/kernels.h:132-173).  Returns false
// when singular.
SFM_HD bool inv4(const float* m, float* out) {
    float a[4][8];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            a[i][j] = m[4 * i + j];
            a[i][4 + j] = (i == j) ? 1.0f : 0.0f;
        }
    bool ok = true;
#pragma unroll
    for (int col = 0; col < 4; col++) {
        int piv = col;
        float best = fabsf(a[col][col]);
#pragma unroll
        for (int r = 0; r < 4; r++)
            if (r > col && fabsf(a[r][col]) > best) { best = fabsf(a[r][col]); piv = r; }
        if (best == 0.0f) ok = false;
#pragma unroll
        for (int r = 0; r < 4; r++)
            if (r == piv && piv != col) {
#pragma unroll
                for (int j = 0; j < 8; j++) { float tmp = a[col][j]; a[col][j] = a[r][j]; a[r][j] = tmp; }
            }
        float ip = 1.0f / a[col][col];
#pragma unroll
        for (int j = 0; j < 8; j++) a[col][j] *= ip;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            if (r == col) continue;
            float f = a[r][col];
#pragma unroll
            for (int j = 0; j < 8; j++) a[r][j] = fmaf(-f, a[col][j], a[r][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) out[4 * i + j] = a[i][4 + j];
    return ok;
}

}  // namespace sfmb200
