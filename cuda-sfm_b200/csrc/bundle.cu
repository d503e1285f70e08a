// Two-view bundle adjustment with inlier re-selection.
//
// The reference stops at linear triangulation and lists bundle adjustment as
// future work (README.md:65-69); SURVEY.md 8f rank 4 names it as the step after
// the path.  This file refines what the path produced - the selected camera-2
// matrix M = P[P_ind] (x1 ~ X, x2 ~ M X, the convention of linear_triangulation,
// sfm.cu:309-336) and the triangulated points - by Levenberg-Marquardt on the
// reprojection error of the inliers, in normalised camera coordinates.
//
// Parameters: camera 2 (rotation by a left-multiplicative so(3) update, translation)
// and one 3-D point per active correspondence; camera 1 is fixed at [I|0]; the
// scale gauge is left to the damping and fixed afterwards to |t| = 1.
// The point blocks are eliminated (Schur complement), so one LM iteration is
//   ba_accumulate_kernel : per point the 2x3 / 2x6 Jacobians, V = Jp^T Jp (3x3),
//                          W = Jc^T Jp (6x3); per CTA the sums of U - W V^-1 W^T (21),
//                          diag U (6), gc - W V^-1 gp (6) and the cost; the last
//                          CTA adds the per-CTA partials in a fixed order (fp64) and
//                          solves the damped 6x6 system by Cholesky;
//   ba_update_kernel     : back-substitution dX = -V^-1 (gp + W^T dc) per point into
//                          the other point buffer, cost of the candidate; the last CTA
//                          accepts (strict decrease, every point still in front of both
//                          cameras) or rejects and rescales lambda.
// Nothing returns to the host between iterations; reductions have a fixed order.
// oracle/oracle.py: bundle_adjust restates the same iteration in fp64.
#include <cooperative_groups.h>

#include "internal.cuh"
#include "sampson.cuh"
#include "smallmat.cuh"

namespace cg = cooperative_groups;

namespace sfmb200 {

constexpr int BA_THREADS = 256;
constexpr int BA_NSUM = 34;        // 21 reduced-camera entries + 6 diag U + 6 gradient + cost

__device__ __forceinline__ int sym6(int i, int j) {     // packed upper triangle, i <= j
    return i * 6 - i * (i - 1) / 2 + (j - i);
}

// Per-point linearisation at (R, t, X).  Returns false when the point is not in front of both cameras.
// Vinv: inverse of the damped 3x3 point block (packed 00 01 02 11 12 22); W [6][3]; gp [3];
// FULL additionally gives the camera-side terms Jc [2][6] and r2 [2].
template <bool FULL>
__device__ __forceinline__ bool ba_linearise(const float* R, const float* t, const float4 p, const float* X, float lam,
                                             float* Vinv, float* W, float* gp, float* Jc, float* r2, float& cost) {
    const float Z = X[2];
    float Q[3], Y[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        Q[i] = fmaf(R[3 * i], X[0], fmaf(R[3 * i + 1], X[1], R[3 * i + 2] * X[2]));
        Y[i] = Q[i] + t[i];
    }
    if (!(Z > 0.0f) || !(Y[2] > 0.0f)) return false;
    const float iz = 1.0f / Z, u = X[0] * iz, v = X[1] * iz;
    const float iz2 = 1.0f / Y[2], u2 = Y[0] * iz2, v2 = Y[1] * iz2;
    const float r1[2] = {u - p.x, v - p.y};
    r2[0] = u2 - p.z;
    r2[1] = v2 - p.w;
    cost = fmaf(r1[0], r1[0], fmaf(r1[1], r1[1], fmaf(r2[0], r2[0], r2[1] * r2[1])));
    // d pi / d X for both views
    const float a[2][3] = {{iz, 0.0f, -u * iz}, {0.0f, iz, -v * iz}};
    const float bp[2][3] = {{iz2, 0.0f, -u2 * iz2}, {0.0f, iz2, -v2 * iz2}};
    float c[2][3];                          // Jp2 = bp * R
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
        for (int k = 0; k < 3; k++) c[r][k] = fmaf(bp[r][0], R[k], fmaf(bp[r][1], R[3 + k], bp[r][2] * R[6 + k]));
    // camera Jacobian: [bp * (-[Q]x) | bp]
    const float N[3][3] = {{0.0f, Q[2], -Q[1]}, {-Q[2], 0.0f, Q[0]}, {Q[1], -Q[0], 0.0f}};
#pragma unroll
    for (int r = 0; r < 2; r++) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            Jc[6 * r + k] = fmaf(bp[r][0], N[0][k], fmaf(bp[r][1], N[1][k], bp[r][2] * N[2][k]));
            Jc[6 * r + 3 + k] = bp[r][k];
        }
    }
    // point block and gradient
    float V[6];
    {
        int q = 0;
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = i; j < 3; j++)
                V[q++] = fmaf(a[0][i], a[0][j], fmaf(a[1][i], a[1][j], fmaf(c[0][i], c[0][j], c[1][i] * c[1][j])));
    }
#pragma unroll
    for (int k = 0; k < 3; k++) gp[k] = fmaf(a[0][k], r1[0], fmaf(a[1][k], r1[1], fmaf(c[0][k], r2[0], c[1][k] * r2[1])));
    const float d = 1.0f + lam;
    const float v00 = V[0] * d, v01 = V[1], v02 = V[2], v11 = V[3] * d, v12 = V[4], v22 = V[5] * d;
    const float c00 = fmaf(v11, v22, -v12 * v12), c01 = fmaf(v02, v12, -v01 * v22), c02 = fmaf(v01, v12, -v02 * v11);
    const float det = fmaf(v00, c00, fmaf(v01, c01, v02 * c02));
    if (!(fabsf(det) > 0.0f)) return false;
    const float id = 1.0f / det;
    Vinv[0] = c00 * id;
    Vinv[1] = c01 * id;
    Vinv[2] = c02 * id;
    Vinv[3] = fmaf(v00, v22, -v02 * v02) * id;
    Vinv[4] = fmaf(v01, v02, -v00 * v12) * id;
    Vinv[5] = fmaf(v00, v11, -v01 * v01) * id;
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int k = 0; k < 3; k++) W[3 * i + k] = fmaf(Jc[i], c[0][k], Jc[6 + i] * c[1][k]);
    (void)FULL;
    return true;
}

__device__ __forceinline__ void sym3_mul(const float* S, const float* x, float* y) {
    y[0] = fmaf(S[0], x[0], fmaf(S[1], x[1], S[2] * x[2]));
    y[1] = fmaf(S[1], x[0], fmaf(S[3], x[1], S[4] * x[2]));
    y[2] = fmaf(S[2], x[0], fmaf(S[4], x[1], S[5] * x[2]));
}

// R <- exp([w]x) R (Rodrigues), t <- t + dt
__device__ void ba_apply_camera(const float* cam, const double* dc, float* out) {
    double w[3] = {dc[0], dc[1], dc[2]};
    double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2);
    double A = th > 1e-12 ? sin(th) / th : 1.0, Bc = th > 1e-12 ? (1.0 - cos(th)) / th2 : 0.0;
    double Kx[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0}, K2[9], Ex[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double acc = 0;
            for (int k = 0; k < 3; k++) acc += Kx[3 * i + k] * Kx[3 * k + j];
            K2[3 * i + j] = acc;
        }
    for (int i = 0; i < 9; i++) Ex[i] = (i % 4 == 0 ? 1.0 : 0.0) + A * Kx[i] + Bc * K2[i];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double acc = 0;
            for (int k = 0; k < 3; k++) acc += Ex[3 * i + k] * (double)cam[3 * k + j];
            out[3 * i + j] = (float)acc;
        }
    for (int i = 0; i < 3; i++) out[9 + i] = (float)((double)cam[9 + i] + dc[3 + i]);
}

// ---- state layout (BAState, internal.cuh) ----
// ctl_i [B][8]: 0 cur buffer, 1 ticket A, 2 ticket B, 3 active points, 4 accepted steps, 5 solve ok, 6 inliers of the adjusted E, 7 committed
// ctl_f [B][8]: 0 lambda, 1 cost (current), 2 cost at entry of this round, 3 spare

// Start of an outer round: camera from P[P_ind], active set = inliers of the current E whose
// triangulated point is finite and in front of both cameras, points copied into buffer 0.
__global__ void __launch_bounds__(BA_THREADS) ba_init_kernel(DeviceState s, BAState ba, float thr, float lambda0, int first_round) {
    const int b = blockIdx.y;
    __shared__ float sM[12];
    __shared__ float sE[9];
    __shared__ int s_cnt;
    if (threadIdx.x < 12) sM[threadIdx.x] = s.P[(size_t)b * 64 + 16 * s.P_ind[b] + threadIdx.x];
    if (threadIdx.x < 9) sE[threadIdx.x] = s.E[(size_t)b * 9 + threadIdx.x];
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    if (blockIdx.x == 0) {
        float* cam = ba.cam + (size_t)b * 24;
        if (threadIdx.x < 9) cam[threadIdx.x] = sM[(threadIdx.x / 3) * 4 + threadIdx.x % 3];
        if (threadIdx.x < 3) cam[9 + threadIdx.x] = sM[4 * threadIdx.x + 3];
        if (threadIdx.x == 0) {
            ba.ctl_i[b * 8 + 0] = 0;
            ba.ctl_i[b * 8 + 4] = 0;
            ba.ctl_f[b * 8 + 0] = lambda0;
            if (first_round) ba.base_count[b] = s.best_count[b];      // the consensus this call must not lose
        }
    }
    const int i = blockIdx.x * BA_THREADS + threadIdx.x;
    bool act = false;
    if (i < s.n) {
        const float4 p = s.corr[(size_t)b * s.n_stride + i];
        const float* in = s.points + (size_t)b * 4 * s.n_stride;
        const float X = in[i], Y = in[(size_t)s.n_stride + i], Z = in[(size_t)2 * s.n_stride + i];
        const float z2 = fmaf(sM[8], X, fmaf(sM[9], Y, fmaf(sM[10], Z, sM[11])));
        act = epipolar_d(s.metric, sE, p.x, p.y, p.z, p.w, -thr) < 0.0f && isfinite(X) && isfinite(Y) && isfinite(Z) && Z > 0.0f && z2 > 0.0f;
        float* out = ba.pts + (size_t)b * 6 * s.n_stride;
        out[i] = X;
        out[(size_t)s.n_stride + i] = Y;
        out[(size_t)2 * s.n_stride + i] = Z;
        ba.active[(size_t)b * s.n_stride + i] = act ? 1 : 0;
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, act);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s_cnt, __popc(m));
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) atomicAdd(&ba.ctl_i[b * 8 + 3], s_cnt);
}

// ---- one LM iteration as four CTA-level phases.  The two-kernel path runs them with "last CTA done" tickets,
// the persistent path (one cooperative launch for all iterations of a round) with grid barriers; both use the
// same code and the same reduction orders, so they give the same bits. ----
struct BAWork {                      // shared memory of one CTA
    float cam[12], cam_new[12], dcf[6];
    double red[(BA_THREADS / 32) * BA_NSUM];
    double tot[BA_NSUM];
    double cand[2];                  // candidate cost, points that left the front of a camera
    int ok;
    int s_last;
};

// Phase A: this CTA's share of the normal equations at (w.cam, pts[cur]) -> part[b][blockIdx.x][34].
__device__ __forceinline__ void ba_cta_accumulate(const DeviceState& s, const BAState& ba, int b, int cur, float lam, BAWork& w) {
    float acc[BA_NSUM];
#pragma unroll
    for (int k = 0; k < BA_NSUM; k++) acc[k] = 0.0f;
    const float* pts = ba.pts + ((size_t)b * 2 + cur) * 3 * s.n_stride;
    const unsigned char* active = ba.active + (size_t)b * s.n_stride;
    for (int i = blockIdx.x * BA_THREADS + threadIdx.x; i < s.n; i += gridDim.x * BA_THREADS) {
        if (!active[i]) continue;
        const float4 p = s.corr[(size_t)b * s.n_stride + i];
        const float X[3] = {pts[i], pts[(size_t)s.n_stride + i], pts[(size_t)2 * s.n_stride + i]};
        float Vinv[6], W[18], gp[3], Jc[12], r2[2], cost;
        if (!ba_linearise<true>(w.cam, w.cam + 9, p, X, lam, Vinv, W, gp, Jc, r2, cost)) continue;
        float Yw[18];                       // W V^-1
#pragma unroll
        for (int r = 0; r < 6; r++) sym3_mul(Vinv, W + 3 * r, Yw + 3 * r);
        int q = 0;
#pragma unroll
        for (int r = 0; r < 6; r++)
#pragma unroll
            for (int c2 = r; c2 < 6; c2++) {
                float Uij = fmaf(Jc[r], Jc[c2], Jc[6 + r] * Jc[6 + c2]);
                float S = fmaf(Yw[3 * r], W[3 * c2], fmaf(Yw[3 * r + 1], W[3 * c2 + 1], Yw[3 * r + 2] * W[3 * c2 + 2]));
                acc[q++] += Uij - S;
            }
#pragma unroll
        for (int r = 0; r < 6; r++) {
            acc[21 + r] += fmaf(Jc[r], Jc[r], Jc[6 + r] * Jc[6 + r]);
            float gc = fmaf(Jc[r], r2[0], Jc[6 + r] * r2[1]);
            acc[27 + r] += gc - fmaf(Yw[3 * r], gp[0], fmaf(Yw[3 * r + 1], gp[1], Yw[3 * r + 2] * gp[2]));
        }
        acc[33] += cost;
    }
    // fixed-order reduction: lanes (fp32 shuffles) -> warps -> CTA partial (fp64 from here on)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();                         // w.red may still be read by the previous phase
#pragma unroll
    for (int k = 0; k < BA_NSUM; k++) {
        float v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
        if (lane == 0) w.red[warp * BA_NSUM + k] = (double)v;
    }
    __syncthreads();
    double* part = ba.part + ((size_t)b * ba.max_blocks + blockIdx.x) * BA_NSUM;
    if (threadIdx.x < BA_NSUM) {
        double v = 0.0;
        for (int q = 0; q < BA_THREADS / 32; q++) v += w.red[q * BA_NSUM + threadIdx.x];
        part[threadIdx.x] = v;
    }
}

// Phase B: every CTA partial of the pair -> w.tot; damped reduced camera system (A + lambda diag U) dc = -g by
// Cholesky in fp64 (thread 0) -> w.dcf, w.ok; candidate camera -> w.cam_new.  Whole CTA; ends synchronised.
__device__ __forceinline__ void ba_cta_reduce_solve(const BAState& ba, int b, float lam, BAWork& w) {
    constexpr int BA_GROUPS = BA_THREADS / BA_NSUM;          // 7 groups of 34 threads, each adds a fixed stripe of CTAs
    __syncthreads();
    {
        const int k = threadIdx.x % BA_NSUM, g = threadIdx.x / BA_NSUM;
        if (g < BA_GROUPS) {
            const double* all = ba.part + (size_t)b * ba.max_blocks * BA_NSUM;
            double v = 0.0;
            for (unsigned q = g; q < gridDim.x; q += BA_GROUPS) v += __ldcg(all + (size_t)q * BA_NSUM + k);
            w.red[g * BA_NSUM + k] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x < BA_NSUM) {
        double v = 0.0;
        for (int g = 0; g < BA_GROUPS; g++) v += w.red[g * BA_NSUM + threadIdx.x];
        w.tot[threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double L[36];
        for (int i = 0; i < 6; i++)
            for (int j = i; j < 6; j++) L[6 * i + j] = L[6 * j + i] = w.tot[sym6(i, j)];
        for (int i = 0; i < 6; i++) L[7 * i] += (double)lam * w.tot[21 + i];
        bool ok = true;
        for (int j = 0; j < 6 && ok; j++) {
            double d = L[7 * j];
            for (int k = 0; k < j; k++) d -= L[6 * j + k] * L[6 * j + k];
            if (!(d > 0.0)) { ok = false; break; }
            d = sqrt(d);
            L[7 * j] = d;
            for (int i = j + 1; i < 6; i++) {
                double v = L[6 * i + j];
                for (int k = 0; k < j; k++) v -= L[6 * i + k] * L[6 * j + k];
                L[6 * i + j] = v / d;
            }
        }
        double dc[6] = {0, 0, 0, 0, 0, 0};
        if (ok) {
            double y[6];
            for (int i = 0; i < 6; i++) {
                double v = -w.tot[27 + i];
                for (int k = 0; k < i; k++) v -= L[6 * i + k] * y[k];
                y[i] = v / L[7 * i];
            }
            for (int i = 5; i >= 0; i--) {
                double v = y[i];
                for (int k = i + 1; k < 6; k++) v -= L[6 * k + i] * dc[k];
                dc[i] = v / L[7 * i];
            }
        }
        for (int i = 0; i < 6; i++) w.dcf[i] = (float)dc[i];
        w.ok = ok ? 1 : 0;
        ba_apply_camera(w.cam, dc, w.cam_new);
    }
    __syncthreads();
}

// Phase C: back-substitution dX = -V^-1 (gp + W^T dc) into the other point buffer, cost of the candidate
// (w.cam_new, new points) -> part2[b][blockIdx.x][2].
__device__ __forceinline__ void ba_cta_update(const DeviceState& s, const BAState& ba, int b, int cur, float lam, BAWork& w) {
    const float* pts = ba.pts + ((size_t)b * 2 + cur) * 3 * s.n_stride;
    float* out = ba.pts + ((size_t)b * 2 + (1 - cur)) * 3 * s.n_stride;
    const unsigned char* active = ba.active + (size_t)b * s.n_stride;
    const float* sCam = w.cam;
    const float* sNew = w.cam_new;
    const float* sDc = w.dcf;
    float cost = 0.0f, bad = 0.0f;
    for (int i = blockIdx.x * BA_THREADS + threadIdx.x; i < s.n; i += gridDim.x * BA_THREADS) {
        if (!active[i]) continue;            // inactive points are never read back from the BA buffers
        float X[3] = {pts[i], pts[(size_t)s.n_stride + i], pts[(size_t)2 * s.n_stride + i]};
        const float4 p = s.corr[(size_t)b * s.n_stride + i];
        float Vinv[6], W[18], gp[3], Jc[12], r2[2], c0;
        if (ba_linearise<false>(sCam, sCam + 9, p, X, lam, Vinv, W, gp, Jc, r2, c0)) {
            float rhs[3], dX[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                float v = gp[k];
#pragma unroll
                for (int r = 0; r < 6; r++) v = fmaf(W[3 * r + k], sDc[r], v);
                rhs[k] = v;
            }
            sym3_mul(Vinv, rhs, dX);
#pragma unroll
            for (int k = 0; k < 3; k++) X[k] -= dX[k];
        }
        float Y[3];
#pragma unroll
        for (int k = 0; k < 3; k++) Y[k] = fmaf(sNew[3 * k], X[0], fmaf(sNew[3 * k + 1], X[1], fmaf(sNew[3 * k + 2], X[2], sNew[9 + k])));
        if (X[2] > 0.0f && Y[2] > 0.0f) {
            const float iz = 1.0f / X[2], iz2 = 1.0f / Y[2];
            const float a0 = fmaf(X[0], iz, -p.x), a1 = fmaf(X[1], iz, -p.y), b0 = fmaf(Y[0], iz2, -p.z), b1 = fmaf(Y[1], iz2, -p.w);
            cost += fmaf(a0, a0, fmaf(a1, a1, fmaf(b0, b0, b1 * b1)));
        } else {
            bad += 1.0f;
        }
        out[i] = X[0];
        out[(size_t)s.n_stride + i] = X[1];
        out[(size_t)2 * s.n_stride + i] = X[2];
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double v0 = (double)cost, v1 = (double)bad;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v0 += __shfl_down_sync(0xFFFFFFFFu, v0, o);
        v1 += __shfl_down_sync(0xFFFFFFFFu, v1, o);
    }
    __syncthreads();
    if (lane == 0) { w.red[warp * 2] = v0; w.red[warp * 2 + 1] = v1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double c = 0.0, bd = 0.0;
        for (int q = 0; q < BA_THREADS / 32; q++) { c += w.red[q * 2]; bd += w.red[q * 2 + 1]; }
        double* part2 = ba.part2 + ((size_t)b * ba.max_blocks + blockIdx.x) * 2;
        part2[0] = c;
        part2[1] = bd;
    }
}

// Phase D: candidate cost and bad-point count of the pair over all CTAs -> w.cand.  Whole CTA; ends synchronised.
__device__ __forceinline__ void ba_cta_reduce_candidate(const BAState& ba, int b, BAWork& w) {
    __syncthreads();
    if (threadIdx.x < 32) {                  // 32 fixed stripes of the per-CTA partials, then one thread adds the stripes
        const double* all = ba.part2 + (size_t)b * ba.max_blocks * 2;
        double c = 0.0, bd = 0.0;
        for (unsigned k = threadIdx.x; k < gridDim.x; k += 32) { c += __ldcg(all + (size_t)k * 2); bd += __ldcg(all + (size_t)k * 2 + 1); }
        w.red[2 * threadIdx.x] = c;
        w.red[2 * threadIdx.x + 1] = bd;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double c = 0.0, bd = 0.0;
        for (int k = 0; k < 32; k++) { c += w.red[2 * k]; bd += w.red[2 * k + 1]; }
        w.cand[0] = c;
        w.cand[1] = bd;
    }
    __syncthreads();
}

// The accept rule, identical in both paths: strict decrease, every point still in front of both cameras, solve ok.
__device__ __forceinline__ bool ba_accept(int ok, double cand_cost, double cand_bad, float cost_now) {
    return ok != 0 && cand_bad == 0.0 && (float)cand_cost < cost_now;
}

__global__ void __launch_bounds__(BA_THREADS) ba_accumulate_kernel(DeviceState s, BAState ba, int first) {
    const int b = blockIdx.y;
    int* ci = ba.ctl_i + b * 8;
    float* cf = ba.ctl_f + b * 8;
    if (ci[3] < 8) return;
    __shared__ BAWork w;
    const int cur = ci[0];
    const float lam = cf[0];
    if (threadIdx.x < 12) w.cam[threadIdx.x] = ba.cam[(size_t)b * 24 + 12 * cur + threadIdx.x];
    __syncthreads();
    ba_cta_accumulate(s, ba, b, cur, lam, w);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) w.s_last = (atomicAdd(&ci[1], 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (!w.s_last) return;
    __threadfence();
    ba_cta_reduce_solve(ba, b, lam, w);
    if (threadIdx.x < 6) ba.dc[b * 6 + threadIdx.x] = (double)w.dcf[threadIdx.x];
    if (threadIdx.x < 12) ba.cam[(size_t)b * 24 + 12 * (1 - cur) + threadIdx.x] = w.cam_new[threadIdx.x];
    if (threadIdx.x == 0) {
        ci[1] = 0;
        ci[5] = w.ok;
        cf[1] = (float)w.tot[33];
        if (first) cf[2] = (float)w.tot[33];
    }
}

__global__ void __launch_bounds__(BA_THREADS) ba_update_kernel(DeviceState s, BAState ba) {
    const int b = blockIdx.y;
    int* ci = ba.ctl_i + b * 8;
    float* cf = ba.ctl_f + b * 8;
    if (ci[3] < 8) return;
    __shared__ BAWork w;
    const int cur = ci[0];
    const float lam = cf[0];
    if (threadIdx.x < 12) {
        w.cam[threadIdx.x] = ba.cam[(size_t)b * 24 + 12 * cur + threadIdx.x];
        w.cam_new[threadIdx.x] = ba.cam[(size_t)b * 24 + 12 * (1 - cur) + threadIdx.x];
    }
    if (threadIdx.x < 6) w.dcf[threadIdx.x] = (float)ba.dc[b * 6 + threadIdx.x];
    __syncthreads();
    ba_cta_update(s, ba, b, cur, lam, w);
    if (threadIdx.x == 0) {
        __threadfence();
        w.s_last = (atomicAdd(&ci[2], 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!w.s_last) return;
    __threadfence();
    ba_cta_reduce_candidate(ba, b, w);
    if (threadIdx.x != 0) return;
    ci[2] = 0;
    if (ba_accept(ci[5], w.cand[0], w.cand[1], cf[1])) {
        ci[0] = 1 - cur;
        ci[4] += 1;
        cf[1] = (float)w.cand[0];
        cf[0] = fmaxf(lam * (1.0f / 3.0f), 1e-9f);
    } else {
        cf[0] = fminf(lam * 4.0f, 1e6f);
    }
}

// All LM iterations of an outer round in ONE cooperative launch: the state (current buffer, lambda, cost, camera)
// is replicated in every CTA and advanced identically, the CTAs meet at two grid barriers per iteration, and CTA 0
// of each pair publishes the result at the end.  ~8 us per iteration at 10,000 correspondences instead of ~20 us
// for the two launches.  Needs the whole grid resident; launch_bundle_adjust falls back to the two-kernel path otherwise.
__global__ void __launch_bounds__(BA_THREADS) ba_persistent_kernel(DeviceState s, BAState ba, int iterations) {
    cg::grid_group grid = cg::this_grid();
    const int b = blockIdx.y;
    int* ci = ba.ctl_i + b * 8;
    float* cf = ba.ctl_f + b * 8;
    __shared__ BAWork w;
    const bool live = ci[3] >= 8;            // pairs with too few active points still take part in the barriers
    int cur = ci[0], accepted = ci[4];
    float lam = cf[0], cost_now = 0.0f, cost_first = 0.0f;
    if (threadIdx.x < 12) w.cam[threadIdx.x] = ba.cam[(size_t)b * 24 + 12 * cur + threadIdx.x];
    __syncthreads();
    for (int it = 0; it < iterations; it++) {
        if (live) ba_cta_accumulate(s, ba, b, cur, lam, w);
        grid.sync();
        if (live) {
            ba_cta_reduce_solve(ba, b, lam, w);
            cost_now = (float)w.tot[33];
            if (it == 0) cost_first = cost_now;
            ba_cta_update(s, ba, b, cur, lam, w);
        }
        grid.sync();
        if (live) {
            ba_cta_reduce_candidate(ba, b, w);
            const bool accept = ba_accept(w.ok, w.cand[0], w.cand[1], cost_now);
            __syncthreads();
            if (accept) {
                cur = 1 - cur;
                accepted += 1;
                cost_now = (float)w.cand[0];
                lam = fmaxf(lam * (1.0f / 3.0f), 1e-9f);
                if (threadIdx.x < 12) w.cam[threadIdx.x] = w.cam_new[threadIdx.x];
            } else {
                lam = fminf(lam * 4.0f, 1e6f);
            }
            __syncthreads();
        }
    }
    if (live && blockIdx.x == 0) {
        if (threadIdx.x < 12) ba.cam[(size_t)b * 24 + 12 * cur + threadIdx.x] = w.cam[threadIdx.x];
        if (threadIdx.x == 0) {
            ci[0] = cur;
            ci[4] = accepted;
            cf[0] = lam;
            cf[1] = cost_now;
            cf[2] = cost_first;
        }
    }
}

// End of an outer round, camera side: gauge |t| = 1, M back into P[P_ind], E = (tx R)^T scaled to
// singular values (1, 1, 0) (x1^T E x2 = 0 with x1 ~ X, x2 ~ R X + t), statistics.
__global__ void ba_finalise_kernel(DeviceState s, BAState ba, float* stats_out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= s.B) return;
    int* ci = ba.ctl_i + b * 8;
    float* cf = ba.ctl_f + b * 8;
    ci[6] = 0;
    ci[7] = 0;
    if (ci[3] >= 8) {
        const float* cam = ba.cam + (size_t)b * 24 + 12 * ci[0];
        float R[9], t[3];
        for (int i = 0; i < 9; i++) R[i] = cam[i];
        for (int i = 0; i < 3; i++) t[i] = cam[9 + i];
        // one Newton step back onto SO(3): R <- (3 I - R R^T) R / 2  (R drifts only by rounding)
        float RRt[9], Rn[9];
        mul33_ABt(R, R, RRt);
        for (int i = 0; i < 9; i++) RRt[i] = (i % 4 == 0 ? 3.0f : 0.0f) - RRt[i];
        mul33(RRt, R, Rn);
        for (int i = 0; i < 9; i++) R[i] = 0.5f * Rn[i];
        const float nt = sqrtf(fmaf(t[0], t[0], fmaf(t[1], t[1], t[2] * t[2])));
        const float sc = nt > 0.0f ? 1.0f / nt : 1.0f;
        cf[3] = sc;
        float* M = ba.cand + (size_t)b * 32;         // candidate camera (16) and essential matrix (9): committed by
        for (int i = 0; i < 3; i++) {                // ba_commit_kernel only if it does not lose inliers
            for (int j = 0; j < 3; j++) M[4 * i + j] = R[3 * i + j];
            M[4 * i + 3] = t[i] * sc;
        }
        M[12] = 0.0f; M[13] = 0.0f; M[14] = 0.0f; M[15] = 1.0f;
        const float tx[9] = {0.0f, -t[2] * sc, t[1] * sc, t[2] * sc, 0.0f, -t[0] * sc, -t[1] * sc, t[0] * sc, 0.0f};
        float F[9];
        mul33(tx, R, F);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) M[16 + 3 * i + j] = F[3 * j + i];
    } else {
        cf[3] = 1.0f;
    }
    if (stats_out) {
        float* o = stats_out + (size_t)b * 8;
        o[0] = (float)ci[3];     // active points
        o[1] = cf[2];            // cost at entry of the round
        o[2] = cf[1];            // cost now
        o[3] = (float)ci[4];     // accepted steps
        o[4] = cf[0];            // lambda
        o[5] = cf[3];            // gauge scale applied
    }
}

// Inliers of the candidate E of each pair (same fp32 test as everywhere) -> ctl_i[6].
__global__ void __launch_bounds__(BA_THREADS) ba_count_kernel(DeviceState s, BAState ba, float thr) {
    const int b = blockIdx.y;
    int* ci = ba.ctl_i + b * 8;
    if (ci[3] < 8) return;
    __shared__ float sE[9];
    __shared__ int s_cnt;
    if (threadIdx.x < 9) sE[threadIdx.x] = ba.cand[(size_t)b * 32 + 16 + threadIdx.x];
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const int i = blockIdx.x * BA_THREADS + threadIdx.x;
    bool inl = false;
    if (i < s.n) {
        const float4 p = s.corr[(size_t)b * s.n_stride + i];
        inl = epipolar_d(s.metric, sE, p.x, p.y, p.z, p.w, -thr) < 0.0f;
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, inl);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s_cnt, __popc(m));
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) atomicAdd(&ci[6], s_cnt);
}

// Guard against losing the consensus: the adjusted camera replaces P[pose_index], E and the best count only if its
// essential matrix still explains at least 95 % of the correspondences the model had when sfmb200_bundle_adjust was
// called (the count may dip by a few borderline points while the pose moves along the rotation / translation
// valley - a strict "never fewer" rule would freeze the adjustment there - but it must not collapse, which happens
// when only a small part of the inliers triangulates in front of both cameras).
__global__ void ba_commit_kernel(DeviceState s, BAState ba, float* stats_out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= s.B) return;
    int* ci = ba.ctl_i + b * 8;
    const bool commit = ci[3] >= 8 && 20LL * ci[6] >= 19LL * ba.base_count[b];
    if (commit) {
        const float* c = ba.cand + (size_t)b * 32;
        float* M = s.P + (size_t)b * 64 + 16 * s.P_ind[b];
        for (int i = 0; i < 16; i++) M[i] = c[i];
        for (int i = 0; i < 9; i++) s.E[(size_t)b * 9 + i] = c[16 + i];
        s.best_count[b] = ci[6];
        ci[7] = 1;
    }
    if (stats_out) {
        stats_out[(size_t)b * 8 + 6] = (float)ci[6];      // inliers of the adjusted model
        stats_out[(size_t)b * 8 + 7] = commit ? 1.0f : 0.0f;
    }
}

// End of an outer round, point side (after the cloud has been re-triangulated with the camera now in
// P[pose_index]): when the adjusted model was committed, its points, in the |t| = 1 gauge, replace the DLT
// points of the active correspondences.  Also clears the per-round state.
__global__ void __launch_bounds__(BA_THREADS) ba_scatter_kernel(DeviceState s, BAState ba) {
    const int b = blockIdx.y;
    const int* ci = ba.ctl_i + b * 8;
    const int i = blockIdx.x * BA_THREADS + threadIdx.x;
    if (i < s.n && ci[7] != 0 && ba.active[(size_t)b * s.n_stride + i]) {
        const float sc = ba.ctl_f[b * 8 + 3];
        const float* pts = ba.pts + ((size_t)b * 2 + ci[0]) * 3 * s.n_stride;
        float* out = s.points + (size_t)b * 4 * s.n_stride;
        out[i] = pts[i] * sc;
        out[(size_t)s.n_stride + i] = pts[(size_t)s.n_stride + i] * sc;
        out[(size_t)2 * s.n_stride + i] = pts[(size_t)2 * s.n_stride + i] * sc;
    }
}

__global__ void ba_publish_kernel(DeviceState s, BAState ba) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= s.B) return;
    ba.ctl_i[b * 8 + 3] = 0;
}

// One outer round = re-triangulate, select the active set, `iterations` LM steps, write back.
// Returns the number of kernel launches.
int launch_bundle_adjust(const DeviceState& s, const BAState& ba, float thr, int iterations, float lambda0,
                         int tri_inliers_only, int first_round, float* d_stats, cudaStream_t st) {
    int launches = 0;
    // init / scatter: one thread per correspondence; LM kernels: at most max_blocks CTAs per pair (about four
    // CTAs per SM over the whole batch, several correspondences per thread) so the per-CTA reductions stay cheap
    const int nb_all = (s.n + BA_THREADS - 1) / BA_THREADS;
    const int nb = nb_all < ba.max_blocks ? nb_all : ba.max_blocks;
    launch_triangulate(s, 1, thr, st);      // only inliers of the current E can become active
    ba_init_kernel<<<dim3(nb_all, s.B), BA_THREADS, 0, st>>>(s, ba, thr, lambda0, first_round);
    launches += 2;
    // one cooperative launch for all iterations when the whole grid is resident, else two launches per iteration
    int resident_per_sm = 0, sms = 0, dev = 0;
    if (ba.persistent) {
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident_per_sm, ba_persistent_kernel, BA_THREADS, 0);
    }
    if (ba.persistent && (long long)nb * s.B <= (long long)resident_per_sm * sms) {
        DeviceState sv = s;
        BAState bv = ba;
        int iters = iterations;
        void* args[] = {&sv, &bv, &iters};
        cudaLaunchCooperativeKernel((void*)ba_persistent_kernel, dim3(nb, s.B), dim3(BA_THREADS), args, 0, st);
        launches += 1;
    } else {
        for (int it = 0; it < iterations; it++) {
            ba_accumulate_kernel<<<dim3(nb, s.B), BA_THREADS, 0, st>>>(s, ba, it == 0);
            ba_update_kernel<<<dim3(nb, s.B), BA_THREADS, 0, st>>>(s, ba);
            launches += 2;
        }
    }
    ba_finalise_kernel<<<(s.B + 63) / 64, 64, 0, st>>>(s, ba, d_stats);
    ba_count_kernel<<<dim3(nb_all, s.B), BA_THREADS, 0, st>>>(s, ba, thr);
    ba_commit_kernel<<<(s.B + 63) / 64, 64, 0, st>>>(s, ba, d_stats);
    launch_triangulate(s, tri_inliers_only, thr, st);   // whole cloud under the camera (and E) now in place
    ba_scatter_kernel<<<dim3(nb_all, s.B), BA_THREADS, 0, st>>>(s, ba);
    ba_publish_kernel<<<(s.B + 63) / 64, 64, 0, st>>>(s, ba);
    launches += 6;
    return launches;
}

}  // namespace sfmb200
