// N-view chaining of consecutive pairs.  The reference shapes
// Image_pair for `image_count` views (sfm.h:23,30-31: vector<float*> U, X) but only ever
// handles two; its README lists more views as future work (README.md:65-69).
//
// Convention: pair b of the handle = (view b, view b+1); correspondence i is the SAME
// track in every pair (index-aligned tracks).  Every pair has been reconstructed in its
// own frame (camera b = [I|0], unit baseline).  This stage puts them in one frame:
//   scale[b] (b >= 1) : robust central value, over the tracks valid in pairs b-1 and b,
//                       of  depth in camera b according to pair b-1 / according to pair b:
//                       median bin of a 2048-bin histogram of the log ratio, refined to
//                       the mean of that bin's entries (integer atomics: deterministic);
//   cameras           : G_0 = [I|0], G_{b+1} = [R_b | S_b t_b] G_b, S_b = prod scale[1..b]
//                       (units of the first baseline);
//   cloud             : per track the mean over the pairs where it is valid of
//                       G_b^-1 (S_b X_b), plus the number of pairs that saw it.
// A track is valid in a pair when it passes the pair's inlier test (same fp32 Sampson
// test as everywhere) and its point is finite and in front of both cameras.
// oracle/oracle.py: chain_scales / chain_cameras / chain_merge restate this in fp64.
#include "internal.cuh"
#include "sampson.cuh"

namespace sfmb200 {

constexpr int CHAIN_THREADS = 256;
constexpr int CHAIN_BINS = 2048;
constexpr float CHAIN_LOG_RANGE = 2.772588722239781f;      // ln 16
constexpr double CHAIN_FIX = 1099511627776.0;              // 2^40 fixed point for the in-bin mean

struct PairView {
    float E[9];
    float M[12];
};
__device__ __forceinline__ void load_pair(const DeviceState& s, int b, PairView& v) {
#pragma unroll
    for (int k = 0; k < 9; k++) v.E[k] = s.E[(size_t)b * 9 + k];
    const float* M = s.P + (size_t)b * 64 + 16 * s.P_ind[b];
#pragma unroll
    for (int k = 0; k < 12; k++) v.M[k] = M[k];
}
// valid + point of track i in pair b; z2 = depth in camera b+1
__device__ __forceinline__ bool track_in_pair(const DeviceState& s, int b, const PairView& v, int i, float thr, float* X, float& z2) {
    const float4 p = s.corr[(size_t)b * s.n_stride + i];
    const float* in = s.points + (size_t)b * 4 * s.n_stride;
    X[0] = in[i]; X[1] = in[(size_t)s.n_stride + i]; X[2] = in[(size_t)2 * s.n_stride + i];
    z2 = fmaf(v.M[8], X[0], fmaf(v.M[9], X[1], fmaf(v.M[10], X[2], v.M[11])));
    return epipolar_d(s.metric, v.E, p.x, p.y, p.z, p.w, -thr) < 0.0f && isfinite(X[0]) && isfinite(X[1]) && isfinite(X[2]) &&
           X[2] > 0.0f && z2 > 0.0f;
}
// log depth ratio of track i between pairs b-1 and b; false when the track does not link them
__device__ __forceinline__ bool link_ratio(const DeviceState& s, int b, const PairView& prev, const PairView& cur, int i, float thr,
                                           float& lr, int& bin) {
    float Xp[3], Xc[3], zp, zc;
    if (!track_in_pair(s, b - 1, prev, i, thr, Xp, zp)) return false;
    if (!track_in_pair(s, b, cur, i, thr, Xc, zc)) return false;
    lr = logf(zp / Xc[2]);
    if (!(fabsf(lr) < CHAIN_LOG_RANGE)) return false;
    int q = (int)((lr + CHAIN_LOG_RANGE) * (CHAIN_BINS / (2.0f * CHAIN_LOG_RANGE)));
    bin = q < CHAIN_BINS - 1 ? q : CHAIN_BINS - 1;
    return true;
}

// pass 0: histogram of the log ratios; pass 1: fixed-point sum and count of the entries of the median bin
__global__ void __launch_bounds__(CHAIN_THREADS) chain_ratio_kernel(DeviceState s, ChainState c, float thr, int pass) {
    const int b = blockIdx.y + 1;
    __shared__ int hist[CHAIN_BINS];
    __shared__ PairView prev, cur;
    if (threadIdx.x == 0) { load_pair(s, b - 1, prev); load_pair(s, b, cur); }
    if (pass == 0)
        for (int k = threadIdx.x; k < CHAIN_BINS; k += CHAIN_THREADS) hist[k] = 0;
    __syncthreads();
    const int mb = pass ? c.median_bin[b] : -1;
    long long sum = 0;
    int cnt = 0;
    for (int i = blockIdx.x * CHAIN_THREADS + threadIdx.x; i < s.n; i += gridDim.x * CHAIN_THREADS) {
        float lr;
        int bin;
        if (!link_ratio(s, b, prev, cur, i, thr, lr, bin)) continue;
        if (pass == 0) atomicAdd(&hist[bin], 1);
        else if (bin == mb) { sum += (long long)((double)lr * CHAIN_FIX); cnt++; }
    }
    if (pass == 0) {
        __syncthreads();
        for (int k = threadIdx.x; k < CHAIN_BINS; k += CHAIN_THREADS)
            if (hist[k]) atomicAdd(&c.hist[(size_t)b * CHAIN_BINS + k], hist[k]);
    } else {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sum += __shfl_down_sync(0xFFFFFFFFu, sum, o);
            cnt += __shfl_down_sync(0xFFFFFFFFu, cnt, o);
        }
        if ((threadIdx.x & 31) == 0 && cnt) {
            atomicAdd(&c.bin_sum[b], (unsigned long long)sum);      // two's complement: signed sums add correctly
            atomicAdd(&c.bin_cnt[b], cnt);
        }
    }
}

// median bin of each link histogram (one CTA per link)
__global__ void __launch_bounds__(CHAIN_THREADS) chain_median_kernel(ChainState c) {
    const int b = blockIdx.x + 1;
    __shared__ int part[CHAIN_THREADS];
    constexpr int PER = CHAIN_BINS / CHAIN_THREADS;
    const int* h = c.hist + (size_t)b * CHAIN_BINS;
    int local[PER], acc = 0;
#pragma unroll
    for (int k = 0; k < PER; k++) { local[k] = h[threadIdx.x * PER + k]; acc += local[k]; }
    part[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int t = 0; t < CHAIN_THREADS; t++) { int v = part[t]; part[t] = run; run += v; }
        c.used[b] = run;
        c.median_bin[b] = -1;
    }
    __syncthreads();
    const int total = c.used[b];
    if (total == 0) return;
    const int target = (total + 1) / 2;          // first bin whose cumulative count reaches it
    int run = part[threadIdx.x];
#pragma unroll
    for (int k = 0; k < PER; k++) {
        if (run < target && run + local[k] >= target) c.median_bin[b] = threadIdx.x * PER + k;
        run += local[k];
    }
}

// scales, cumulative scales and global cameras (one thread; pairs are few)
__global__ void chain_compose_kernel(DeviceState s, ChainState c) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double G[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    double S = 1.0;
    for (int k = 0; k < 12; k++) c.cameras[k] = (float)G[k];
    c.used[0] = 0;
    for (int b = 0; b < s.B; b++) {
        double sc = 1.0;
        if (b > 0 && c.bin_cnt[b] > 0) sc = exp((double)(long long)c.bin_sum[b] / CHAIN_FIX / (double)c.bin_cnt[b]);
        S *= sc;
        c.scales[b] = (float)sc;
        c.cum_scales[b] = (float)S;
        const float* M = s.P + (size_t)b * 64 + 16 * s.P_ind[b];
        double T[12], N[12];
        for (int r = 0; r < 3; r++) {
            for (int q = 0; q < 3; q++) T[4 * r + q] = M[4 * r + q];
            T[4 * r + 3] = S * (double)M[4 * r + 3];
        }
        for (int r = 0; r < 3; r++)
            for (int q = 0; q < 4; q++) {
                double a = T[4 * r] * G[q] + T[4 * r + 1] * G[4 + q] + T[4 * r + 2] * G[8 + q];
                N[4 * r + q] = a + (q == 3 ? T[4 * r + 3] : 0.0);
            }
        for (int k = 0; k < 12; k++) { G[k] = N[k]; c.cameras[(size_t)(b + 1) * 12 + k] = (float)N[k]; }
    }
}

__global__ void __launch_bounds__(CHAIN_THREADS) chain_merge_kernel(DeviceState s, ChainState c, float thr, float* cloud, int* count) {
    extern __shared__ float sh[];                 // per pair: E 9, M 12, G 12, S 1 = 34 floats
    for (int t = threadIdx.x; t < s.B * 34; t += CHAIN_THREADS) {
        int b = t / 34, k = t % 34;
        float v;
        if (k < 9) v = s.E[(size_t)b * 9 + k];
        else if (k < 21) v = s.P[(size_t)b * 64 + 16 * s.P_ind[b] + (k - 9)];
        else if (k < 33) v = c.cameras[(size_t)b * 12 + (k - 21)];
        else v = c.cum_scales[b];
        sh[t] = v;
    }
    __syncthreads();
    const int i = blockIdx.x * CHAIN_THREADS + threadIdx.x;
    if (i >= s.n) return;
    float ax = 0.0f, ay = 0.0f, az = 0.0f;
    int cnt = 0;
    for (int b = 0; b < s.B; b++) {
        const float* pv = sh + b * 34;
        PairView v;
#pragma unroll
        for (int k = 0; k < 9; k++) v.E[k] = pv[k];
#pragma unroll
        for (int k = 0; k < 12; k++) v.M[k] = pv[9 + k];
        float X[3], z2;
        if (!track_in_pair(s, b, v, i, thr, X, z2)) continue;
        const float* G = pv + 21;
        const float S = pv[33];
        const float d0 = fmaf(S, X[0], -G[3]), d1 = fmaf(S, X[1], -G[7]), d2 = fmaf(S, X[2], -G[11]);
        ax += fmaf(G[0], d0, fmaf(G[4], d1, G[8] * d2));      // R^T (S X - t)
        ay += fmaf(G[1], d0, fmaf(G[5], d1, G[9] * d2));
        az += fmaf(G[2], d0, fmaf(G[6], d1, G[10] * d2));
        cnt++;
    }
    const float inv = cnt ? 1.0f / (float)cnt : 0.0f;
    if (cloud) {
        cloud[i] = ax * inv;
        cloud[(size_t)s.n + i] = ay * inv;
        cloud[(size_t)2 * s.n + i] = az * inv;
        cloud[(size_t)3 * s.n + i] = 1.0f;
    }
    if (count) count[i] = cnt;
}

int launch_chain(const DeviceState& s, const ChainState& c, float thr, float* d_cloud, int* d_count, cudaStream_t st) {
    int launches = 0;
    const int nb_all = (s.n + CHAIN_THREADS - 1) / CHAIN_THREADS;
    int nb = (592 + s.B - 1) / s.B;
    nb = nb < 1 ? 1 : (nb > nb_all ? nb_all : nb);
    cudaMemsetAsync(c.hist, 0, (size_t)s.B * CHAIN_BINS * sizeof(int), st);
    cudaMemsetAsync(c.bin_sum, 0, (size_t)s.B * sizeof(unsigned long long), st);
    cudaMemsetAsync(c.bin_cnt, 0, (size_t)s.B * sizeof(int), st);
    if (s.B > 1) {
        chain_ratio_kernel<<<dim3(nb, s.B - 1), CHAIN_THREADS, 0, st>>>(s, c, thr, 0);
        chain_median_kernel<<<s.B - 1, CHAIN_THREADS, 0, st>>>(c);
        chain_ratio_kernel<<<dim3(nb, s.B - 1), CHAIN_THREADS, 0, st>>>(s, c, thr, 1);
        launches += 3;
    }
    chain_compose_kernel<<<1, 32, 0, st>>>(s, c);
    launches++;
    if (d_cloud || d_count) {
        chain_merge_kernel<<<nb_all, CHAIN_THREADS, (size_t)s.B * 34 * sizeof(float), st>>>(s, c, thr, d_cloud, d_count);
        launches++;
    }
    return launches;
}

// ===========================================================================
// Global bundle adjustment over the chained reconstruction ("bundle adjustment" and more views are the
// reference's listed future work, README.md:65-69; sfm.h:23,30-31 shapes Image_pair for image_count views).
//
// After sfmb200_chain_views the V = pairs + 1 cameras G_k (world = camera 0 -> camera k) and one point per track sit in one
// frame.  This stage refines ALL of them together by Levenberg-Marquardt on the reprojection error
//     sum_k sum_{i observed by k} | pi(R_k X_i + t_k) - x_{k,i} |^2          (normalised camera coordinates)
// Observation (k, i) = the coordinates of track i in view k: pair 0's first image for k = 0, else pair k-1's second image;
// it takes part when the track is valid (inlier of the pair's E, point finite and in front of both cameras) in a pair that
// contains view k.  Parameters: cameras 1..V-1 (rotation by a left-multiplicative so(3) update, translation; camera 0 stays
// [I|0]) and one 3-D point per track seen by >= 2 views; the scale gauge is left to the damping and restored afterwards
// (|t_1| keeps its length).  The point blocks are eliminated: the reduced camera system S dc = -rhs has 6 (V-1) unknowns,
//     S = U* - sum_i W_i V_i*^-1 W_i^T,   rhs = g_c - sum_i W_i V_i*^-1 g_p,i       (* = LM damping on the diagonals)
// and is assembled block by block (6x6 per camera pair), per-CTA partial sums in a fixed order (fp32 per thread -> fp64
// across threads and CTAs), solved by the last CTA with a Cholesky factorisation in shared memory (fp64, <= 96 x 96).
//   gba_prepare_kernel     observation masks, start values
//   gba_accumulate_kernel  per track V_i*^-1, g_p,i, cost; per camera pair the 6x6 block + rhs; last CTA: solve, candidate cameras
//   gba_update_kernel      back-substitution dX_i = -V_i*^-1 (g_p,i + W_i^T dc), candidate cost; last CTA accepts (strict
//                          decrease, every observation still in front of its camera: lambda / 3) or rejects (4 lambda)
//   gba_finish_kernel      gauge, cameras back into the chain's array, cloud back into the caller's buffer
// Two launches per iteration, nothing returns to the host in between.  oracle/oracle.py: bundle_adjust_global restates the
// iteration in fp64 (explicit Jacobian, damped normal equations - the same step).
// ===========================================================================
constexpr int GBA_THREADS = 256;
constexpr int GBA_SLOTS = 42;               // 36 block entries + 6 right-hand-side entries
constexpr float GBA_GATE = 25.0f;           // observation gate: squared reprojection error at the start <= 25 x threshold

__device__ __forceinline__ float2 gba_obs(const DeviceState& s, int k, int i) {
    if (k == 0) {
        const float4 p = s.corr[i];
        return make_float2(p.x, p.y);
    }
    const float4 p = s.corr[(size_t)(k - 1) * s.n_stride + i];
    return make_float2(p.z, p.w);
}
// residual and Jacobians of one observation; cam = R (9, row-major) | t (3).  false: the point is not in front of the camera
template <bool WANT_JC>
__device__ __forceinline__ bool gba_linearise(const float* cam, float2 uv, const float* X, float Jp[2][3], float Jc[2][6], float r[2]) {
    float Q[3], Y[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        Q[a] = fmaf(cam[3 * a], X[0], fmaf(cam[3 * a + 1], X[1], cam[3 * a + 2] * X[2]));
        Y[a] = Q[a] + cam[9 + a];
    }
    if (!(Y[2] > 0.0f)) return false;
    const float iz = 1.0f / Y[2], u = Y[0] * iz, v = Y[1] * iz;
    r[0] = u - uv.x;
    r[1] = v - uv.y;
    const float bp[2][3] = {{iz, 0.0f, -u * iz}, {0.0f, iz, -v * iz}};
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int c = 0; c < 3; c++) Jp[a][c] = fmaf(bp[a][0], cam[c], fmaf(bp[a][1], cam[3 + c], bp[a][2] * cam[6 + c]));
    if (WANT_JC) {
        const float N[3][3] = {{0.0f, Q[2], -Q[1]}, {-Q[2], 0.0f, Q[0]}, {Q[1], -Q[0], 0.0f}};       // -[Q]x
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                Jc[a][c] = fmaf(bp[a][0], N[0][c], fmaf(bp[a][1], N[1][c], bp[a][2] * N[2][c]));
                Jc[a][3 + c] = bp[a][c];
            }
    }
    return true;
}

__global__ void __launch_bounds__(GBA_THREADS) gba_prepare_kernel(DeviceState s, ChainState c, GbaState g, float thr, const float* cloud,
                                                                  const int* count) {
    const int V = s.B + 1;
    if (blockIdx.x == 0) {
        for (int t = threadIdx.x; t < V * 12; t += GBA_THREADS) {
            const int k = t / 12, q = t % 12;
            const float* G = c.cameras + (size_t)k * 12;                    // 3x4 row-major [R | t]
            const float v = q < 9 ? G[4 * (q / 3) + (q % 3)] : G[4 * (q - 9) + 3];
            g.cam[t] = v;
            g.cam[V * 12 + t] = v;
        }
        if (threadIdx.x == 0) {
            g.ctl_i[0] = 0;      // current buffer
            g.ctl_i[1] = 0;      // accumulate ticket
            g.ctl_i[2] = 0;      // update ticket
            g.ctl_i[3] = 0;      // accepted steps
            g.ctl_i[4] = 1;      // last solve succeeded
            g.ctl_i[5] = 0;      // iterations run
            g.ctl_f[0] = 1e-3f;  // lambda
            g.ctl_f[1] = 0.0f;   // cost at the current state
            g.ctl_f[2] = -1.0f;  // cost at entry
            const float* G1 = c.cameras + 12;
            g.ctl_f[3] = sqrtf(G1[3] * G1[3] + G1[7] * G1[7] + G1[11] * G1[11]);     // |t_1|: the gauge
        }
    }
    const int i = blockIdx.x * GBA_THREADS + threadIdx.x;
    if (i >= s.n) return;
    unsigned int mask = 0u;
    for (int b = 0; b < s.B; b++) {
        PairView v;
        load_pair(s, b, v);
        float X[3], z2;
        if (track_in_pair(s, b, v, i, thr, X, z2)) mask |= (3u << b);        // views b and b + 1
    }
    const float x = cloud[i], y = cloud[(size_t)s.n + i], z = cloud[(size_t)2 * s.n + i];
    if (!(count[i] > 0) || !isfinite(x) || !isfinite(y) || !isfinite(z)) mask = 0u;
    // Gate: a match can pass a pair's epipolar test and still be wrong ALONG the epipolar line; in the joint problem such an
    // observation is a gross outlier that a plain least-squares cost would follow.  Observations whose reprojection error at
    // the chained start exceeds GBA_GATE x the inlier threshold (12 px at the reference's focal length) are left out.
    for (int k = 0; k < V && mask; k++) {
        if (!((mask >> k) & 1u)) continue;
        const float* G = c.cameras + (size_t)k * 12;
        const float Yx = fmaf(G[0], x, fmaf(G[1], y, fmaf(G[2], z, G[3])));
        const float Yy = fmaf(G[4], x, fmaf(G[5], y, fmaf(G[6], z, G[7])));
        const float Yz = fmaf(G[8], x, fmaf(G[9], y, fmaf(G[10], z, G[11])));
        const float2 uv = gba_obs(s, k, i);
        const float du = Yx / Yz - uv.x, dv = Yy / Yz - uv.y;
        if (!(Yz > 0.0f) || !(du * du + dv * dv <= GBA_GATE * thr)) mask &= ~(1u << k);
    }
    if (__popc(mask) < 2) mask = 0u;
    g.obs[i] = mask;
    for (int buf = 0; buf < 2; buf++) {
        float* P = g.pts + (size_t)buf * 3 * s.n;
        P[i] = x; P[(size_t)s.n + i] = y; P[(size_t)2 * s.n + i] = z;
    }
}

// deterministic CTA sum of one double per thread into out[0] by thread 0 (red: GBA_THREADS / 32 doubles of shared memory)
__device__ __forceinline__ double gba_block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < GBA_THREADS / 32; w++) t += red[w];
    return t;
}

__global__ void __launch_bounds__(GBA_THREADS) gba_accumulate_kernel(DeviceState s, GbaState g) {
    extern __shared__ double gba_sh[];           // last CTA: S [D][D], rhs [D]
    __shared__ double red[GBA_THREADS / 32];
    __shared__ float scam[17 * 12];
    __shared__ int s_last;
    const int V = s.B + 1, M = s.B, n = s.n;
    const int cur = g.ctl_i[0];
    const float lam = g.ctl_f[0];
    for (int t = threadIdx.x; t < V * 12; t += GBA_THREADS) scam[t] = g.cam[(size_t)cur * V * 12 + t];
    __syncthreads();
    const float* P = g.pts + (size_t)cur * 3 * n;
    const int nblocks = M * (M + 1) / 2;
    const int stride = gridDim.x * GBA_THREADS;
    // ---- pass A: per track V*^-1, g_p, cost ----
    float cost = 0.0f;
    for (int i = blockIdx.x * GBA_THREADS + threadIdx.x; i < n; i += stride) {
        const unsigned int mask = g.obs[i];
        if (!mask) continue;
        const float X[3] = {P[i], P[(size_t)n + i], P[(size_t)2 * n + i]};
        float Vm[6] = {0, 0, 0, 0, 0, 0}, gp[3] = {0, 0, 0};
        for (int k = 0; k < V; k++) {
            if (!((mask >> k) & 1u)) continue;
            float Jp[2][3], Jc[2][6], r[2];
            if (!gba_linearise<false>(scam + 12 * k, gba_obs(s, k, i), X, Jp, Jc, r)) continue;
            cost += r[0] * r[0] + r[1] * r[1];
            int q = 0;
#pragma unroll
            for (int a = 0; a < 3; a++) {
#pragma unroll
                for (int b = a; b < 3; b++) Vm[q++] += Jp[0][a] * Jp[0][b] + Jp[1][a] * Jp[1][b];
                gp[a] += Jp[0][a] * r[0] + Jp[1][a] * r[1];
            }
        }
        const float d = 1.0f + lam;
        const float v00 = Vm[0] * d, v01 = Vm[1], v02 = Vm[2], v11 = Vm[3] * d, v12 = Vm[4], v22 = Vm[5] * d;
        const float c00 = v11 * v22 - v12 * v12, c01 = v02 * v12 - v01 * v22, c02 = v01 * v12 - v02 * v11;
        const float det = v00 * c00 + v01 * c01 + v02 * c02;
        const float id = fabsf(det) > 0.0f ? 1.0f / det : 0.0f;
        float* L = g.lin + (size_t)i * 9;
        L[0] = c00 * id; L[1] = c01 * id; L[2] = c02 * id;
        L[3] = (v00 * v22 - v02 * v02) * id; L[4] = (v01 * v02 - v00 * v12) * id; L[5] = (v00 * v11 - v01 * v01) * id;
        L[6] = gp[0]; L[7] = gp[1]; L[8] = gp[2];
    }
    {
        const double t = gba_block_sum((double)cost, red);
        if (threadIdx.x == 0) g.part[((size_t)blockIdx.x * (nblocks + 1) + nblocks) * GBA_SLOTS] = t;
    }
    // ---- pass B: the 6x6 blocks of the reduced camera system (same thread -> same tracks: no grid barrier needed) ----
    int blk = 0;
    for (int m = 0; m < M; m++)
        for (int l = 0; l <= m; l++, blk++) {
            float acc[GBA_SLOTS];
#pragma unroll
            for (int q = 0; q < GBA_SLOTS; q++) acc[q] = 0.0f;
            for (int i = blockIdx.x * GBA_THREADS + threadIdx.x; i < n; i += stride) {
                const unsigned int mask = g.obs[i];
                if (!((mask >> (m + 1)) & 1u) || !((mask >> (l + 1)) & 1u)) continue;
                const float X[3] = {P[i], P[(size_t)n + i], P[(size_t)2 * n + i]};
                const float* L = g.lin + (size_t)i * 9;
                const float Vi[6] = {L[0], L[1], L[2], L[3], L[4], L[5]};
                float Jp[2][3], Jc[2][6], r[2];
                if (!gba_linearise<true>(scam + 12 * (m + 1), gba_obs(s, m + 1, i), X, Jp, Jc, r)) continue;
                float Wm[6][3], T[6][3];
#pragma unroll
                for (int a = 0; a < 6; a++)
#pragma unroll
                    for (int c = 0; c < 3; c++) Wm[a][c] = Jc[0][a] * Jp[0][c] + Jc[1][a] * Jp[1][c];
#pragma unroll
                for (int a = 0; a < 6; a++) {                    // T = W_m V^-1
                    T[a][0] = Wm[a][0] * Vi[0] + Wm[a][1] * Vi[1] + Wm[a][2] * Vi[2];
                    T[a][1] = Wm[a][0] * Vi[1] + Wm[a][1] * Vi[3] + Wm[a][2] * Vi[4];
                    T[a][2] = Wm[a][0] * Vi[2] + Wm[a][1] * Vi[4] + Wm[a][2] * Vi[5];
                }
                if (l == m) {
#pragma unroll
                    for (int a = 0; a < 6; a++) {
#pragma unroll
                        for (int b = 0; b < 6; b++) {
                            float u = Jc[0][a] * Jc[0][b] + Jc[1][a] * Jc[1][b];
                            if (a == b) u *= 1.0f + lam;
                            acc[6 * a + b] += u - (T[a][0] * Wm[b][0] + T[a][1] * Wm[b][1] + T[a][2] * Wm[b][2]);
                        }
                        acc[36 + a] += Jc[0][a] * r[0] + Jc[1][a] * r[1] - (T[a][0] * L[6] + T[a][1] * L[7] + T[a][2] * L[8]);
                    }
                } else {
                    float Jp2[2][3], Jc2[2][6], r2[2];
                    if (!gba_linearise<true>(scam + 12 * (l + 1), gba_obs(s, l + 1, i), X, Jp2, Jc2, r2)) continue;
#pragma unroll
                    for (int b = 0; b < 6; b++) {
                        const float w0 = Jc2[0][b] * Jp2[0][0] + Jc2[1][b] * Jp2[1][0];
                        const float w1 = Jc2[0][b] * Jp2[0][1] + Jc2[1][b] * Jp2[1][1];
                        const float w2 = Jc2[0][b] * Jp2[0][2] + Jc2[1][b] * Jp2[1][2];
#pragma unroll
                        for (int a = 0; a < 6; a++) acc[6 * a + b] -= T[a][0] * w0 + T[a][1] * w1 + T[a][2] * w2;
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < GBA_SLOTS; q++) {
                const double t = gba_block_sum((double)acc[q], red);
                if (threadIdx.x == 0) g.part[((size_t)blockIdx.x * (nblocks + 1) + blk) * GBA_SLOTS + q] = t;
            }
        }
    // ---- last CTA: fixed-order sum over the CTAs, damped Cholesky solve, candidate cameras ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&g.ctl_i[1], 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int D = 6 * M;
    double* S = gba_sh;
    double* rhs = gba_sh + (size_t)D * D;
    for (int e = threadIdx.x; e < nblocks * GBA_SLOTS + 1; e += GBA_THREADS) {
        const int b = e / GBA_SLOTS, q = e % GBA_SLOTS;
        double t = 0.0;
        const size_t off = e < nblocks * GBA_SLOTS ? (size_t)b * GBA_SLOTS + q : (size_t)nblocks * GBA_SLOTS;
        for (int cta = 0; cta < (int)gridDim.x; cta++) t += __ldcg(&g.part[(size_t)cta * (nblocks + 1) * GBA_SLOTS + off]);
        if (e == nblocks * GBA_SLOTS) {
            g.ctl_f[1] = (float)t;
            if (g.ctl_f[2] < 0.0f) g.ctl_f[2] = (float)t;
            continue;
        }
        // block index -> (m, l)
        int m = 0, base = 0;
        while (base + m + 1 <= b) { base += m + 1; m++; }
        const int l = b - base;
        if (q < 36) {
            const int a = q / 6, c = q % 6;
            S[(size_t)(6 * m + a) * D + 6 * l + c] = t;
            if (l != m) S[(size_t)(6 * l + c) * D + 6 * m + a] = t;
        } else if (l == m) {
            rhs[6 * m + (q - 36)] = t;
        }
    }
    __syncthreads();
    // symmetrise the diagonal blocks (their two triangles were summed separately), then Cholesky S = L L^T in place
    for (int e = threadIdx.x; e < D * D; e += GBA_THREADS) {
        const int a = e / D, c = e % D;
        if (a < c && a / 6 == c / 6) {
            const double v = 0.5 * (S[(size_t)a * D + c] + S[(size_t)c * D + a]);
            S[(size_t)a * D + c] = v;
            S[(size_t)c * D + a] = v;
        }
    }
    __syncthreads();
    __shared__ int s_ok;
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    for (int j = 0; j < D; j++) {
        if (threadIdx.x == 0) {
            const double d = S[(size_t)j * D + j];
            if (!(d > 0.0)) s_ok = 0;
            S[(size_t)j * D + j] = sqrt(d > 0.0 ? d : 1.0);
        }
        __syncthreads();
        const double ljj = S[(size_t)j * D + j];
        for (int a = j + 1 + threadIdx.x; a < D; a += GBA_THREADS) S[(size_t)a * D + j] /= ljj;
        __syncthreads();
        for (int e = threadIdx.x; e < (D - j - 1) * (D - j - 1); e += GBA_THREADS) {
            const int a = j + 1 + e / (D - j - 1), c = j + 1 + e % (D - j - 1);
            if (c <= a) S[(size_t)a * D + c] -= S[(size_t)a * D + j] * S[(size_t)c * D + j];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        // L y = -rhs, L^T dc = y
        for (int a = 0; a < D; a++) {
            double v = -rhs[a];
            for (int c = 0; c < a; c++) v -= S[(size_t)a * D + c] * rhs[c];
            rhs[a] = v / S[(size_t)a * D + a];
        }
        for (int a = D - 1; a >= 0; a--) {
            double v = rhs[a];
            for (int c = a + 1; c < D; c++) v -= S[(size_t)c * D + a] * rhs[c];
            rhs[a] = v / S[(size_t)a * D + a];
        }
        bool fin = s_ok != 0;
        for (int a = 0; a < D; a++) fin = fin && isfinite(rhs[a]);
        for (int a = 0; a < D; a++) g.dc[a] = fin ? rhs[a] : 0.0;
        g.ctl_i[4] = fin ? 1 : 0;
        g.ctl_i[1] = 0;          // ticket for the next iteration
    }
    __syncthreads();
    // candidate cameras into the other buffer: R <- exp([w]x) R, t <- t + dt (camera 0 is copied)
    float* out = g.cam + (size_t)(1 - cur) * V * 12;
    for (int k = threadIdx.x; k < V; k += GBA_THREADS) {
        if (k == 0) {
            for (int q = 0; q < 12; q++) out[q] = scam[q];
            continue;
        }
        const double* dc = g.dc + 6 * (k - 1);
        const double w[3] = {dc[0], dc[1], dc[2]};
        const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2);
        const double A = th > 1e-12 ? sin(th) / th : 1.0, Bc = th > 1e-12 ? (1.0 - cos(th)) / th2 : 0.5;
        const double Kx[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
        double Ex[9];
        for (int a = 0; a < 3; a++)
            for (int c = 0; c < 3; c++) {
                double k2 = 0.0;
                for (int e = 0; e < 3; e++) k2 += Kx[3 * a + e] * Kx[3 * e + c];
                Ex[3 * a + c] = (a == c ? 1.0 : 0.0) + A * Kx[3 * a + c] + Bc * k2;
            }
        const float* cam = scam + 12 * k;
        for (int a = 0; a < 3; a++)
            for (int c = 0; c < 3; c++) {
                double v = 0.0;
                for (int e = 0; e < 3; e++) v += Ex[3 * a + e] * (double)cam[3 * e + c];
                out[12 * k + 3 * a + c] = (float)v;
            }
        for (int a = 0; a < 3; a++) out[12 * k + 9 + a] = (float)((double)cam[9 + a] + dc[3 + a]);
    }
}

__global__ void __launch_bounds__(GBA_THREADS) gba_update_kernel(DeviceState s, GbaState g) {
    __shared__ double red[GBA_THREADS / 32];
    __shared__ float scam[17 * 12], ncam[17 * 12];
    __shared__ float sdc[96];
    __shared__ int s_last;
    const int V = s.B + 1, M = s.B, n = s.n;
    const int cur = g.ctl_i[0];
    for (int t = threadIdx.x; t < V * 12; t += GBA_THREADS) {
        scam[t] = g.cam[(size_t)cur * V * 12 + t];
        ncam[t] = g.cam[(size_t)(1 - cur) * V * 12 + t];
    }
    for (int t = threadIdx.x; t < 6 * M; t += GBA_THREADS) sdc[t] = (float)g.dc[t];
    __syncthreads();
    const float* P = g.pts + (size_t)cur * 3 * n;
    float* Pn = g.pts + (size_t)(1 - cur) * 3 * n;
    float cost = 0.0f;
    int bad = 0;
    for (int i = blockIdx.x * GBA_THREADS + threadIdx.x; i < n; i += gridDim.x * GBA_THREADS) {
        const unsigned int mask = g.obs[i];
        if (!mask) continue;
        const float X[3] = {P[i], P[(size_t)n + i], P[(size_t)2 * n + i]};
        const float* L = g.lin + (size_t)i * 9;
        float sv[3] = {L[6], L[7], L[8]};                               // g_p + sum_k W_k^T dc_k
        for (int k = 1; k < V; k++) {
            if (!((mask >> k) & 1u)) continue;
            float Jp[2][3], Jc[2][6], r[2];
            if (!gba_linearise<true>(scam + 12 * k, gba_obs(s, k, i), X, Jp, Jc, r)) continue;
            float jd[2] = {0.0f, 0.0f};                                  // Jc dc_k
#pragma unroll
            for (int a = 0; a < 6; a++) {
                jd[0] += Jc[0][a] * sdc[6 * (k - 1) + a];
                jd[1] += Jc[1][a] * sdc[6 * (k - 1) + a];
            }
#pragma unroll
            for (int c = 0; c < 3; c++) sv[c] += Jp[0][c] * jd[0] + Jp[1][c] * jd[1];
        }
        float Xn[3];
        Xn[0] = X[0] - (L[0] * sv[0] + L[1] * sv[1] + L[2] * sv[2]);
        Xn[1] = X[1] - (L[1] * sv[0] + L[3] * sv[1] + L[4] * sv[2]);
        Xn[2] = X[2] - (L[2] * sv[0] + L[4] * sv[1] + L[5] * sv[2]);
        Pn[i] = Xn[0]; Pn[(size_t)n + i] = Xn[1]; Pn[(size_t)2 * n + i] = Xn[2];
        for (int k = 0; k < V; k++) {
            if (!((mask >> k) & 1u)) continue;
            float Jp[2][3], Jc[2][6], r[2];
            if (!gba_linearise<false>(ncam + 12 * k, gba_obs(s, k, i), Xn, Jp, Jc, r)) { bad++; continue; }
            cost += r[0] * r[0] + r[1] * r[1];
        }
    }
    const double tc = gba_block_sum((double)cost, red);
    const double tb = gba_block_sum((double)bad, red);
    if (threadIdx.x == 0) {
        g.part2[2 * blockIdx.x] = tc;
        g.part2[2 * blockIdx.x + 1] = tb;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&g.ctl_i[2], 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (!s_last || threadIdx.x != 0) return;
    __threadfence();
    double c2 = 0.0, b2 = 0.0;
    for (int cta = 0; cta < (int)gridDim.x; cta++) {
        c2 += __ldcg(&g.part2[2 * cta]);
        b2 += __ldcg(&g.part2[2 * cta + 1]);
    }
    const bool accept = g.ctl_i[4] != 0 && b2 == 0.0 && c2 < (double)g.ctl_f[1];
    if (accept) {
        g.ctl_i[0] = 1 - cur;
        g.ctl_i[3] += 1;
        g.ctl_f[0] = fmaxf(g.ctl_f[0] * (1.0f / 3.0f), 1e-9f);
        g.ctl_f[4] = (float)c2;                                      // cost of the accepted state (stats)
    } else {
        g.ctl_f[0] = fminf(g.ctl_f[0] * 4.0f, 1e6f);
    }
    g.ctl_i[5] += 1;
    g.ctl_i[2] = 0;
}

// gauge (|t_1| keeps the length it had), cameras back into the chain's [V][12] array, cloud back into the caller's buffer
__global__ void __launch_bounds__(GBA_THREADS) gba_finish_kernel(DeviceState s, ChainState c, GbaState g, float* cloud, float* stats) {
    const int V = s.B + 1, n = s.n;
    const int cur = g.ctl_i[0];
    const float* cam = g.cam + (size_t)cur * V * 12;
    const float t1 = sqrtf(cam[12 + 9] * cam[12 + 9] + cam[12 + 10] * cam[12 + 10] + cam[12 + 11] * cam[12 + 11]);
    const float sc = (t1 > 0.0f && g.ctl_f[3] > 0.0f) ? g.ctl_f[3] / t1 : 1.0f;
    if (blockIdx.x == 0) {
        for (int t = threadIdx.x; t < V * 12; t += GBA_THREADS) {
            const int k = t / 12, q = t % 12, row = q / 4, col = q % 4;
            c.cameras[t] = col < 3 ? cam[12 * k + 3 * row + col] : cam[12 * k + 9 + row] * sc;
        }
        if (threadIdx.x == 0 && stats) {
            stats[0] = g.ctl_f[2];                                    // cost at entry
            stats[1] = g.ctl_i[3] > 0 ? g.ctl_f[4] : g.ctl_f[2];      // cost at exit
            stats[2] = (float)g.ctl_i[3];                             // accepted steps
            stats[3] = g.ctl_f[0];                                    // lambda
            stats[4] = sc;                                            // gauge scale applied
            stats[5] = (float)g.ctl_i[5];                             // iterations run
        }
    }
    const float* P = g.pts + (size_t)cur * 3 * n;
    for (int i = blockIdx.x * GBA_THREADS + threadIdx.x; i < n; i += gridDim.x * GBA_THREADS) {
        if (!g.obs[i]) continue;                                      // tracks outside the adjustment keep the chain's value
        cloud[i] = P[i] * sc;
        cloud[(size_t)n + i] = P[(size_t)n + i] * sc;
        cloud[(size_t)2 * n + i] = P[(size_t)2 * n + i] * sc;
    }
}

size_t gba_part_doubles(int pairs, int nb) { return (size_t)nb * (pairs * (pairs + 1) / 2 + 1) * GBA_SLOTS; }

int launch_global_ba(const DeviceState& s, const ChainState& c, const GbaState& g, float thr, int iterations, float* d_cloud,
                     const int* d_count, float* d_stats, cudaStream_t st) {
    const int nb_all = (s.n + GBA_THREADS - 1) / GBA_THREADS;
    const int nb = nb_all < g.nb ? nb_all : g.nb;
    const int D = 6 * s.B;
    const size_t smem = ((size_t)D * D + D) * sizeof(double);
    cudaFuncSetAttribute(gba_accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    gba_prepare_kernel<<<nb_all, GBA_THREADS, 0, st>>>(s, c, g, thr, d_cloud, d_count);
    for (int it = 0; it < iterations; it++) {
        gba_accumulate_kernel<<<nb, GBA_THREADS, smem, st>>>(s, g);
        gba_update_kernel<<<nb, GBA_THREADS, 0, st>>>(s, g);
    }
    gba_finish_kernel<<<nb, GBA_THREADS, 0, st>>>(s, c, g, d_cloud, d_stats);
    return 2 + 2 * iterations;
}

}  // namespace sfmb200
