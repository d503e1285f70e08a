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

}  // namespace sfmb200
