// Local optimisation of the RANSAC winner (LO-RANSAC refit on the inlier set).
// The reference stops at the best minimal-sample hypothesis and lists iterating
// further as future work (README.md:65-69); SURVEY.md 8f ranks this as the
// next component after the hot path: normalised 8-point on ALL inliers through
// the same 9x9 symmetric Jacobi eigensolve, re-score, keep if better.
//
// A chain of models m_0 = RANSAC winner, m_{k+1} = fit(inliers of m_k at
// mult_k * thr) with mult_k = 4, 2, 1, 1, ... (the usual LO-RANSAC shrinking
// threshold); the incumbent E is the chain member with the most inliers at thr.
// One iteration = two launches per batch, no host round trip:
//   refit_eval_kernel  : for the chain model: inlier count at thr (same fp32
//                        Sampson test as the scoring kernel) and the first/second
//                        moments of its inliers at the fitting threshold; the
//                        last CTA to finish adopts the model as E iff it has MORE
//                        inliers than the incumbent and derives the Hartley
//                        similarity (centroid, sqrt(2)/RMS scale) of both images.
//   refit_solve_kernel : 9x9 Gram matrix of the normalised design rows of the
//                        chain model's fitting inliers (45 sums); the last CTA reduces the
//                        per-CTA partials in a fixed order (fp64), runs a
//                        warp-cooperative two-sided Jacobi (one lane per row of G
//                        and V, shared memory), de-normalises, projects to rank 2
//                        and publishes the next candidate.
// Everything is deterministic: fixed grid, fixed-order reductions, no float atomics.
#include "internal.cuh"
#include "sampson.cuh"
#include "smallmat.cuh"

namespace sfmb200 {

constexpr int REFIT_THREADS = 256;

// Fixed-order block reduction of NV values per thread; result valid in thread 0.
template <int NV>
__device__ __forceinline__ void block_reduce(float* v, float* smem /* [warps][NV] */) {
#pragma unroll
    for (int i = 0; i < NV; i++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_down_sync(0xFFFFFFFFu, v[i], o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) smem[warp * NV + i] = v[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) {
            float acc = smem[i];
            for (int w = 1; w < REFIT_THREADS / 32; w++) acc += smem[w * NV + i];
            v[i] = acc;
        }
    }
    __syncthreads();
}

// moments: [0] count, [1..2] sum x1,y1, [3..4] sum x2,y2, [5] sum x1^2+y1^2, [6] sum x2^2+y2^2
constexpr int NMOM = 7;

__global__ void __launch_bounds__(REFIT_THREADS) refit_eval_kernel(DeviceState s, RefitState r, float thr, float thr_fit, int first) {
    const int b = blockIdx.y;
    __shared__ float red[(REFIT_THREADS / 32) * (NMOM + 1)];
    __shared__ float sE[9];
    __shared__ int s_last;
    int* flags = r.flags + (size_t)b * 4;             // [0] active, [1] incumbent count, [2] eval ticket, [3] solve ticket
    if (!first && flags[0] == 0) return;              // chain ended (fewer than 8 fitting inliers)
    const float* cand = r.cand + (size_t)b * 9;
    if (threadIdx.x < 9) sE[threadIdx.x] = cand[threadIdx.x];
    __syncthreads();
    float m[NMOM + 1];                                // [NMOM] = count at thr (acceptance)
#pragma unroll
    for (int i = 0; i <= NMOM; i++) m[i] = 0.0f;
    const float4* corr = s.corr + (size_t)b * s.n_stride;
    for (int i = blockIdx.x * REFIT_THREADS + threadIdx.x; i < s.n; i += gridDim.x * REFIT_THREADS) {
        float4 p = corr[i];
        if (epipolar_d(s.metric, sE, p.x, p.y, p.z, p.w, -thr) < 0.0f) m[NMOM] += 1.0f;
        if (epipolar_d(s.metric, sE, p.x, p.y, p.z, p.w, -thr_fit) < 0.0f) {
            m[0] += 1.0f;
            m[1] += p.x; m[2] += p.y; m[3] += p.z; m[4] += p.w;
            m[5] += fmaf(p.x, p.x, p.y * p.y);
            m[6] += fmaf(p.z, p.z, p.w * p.w);
        }
    }
    block_reduce<NMOM + 1>(m, red);
    float* part = r.mom_part + ((size_t)b * gridDim.x + blockIdx.x) * (NMOM + 1);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i <= NMOM; i++) part[i] = m[i];
        __threadfence();
        s_last = (atomicAdd(&flags[2], 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last || threadIdx.x != 0) return;
    __threadfence();
    flags[2] = 0;                                      // ticket reset for the next launch
    double tot[NMOM + 1];
    for (int i = 0; i <= NMOM; i++) tot[i] = 0.0;
    const float* all = r.mom_part + (size_t)b * gridDim.x * (NMOM + 1);
    for (unsigned k = 0; k < gridDim.x; k++)
        for (int i = 0; i <= NMOM; i++) tot[i] += (double)__ldcg(all + k * (NMOM + 1) + i);
    const int count = (int)(tot[NMOM] + 0.5);          // inliers at thr: the acceptance score
    if (count > flags[1]) {                            // strictly better than the incumbent (first call: always)
        flags[1] = count;
#pragma unroll
        for (int i = 0; i < 9; i++) s.E[(size_t)b * 9 + i] = sE[i];
        s.best_count[b] = count;
        if (!first) r.iters_done[b] += 1;
    }
    // Hartley similarity of the fitting inliers from their moments: centroid, scale = sqrt(2) / RMS distance
    const double n = tot[0];
    if (n >= 8.0) {
        double c1x = tot[1] / n, c1y = tot[2] / n, c2x = tot[3] / n, c2y = tot[4] / n;
        double v1 = tot[5] / n - (c1x * c1x + c1y * c1y), v2 = tot[6] / n - (c2x * c2x + c2y * c2y);
        float* T = r.T + (size_t)b * 8;
        T[0] = (float)(v1 > 0 ? sqrt(2.0 / v1) : 1.0); T[1] = (float)c1x; T[2] = (float)c1y;
        T[3] = (float)(v2 > 0 ? sqrt(2.0 / v2) : 1.0); T[4] = (float)c2x; T[5] = (float)c2y;
        flags[0] = 1;
    } else {
        flags[0] = 0;
    }
}

constexpr int NG = 45;

__global__ void __launch_bounds__(REFIT_THREADS) refit_solve_kernel(DeviceState s, RefitState r, float thr_fit) {
    const int b = blockIdx.y;
    __shared__ float red[(REFIT_THREADS / 32) * NG];
    __shared__ float sE[9], sT[6];
    __shared__ float G[9][9], V[9][9];
    __shared__ int s_last;
    int* flags = r.flags + (size_t)b * 4;
    if (flags[0] == 0) return;
    if (threadIdx.x < 9) sE[threadIdx.x] = r.cand[(size_t)b * 9 + threadIdx.x];
    if (threadIdx.x < 6) sT[threadIdx.x] = r.T[(size_t)b * 8 + threadIdx.x];
    __syncthreads();
    float g[NG];
#pragma unroll
    for (int i = 0; i < NG; i++) g[i] = 0.0f;
    const float4* corr = s.corr + (size_t)b * s.n_stride;
    for (int i = blockIdx.x * REFIT_THREADS + threadIdx.x; i < s.n; i += gridDim.x * REFIT_THREADS) {
        float4 p = corr[i];
        if (epipolar_d(s.metric, sE, p.x, p.y, p.z, p.w, -thr_fit) < 0.0f) {
            float x1 = sT[0] * (p.x - sT[1]), y1 = sT[0] * (p.y - sT[2]);
            float x2 = sT[3] * (p.z - sT[4]), y2 = sT[3] * (p.w - sT[5]);
            float a[9] = {x1 * x2, x1 * y2, x1, y1 * x2, y1 * y2, y1, x2, y2, 1.0f};
            int k = 0;
#pragma unroll
            for (int i2 = 0; i2 < 9; i2++)
#pragma unroll
                for (int j2 = i2; j2 < 9; j2++) { g[k] = fmaf(a[i2], a[j2], g[k]); k++; }
        }
    }
    block_reduce<NG>(g, red);
    float* part = r.gram_part + ((size_t)b * gridDim.x + blockIdx.x) * NG;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NG; i++) part[i] = g[i];
        __threadfence();
        s_last = (atomicAdd(&flags[3], 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    // ---- last CTA: fixed-order final reduction, then the eigensolve in warp 0 ----
    __threadfence();
    if (threadIdx.x < NG) {
        double acc = 0.0;
        const float* all = r.gram_part + (size_t)b * gridDim.x * NG;
        for (unsigned k = 0; k < gridDim.x; k++) acc += (double)__ldcg(all + k * NG + threadIdx.x);
        // unpack upper-triangular index -> (i, j)
        int idx = threadIdx.x, i = 0;
        while (idx >= 9 - i) { idx -= 9 - i; i++; }
        int j = i + idx;
        G[i][j] = (float)acc;
        G[j][i] = (float)acc;
    }
    if (threadIdx.x < 81) V[threadIdx.x / 9][threadIdx.x % 9] = (threadIdx.x / 9 == threadIdx.x % 9) ? 1.0f : 0.0f;
    if (threadIdx.x == 0) flags[3] = 0;
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    // Warp-cooperative two-sided Jacobi in PARALLEL ORDER: a sweep is 9 rounds of a round-robin tournament over
    // the 9 indices (+ one bye), each round rotating 4 disjoint index pairs at once - the four rotations commute,
    // and none touches another's (p,p), (q,q), (p,q), so the angles taken before the round are the ones the
    // sequential method would use.  Lanes 0..3 compute the angles, lane k owns row k (then column k) of G and V.
    // 72 rounds instead of 288 sequential rotations: the stage drops from ~100 us to ~15 us.
    __shared__ float rot[4][2];
    __shared__ int rpq[4][2];
    for (int sw = 0; sw < 8; sw++) {
        // converged when the off-diagonal mass is at fp32 rounding level of the diagonal (usually after 3-4 sweeps)
        __syncwarp();
        float off = 0.0f, dia = 0.0f;
        if (lane < 9) {
            for (int j = 0; j < 9; j++) {
                const float v = G[lane][j];
                if (j == lane) dia = v * v; else off = fmaf(v, v, off);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            off += __shfl_xor_sync(0xFFFFFFFFu, off, o);
            dia += __shfl_xor_sync(0xFFFFFFFFu, dia, o);
        }
        if (off <= 1e-12f * dia) break;
        for (int rd = 0; rd < 9; rd++) {
            __syncwarp();
            if (lane < 4) {
                int p = (rd + lane + 1) % 9, q = (rd + 9 - (lane + 1)) % 9;
                if (p > q) { int t2 = p; p = q; q = t2; }
                float c, sn, t;
                jacobi_angle(G[p][p], G[q][q], G[p][q], c, sn, t);
                rot[lane][0] = c; rot[lane][1] = sn;
                rpq[lane][0] = p; rpq[lane][1] = q;
            }
            __syncwarp();
            if (lane < 9) {                            // G <- G J, V <- V J (columns p, q of the four pairs)
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int p = rpq[k][0], q = rpq[k][1];
                    const float c = rot[k][0], sn = rot[k][1];
                    const float gkp = G[lane][p], gkq = G[lane][q], vkp = V[lane][p], vkq = V[lane][q];
                    G[lane][p] = fmaf(c, gkp, -sn * gkq); G[lane][q] = fmaf(sn, gkp, c * gkq);
                    V[lane][p] = fmaf(c, vkp, -sn * vkq); V[lane][q] = fmaf(sn, vkp, c * vkq);
                }
            }
            __syncwarp();
            if (lane < 9) {                            // G <- J^T G (rows p, q of the four pairs)
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int p = rpq[k][0], q = rpq[k][1];
                    const float c = rot[k][0], sn = rot[k][1];
                    const float gpj = G[p][lane], gqj = G[q][lane];
                    G[p][lane] = fmaf(c, gpj, -sn * gqj); G[q][lane] = fmaf(sn, gpj, c * gqj);
                }
            }
        }
    }
    __syncwarp();
    if (lane != 0) return;
    int m = 0;
    for (int i = 1; i < 9; i++)
        if (G[i][i] < G[m][m]) m = i;
    float e[9];
    for (int k = 0; k < 9; k++) e[k] = V[k][m];
    // de-normalise E = T1^T Eh T2 (same algebra as hyp_solver.cuh), project, publish
    const float s1 = sT[0], c1x = sT[1], c1y = sT[2], s2 = sT[3], c2x = sT[4], c2y = sT[5];
    float M[9], E[9];
    for (int i = 0; i < 3; i++) {
        M[3 * i + 0] = e[3 * i + 0] * s2;
        M[3 * i + 1] = e[3 * i + 1] * s2;
        M[3 * i + 2] = fmaf(-s2 * c2x, e[3 * i + 0], fmaf(-s2 * c2y, e[3 * i + 1], e[3 * i + 2]));
    }
    for (int j = 0; j < 3; j++) {
        E[0 + j] = s1 * M[0 + j];
        E[3 + j] = s1 * M[3 + j];
        E[6 + j] = fmaf(-s1 * c1x, M[0 + j], fmaf(-s1 * c1y, M[3 + j], M[6 + j]));
    }
    project_essential(E);
    bool finite = true;
    for (int i = 0; i < 9; i++) finite = finite && (fabsf(E[i]) <= 3.0e38f);
    for (int i = 0; i < 9; i++) r.cand[(size_t)b * 9 + i] = finite ? E[i] : 0.0f;
}

__global__ void refit_reset_kernel(DeviceState s, RefitState r) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= s.B) return;
    r.flags[(size_t)b * 4 + 0] = 0;
    r.flags[(size_t)b * 4 + 1] = -1;
    r.flags[(size_t)b * 4 + 2] = 0;
    r.flags[(size_t)b * 4 + 3] = 0;
    r.iters_done[b] = 0;
    for (int i = 0; i < 9; i++) r.cand[(size_t)b * 9 + i] = s.E[(size_t)b * 9 + i];   // chain starts at the RANSAC winner
}

// fitting-threshold multiplier of chain step k: 4, 2, 1, 1, ...
static float refit_mult(int k) { return k == 0 ? 4.0f : (k == 1 ? 2.0f : 1.0f); }

int launch_refit(const DeviceState& s, const RefitState& r, float thr, int iterations, cudaStream_t st) {
    int blocks = (s.n + REFIT_THREADS * 8 - 1) / (REFIT_THREADS * 8);
    if (blocks < 1) blocks = 1;
    if (blocks > r.max_blocks) blocks = r.max_blocks;
    dim3 grid(blocks, s.B);
    refit_reset_kernel<<<(s.B + 127) / 128, 128, 0, st>>>(s, r);
    int launches = 1;
    refit_eval_kernel<<<grid, REFIT_THREADS, 0, st>>>(s, r, thr, thr * refit_mult(0), 1);
    launches++;
    for (int it = 0; it < iterations; it++) {
        refit_solve_kernel<<<grid, REFIT_THREADS, 0, st>>>(s, r, thr * refit_mult(it));
        refit_eval_kernel<<<grid, REFIT_THREADS, 0, st>>>(s, r, thr, thr * refit_mult(it + 1), 0);
        launches += 2;
    }
    return launches;
}

}  // namespace sfmb200
