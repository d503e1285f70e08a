// Inlier scoring: hypotheses x correspondences Sampson test with a fused
// (count, index) arg-max.  Replaces the reference's calculateInliers
// (SfM/sfm.cu:155-236: six cublasSgemmStridedBatched + element-wise kernels +
// vecnorm + threshold_count over 88*N*H bytes of temporaries) and the
// thrust::max_element selection (sfm.cu:135-137).
//
// Mapping.  A CTA owns a tile of hypotheses (threads own 2 or 4 hypotheses: E
// lives in registers for the whole kernel) and a contiguous range of points.
// Points stream through shared memory in 512-point stages filled by 1-D TMA
// bulk copies (cp.async.bulk + mbarrier complete_tx, double buffered); every
// thread reads the same point per step, so the LDS is a broadcast.  Inlier
// counts stay in per-thread registers: the sign bit of
//     d = fma(den, -thr, num*num)        (d < 0  <=>  num^2 < thr*den)
// is added with one integer instruction, so there is no ballot/popc per step.
// Per evaluation: 16 FFMA + 2 FMUL (FP32 pipe) + 1 integer add.
//
// Two code paths with bit-identical results (same fma tree per lane):
//   scalar : FFMA, 2 hypotheses per thread;
//   packed : FFMA2 (fma.rn.f32x2, new on sm_100) on hypothesis pairs,
//            4 hypotheses per thread, points staged as (x1,x1,y1,y1),(x2,x2,y2,y2).
//
// Epilogue.  splits == 1: counts written directly, block arg-max from
// registers.  splits > 1: partial counts are atomically added to counts[];
// the last CTA to finish a hypothesis tile (ticket counter) reads the totals
// and does the tile's arg-max.  Either way one atomicMax per tile on the packed
// key (count << 32) | (0xFFFFFFFF - global index): highest count, lowest index
// on ties = thrust::max_element semantics.  No separate arg-max kernel.
//
// Roofline: FP32 pipe; algorithmic work 34 FLOP per evaluation.
#include "internal.cuh"

namespace sfmb200 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier.
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// d = num^2 - thr*den for one hypothesis and one point (z = 1 in both views).
// The oracle's fp32 port (oracle/oracle_c.c: sampson_d_f32) mirrors this tree.
__device__ __forceinline__ float sampson_d(const float* e, float x1, float y1, float x2, float y2, float nthr) {
    float l0 = fmaf(e[0], x2, fmaf(e[1], y2, e[2]));
    float l1 = fmaf(e[3], x2, fmaf(e[4], y2, e[5]));
    float l2 = fmaf(e[6], x2, fmaf(e[7], y2, e[8]));
    float num = fmaf(x1, l0, fmaf(y1, l1, l2));
    float m0 = fmaf(e[0], x1, fmaf(e[3], y1, e[6]));
    float m1 = fmaf(e[1], x1, fmaf(e[4], y1, e[7]));
    float den = fmaf(l0, l0, fmaf(l1, l1, fmaf(m0, m0, m1 * m1)));
    return fmaf(den, nthr, num * num);
}
__device__ __forceinline__ float2 sampson_d2(const float2* e, float2 x1, float2 y1, float2 x2, float2 y2, float2 nthr) {
    float2 l0 = __ffma2_rn(e[0], x2, __ffma2_rn(e[1], y2, e[2]));
    float2 l1 = __ffma2_rn(e[3], x2, __ffma2_rn(e[4], y2, e[5]));
    float2 l2 = __ffma2_rn(e[6], x2, __ffma2_rn(e[7], y2, e[8]));
    float2 num = __ffma2_rn(x1, l0, __ffma2_rn(y1, l1, l2));
    float2 m0 = __ffma2_rn(e[0], x1, __ffma2_rn(e[3], y1, e[6]));
    float2 m1 = __ffma2_rn(e[1], x1, __ffma2_rn(e[4], y1, e[7]));
    float2 den = __ffma2_rn(l0, l0, __ffma2_rn(l1, l1, __ffma2_rn(m0, m0, __fmul2_rn(m1, m1))));
    return __ffma2_rn(den, nthr, __fmul2_rn(num, num));
}

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long w = __shfl_xor_sync(0xFFFFFFFFu, v, o);
        v = w > v ? w : v;
    }
    return v;
}

template <int HPT, bool PACKED, int THREADS>
__global__ void __launch_bounds__(THREADS)
score_kernel(DeviceState s, int H, int h_offset, int splits, int pts_per_split, float thr) {
    constexpr int SCORE_THREADS = THREADS;
    constexpr int HPC = HPT * SCORE_THREADS;
    constexpr int F4_PER_PT = PACKED ? 2 : 1;
    __shared__ __align__(128) float4 buf[2][SCORE_CHUNK * F4_PER_PT];
    __shared__ __align__(8) uint64_t full[2];
    __shared__ unsigned long long red[THREADS / 32];
    __shared__ int s_ticket;

    const int tid = threadIdx.x;
    const int tile = blockIdx.x, split = blockIdx.y, b = blockIdx.z;
    const int p0 = split * pts_per_split;
    const int p1 = min(s.n, p0 + pts_per_split);
    const int npts = max(0, p1 - p0);
    const int nchunks = (npts + SCORE_CHUNK - 1) / SCORE_CHUNK;
    const float4* src = (PACKED ? s.corr_dup + (size_t)b * s.n_stride * 2 + (size_t)p0 * 2
                                : s.corr + (size_t)b * s.n_stride + p0);

    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int c) {
        int cnt = min(SCORE_CHUNK, npts - c * SCORE_CHUNK);
        uint32_t bytes = (uint32_t)cnt * 16u * F4_PER_PT;
        mbar_expect_tx(&full[c & 1], bytes);
        tma_load_1d(&buf[c & 1][0], src + (size_t)c * SCORE_CHUNK * F4_PER_PT, bytes, &full[c & 1]);
    };
    if (tid == 0 && nchunks > 0) issue(0);

    // Essential matrices of this thread's hypotheses -> registers.
    const float* Eb = s.Ecand + (size_t)b * 9 * s.h_stride;
    float e[HPT][9];
    int hl[HPT];
#pragma unroll
    for (int j = 0; j < HPT; j++) {
        hl[j] = tile * HPC + j * SCORE_THREADS + tid;
        bool valid = hl[j] < H;
#pragma unroll
        for (int k = 0; k < 9; k++) e[j][k] = valid ? __ldg(Eb + (size_t)k * s.h_stride + hl[j]) : 0.0f;
    }
    unsigned int cnt[HPT];
#pragma unroll
    for (int j = 0; j < HPT; j++) cnt[j] = 0u;
    const float nthr = -thr;

    if constexpr (PACKED) {
        static_assert(!PACKED || HPT % 2 == 0, "packed path works on hypothesis pairs");
        float2 e2[HPT / 2][9];
#pragma unroll
        for (int j = 0; j < HPT / 2; j++)
#pragma unroll
            for (int k = 0; k < 9; k++) e2[j][k] = make_float2(e[2 * j][k], e[2 * j + 1][k]);
        const float2 nthr2 = make_float2(nthr, nthr);
        for (int c = 0; c < nchunks; c++) {
            if (tid == 0 && c + 1 < nchunks) issue(c + 1);
            mbar_wait(&full[c & 1], (c >> 1) & 1);
            const float4* pb = buf[c & 1];
            const int n_here = min(SCORE_CHUNK, npts - c * SCORE_CHUNK);
#pragma unroll 4
            for (int i = 0; i < n_here; i++) {
                float4 a = pb[2 * i], q = pb[2 * i + 1];
                float2 x1 = make_float2(a.x, a.y), y1 = make_float2(a.z, a.w);
                float2 x2 = make_float2(q.x, q.y), y2 = make_float2(q.z, q.w);
#pragma unroll
                for (int j = 0; j < HPT / 2; j++) {
                    float2 d = sampson_d2(e2[j], x1, y1, x2, y2, nthr2);
                    cnt[2 * j] += __float_as_uint(d.x) >> 31;
                    cnt[2 * j + 1] += __float_as_uint(d.y) >> 31;
                }
            }
            __syncthreads();
        }
    } else {
        for (int c = 0; c < nchunks; c++) {
            if (tid == 0 && c + 1 < nchunks) issue(c + 1);
            mbar_wait(&full[c & 1], (c >> 1) & 1);
            const float4* pb = buf[c & 1];
            const int n_here = min(SCORE_CHUNK, npts - c * SCORE_CHUNK);
#pragma unroll 4
            for (int i = 0; i < n_here; i++) {
                float4 p = pb[i];
#pragma unroll
                for (int j = 0; j < HPT; j++) {
                    float d = sampson_d(e[j], p.x, p.y, p.z, p.w, nthr);
                    cnt[j] += __float_as_uint(d) >> 31;
                }
            }
            __syncthreads();
        }
    }

    // ---- epilogue: counts + fused arg-max ----
    int* counts = s.counts + (size_t)b * s.h_stride;
    unsigned long long key = 0ull;
    if (splits == 1) {
#pragma unroll
        for (int j = 0; j < HPT; j++)
            if (hl[j] < H) {
                counts[hl[j]] = (int)cnt[j];
                unsigned long long k =
                    ((unsigned long long)cnt[j] << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)(h_offset + hl[j]));
                key = k > key ? k : key;
            }
    } else {
#pragma unroll
        for (int j = 0; j < HPT; j++)
            if (hl[j] < H && cnt[j] != 0u) atomicAdd(&counts[hl[j]], (int)cnt[j]);
        __threadfence();
        __syncthreads();
        if (tid == 0) s_ticket = atomicAdd(&s.tile_done[(size_t)b * s.tiles_max + tile], 1);
        __syncthreads();
        if (s_ticket != splits - 1) return;
        __threadfence();
#pragma unroll
        for (int j = 0; j < HPT; j++)
            if (hl[j] < H) {
                unsigned int total = (unsigned int)__ldcg(&counts[hl[j]]);
                unsigned long long k =
                    ((unsigned long long)total << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)(h_offset + hl[j]));
                key = k > key ? k : key;
            }
    }
    key = warp_max_u64(key);
    if ((tid & 31) == 0) red[tid >> 5] = key;
    __syncthreads();
    if (tid < 32) {
        unsigned long long v = tid < SCORE_THREADS / 32 ? red[tid] : 0ull;
        v = warp_max_u64(v);
        if (tid == 0 && v != 0ull) atomicMax(&s.best[b], v);
    }
}

// Kernel family.  Register-bank pressure is what bounds the FP32 pipe here: an
// FFMA reads three registers and an FFMA2 three register PAIRS per issue, and
// the only operand that can come from the operand-reuse cache is the point
// coordinate shared by the HPT hypotheses of a thread, so more hypotheses per
// thread = more reuse, at the price of registers (occupancy) and coarser tiles.
struct ScoreVariant { int hpt; int packed; int threads; };
static const ScoreVariant kVariants[] = {
    {2, 0, 256}, {4, 1, 256}, {4, 0, 256}, {8, 0, 256}, {8, 1, 256}, {4, 0, 128}, {4, 1, 128}, {8, 0, 128}, {8, 1, 128},
};
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);
constexpr int kDefaultVariant = 1;

template <int HPT, bool PACKED, int THREADS>
static void launch_one(const DeviceState& s, dim3 grid, int H, int h_offset, int splits, int pps, float thr, cudaStream_t st) {
    score_kernel<HPT, PACKED, THREADS><<<grid, THREADS, 0, st>>>(s, H, h_offset, splits, pps, thr);
}
template <int HPT, bool PACKED, int THREADS>
static int occupancy_one() {
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, score_kernel<HPT, PACKED, THREADS>, THREADS, 0);
    return n > 0 ? n : 1;
}
#define SFM_FOR_VARIANT(v, CALL)                  \
    switch (v) {                                  \
        case 0: CALL(2, false, 256); break;       \
        case 1: CALL(4, true, 256); break;        \
        case 2: CALL(4, false, 256); break;       \
        case 3: CALL(8, false, 256); break;       \
        case 4: CALL(8, true, 256); break;        \
        case 5: CALL(4, false, 128); break;       \
        case 6: CALL(4, true, 128); break;        \
        case 7: CALL(8, false, 128); break;       \
        default: CALL(8, true, 128); break;       \
    }

static int variant_occupancy(int v) {
    static int cache[kNumVariants] = {0};
    if (cache[v] == 0) {
#define SFM_OCC(h, p, t) cache[v] = occupancy_one<h, p, t>()
        SFM_FOR_VARIANT(v, SFM_OCC)
#undef SFM_OCC
    }
    return cache[v];
}

int score_num_variants() { return kNumVariants; }

ScorePlan make_score_plan(int B, int n, int H, int variant_override) {
    ScorePlan p;
    p.variant = (variant_override >= 0 && variant_override < kNumVariants) ? variant_override : kDefaultVariant;
    const ScoreVariant& v = kVariants[p.variant];
    p.hyp_per_cta = v.hpt * v.threads;
    p.tiles = (H + p.hyp_per_cta - 1) / p.hyp_per_cta;
    if (p.tiles < 1) p.tiles = 1;
    int sms = 148;
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const long long slots = (long long)sms * variant_occupancy(p.variant);
    int chunks = (n + SCORE_CHUNK - 1) / SCORE_CHUNK;
    if (chunks < 1) chunks = 1;
    // Pick the number of point splits that minimises (waves of resident CTAs) x
    // (TMA stages per CTA + a fixed per-CTA cost): the grid's tail wave is what
    // costs most at config-2 sizes.  Ties go to fewer splits (fewer atomics).
    const long long base = (long long)p.tiles * B;
    double best_cost = 1e300;
    int best_s = 1;
    for (int sp = 1; sp <= chunks; sp++) {
        int cps = (chunks + sp - 1) / sp;
        int real = (chunks + cps - 1) / cps;
        if (real != sp) continue;
        long long ctas = base * sp;
        long long waves = (ctas + slots - 1) / slots;
        double cost = (double)waves * (cps + 0.15);
        if (cost < best_cost * 0.999) { best_cost = cost; best_s = sp; }
        if (ctas > slots * 64) break;
    }
    int cps = (chunks + best_s - 1) / best_s;
    p.pts_per_split = cps * SCORE_CHUNK;
    p.splits = (chunks + cps - 1) / cps;
    return p;
}

void launch_score(const DeviceState& s, const ScorePlan& plan, int H, int h_offset, float thr, cudaStream_t st) {
    dim3 grid(plan.tiles, plan.splits, s.B);
#define SFM_LAUNCH(h, p, t) launch_one<h, p, t>(s, grid, H, h_offset, plan.splits, plan.pts_per_split, thr, st)
    SFM_FOR_VARIANT(plan.variant, SFM_LAUNCH)
#undef SFM_LAUNCH
}

// Selected hypothesis -> E, index, count (per pair).
__global__ void select_kernel(DeviceState s, int h_offset) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= s.B) return;
    unsigned long long packed = s.best[b];
    unsigned int hg = 0xFFFFFFFFu - (unsigned int)(packed & 0xFFFFFFFFull);
    int local = (int)hg - h_offset;
    s.best_idx[b] = (int)hg;
    s.best_count[b] = (int)(packed >> 32);
    const float* Eb = s.Ecand + (size_t)b * 9 * s.h_stride;
#pragma unroll
    for (int k = 0; k < 9; k++)
        s.E[(size_t)b * 9 + k] = (local >= 0 && local < s.h_stride) ? Eb[(size_t)k * s.h_stride + local] : 0.0f;
}

void launch_select(const DeviceState& s, int h_offset, cudaStream_t st) {
    select_kernel<<<(s.B + 127) / 128, 128, 0, st>>>(s, h_offset);
}

// ---------------------------------------------------------------------------
// FP32-pipe probe: the roofline denominator for hypothesis generation and
// scoring is not in MEASURED_PEAKS.json (it has HBM and bf16 only), so bench.py
// measures it live: dependent-chain-free FFMA / FFMA2 streams, all SMs busy.
// mode 0: scalar FFMA with three distinct register operands
// mode 1: packed FFMA2
// ---------------------------------------------------------------------------
constexpr int PROBE_THREADS = 256;
constexpr int PROBE_CHAINS = 16;
constexpr int PROBE_INNER = 64;

template <int MODE>
__global__ void __launch_bounds__(PROBE_THREADS) fma_probe_kernel(int iters, float* sink, float a0, float b0) {
    if constexpr (MODE == 0) {
        float acc[PROBE_CHAINS];
        float a = a0 + threadIdx.x * 1e-9f, b = b0;
#pragma unroll
        for (int i = 0; i < PROBE_CHAINS; i++) acc[i] = (float)i;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int r = 0; r < PROBE_INNER; r++)
#pragma unroll
                for (int i = 0; i < PROBE_CHAINS; i++) acc[i] = fmaf(acc[i], a, b);
        }
        float t = 0.0f;
#pragma unroll
        for (int i = 0; i < PROBE_CHAINS; i++) t += acc[i];
        if (t == 123.456f) sink[0] = t;
    } else {
        float2 acc[PROBE_CHAINS];
        float2 a = make_float2(a0 + threadIdx.x * 1e-9f, a0), b = make_float2(b0, b0 * 0.5f);
#pragma unroll
        for (int i = 0; i < PROBE_CHAINS; i++) acc[i] = make_float2((float)i, (float)-i);
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int r = 0; r < PROBE_INNER; r++)
#pragma unroll
                for (int i = 0; i < PROBE_CHAINS; i++) acc[i] = __ffma2_rn(acc[i], a, b);
        }
        float t = 0.0f;
#pragma unroll
        for (int i = 0; i < PROBE_CHAINS; i++) t += acc[i].x + acc[i].y;
        if (t == 123.456f) sink[0] = t;
    }
}

double launch_fma_probe(int mode, int iters, cudaStream_t st, float* d_sink) {
    const int ctas = 148 * 8;
    if (mode == 0)
        fma_probe_kernel<0><<<ctas, PROBE_THREADS, 0, st>>>(iters, d_sink, 0.999f, 0.001f);
    else
        fma_probe_kernel<1><<<ctas, PROBE_THREADS, 0, st>>>(iters, d_sink, 0.999f, 0.001f);
    double per_thread = (double)iters * PROBE_INNER * PROBE_CHAINS * (mode == 0 ? 1.0 : 2.0);
    return per_thread * PROBE_THREADS * ctas;
}

}  // namespace sfmb200
