// Prototype: scoring with the point batch in __constant__ memory (operands come from uniform registers).
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
__constant__ float4 cpts[4000];
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
template <int HPT, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k(const float* __restrict__ E, int* counts, int npts, int hs, float thr) {
    float e[HPT][9]; unsigned cnt[HPT];
    int tid = blockIdx.x * THREADS * HPT + threadIdx.x;
#pragma unroll
    for (int j = 0; j < HPT; j++) { cnt[j] = 0;
#pragma unroll
        for (int q = 0; q < 9; q++) e[j][q] = E[(size_t)q * hs + tid + j * THREADS]; }
#pragma unroll 4
    for (int i = 0; i < npts; i++) {
        float4 p = cpts[i];
#pragma unroll
        for (int j = 0; j < HPT; j++) { float d = sampson_d(e[j], p.x, p.y, p.z, p.w, -thr); cnt[j] += __float_as_uint(d) >> 31; }
    }
#pragma unroll
    for (int j = 0; j < HPT; j++) atomicAdd(&counts[tid + j * THREADS], (int)cnt[j]);
}
template <int HPT, int THREADS, int MINB>
void run(const char* name, float* dE, int* dC, int hs, int npts) {
    int ctas = hs / (HPT * THREADS);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int w = 0; w < 2; w++) k<HPT, THREADS, MINB><<<ctas, THREADS>>>(dE, dC, npts, hs, 1e-6f);
    cudaEventRecord(a);
    const int reps = 5;
    for (int r = 0; r < reps; r++) k<HPT, THREADS, MINB><<<ctas, THREADS>>>(dE, dC, npts, hs, 1e-6f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= reps;
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<HPT, THREADS, MINB>, THREADS, 0);
    printf("%-22s ctas %5d occ %d  %.3f ms  %.3e evals/s  %.1f TFLOP/s(34)  err=%s\n", name, ctas, occ, ms, (double)hs * npts / (ms * 1e-3),
           34.0 * hs * npts / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    const int npts = 4000;
    const int hs = 148 * 2048 * 4;   // multiple of every tile size used below
    std::vector<float4> pts(npts);
    for (int i = 0; i < npts; i++) pts[i] = make_float4(0.1f * sinf(i), 0.1f * cosf(i), 0.1f * sinf(2 * i), 0.1f * cosf(3 * i));
    cudaMemcpyToSymbol(cpts, pts.data(), sizeof(float4) * npts);
    std::vector<float> E((size_t)9 * hs);
    for (size_t i = 0; i < E.size(); i++) E[i] = sinf(0.37f * i);
    float* dE; int* dC;
    cudaMalloc(&dE, E.size() * 4); cudaMalloc(&dC, hs * 4); cudaMemset(dC, 0, hs * 4);
    cudaMemcpy(dE, E.data(), E.size() * 4, cudaMemcpyHostToDevice);
    run<8, 256, 1>("HPT8 T256 minb1", dE, dC, hs, npts);
    run<8, 256, 2>("HPT8 T256 minb2", dE, dC, hs, npts);
    run<8, 128, 4>("HPT8 T128 minb4", dE, dC, hs, npts);
    run<4, 256, 2>("HPT4 T256 minb2", dE, dC, hs, npts);
    run<4, 256, 3>("HPT4 T256 minb3", dE, dC, hs, npts);
    run<4, 128, 4>("HPT4 T128 minb4", dE, dC, hs, npts);
    run<2, 256, 4>("HPT2 T256 minb4", dE, dC, hs, npts);
    run<16, 128, 2>("HPT16 T128 minb2", dE, dC, hs, npts);
    run<16, 256, 1>("HPT16 T256 minb1", dE, dC, hs, npts);
    return 0;
}
