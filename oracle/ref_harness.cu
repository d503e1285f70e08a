// Reference-oracle harness (TEST INFRASTRUCTURE ONLY - never linked into the
// product).  A unity translation unit that #includes the UNMODIFIED reference
// source /root/reference/SfM/sfm.cu where it lies (path given by -I at build
// time, see oracle/Makefile) and exposes its stages through a small C ABI so
// tests can (1) run the as-built pipeline for timing and (2) inject identical
// inputs stage by stage for parity.
//
// Three preprocessor stand-ins, no source edits:
//   * `class` -> `struct` while the reference headers are parsed, so that
//     Image_pair's private members (sfm.h:21-41) are reachable;
//   * std::random_device -> a fixed-seed functor, so estimateE's shuffle
//     (sfm.cu:102-104) is reproducible.
//   * cudaMalloc -> the same call with padding behind the buffer.  linear_triangulation() hands cuSOLVER's batched SVD
//     of num_points matrices an info buffer of 4 ints (sfm.cu:326-328; right for choosePose's batch of 4, reused for
//     the batch of N): 4*N bytes are written behind a 16-byte allocation.  Whether that faults depends on what the
//     allocator has mapped behind it - on the B200 boxes of round 2 the unpadded build died with "illegal memory access"
//     (reported at sfm.cu:330) in up to 6 of 6 processes.  With every allocation of the reference padded by
//     ref_malloc_pad bytes (4 * num_points + 4 KB, set in ref_create) the overflow lands in the reference's own slack
//     and the as-built path runs deterministically; its arithmetic and its timing (21 allocations per call) are untouched.
// Every system / CUDA / Thrust header the reference uses is included first so
// the stand-ins never touch them.
#include <assert.h>
#include <cublas_v2.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include <curand.h>
#include <cusolverDn.h>
#include <thrust/device_vector.h>
#include <thrust/extrema.h>
#include <thrust/host_vector.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

namespace std {
struct sfm_fixed_random_device {
    static unsigned& seed() { static unsigned s = 12345u; return s; }
    unsigned operator()() { return seed(); }
};
}  // namespace std
static size_t ref_malloc_pad = 64 * 1024;
template <class T>
static inline cudaError_t sfm_padded_cudaMalloc(T** p, size_t bytes) {
    return cudaMalloc(p, bytes + ref_malloc_pad);
}
#define random_device sfm_fixed_random_device
#define class struct
#define cudaMalloc sfm_padded_cudaMalloc
#include "SfM/sfm.cu"
#undef cudaMalloc
#undef class
#undef random_device

using SfM::Image_pair;

extern "C" {

void* ref_create(const float K[9], const float Kinv[9], int n) {
    float k[9], ki[9];
    memcpy(k, K, sizeof(k));
    memcpy(ki, Kinv, sizeof(ki));
    ref_malloc_pad = (size_t)4 * (size_t)(n > 0 ? n : 0) + 4096;      // see the third stand-in above
    return new Image_pair(k, ki, 2, n);
}
void ref_destroy(void* p) { delete (Image_pair*)p; }
void ref_set_seed(unsigned s) { std::sfm_fixed_random_device::seed() = s; }

// fillXU from host pixel correspondences (u1,v1,u2,v2)[n]: builds the SiftPoint
// array the reference expects (only xpos, ypos, match_xpos, match_ypos are read).
int ref_fillXU(void* p, const float* h_px) {
    Image_pair* ip = (Image_pair*)p;
    int n = ip->num_points;
    std::vector<SiftPoint> pts(n);
    memset(pts.data(), 0, sizeof(SiftPoint) * n);
    for (int i = 0; i < n; i++) {
        pts[i].xpos = h_px[4 * i + 0];
        pts[i].ypos = h_px[4 * i + 1];
        pts[i].match_xpos = h_px[4 * i + 2];
        pts[i].match_ypos = h_px[4 * i + 3];
    }
    SiftPoint* d = nullptr;
    cudaMalloc(&d, sizeof(SiftPoint) * n);
    cudaMemcpy(d, pts.data(), sizeof(SiftPoint) * n, cudaMemcpyHostToDevice);
    ip->fillXU(d);
    cudaDeviceSynchronize();
    cudaFree(d);
    return (int)cudaGetLastError();
}
// X[image] (3 x n row-major) to host.
int ref_get_X(void* p, int image, float* h_X) {
    Image_pair* ip = (Image_pair*)p;
    cudaMemcpy(h_X, ip->X[image], sizeof(float) * 3 * ip->num_points, cudaMemcpyDeviceToHost);
    return (int)cudaGetLastError();
}
int ref_set_X(void* p, int image, const float* h_X) {
    Image_pair* ip = (Image_pair*)p;
    cudaMemcpy(ip->X[image], h_X, sizeof(float) * 3 * ip->num_points, cudaMemcpyHostToDevice);
    return (int)cudaGetLastError();
}

// Stage injection: the reference's own K3 -> K4/K5 -> K6 -> K7 chain
// (sfm.cu:107-129) on caller-supplied sample rows idx[H][8].
// out: h_E [H][9]; optional h_A [H][72] design matrices; optional h_V [H][81].
int ref_e_candidates(void* p, const int* h_idx, int H, float* h_E, float* h_A, float* h_V) {
    Image_pair* ip = (Image_pair*)p;
    int* d_indices;
    cudaMalloc((void**)&d_indices, sizeof(int) * 8 * (H + 1));
    cudaMemset(d_indices, 0, sizeof(int) * 8 * (H + 1));      // row H: the reference's off-by-one thread (kernels.h:242)
    cudaMemcpy(d_indices, h_idx, sizeof(int) * 8 * H, cudaMemcpyHostToDevice);
    float *d_A, *d_E_candidate, *d_ut, *d_vt, *d_s;
    cudaMalloc((void**)&d_A, sizeof(float) * 72 * (H + 1));   // + the one matrix written out of bounds
    cudaMalloc((void**)&d_E_candidate, sizeof(float) * 9 * H);
    cudaMalloc((void**)&d_ut, sizeof(float) * 64 * H);
    cudaMalloc((void**)&d_vt, sizeof(float) * 81 * H);
    cudaMalloc((void**)&d_s, sizeof(float) * 8 * H);
    int* d_info = NULL;
    cudaMalloc((void**)&d_info, sizeof(int) * (H > 4 ? H : 4));
    int grids = ceil((H + cuda_block_size - 1) / cuda_block_size);
    kernels::kernels<<<grids, cuda_block_size>>>(ip->X[0], ip->X[1], d_A, d_indices, H, ip->num_points);
    kernels::regular_svd(d_A, d_ut, d_s, d_vt, 8, 9, H, d_info, ip->cusolverH, ip->gesvdj_params);
    kernels::row_extraction_kernel<<<grids, cuda_block_size>>>(d_vt, d_E_candidate, H);
    kernels::normalizeE<<<grids, cuda_block_size>>>(d_E_candidate, H);
    cudaDeviceSynchronize();
    if (h_E) cudaMemcpy(h_E, d_E_candidate, sizeof(float) * 9 * H, cudaMemcpyDeviceToHost);
    if (h_A) cudaMemcpy(h_A, d_A, sizeof(float) * 72 * H, cudaMemcpyDeviceToHost);
    if (h_V) cudaMemcpy(h_V, d_vt, sizeof(float) * 81 * H, cudaMemcpyDeviceToHost);
    cudaFree(d_A); cudaFree(d_E_candidate); cudaFree(d_ut); cudaFree(d_vt); cudaFree(d_s); cudaFree(d_info);
    cudaFree(d_indices);
    return (int)cudaGetLastError();
}

// The reference's estimateE body (sfm.cu:107-151) with H and the sample rows
// injected instead of H = N/8 from the host shuffle; everything else - the
// kernels, cuSOLVER / cuBLAS calls, per-call cudaMalloc/cudaFree, thrust
// arg-max with its off-by-one - is the reference's own code.  Timing baseline.
// Returns the wall time in ms; out_best = the index the reference selects.
float ref_estimateE_injected(void* p, const int* h_idx, int H, int* out_best) {
    Image_pair* ip = (Image_pair*)p;
    cudaDeviceSynchronize();
    auto t0 = std::chrono::high_resolution_clock::now();
    const int ransac_count = H;
    int* d_indices;
    cudaMalloc((void**)&d_indices, sizeof(int) * 8 * (H + 1));
    cudaMemset(d_indices, 0, sizeof(int) * 8 * (H + 1));      // row H: the reference's off-by-one thread (kernels.h:242)
    cudaMemcpy(d_indices, h_idx, sizeof(int) * 8 * H, cudaMemcpyHostToDevice);
    float* d_A;
    cudaMalloc((void**)&d_A, 8 * 9 * (size_t)(ransac_count + 1) * sizeof(float));
    int grids = ceil((ransac_count + cuda_block_size - 1) / cuda_block_size);
    kernels::kernels<<<grids, cuda_block_size>>>(ip->X[0], ip->X[1], d_A, d_indices, ransac_count, ip->num_points);
    float* d_E_candidate;
    cudaMalloc((void**)&d_E_candidate, 3 * 3 * (size_t)ransac_count * sizeof(float));
    float *d_ut, *d_vt, *d_s;
    cudaMalloc((void**)&d_ut, 8 * 8 * (size_t)ransac_count * sizeof(float));
    cudaMalloc((void**)&d_vt, 9 * 9 * (size_t)ransac_count * sizeof(float));
    cudaMalloc((void**)&d_s, 8 * (size_t)ransac_count * sizeof(float));
    int* d_info = NULL;
    cudaMalloc((void**)&d_info, sizeof(int) * (H > 4 ? H : 4));
    kernels::regular_svd(d_A, d_ut, d_s, d_vt, 8, 9, ransac_count, d_info, ip->cusolverH, ip->gesvdj_params);
    int blocks = ceil((ransac_count + cuda_block_size - 1) / cuda_block_size);
    kernels::row_extraction_kernel<<<blocks, cuda_block_size>>>(d_vt, d_E_candidate, ransac_count);
    kernels::normalizeE<<<grids, cuda_block_size>>>(d_E_candidate, ransac_count);
    int* d_inliers = ip->calculateInliers(d_E_candidate, ransac_count);
    thrust::device_ptr<int> dv_in(d_inliers);
    auto iter = thrust::max_element(dv_in, dv_in + ransac_count);
    int best_pos = (iter - dv_in) - 1;
    if (best_pos < 0) best_pos = 0;   // the reference would read before the buffer
    cudaMemcpy(ip->d_E, &(d_E_candidate[9 * best_pos]), 3 * 3 * sizeof(float), cudaMemcpyDeviceToDevice);
    cudaFree(d_A); cudaFree(d_ut); cudaFree(d_s); cudaFree(d_vt); cudaFree(d_info); cudaFree(d_indices);
    cudaFree(d_inliers); cudaFree(d_E_candidate);
    cudaDeviceSynchronize();
    auto t1 = std::chrono::high_resolution_clock::now();
    if (out_best) *out_best = best_pos;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { fprintf(stderr, "ref_estimateE_injected: %s\n", cudaGetErrorString(e)); return -1.0f; }
    return std::chrono::duration<float, std::milli>(t1 - t0).count();
}

// As-built stages, each returning wall ms (device-synchronised both sides).
#define TIMED(call)                                                              \
    cudaDeviceSynchronize();                                                     \
    auto t0 = std::chrono::high_resolution_clock::now();                         \
    call;                                                                        \
    cudaDeviceSynchronize();                                                     \
    auto t1 = std::chrono::high_resolution_clock::now();                         \
    return std::chrono::duration<float, std::milli>(t1 - t0).count();
float ref_estimateE(void* p) { TIMED(((Image_pair*)p)->estimateE()) }
float ref_computePosecandidates(void* p) { TIMED(((Image_pair*)p)->computePosecandidates()) }
float ref_choosePose(void* p) { TIMED(((Image_pair*)p)->choosePose()) }
float ref_linear_triangulation(void* p) { TIMED(((Image_pair*)p)->linear_triangulation()) }

int ref_set_E(void* p, const float* h_E) {
    cudaMemcpy(((Image_pair*)p)->d_E, h_E, sizeof(float) * 9, cudaMemcpyHostToDevice);
    return (int)cudaGetLastError();
}
int ref_get_E(void* p, float* h_E) {
    cudaMemcpy(h_E, ((Image_pair*)p)->d_E, sizeof(float) * 9, cudaMemcpyDeviceToHost);
    return (int)cudaGetLastError();
}
int ref_get_P(void* p, float* h_P) {
    cudaMemcpy(h_P, ((Image_pair*)p)->d_P, sizeof(float) * 64, cudaMemcpyDeviceToHost);
    return (int)cudaGetLastError();
}
int ref_set_P(void* p, const float* h_P) {
    cudaMemcpy(((Image_pair*)p)->d_P, h_P, sizeof(float) * 64, cudaMemcpyHostToDevice);
    return (int)cudaGetLastError();
}
int ref_get_P_ind(void* p) { return ((Image_pair*)p)->P_ind; }
int ref_get_points(void* p, float* h_points) {
    Image_pair* ip = (Image_pair*)p;
    cudaMemcpy(h_points, ip->d_final_points, sizeof(float) * 4 * ip->num_points, cudaMemcpyDeviceToHost);
    return (int)cudaGetLastError();
}
// copyBoidsToVBO into plain device buffers (no GL needed), copied back.
int ref_vbo(void* p, float* h_pos, float* h_col) {
    Image_pair* ip = (Image_pair*)p;
    size_t bytes = sizeof(float) * 4 * ip->num_points;
    float *d_pos, *d_col;
    cudaMalloc(&d_pos, bytes);
    cudaMalloc(&d_col, bytes);
    ip->copyBoidsToVBO(d_pos, d_col);
    cudaMemcpy(h_pos, d_pos, bytes, cudaMemcpyDeviceToHost);
    cudaMemcpy(h_col, d_col, bytes, cudaMemcpyDeviceToHost);
    cudaFree(d_pos);
    cudaFree(d_col);
    return (int)cudaGetLastError();
}
// host svd() / det() of svd.h, for CPU-side pinning of the 3x3 contract.
void ref_host_svd(const float* a, float* u, float* s, float* v) { svd(a, u, s, v); }
float ref_host_det(const float* a) { return det(a); }

}  // extern "C"
