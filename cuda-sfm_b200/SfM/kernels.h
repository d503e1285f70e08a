// Source-compatible replacement of the host-callable part of the reference's
// SfM/kernels.h (namespace kernels): the cuBLAS / cuSOLVER / Thrust wrappers as
// thin calls into libsfmb200 (include/sfmb200_la.h), cuda_alloc_copy, the debug
// printers and timer().  Library handles are accepted and ignored (template
// parameters, so neither cuBLAS nor cuSOLVER headers are needed).  Like the
// reference's file this header DEFINES non-inline-looking symbols; unlike it,
// everything is `inline`, so it may be included from several translation units
// (SURVEY Q3).
//
// The reference's __global__ kernels of this header (copy_point, kernels,
// normalizeE, threshold_count, ...) are implementation details of its sfm.cu;
// their B200 equivalents live inside libsfmb200 and are reached through
// SfM::Image_pair / the C ABI.
#pragma once

#include <cassert>
#include <iostream>
#include <string>

#include "../../include/sfmb200_la.h"
#include "common.h"
#include "sfm.h"
#include "svd.h"

namespace kernels {
#ifndef enable_debug
#define enable_debug false
#endif

template <typename T>
void printVector(const T* a1, int n, std::string name) {
    if (!enable_debug) return;
    T* host = new T[n];
    cudaMemcpy(host, a1, n * sizeof(T), cudaMemcpyDeviceToHost);
    std::cout << name << "\n{\n";
    for (int i = 0; i < n; i++) std::cout << "\t" << host[i] << "\n";
    std::cout << "}" << std::endl;
    delete[] host;
}
template <typename T>
void printMatrix(const T* A, int row, int col, int print_col, const char* name) {
    if (!enable_debug) return;
    T* host = new T[(size_t)row * col];
    cudaMemcpy(host, A, (size_t)row * col * sizeof(T), cudaMemcpyDeviceToHost);
    std::cout << name << "\n{\n";
    for (int i = 0; i < row; i++) {
        for (int j = 0; j < col; j++) {
            if (j < print_col || j > col - print_col - 1) std::cout << "\t" << host[access2(i, j, col)];
            else if (j == print_col) std::cout << "\t....";
        }
        std::cout << "\n";
    }
    std::cout << "}" << std::endl;
    delete[] host;
}
template <typename T>
void print3DSlice(const T* A, int row, int col, int slice, int print_col, const char* name) {
    printMatrix(A + (size_t)row * col * slice, row, col, print_col, name);
}

inline void la_check(int rc, const char* what) {
    if (rc == 0) return;
    fprintf(stderr, "%s failed (%d)%s\n", what, rc, rc == -5 ? ": matrix is singular" : "");
    exit(EXIT_FAILURE);      // the reference exits on its failures too (kernels.h:144-161)
}

// C(m,n) = A(m,k) * B(k,n)
template <class Handle>
void gpu_blas_mmul(const float* A, const float* B, float* C, const int m, const int k, const int n, Handle) {
    la_check(sfmb200_la_mmul(A, B, C, m, k, n, nullptr), "gpu_blas_mmul");
}
template <class Handle>
void gpu_blas_mmul_batched(const float* A, const float* B, float* C, const int m, const int k, const int n,
                           const int stride_A, const int stride_B, const int stride_C, const int batches, Handle) {
    assert(stride_A == 0 || stride_A == m * k);
    assert(stride_B == 0 || stride_B == n * k);
    assert(stride_C == 0 || stride_C == m * n);
    la_check(sfmb200_la_mmul_batched(A, B, C, m, k, n, stride_A, stride_B, stride_C, batches, nullptr), "gpu_blas_mmul_batched");
}
template <class Handle>
void gpu_blas_mmul_transpose_batched(const float* A, const float* B, float* C, const int m, const int k, const int n,
                                     const int stride_A, const int stride_B, const int stride_C, const int batches, Handle) {
    la_check(sfmb200_la_mmul_transpose_batched(A, B, C, m, k, n, stride_A, stride_B, stride_C, batches, nullptr),
             "gpu_blas_mmul_transpose_batched");
}
template <class Handle>
void invert_device(float* src, float* dst, int n, int batchSize, Handle) {
    la_check(sfmb200_la_invert(src, dst, n, batchSize, nullptr), "invert");
}
template <class Handle>
void invert(float* s, float* d, int n, int batch, Handle h) {
    invert_device(s, d, n, batch, h);     // in-place safe: the kernel reads the whole matrix first
}
// cusolverDnSgesvdjBatched semantics: src column-major m x n; outputs in the
// reference's (confusing) argument order: its `VT` receives U, its `U` receives V.
template <class SolverHandle, class Stream, class Params>
void svd_square(float* src, float* VT, float* S, float* U, int m, int n, const int batchSize, int* /*d_info*/, SolverHandle,
                Stream, Params) {
    assert(m == n);
    la_check(sfmb200_la_svd_batched(src, S, VT, U, m, n, batchSize, nullptr), "svd_square");
    cudaDeviceSynchronize();
}
// row-major 8 x 9 matrices in, like the reference (it transposes to column-major first)
template <class SolverHandle, class Params>
void regular_svd(float* src, float* UT, float* S, float* VT, int m, int n, const int batchSize, int* /*d_info*/, SolverHandle,
                 Params) {
    float* colmajor = nullptr;
    cudaMalloc((void**)&colmajor, sizeof(float) * m * n * batchSize);
    la_check(sfmb200_la_transpose_batched(src, colmajor, m, n, batchSize, nullptr), "regular_svd");   // one launch, not one per matrix
    la_check(sfmb200_la_svd_batched(colmajor, S, UT, VT, m, n, batchSize, nullptr), "regular_svd");
    cudaDeviceSynchronize();
    cudaFree(colmajor);
}
template <typename T>
T* cuda_alloc_copy(const T* host, int size) {
    T* data;
    cudaMalloc((void**)&data, size * sizeof(T));
    cudaMemcpy(data, host, size * sizeof(T), cudaMemcpyHostToDevice);
    return data;
}
// host-callable forms of the element-wise helpers (the reference exposes them as kernels)
inline void element_wise_mult(float* A, float* B, int size) { la_check(sfmb200_la_elementwise(0, A, B, size, nullptr), "element_wise_mult"); }
inline void element_wise_div(float* A, float* B, int size) { la_check(sfmb200_la_elementwise(1, A, B, size, nullptr), "element_wise_div"); }
inline void element_wise_sum(float* A, float* B, int size) { la_check(sfmb200_la_elementwise(2, A, B, size, nullptr), "element_wise_sum"); }
inline void vecnorm(float* A, float* res, int row, int col, float exp, float final_pow) {
    la_check(sfmb200_la_vecnorm(A, res, row, col, exp, final_pow, nullptr), "vecnorm");
}
inline void threshold_count(float* A, int* count_res, int batch_size, int ransac_count, float threshold) {
    la_check(sfmb200_la_threshold_count(A, count_res, batch_size, ransac_count, threshold, nullptr), "threshold_count");
}
inline void row_extraction_kernel(float* d_vt, float* d_E, int number_points) {
    la_check(sfmb200_la_row_extraction(d_vt, d_E, number_points, nullptr), "row_extraction_kernel");
}
// host-callable forms of the reference's remaining hot-path kernels, same names and argument lists
// (kernels.h:196-209, 236-295, 357-450, 471-495); call them as functions, without <<< >>>
inline void transpose(float* odata, float* idata, int width, int height) {      // one height x width matrix
    la_check(sfmb200_la_transpose_batched(idata, odata, height, width, 1, nullptr), "transpose");
}
inline void kernels(float* d1, float* d2, float* A, int* indices, const int ransac_iterations, int num_points) {
    la_check(sfmb200_la_design_matrix(d1, d2, A, indices, ransac_iterations, num_points, nullptr), "kernels");
}
inline void copy_point(SiftPoint* data, int numPoints, float* U1, float* U2) {
    la_check(sfmb200_la_copy_point(data, numPoints, U1, U2, nullptr), "copy_point");
}
inline void normalizeE(float* E, int ransac_iterations) { la_check(sfmb200_la_normalize_E(E, ransac_iterations, nullptr), "normalizeE"); }
inline void candidate_kernels(float* d_P, const float* u, const float* v) {
    la_check(sfmb200_la_candidate_poses(d_P, u, v, nullptr), "candidate_kernels");
}
inline void compute_linear_triangulation_A(float* A, const float* pt1, const float* pt2, const int count, const int num_points,
                                           const float* m1, const float* m2, int P_ind, bool candidate_m2) {
    la_check(sfmb200_la_triangulation_A(A, pt1, pt2, count, num_points, m1, m2, P_ind, candidate_m2, nullptr),
             "compute_linear_triangulation_A");
}
inline void normalize_pt_kernal(float* v, float* converted_pt, int number_points) {
    la_check(sfmb200_la_normalize_pt(v, converted_pt, number_points, nullptr), "normalize_pt_kernal");
}
inline void kernCopyPositionsToVBO(int N, float* pos, float* vbo, float s_scale) {
    la_check(sfmb200_la_copy_to_vbo(N, pos, vbo, s_scale, nullptr), "kernCopyPositionsToVBO");
}
inline void kernCopyVelocitiesToVBO(int N, float* vbo, float s_scale) {
    (void)s_scale;
    la_check(sfmb200_la_copy_to_vbo(N, nullptr, vbo, 1.0f, nullptr), "kernCopyVelocitiesToVBO");
}
inline int max_element_index(const int* d_v, int n) {      // thrust::max_element(dv, dv + n) - dv
    int32_t idx = 0;
    la_check(sfmb200_la_argmax_first(d_v, n, &idx, nullptr), "max_element");
    return idx;
}
}  // namespace kernels

// ---- SfM::Image_pair self tests (declared in sfm.h) with the reference's literals ----
namespace SfM {
namespace detail {
inline bool close(const float* got, const float* want, int n, float tol) {
    for (int i = 0; i < n; i++)
        if (!(std::fabs(got[i] - want[i]) <= tol * (1.0f + std::fabs(want[i])))) return false;
    return true;
}
template <typename T>
std::vector<T> download(const T* d, int n) {
    std::vector<T> h(n);
    cudaMemcpy(h.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost);
    return h;
}
}  // namespace detail

inline bool Image_pair::testBatchedmult() {      // sfm.cu:389-423: 6 batches of (3x1)(1x3), B shared
    float A[] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 0, 1, 2, 3, 4, 5, 6, 7, 8};
    float B[] = {1, 2, 3, 4, 5, 6, 7, 8, 9};
    float* d_A = kernels::cuda_alloc_copy(A, 18);
    float* d_B = kernels::cuda_alloc_copy(B, 9);
    float* d_C;
    cudaMalloc((void**)&d_C, 6 * 9 * sizeof(float));
    // the reference's raw call is C_b(1x3) = A_b(1x1) * B(1x3) with strides 3 / 0 / 3 over 6 batches
    kernels::la_check(sfmb200_la_mmul_batched(d_A, d_B, d_C, 1, 1, 3, 3, 0, 3, 6, nullptr), "testBatchedmult");
    auto C = detail::download(d_C, 18);
    bool ok = true;
    for (int b = 0; b < 6; b++)
        for (int j = 0; j < 3; j++) ok = ok && C[3 * b + j] == A[3 * b] * B[j];
    cudaFree(d_A); cudaFree(d_B); cudaFree(d_C);
    return ok;
}
inline bool Image_pair::testSVD() {              // sfm.cu:424-441: two 4x4, reconstruct from U S V^T
    float b[32] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 1, 2, 3, 4, 5, 6, 7, 8, 10, 11, 12, 14};
    float* d_b = kernels::cuda_alloc_copy(b, 32);
    float *d_VT, *d_S, *d_U;
    cudaMalloc((void**)&d_VT, sizeof(float) * 32);
    cudaMalloc((void**)&d_U, sizeof(float) * 32);
    cudaMalloc((void**)&d_S, sizeof(float) * 8);
    kernels::svd_square(d_b, d_VT, d_S, d_U, 4, 4, 2, (int*)nullptr, 0, 0, 0);
    auto Uc = detail::download(d_VT, 32), S = detail::download(d_S, 8), Vc = detail::download(d_U, 32);
    bool ok = true;
    for (int k = 0; k < 2; k++)
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++) {          // column-major A(r,c) = b[k*16 + c*4 + r]
                float acc = 0;
                for (int i = 0; i < 4; i++) acc += Uc[k * 16 + i * 4 + r] * S[k * 4 + i] * Vc[k * 16 + i * 4 + c];
                ok = ok && std::fabs(acc - b[k * 16 + c * 4 + r]) < 1e-3f;
            }
    for (int k = 0; k < 2; k++)
        for (int i = 0; i < 3; i++) ok = ok && S[k * 4 + i] >= S[k * 4 + i + 1];
    cudaFree(d_b); cudaFree(d_VT); cudaFree(d_S); cudaFree(d_U);
    return ok;
}
inline bool Image_pair::testInverse() {          // sfm.cu:442-454: inverse of [1 2 0; 0 2 0; 1 2 1]
    float a[] = {0.9649f, 0.9572f, 0.1419f, 0.1576f, 0.4854f, 0.4218f, 0.9706f, 0.8003f, 0.9157f, 1, 2, 0, 0, 2, 0, 1, 2, 1};
    float* d_A = kernels::cuda_alloc_copy(a, 18);
    float* d_b;
    cudaMalloc((void**)&d_b, 18 * sizeof(float));
    kernels::invert(d_A + 9, d_b, 3, 1, 0);
    auto inv = detail::download(d_b, 9);
    const float want[9] = {1, -1, 0, 0, 0.5f, 0, -1, 0, 1};
    cudaFree(d_A); cudaFree(d_b);
    return detail::close(inv.data(), want, 9, 1e-5f);
}
inline bool Image_pair::testThrust_max() {       // sfm.cu:455-466: "maximum value is 6 at position 5"
    int a[] = {1, 2, 3, 4, 5, 6, 4, 1, 3};
    int* d_A = kernels::cuda_alloc_copy<int>(a, 7);
    int pos = kernels::max_element_index(d_A, 6);
    cudaFree(d_A);
    return pos == 5;
}
inline bool Image_pair::testBatchedmultTranspose() {   // sfm.cu:467-489: (stored 3x4)^T (4x3) * B(3x3), 2 batches
    float A[] = {1, 2, 3, 1, 4, 5, 6, 1, 7, 8, 9, 1, 0, 1, 2, 1, 3, 4, 5, 1, 6, 7, 8, 1};
    float B[] = {1, 2, 3, 4, 5, 6, 7, 8, 9};
    float* d_A = kernels::cuda_alloc_copy(A, 24);
    float* d_B = kernels::cuda_alloc_copy(B, 9);
    float* d_C;
    cudaMalloc((void**)&d_C, 24 * sizeof(float));
    kernels::gpu_blas_mmul_transpose_batched(d_A, d_B, d_C, 4, 3, 3, 12, 0, 12, 2, 0);
    auto C = detail::download(d_C, 24);
    bool ok = true;
    for (int b = 0; b < 2; b++)
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 3; j++) {
                float acc = 0;
                for (int l = 0; l < 3; l++) acc += A[b * 12 + l * 4 + i] * B[l * 3 + j];
                ok = ok && C[b * 12 + i * 3 + j] == acc;
            }
    cudaFree(d_A); cudaFree(d_B); cudaFree(d_C);
    return ok;
}
inline bool Image_pair::testRow_extraction_kernel() {  // sfm.cu:490-502: block b -> 81 b + 72 .. 81 b + 80
    std::vector<float> data(729);
    for (int i = 0; i < 729; i++) data[i] = (float)i;
    float* d_d = kernels::cuda_alloc_copy(data.data(), 729);
    float* res;
    cudaMalloc((void**)&res, 81 * sizeof(float));
    kernels::row_extraction_kernel(d_d, res, 9);
    auto r = detail::download(res, 81);
    bool ok = true;
    for (int b = 0; b < 9; b++)
        for (int k = 0; k < 9; k++) ok = ok && r[9 * b + k] == (float)(81 * b + 72 + k);
    cudaFree(res); cudaFree(d_d);
    return ok;
}
inline bool Image_pair::testVecnorm() {          // sfm.cu:503-510: columns of [1 2 3;4 5 6;7 8 9] -> 8.124, 9.644, 11.225
    float test[9] = {1, 2, 3, 4, 5, 6, 7, 8, 9};
    float* d = kernels::cuda_alloc_copy(test, 9);
    float* norm;
    cudaMalloc((void**)&norm, 3 * sizeof(float));
    kernels::vecnorm(d, norm, 3, 3, 2, 1);
    auto r = detail::download(norm, 3);
    const float want[3] = {8.1240384f, 9.6436508f, 11.224972f};
    cudaFree(d); cudaFree(norm);
    return detail::close(r.data(), want, 3, 1e-5f);
}
}  // namespace SfM
