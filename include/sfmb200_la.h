/* sfmb200_la - C ABI of the small linear-algebra helpers behind the kernels.h
 * facade.  Each one replaces a free function of the reference's namespace
 * kernels (SfM/kernels.h) that wrapped cuBLAS / cuSOLVER / Thrust; argument
 * meaning is the reference's.  All matrix pointers are DEVICE pointers; `stream`
 * is a cudaStream_t (NULL = legacy default stream, like the reference).
 * Return: 0 ok, -1 bad argument, -2 CUDA error, -5 singular matrix.
 */
#ifndef SFMB200_LA_H
#define SFMB200_LA_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* gpu_blas_mmul (kernels.h:102-109): C(m,n) = A(m,k) B(k,n), row-major. */
int sfmb200_la_mmul(const float* A, const float* B, float* C, int m, int k, int n, void* stream);
/* gpu_blas_mmul_batched (kernels.h:111-121): strides in elements, 0 = shared operand. */
int sfmb200_la_mmul_batched(const float* A, const float* B, float* C, int m, int k, int n, int stride_A, int stride_B,
                            int stride_C, int batches, void* stream);
/* gpu_blas_mmul_transpose_batched (kernels.h:123-130): A is stored k x m; C(m,n) = A^T B. */
int sfmb200_la_mmul_transpose_batched(const float* A, const float* B, float* C, int m, int k, int n, int stride_A,
                                      int stride_B, int stride_C, int batches, void* stream);
/* invert (kernels.h:132-173): batched n x n inverse, n <= 9; -5 if the first matrix is singular. */
int sfmb200_la_invert(const float* src, float* dst, int n, int batch, void* stream);
/* svd_square / regular_svd (kernels.h:175-234) = cusolverDnSgesvdjBatched: column-major A (m x n),
 * S min(m,n) descending, U m x m and V n x n column-major; m, n <= 9. */
int sfmb200_la_svd_batched(const float* A, float* S, float* U, float* V, int m, int n, int batch, void* stream);
/* transpose (kernels.h:196-209, launched once per matrix by regular_svd, 214-218): batched
 * row-major rows x cols -> column-major, one launch for the whole batch. */
int sfmb200_la_transpose_batched(const float* in, float* out, int rows, int cols, int batch, void* stream);
/* vecnorm (kernels.h:325-341) */
int sfmb200_la_vecnorm(const float* A, float* res, int row, int col, float exp, float final_pow, void* stream);
/* element_wise_mult / _div / _sum (kernels.h:297-323): op 0 / 1 / 2, in place on A. */
int sfmb200_la_elementwise(int op, float* A, const float* B, int size, void* stream);
/* threshold_count (kernels.h:343-355) */
int sfmb200_la_threshold_count(const float* A, int32_t* count, int batch_size, int ransac_count, float threshold,
                               void* stream);
/* row_extraction_kernel (kernels.h:452-458): floats 72..80 of each 81-float block. */
int sfmb200_la_row_extraction(const float* d_vt, float* d_E, int count, void* stream);
/* thrust::max_element (sfm.cu:136): index of the first maximum, to the host. */
int sfmb200_la_argmax_first(const int32_t* d_v, int n, int32_t* h_index, void* stream);

/* ---- the reference's hot-path kernels by name, as stand-alone entry points (the product path fuses them away) ---- */
/* copy_point (kernels.h:261-279): device SiftPoint[n] -> U1, U2 3 x n SoA pixel arrays with a row of ones. */
int sfmb200_la_copy_point(const void* d_sift, int n, float* U1, float* U2, void* stream);
/* kernels::kernels (kernels.h:236-259): A[h] (8x9 row-major) = rows kron(x1, x2) of the correspondences
 * indices[8h .. 8h+7]; d1, d2 are 3 x n SoA.  Guard h < H (the reference lets index == H through). */
int sfmb200_la_design_matrix(const float* d1, const float* d2, float* A, const int32_t* indices, int H, int n, void* stream);
/* normalizeE (kernels.h:281-295): in place E <- U diag(1,1,0) V^T for H row-major 3x3. */
int sfmb200_la_normalize_E(float* E, int H, void* stream);
/* candidate_kernels (kernels.h:357-385): d_P[4][16] from device 3x3 row-major u, v. */
int sfmb200_la_candidate_poses(float* d_P, const float* d_u, const float* d_v, void* stream);
/* compute_linear_triangulation_A (kernels.h:387-431): same arguments; candidate_m2 requires count == 4. */
int sfmb200_la_triangulation_A(float* A, const float* pt1, const float* pt2, int count, int num_points, const float* m1,
                               const float* m2, int P_ind, int candidate_m2, void* stream);
/* normalize_pt_kernal (kernels.h:433-450): v = n blocks of 16 floats, null vector in floats 12..15;
 * out = 4 x n SoA (x, y, z, 1), zeros when w == 0 or |w| > 5. */
int sfmb200_la_normalize_pt(const float* v, float* out, int n, void* stream);
/* kernCopyPositionsToVBO (kernels.h:471-483): 4 x n SoA -> n x 4 AoS (x,y,z,1)*scale; with pos_4xN == NULL
 * it is kernCopyVelocitiesToVBO (485-495): all ones. */
int sfmb200_la_copy_to_vbo(int n, const float* pos_4xN, float* vbo, float scale, void* stream);
#ifdef __cplusplus
}
#endif
#endif
