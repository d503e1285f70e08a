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
#ifdef __cplusplus
}
#endif
#endif
