// Small linear-algebra entry points behind the kernels.h facade
// (reference: SfM/kernels.h:98-234, 297-355, 452-458 and thrust::max_element at
// SfM/sfm.cu:136).  The hot path does not use them - it never materialises the
// per-(hypothesis, point) temporaries these operate on - but the reference
// exposes them as free functions in namespace kernels, so a drop-in keeps them,
// implemented as plain kernels instead of cuBLAS / cuSOLVER / Thrust calls.
// Same argument meaning and row-major / column-major conventions as the
// reference call sites; all pointers are device pointers.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sfmb200_la.h"
#include "smallmat.cuh"

namespace {

// C(m,n) = A(m,k) * B(k,n), row-major, batch = blockIdx.y with element strides.
__global__ void mmul_kernel(const float* A, const float* B, float* C, int m, int k, int n, long long sA, long long sB,
                            long long sC, bool a_transposed) {
    const float* a = A + sA * blockIdx.y;
    const float* b = B + sB * blockIdx.y;
    float* c = C + sC * blockIdx.y;
    long long total = (long long)m * n;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        int i = (int)(t / n), j = (int)(t % n);
        float acc = 0.0f;
        for (int l = 0; l < k; l++) {
            float av = a_transposed ? a[(long long)l * m + i] : a[(long long)i * k + l];
            acc = fmaf(av, b[(long long)l * n + j], acc);
        }
        c[t] = acc;
    }
}

constexpr int LA_MAX = 9;

// Gauss-Jordan inverse with partial pivoting, one thread per matrix (n <= 9).
__global__ void invert_kernel(const float* src, float* dst, int n, int batch, int* info) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    float a[LA_MAX][2 * LA_MAX];
    const float* s = src + (size_t)b * n * n;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) { a[i][j] = s[i * n + j]; a[i][n + j] = (i == j) ? 1.0f : 0.0f; }
    int bad = 0;
    for (int col = 0; col < n; col++) {
        int piv = col;
        float best = fabsf(a[col][col]);
        for (int r = col + 1; r < n; r++)
            if (fabsf(a[r][col]) > best) { best = fabsf(a[r][col]); piv = r; }
        if (best == 0.0f) { bad = col + 1; break; }
        if (piv != col)
            for (int j = 0; j < 2 * n; j++) { float t = a[col][j]; a[col][j] = a[piv][j]; a[piv][j] = t; }
        float ip = 1.0f / a[col][col];
        for (int j = 0; j < 2 * n; j++) a[col][j] *= ip;
        for (int r = 0; r < n; r++) {
            if (r == col) continue;
            float f = a[r][col];
            for (int j = 0; j < 2 * n; j++) a[r][j] = fmaf(-f, a[col][j], a[r][j]);
        }
    }
    float* d = dst + (size_t)b * n * n;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) d[i * n + j] = bad ? 0.0f : a[i][n + j];
    if (info) info[b] = bad;
}

// Batched SVD with cusolverDnSgesvdjBatched's interface: column-major A (m x n,
// lda = m), outputs S (min(m,n), descending), U (m x m) and V (n x n), both
// column-major.  One-sided Jacobi, one thread per matrix, m, n <= 9.
__global__ void svd_batched_kernel(const float* A, float* S, float* U, float* V, int m, int n, int batch) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    float W[LA_MAX][LA_MAX];   // W[r][c] = working A V, r < m, c < n
    float Vm[LA_MAX][LA_MAX];
    const float* a = A + (size_t)b * m * n;
    for (int c = 0; c < n; c++)
        for (int r = 0; r < m; r++) W[r][c] = a[(size_t)c * m + r];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) Vm[i][j] = (i == j) ? 1.0f : 0.0f;
    for (int sw = 0; sw < 15; sw++) {
        bool rotated = false;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                float app = 0, aqq = 0, apq = 0;
                for (int r = 0; r < m; r++) {
                    app = fmaf(W[r][p], W[r][p], app);
                    aqq = fmaf(W[r][q], W[r][q], aqq);
                    apq = fmaf(W[r][p], W[r][q], apq);
                }
                if (fabsf(apq) <= 1e-7f * sqrtf(app * aqq) || apq == 0.0f) continue;
                rotated = true;
                float c, s, t;
                sfmb200::jacobi_angle(app, aqq, apq, c, s, t);
                for (int r = 0; r < m; r++) {
                    float x = W[r][p], y = W[r][q];
                    W[r][p] = fmaf(c, x, -s * y); W[r][q] = fmaf(s, x, c * y);
                }
                for (int r = 0; r < n; r++) {
                    float x = Vm[r][p], y = Vm[r][q];
                    Vm[r][p] = fmaf(c, x, -s * y); Vm[r][q] = fmaf(s, x, c * y);
                }
            }
        if (!rotated) break;
    }
    // column norms, sorted descending (selection sort on an index array)
    float nrm[LA_MAX];
    int ord[LA_MAX];
    for (int c = 0; c < n; c++) {
        float s2 = 0;
        for (int r = 0; r < m; r++) s2 = fmaf(W[r][c], W[r][c], s2);
        nrm[c] = sqrtf(s2);
        ord[c] = c;
    }
    for (int i = 0; i < n - 1; i++)
        for (int j = i + 1; j < n; j++)
            if (nrm[ord[j]] > nrm[ord[i]]) { int t = ord[i]; ord[i] = ord[j]; ord[j] = t; }
    int mn = m < n ? m : n;
    float* So = S + (size_t)b * mn;
    float* Uo = U + (size_t)b * m * m;
    float* Vo = V + (size_t)b * n * n;
    for (int c = 0; c < n; c++)
        for (int r = 0; r < n; r++) Vo[(size_t)c * n + r] = Vm[r][ord[c]];
    // U: normalised leading columns, completed by Gram-Schmidt on unit vectors
    float Um[LA_MAX][LA_MAX];
    int have = 0;
    for (int c = 0; c < mn; c++) {
        So[c] = nrm[ord[c]];
        if (nrm[ord[c]] > 1e-30f) {
            for (int r = 0; r < m; r++) Um[r][have] = W[r][ord[c]] / nrm[ord[c]];
            have++;
        } else {
            break;
        }
    }
    for (int cand = 0; cand < m && have < m; cand++) {
        float v[LA_MAX];
        for (int r = 0; r < m; r++) v[r] = (r == cand) ? 1.0f : 0.0f;
        for (int pass = 0; pass < 2; pass++)
            for (int c = 0; c < have; c++) {
                float d = 0;
                for (int r = 0; r < m; r++) d = fmaf(Um[r][c], v[r], d);
                for (int r = 0; r < m; r++) v[r] = fmaf(-d, Um[r][c], v[r]);
            }
        float s2 = 0;
        for (int r = 0; r < m; r++) s2 = fmaf(v[r], v[r], s2);
        if (s2 < 1e-6f) continue;
        float inv = 1.0f / sqrtf(s2);
        for (int r = 0; r < m; r++) Um[r][have] = v[r] * inv;
        have++;
    }
    for (int c = 0; c < m; c++)
        for (int r = 0; r < m; r++) Uo[(size_t)c * m + r] = (c < have) ? Um[r][c] : 0.0f;
}

__global__ void vecnorm_kernel(const float* A, float* res, int row, int col, float ex, float final_pow) {
    int index = blockIdx.x * blockDim.x + threadIdx.x;
    if (index >= col) return;
    float acc = 0.0f;
    for (int i = 0; i < row; i++) acc += powf(A[(size_t)i * col + index], ex);
    res[index] = (ex == final_pow) ? acc : powf(acc, final_pow / ex);
}

__global__ void elementwise_kernel(int op, float* A, const float* B, int size) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= size) return;
    float b = B[i];
    if (op == 0) A[i] *= b;
    else if (op == 1) A[i] = (b == 0.0f) ? 0.0f : A[i] / b;
    else A[i] += b;
}

// one warp per hypothesis, coalesced (the reference walks N residuals per thread)
__global__ void threshold_count_kernel(const float* A, int* count, int batch_size, int ransac_count, float thr) {
    int w = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x & 31;
    if (w >= ransac_count) return;
    int c = 0;
    for (int i = lane; i < batch_size; i += 32) c += A[(size_t)w * batch_size + i] < thr;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if (lane == 0) count[w] = c;
}

__global__ void row_extraction_kernel(const float* vt, float* E, int count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count * 9) return;
    int h = i / 9, k = i % 9;
    E[i] = vt[(size_t)h * 81 + 72 + k];
}

__global__ void argmax_first_kernel(const int* v, int n, unsigned long long* out) {
    __shared__ unsigned long long red[32];
    unsigned long long key = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        // order-preserving map of int32 to uint32, then lowest index wins ties
        unsigned long long k = ((unsigned long long)((unsigned)v[i] ^ 0x80000000u) << 32) | (0xFFFFFFFFu - (unsigned)i);
        key = k > key ? k : key;
    }
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long w = __shfl_xor_sync(0xFFFFFFFFu, key, o);
        key = w > key ? w : key;
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = key;
    __syncthreads();
    if (threadIdx.x < 32) {
        key = threadIdx.x < blockDim.x / 32 ? red[threadIdx.x] : 0ull;
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long w = __shfl_xor_sync(0xFFFFFFFFu, key, o);
            key = w > key ? w : key;
        }
        if (threadIdx.x == 0) *out = key;
    }
}

// out[b][c][r] = in[b][r][c]: row-major rows x cols -> column-major (what cuSOLVER wants).  The
// reference does this with one <<<1,(10,10)>>> launch PER MATRIX (kernels.h:214-218).
__global__ void transpose_batched_kernel(const float* in, float* out, int rows, int cols, long long total) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        int per = rows * cols;
        long long b = t / per;
        int e = (int)(t - b * per);
        int c = e / rows, r = e - c * rows;          // output element (c, r) of matrix b
        out[t] = in[b * per + (long long)r * cols + c];
    }
}

int check() { return cudaGetLastError() == cudaSuccess ? 0 : -2; }

}  // namespace

extern "C" {

int sfmb200_la_mmul(const float* A, const float* B, float* C, int m, int k, int n, void* stream) {
    if (!A || !B || !C || m < 1 || k < 1 || n < 1) return -1;
    long long total = (long long)m * n;
    int blocks = (int)((total + 255) / 256 > 65535 ? 65535 : (total + 255) / 256);
    mmul_kernel<<<dim3(blocks, 1), 256, 0, (cudaStream_t)stream>>>(A, B, C, m, k, n, 0, 0, 0, false);
    return check();
}
int sfmb200_la_mmul_batched(const float* A, const float* B, float* C, int m, int k, int n, int sA, int sB, int sC,
                            int batches, void* stream) {
    if (!A || !B || !C || m < 1 || k < 1 || n < 1 || batches < 1 || batches > 65535) return -1;
    long long total = (long long)m * n;
    int blocks = (int)((total + 255) / 256 > 1024 ? 1024 : (total + 255) / 256);
    mmul_kernel<<<dim3(blocks, batches), 256, 0, (cudaStream_t)stream>>>(A, B, C, m, k, n, sA, sB, sC, false);
    return check();
}
int sfmb200_la_mmul_transpose_batched(const float* A, const float* B, float* C, int m, int k, int n, int sA, int sB,
                                      int sC, int batches, void* stream) {
    if (!A || !B || !C || m < 1 || k < 1 || n < 1 || batches < 1 || batches > 65535) return -1;
    long long total = (long long)m * n;
    int blocks = (int)((total + 255) / 256 > 1024 ? 1024 : (total + 255) / 256);
    mmul_kernel<<<dim3(blocks, batches), 256, 0, (cudaStream_t)stream>>>(A, B, C, m, k, n, sA, sB, sC, true);
    return check();
}
int sfmb200_la_invert(const float* src, float* dst, int n, int batch, void* stream) {
    if (!src || !dst || n < 1 || n > LA_MAX || batch < 1) return -1;
    int* info = nullptr;
    if (cudaMalloc(&info, sizeof(int) * batch) != cudaSuccess) return -2;
    invert_kernel<<<(batch + 63) / 64, 64, 0, (cudaStream_t)stream>>>(src, dst, n, batch, info);
    int rc = check();
    int first = 0;
    if (rc == 0 && cudaMemcpyAsync(&first, info, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess) rc = -2;
    if (rc == 0 && cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) rc = -2;
    cudaFree(info);
    if (rc == 0 && first != 0) return -5;   // singular (the reference exit()s here, kernels.h:144-161)
    return rc;
}
int sfmb200_la_svd_batched(const float* A, float* S, float* U, float* V, int m, int n, int batch, void* stream) {
    if (!A || !S || !U || !V || m < 1 || n < 1 || m > LA_MAX || n > LA_MAX || batch < 1) return -1;
    svd_batched_kernel<<<(batch + 63) / 64, 64, 0, (cudaStream_t)stream>>>(A, S, U, V, m, n, batch);
    return check();
}
int sfmb200_la_transpose_batched(const float* in, float* out, int rows, int cols, int batch, void* stream) {
    if (!in || !out || rows < 1 || cols < 1 || batch < 1) return -1;
    long long total = (long long)rows * cols * batch;
    int blocks = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
    transpose_batched_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(in, out, rows, cols, total);
    return check();
}
int sfmb200_la_vecnorm(const float* A, float* res, int row, int col, float ex, float final_pow, void* stream) {
    if (!A || !res || row < 1 || col < 1) return -1;
    vecnorm_kernel<<<(col + 255) / 256, 256, 0, (cudaStream_t)stream>>>(A, res, row, col, ex, final_pow);
    return check();
}
int sfmb200_la_elementwise(int op, float* A, const float* B, int size, void* stream) {
    if (!A || !B || size < 1 || op < 0 || op > 2) return -1;
    elementwise_kernel<<<(size + 255) / 256, 256, 0, (cudaStream_t)stream>>>(op, A, B, size);
    return check();
}
int sfmb200_la_threshold_count(const float* A, int32_t* count, int batch_size, int ransac_count, float thr,
                               void* stream) {
    if (!A || !count || batch_size < 1 || ransac_count < 1) return -1;
    long long threads = (long long)ransac_count * 32;
    threshold_count_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(A, count, batch_size,
                                                                                                ransac_count, thr);
    return check();
}
int sfmb200_la_row_extraction(const float* vt, float* E, int count, void* stream) {
    if (!vt || !E || count < 1) return -1;
    row_extraction_kernel<<<(count * 9 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(vt, E, count);
    return check();
}
int sfmb200_la_argmax_first(const int32_t* d_v, int n, int32_t* h_index, void* stream) {
    if (!d_v || !h_index || n < 1) return -1;
    unsigned long long* d_out = nullptr;
    if (cudaMalloc(&d_out, 8) != cudaSuccess) return -2;
    argmax_first_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(d_v, n, d_out);
    unsigned long long key = 0;
    int rc = check();
    if (rc == 0 && cudaMemcpyAsync(&key, d_out, 8, cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess) rc = -2;
    if (rc == 0 && cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) rc = -2;
    cudaFree(d_out);
    *h_index = (int32_t)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
    return rc;
}

}  // extern "C"
