// Host-callable entry points over the shared small-matrix code.  They back the
// svd.h facade (reference: SfM/svd.h) and let the CPU-only test tier check the
// per-hypothesis math against the fp64 oracle without a GPU.  They are NOT a
// CPU fallback for the hot path: nothing in the product path calls them.
#include "hyp_solver.cuh"
#include "../../include/sfmb200.h"

using namespace sfmb200;

extern "C" {

void sfmb200_host_svd3(const float a[9], float u[9], float s[9], float v[9]) { svd3<5>(a, u, s, v); }

void sfmb200_host_svd3_reference_orientation(const float a[9], float u[9], float s[9], float v[9]) {
    svd3_reference_orientation(a, u, s, v);
}

void sfmb200_host_solve_hypothesis(const float pts[32], float E[9]) {
    Corr c[8];
    for (int i = 0; i < 8; i++) c[i] = Corr{pts[4 * i], pts[4 * i + 1], pts[4 * i + 2], pts[4 * i + 3]};
    solve_hypothesis<0>(c, E);
}

void sfmb200_host_solve_hypothesis_projector(const float pts[32], float E[9]) {
    Corr c[8];
    for (int i = 0; i < 8; i++) c[i] = Corr{pts[4 * i], pts[4 * i + 1], pts[4 * i + 2], pts[4 * i + 3]};
    solve_hypothesis_projector(c, E);
}

void sfmb200_host_null4(const float A[16], float x[4]) { null4<5>(A, x); }

int sfmb200_host_null4_fast(const float A[16], float x[4]) { return null4_inverse_iteration<5>(A, x) ? 0 : 1; }

int sfmb200_host_dlt_null(const float A[16], float x[4]) { return dlt_null_adjugate1(A, x) ? 0 : 1; }

void sfmb200_host_dlt_null_power4(const float A[16], float x[4]) {
    const float x1[1] = {A[2]}, y1[1] = {A[6]};
    float a[1][4], b[1][4], out[1][4];
    for (int c = 0; c < 4; c++) { a[0][c] = A[8 + c]; b[0][c] = A[12 + c]; }
    dlt_null_power4<1>(x1, y1, a, b, out);
    for (int c = 0; c < 4; c++) x[c] = out[0][c];
}

int sfmb200_host_inv4(const float m[16], float out[16]) { return inv4(m, out) ? 0 : -1; }

void sfmb200_host_sample_indices(uint64_t seed, uint64_t h, int n, int32_t idx[8]) {
    int tmp[8];
    sample_indices(seed, h, n, tmp);
    for (int i = 0; i < 8; i++) idx[i] = tmp[i];
}

void sfmb200_host_sample_indices_disjoint(uint64_t seed, uint64_t h, int n, int32_t idx[8]) {
    int tmp[8];
    sample_indices_disjoint(seed, h, n, tmp);
    for (int i = 0; i < 8; i++) idx[i] = tmp[i];
}

}  // extern "C"
