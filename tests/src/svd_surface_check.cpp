// Exercises the svd.h facade surface (reference: SfM/svd.h:33-501) on the host: every name of the reference,
// including its internal steps, with the meaning the reference gives it.  Built by tests/test_cpu_hostlogic.py
// with g++ (and, as a .cu, by nvcc with a kernel that calls svd() on the device - compile only on the CPU tier).
#include <cstdio>
#include <cstdlib>

#include "svd.h"

#ifdef __CUDACC__
__global__ void svd_on_device(const float* a, float* out) {
    float u[9], s[9], v[9], q[4], m[9], ata[9];
    svd(a, u, s, v);
    multAtB(a, a, ata);
    jacobiEigenanlysis(ata, q);
    quatToMat3(q, m);
    float r[9], qq[9];
    QRDecomposition(a, qq, r);
    sortSingularValues(m, v);
    float inv[16], id[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    InvertMatrix4x4(id, inv);
    for (int i = 0; i < 9; i++) out[i] = u[i] + s[i] + v[i] + m[i] + r[i] + qq[i] + det(a) + inv[0];
}
#endif

static float maxabs(const float* a, int n) {
    float m = 0;
    for (int i = 0; i < n; i++) m = fmaxf(m, fabsf(a[i]));
    return m;
}

int main() {
    srand(5);
    float worst_diag = 0, worst_qr = 0, worst_svd = 0, worst_orth = 0, worst_pd = 0;
    for (int trial = 0; trial < 200; trial++) {
        float a[9];
        for (int i = 0; i < 9; i++) a[i] = (float)rand() / RAND_MAX * 2 - 1;
        // jacobiEigenanlysis: V^T (A^T A) V diagonal, V a rotation
        float ata[9], work[9], q[4], V[9], t[9], d[9];
        multAtB(a, a, ata);
        for (int i = 0; i < 9; i++) work[i] = ata[i];
        jacobiEigenanlysis(work, q);
        quatToMat3(q, V);
        multAtB(V, ata, t);
        multAB(t, V, d);
        float off = fmaxf(fmaxf(fabsf(d[1]), fabsf(d[2])), fabsf(d[5]));
        worst_diag = fmaxf(worst_diag, off / maxabs(ata, 9));
        float vtv[9];
        multAtB(V, V, vtv);
        vtv[0] -= 1; vtv[4] -= 1; vtv[8] -= 1;
        worst_orth = fmaxf(worst_orth, maxabs(vtv, 9));
        if (fabsf(det(V) - 1) > 1e-4f) return 2;
        // sortSingularValues: column norms of b descending, b V^T unchanged
        float b[9], bv0[9], bv1[9], Vs[9];
        multAB(a, V, b);
        multABt(b, V, bv0);
        for (int i = 0; i < 9; i++) Vs[i] = V[i];
        sortSingularValues(b, Vs);
        multABt(b, Vs, bv1);
        for (int i = 0; i < 9; i++) if (fabsf(bv0[i] - bv1[i]) > 1e-5f) return 3;
        float n0 = dist2(b[0], b[3], b[6]), n1 = dist2(b[1], b[4], b[7]), n2 = dist2(b[2], b[5], b[8]);
        if (!(n0 >= n1 && n1 >= n2)) return 4;
        // QRDecomposition: b = q r, r upper triangular, q a rotation
        float qm[9], r[9], rec[9];
        QRDecomposition(b, qm, r);
        multAB(qm, r, rec);
        for (int i = 0; i < 9; i++) worst_qr = fmaxf(worst_qr, fabsf(rec[i] - b[i]));
        if (fmaxf(fmaxf(fabsf(r[3]), fabsf(r[6])), fabsf(r[7])) > 1e-5f) return 5;
        // QRGivensQuaternion: half-angle of the rotation that annihilates a2
        float ch, sh;
        QRGivensQuaternion(a[0], a[3], ch, sh);
        float c = ch * ch - sh * sh, s = 2 * ch * sh;
        if (fabsf(-s * a[0] + c * a[3]) > 1e-5f || fabsf(ch * ch + sh * sh - 1) > 1e-5f) return 6;
        approximateGivensQuaternion(ata[0], ata[3], ata[4], ch, sh);
        c = ch * ch - sh * sh; s = 2 * ch * sh;
        if (fabsf(ata[3] * (c * c - s * s) - (ata[0] - ata[4]) * s * c) > 1e-5f * (1 + maxabs(ata, 9))) return 7;
        // svd: a = u s v^T, sigma sorted, u and v rotations
        float u[9], sg[9], v[9], us[9];
        svd(a, u, sg, v);
        multAB(u, sg, us);
        multABt(us, v, rec);
        for (int i = 0; i < 9; i++) worst_svd = fmaxf(worst_svd, fabsf(rec[i] - a[i]));
        if (!(sg[0] >= sg[4] && sg[4] >= fabsf(sg[8]))) return 8;
        if (fabsf(det(u) - 1) > 1e-4f || fabsf(det(v) - 1) > 1e-4f) return 9;
        // pd: a = U P, P symmetric
        float U[9], P[9];
        pd(a, U, P);
        multAB(U, P, rec);
        for (int i = 0; i < 9; i++) worst_pd = fmaxf(worst_pd, fabsf(rec[i] - a[i]));
        if (fabsf(P[1] - P[3]) > 1e-5f || fabsf(P[2] - P[6]) > 1e-5f) return 10;
    }
    float x = 1, y = 2;
    condSwap(true, x, y);
    if (x != 2 || y != 1) return 11;
    condNegSwap(true, x, y);
    if (x != 1 || y != -2) return 12;
    float m16[16] = {1, 0, 0, 1, 0, 2, 0, 2, 0, 0, 4, 3, 0, 0, 0, 1}, inv[16];
    if (!InvertMatrix4x4(m16, inv) || fabsf(inv[5] - 0.5f) > 1e-6f) return 13;
    printf("svd.h surface ok: diag %.2e orth %.2e qr %.2e svd %.2e pd %.2e\n", worst_diag, worst_orth, worst_qr, worst_svd, worst_pd);
    return (worst_diag < 1e-5f && worst_orth < 1e-5f && worst_qr < 1e-5f && worst_svd < 1e-5f && worst_pd < 1e-5f) ? 0 : 1;
}
