/* Pure C caller of the C ABI (no Python, no torch): allocates device buffers with the CUDA runtime, runs the selective-scan
 * forward + backward of include/ss2d_b200.h on a small problem and checks them against the plain-C oracle
 * (oracle/ss2d_oracle.c, linked in as test infrastructure).  Built and run by tests/test_c_abi_gpu.py:
 *   gcc -O2 -std=c11 scan_smoke.c ../../oracle/ss2d_oracle.c -I../../include -I/usr/local/cuda/include \
 *       -L../../focalnet_b200/lib -lss2d_b200 -L/usr/local/cuda/lib64 -lcudart -lm -fopenmp -o scan_smoke            */
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "ss2d_b200.h"

int ss2d_oracle_scan_fwd(const float *u, const float *delta, const float *A, const float *B, const float *C, const float *D,
                         const float *z, const float *delta_bias, int softplus, int64_t batch, int64_t dim, int64_t L, int64_t N,
                         int64_t G, float *out, float *x, float *last);
int ss2d_oracle_scan_bwd(const float *u, const float *delta, const float *A, const float *B, const float *C, const float *D,
                         const float *z, const float *delta_bias, const float *dout, int softplus, int64_t batch, int64_t dim,
                         int64_t L, int64_t N, int64_t G, float *du, float *ddelta, float *dA, float *dB, float *dC, float *dD,
                         float *dbias, float *dz);

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 2; } } while (0)

static float frand(uint32_t *s) { *s = *s * 1664525u + 1013904223u; return (float)((*s >> 8) & 0xffffff) / 16777216.f; }
static float *dev_copy(const float *h, size_t n) { float *d = NULL; if (cudaMalloc((void **)&d, n * 4) != cudaSuccess) return NULL; cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice); return d; }
static float *dev_zero(size_t n) { float *d = NULL; if (cudaMalloc((void **)&d, n * 4) != cudaSuccess) return NULL; cudaMemset(d, 0, n * 4); return d; }
static double rel_err(const float *a, const float *b, size_t n) {
    double mx = 1e-3, err = 0;
    for (size_t i = 0; i < n; ++i) { if (fabs(b[i]) > mx) mx = fabs(b[i]); if (fabs((double)a[i] - b[i]) > err) err = fabs((double)a[i] - b[i]); }
    return err / mx;
}

int main(void) {
    const int64_t Bn = 2, Dm = 32, N = 16, L = 1040, G = 4;  /* dstate 16: the state-lanes kernels are reachable by pinning */
    const size_t nu = Bn * Dm * L, nbc = Bn * G * N * L, nA = Dm * N, nx = Bn * Dm * 1 * 2 * N;
    uint32_t seed = 12345u;
    float *u = malloc(nu * 4), *dl = malloc(nu * 4), *dout = malloc(nu * 4), *A = malloc(nA * 4), *Bm = malloc(nbc * 4), *Cm = malloc(nbc * 4),
          *D = malloc(Dm * 4), *bias = malloc(Dm * 4);
    for (size_t i = 0; i < nu; ++i) { u[i] = 2 * frand(&seed) - 1; dl[i] = 0.5f * frand(&seed); dout[i] = 2 * frand(&seed) - 1; }
    for (size_t i = 0; i < nbc; ++i) { Bm[i] = 2 * frand(&seed) - 1; Cm[i] = 2 * frand(&seed) - 1; }
    for (size_t i = 0; i < nA; ++i) A[i] = -0.5f * frand(&seed);
    for (int i = 0; i < Dm; ++i) { D[i] = 2 * frand(&seed) - 1; bias[i] = 0.5f * frand(&seed); }
    if (ss2d_abi_version() != SS2D_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 1; }
    printf("%s\n", ss2d_build_info());

    int worst_fail = 0;
    for (int family = SS2D_FAMILY_STATELANES; family <= SS2D_FAMILY_WARPSCAN; ++family) {
        ss2d_scan_fwd_params p = {0};
        p.batch = Bn; p.dim = Dm; p.seqlen = L; p.dstate = N; p.ngroups = G;
        p.in_dtype = SS2D_F32; p.out_dtype = SS2D_F32; p.delta_softplus = 1; p.family = family;
        p.u = dev_copy(u, nu); p.delta = dev_copy(dl, nu); p.A = dev_copy(A, nA); p.B = dev_copy(Bm, nbc); p.C = dev_copy(Cm, nbc);
        p.D = dev_copy(D, Dm); p.delta_bias = dev_copy(bias, Dm);
        p.u_bstride = p.delta_bstride = Dm * L; p.u_dstride = p.delta_dstride = L;
        p.B_bstride = p.C_bstride = G * N * L; p.B_gstride = p.C_gstride = N * L; p.B_nstride = p.C_nstride = L;
        float *d_out = dev_zero(nu), *d_x = dev_zero(nx), *d_ck = dev_zero((size_t)ss2d_scan_ckpt_floats(Bn, Dm, L, N));
        p.out = d_out; p.out_bstride = Dm * L; p.out_dstride = L; p.x = d_x; p.ckpt = d_ck;
        if (ss2d_scan_family(&p) != family) { fprintf(stderr, "family %d not honoured\n", family); return 1; }
        int rc = ss2d_selective_scan_fwd(&p, NULL);
        if (rc) { fprintf(stderr, "fwd: %s\n", ss2d_error_string(rc)); return 1; }
        ss2d_scan_bwd_params q = {0};
        q.f = p;
        q.dout = dev_copy(dout, nu); q.dout_bstride = Dm * L; q.dout_dstride = L;
        float *d_du = dev_zero(nu), *d_dd = dev_zero(nu), *d_dA = dev_zero(nA), *d_dB = dev_zero(nbc), *d_dC = dev_zero(nbc), *d_dD = dev_zero(Dm),
              *d_db = dev_zero(Dm);
        q.du = d_du; q.ddelta = d_dd; q.dA = d_dA; q.dB = d_dB; q.dC = d_dC; q.dD = d_dD; q.ddelta_bias = d_db;
        rc = ss2d_selective_scan_bwd(&q, NULL);
        if (rc) { fprintf(stderr, "bwd: %s\n", ss2d_error_string(rc)); return 1; }
        CK(cudaDeviceSynchronize());

        float *out = malloc(nu * 4), *du = malloc(nu * 4), *dd = malloc(nu * 4), *dA = malloc(nA * 4), *dB = malloc(nbc * 4), *dC = malloc(nbc * 4);
        CK(cudaMemcpy(out, d_out, nu * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(du, d_du, nu * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(dd, d_dd, nu * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(dA, d_dA, nA * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(dB, d_dB, nbc * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(dC, d_dC, nbc * 4, cudaMemcpyDeviceToHost));
        float *o_out = malloc(nu * 4), *o_x = malloc(nx * 4), *o_last = malloc(Bn * Dm * N * 4), *o_du = malloc(nu * 4), *o_dd = malloc(nu * 4),
              *o_dA = malloc(nA * 4), *o_dB = malloc(nbc * 4), *o_dC = malloc(nbc * 4), *o_dD = malloc(Dm * 4), *o_db = malloc(Dm * 4);
        ss2d_oracle_scan_fwd(u, dl, A, Bm, Cm, D, NULL, bias, 1, Bn, Dm, L, N, G, o_out, o_x, o_last);
        ss2d_oracle_scan_bwd(u, dl, A, Bm, Cm, D, NULL, bias, dout, 1, Bn, Dm, L, N, G, o_du, o_dd, o_dA, o_dB, o_dC, o_dD, o_db, NULL);
        const double e[6] = {rel_err(out, o_out, nu), rel_err(du, o_du, nu), rel_err(dd, o_dd, nu), rel_err(dA, o_dA, nA), rel_err(dB, o_dB, nbc),
                             rel_err(dC, o_dC, nbc)};
        printf("family %d: rel err out %.2e du %.2e ddelta %.2e dA %.2e dB %.2e dC %.2e\n", family, e[0], e[1], e[2], e[3], e[4], e[5]);
        for (int i = 0; i < 6; ++i) if (!(e[i] < 1e-3)) worst_fail = 1;
    }
    /* error behaviour: invalid arguments are refused before anything is launched */
    ss2d_scan_fwd_params bad = {0};
    if (ss2d_selective_scan_fwd(&bad, NULL) >= 0) { fprintf(stderr, "empty params not refused\n"); return 1; }
    if (worst_fail) { fprintf(stderr, "FAIL\n"); return 1; }
    printf("C ABI smoke: OK\n");
    return 0;
}
