// ss2d_optim.cu — optimizer side of the data-parallel training step, sm_100a.
//
// Replaces, for ONE flat fp32 bucket holding every parameter gradient of the model (2,541,673 values for the ITS
// MIMOUNet), the three separate multi-tensor sweeps of the reference's step (ITS/train.py:89-91):
//     torch.nn.utils.clip_grad_norm_(model.parameters(), 0.001)   -> per-tensor norms, stack, norm, per-tensor mul_
//     optimizer.step()                                            -> Adam(lr, betas=(0.9,0.999), eps=1e-8), train.py:16
// and optimizer.zero_grad() (train.py:61), by two launches over the bucket:
//   1. ss2d_optim_sumsq: per-block partial sums of (grad_scale * g)^2 (grad_scale = 1 / world size turns the all-reduce
//      SUM into the average), fixed block -> partial mapping, so the norm is bit-reproducible;
//   2. ss2d_optim_clip_adam: every block re-reduces the partials (a few hundred floats from L2) in a fixed order, forms
//      clip = min(1, max_norm / (norm + 1e-6)) like clip_grad_norm_, applies Adam and zeroes the gradient for the next
//      step.  HBM-bound: 4 arrays read + 4 written, 32 bytes per parameter (81 MB per step), 128-bit accesses.
#include "ss2d_common.cuh"
#include "../../include/ss2d_b200.h"

namespace ss2d {

constexpr int kOptThreads = 256;

__global__ void __launch_bounds__(kOptThreads) optim_sumsq_kernel(const float *__restrict__ g, int64_t n, float scale,
                                                                   float *__restrict__ partials) {
    float acc = 0.f;
    const int64_t n4 = n / 4;
    const float4 *g4 = reinterpret_cast<const float4 *>(g);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = g4[i];
        const float a = v.x * scale, b = v.y * scale, c = v.z * scale, d = v.w * scale;
        acc += a * a + b * b + c * c + d * d;
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n - n4 * 4)) {
        const float a = g[n4 * 4 + threadIdx.x] * scale;
        acc += a * a;
    }
    __shared__ float red[kOptThreads / kWarp];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kOptThreads / kWarp; ++w) s += red[w];
        partials[blockIdx.x] = s;
    }
}

struct AdamArgs {
    float lr, beta1, beta2, eps, bc1, bc2_sqrt, max_norm, grad_scale;
};

__device__ __forceinline__ void adam1(float &p, float &g, float &m, float &v, const AdamArgs &a, float coef) {
    const float gr = g * coef;
    m = a.beta1 * m + (1.f - a.beta1) * gr;                 // exp_avg.lerp_(grad, 1 - beta1)
    v = a.beta2 * v + (1.f - a.beta2) * gr * gr;            // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
    p -= (a.lr / a.bc1) * (m / denom);
    g = 0.f;
}

__global__ void __launch_bounds__(kOptThreads) optim_clip_adam_kernel(float *__restrict__ p, float *__restrict__ g,
                                                                      float *__restrict__ m, float *__restrict__ v, int64_t n,
                                                                      const float *__restrict__ partials, int npart,
                                                                      float *__restrict__ norm_out, AdamArgs a) {
    // total norm: the same fixed-order reduction in every block
    __shared__ float red[kOptThreads / kWarp];
    __shared__ float s_coef;
    float acc = 0.f;
    for (int i = threadIdx.x; i < npart; i += kOptThreads) acc += partials[i];
#pragma unroll
    for (int k = 16; k >= 1; k >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, k);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kOptThreads / kWarp; ++w) s += red[w];
        const float norm = sqrtf(s);
        if (blockIdx.x == 0 && norm_out) *norm_out = norm;
        float clip = a.max_norm > 0.f ? a.max_norm / (norm + 1e-6f) : 1.f;  // clip_grad_norm_: clamp(max_norm / (norm + 1e-6), max=1)
        s_coef = a.grad_scale * fminf(clip, 1.f);
    }
    __syncthreads();
    const float coef = s_coef;
    const int64_t n4 = n / 4;
    float4 *p4 = reinterpret_cast<float4 *>(p), *g4 = reinterpret_cast<float4 *>(g), *m4 = reinterpret_cast<float4 *>(m),
           *v4 = reinterpret_cast<float4 *>(v);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 P = p4[i], G = g4[i], M = m4[i], V = v4[i];
        adam1(P.x, G.x, M.x, V.x, a, coef);
        adam1(P.y, G.y, M.y, V.y, a, coef);
        adam1(P.z, G.z, M.z, V.z, a, coef);
        adam1(P.w, G.w, M.w, V.w, a, coef);
        p4[i] = P; g4[i] = G; m4[i] = M; v4[i] = V;
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n - n4 * 4)) {
        const int64_t i = n4 * 4 + threadIdx.x;
        adam1(p[i], g[i], m[i], v[i], a, coef);
    }
}

}  // namespace ss2d

extern "C" int64_t ss2d_optim_partials(void) { return SS2D_OPTIM_PARTIALS; }

extern "C" int ss2d_optim_clip_adam(float *param, float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float *partials,
                                    float *norm_out, float lr, float beta1, float beta2, float eps, int64_t step, float max_norm,
                                    float grad_scale, void *stream) {
    using namespace ss2d;
    if (!param || !grad || !exp_avg || !exp_avg_sq || !partials || n <= 0 || step <= 0) return SS2D_EINVAL;
    for (const void *q : {(const void *)param, (const void *)grad, (const void *)exp_avg, (const void *)exp_avg_sq})
        if (reinterpret_cast<uintptr_t>(q) & 15) return SS2D_ESTRIDE;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int64_t want = (n / 4 + kOptThreads - 1) / kOptThreads;
    const int grid = (int)(want < 1 ? 1 : (want > SS2D_OPTIM_PARTIALS ? SS2D_OPTIM_PARTIALS : want));
    optim_sumsq_kernel<<<grid, kOptThreads, 0, s>>>(grad, n, grad_scale, partials);
    AdamArgs a;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.max_norm = max_norm; a.grad_scale = grad_scale;
    a.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    a.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    optim_clip_adam_kernel<<<grid, kOptThreads, 0, s>>>(param, grad, exp_avg, exp_avg_sq, n, partials, grid, norm_out, a);
    return (int)cudaGetLastError();
}
