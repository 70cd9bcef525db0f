// ss2d_dtproj.cu — the low-rank dt projection of the SS2D core, hand-written (sm_100a).
//
// Replaces `dts = F.conv1d(dts.contiguous().view(B, -1, L), dt_projs_weight.view(K * D, -1, 1), groups=K)` of
// cross_selective_scan (reference: ITS/models/vmamba_layers.py:264; einsum twin :270) and its autograd backward:
//     delta[b, k*D + d, l] = sum_r W[k, d, r] * dtlr[b, k, r, l]            R = dt_rank (6 in the ITS model), K = 4, D = 192
// The op writes the 768-row delta from 24 rows — it is a pure HBM stream (6 MACs per output), but the library path runs it
// as a grouped convolution with an NCHW<->NHWC conversion on each side and, in the backward, reads ddelta twice more
// (dgrad, wgrad) through the same conversions.  Here:
//   fwd : one pass, every thread owns 4 consecutive steps (128-bit accesses), keeps the R low-rank rows in registers and
//         streams the D output rows of its slice; W of the direction sits in shared memory.  Bytes: 4*B*K*L*(R + D).
//   bwd : ddelta is read twice — once for d_dtlr[b,k,r,l] = sum_d W[k,d,r] ddelta[b,k*D+d,l] (same thread mapping, R
//         128-bit accumulators per thread) and once for dW[k,d,r] = sum_{b,l} ddelta * dtlr (a warp per (batch row, k, 4 consecutive
//         d), lanes across l, one atomicAdd per (k,d,r) and warp).  Bytes: 4*B*K*L*(2*D + 2*R).
// `dtlr` is addressed through element strides so that it can be the dt rows of the permuted x_dbl (no .contiguous() copy).
#include "ss2d_common.cuh"
#include "../../include/ss2d_b200.h"

namespace ss2d {

constexpr int kDtThreads = 256;
constexpr int kDtMaxD = 512;  // W of one direction in shared memory: D * R floats

struct DtGeom {
    int64_t B, K, D, R, L, sb, sk, sr;
    int dslices, dper;
};

template <int R>
__global__ void __launch_bounds__(kDtThreads) dt_proj_fwd_kernel(const float *__restrict__ dtlr, const float *__restrict__ W,
                                                                float *__restrict__ out, const DtGeom g) {
    extern __shared__ float sW[];  // [dper][R]
    const int k = blockIdx.y / g.dslices, ds = blockIdx.y % g.dslices, b = blockIdx.z;
    const int d0 = ds * g.dper, dn = min(g.dper, (int)g.D - d0);
    for (int i = threadIdx.x; i < dn * R; i += kDtThreads) sW[i] = __ldg(W + ((int64_t)k * g.D + d0) * R + i);
    __syncthreads();
    const int64_t l = ((int64_t)blockIdx.x * kDtThreads + threadIdx.x) * 4;
    if (l >= g.L) return;
    float4 x[R];
#pragma unroll
    for (int r = 0; r < R; ++r) x[r] = __ldg(reinterpret_cast<const float4 *>(dtlr + b * g.sb + k * g.sk + r * g.sr + l));
    float *o = out + (((int64_t)b * g.K + k) * g.D + d0) * g.L + l;
#pragma unroll 4
    for (int d = 0; d < dn; ++d) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float w = sW[d * R + r];
            acc.x = fmaf(w, x[r].x, acc.x); acc.y = fmaf(w, x[r].y, acc.y);
            acc.z = fmaf(w, x[r].z, acc.z); acc.w = fmaf(w, x[r].w, acc.w);
        }
        *reinterpret_cast<float4 *>(o + (int64_t)d * g.L) = acc;
    }
}

// d_dtlr[b, k, r, l] = sum_d W[k, d, r] * dout[b, k*D + d, l]   (contiguous (B, K, R, L) output)
template <int R>
__global__ void __launch_bounds__(kDtThreads) dt_proj_bwd_x_kernel(const float *__restrict__ dout, const float *__restrict__ W,
                                                                  float *__restrict__ dx, const DtGeom g) {
    extern __shared__ float sW[];  // [D][R]
    const int k = blockIdx.y, b = blockIdx.z;
    for (int i = threadIdx.x; i < (int)g.D * R; i += kDtThreads) sW[i] = __ldg(W + (int64_t)k * g.D * R + i);
    __syncthreads();
    const int64_t l = ((int64_t)blockIdx.x * kDtThreads + threadIdx.x) * 4;
    if (l >= g.L) return;
    float4 acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float *gp = dout + ((int64_t)b * g.K + k) * g.D * g.L + l;
#pragma unroll 4
    for (int d = 0; d < (int)g.D; ++d) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(gp + (int64_t)d * g.L));
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float w = sW[d * R + r];
            acc[r].x = fmaf(w, v.x, acc[r].x); acc[r].y = fmaf(w, v.y, acc[r].y);
            acc[r].z = fmaf(w, v.z, acc[r].z); acc[r].w = fmaf(w, v.w, acc[r].w);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) *reinterpret_cast<float4 *>(dx + (((int64_t)b * g.K + k) * R + r) * g.L + l) = acc[r];
}

// dW[k, d, r] += sum_l dout[b, k*D + d, l] * dtlr[b, k, r, l]   — one warp per (b, k, kDtRows consecutive d), lanes across l:
// the R low-rank rows are loaded once per kDtRows rows of dout (they come from L1 / L2: 192 rows of a direction share them)
constexpr int kDtRows = 4;
template <int R>
__global__ void __launch_bounds__(kDtThreads) dt_proj_bwd_w_kernel(const float *__restrict__ dout, const float *__restrict__ dtlr,
                                                                  float *__restrict__ dW, const DtGeom g) {
    const int lane = threadIdx.x & 31;
    const int64_t dgroups = (g.D + kDtRows - 1) / kDtRows;
    const int64_t unit = (int64_t)blockIdx.x * (kDtThreads / kWarp) + (threadIdx.x >> 5);  // (b, k, d group)
    if (unit >= g.B * g.K * dgroups) return;
    const int d0 = (int)(unit % dgroups) * kDtRows, k = (int)((unit / dgroups) % g.K), b = (int)(unit / (dgroups * g.K));
    const float *gp = dout + (((int64_t)b * g.K + k) * g.D + d0) * g.L;
    const float *xp = dtlr + b * g.sb + k * g.sk;
    float acc[kDtRows][R];
#pragma unroll
    for (int i = 0; i < kDtRows; ++i)
#pragma unroll
        for (int r = 0; r < R; ++r) acc[i][r] = 0.f;
    for (int64_t l = lane * 4; l < g.L; l += kWarp * 4) {
        float4 v[kDtRows];
#pragma unroll
        for (int i = 0; i < kDtRows; ++i)
            v[i] = d0 + i < g.D ? __ldg(reinterpret_cast<const float4 *>(gp + (int64_t)i * g.L + l)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float4 x = __ldg(reinterpret_cast<const float4 *>(xp + r * g.sr + l));
#pragma unroll
            for (int i = 0; i < kDtRows; ++i)
                acc[i][r] = fmaf(v[i].x, x.x, fmaf(v[i].y, x.y, fmaf(v[i].z, x.z, fmaf(v[i].w, x.w, acc[i][r]))));
        }
    }
#pragma unroll
    for (int i = 0; i < kDtRows; ++i)
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) acc[i][r] += __shfl_xor_sync(0xffffffffu, acc[i][r], m);
        }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < kDtRows; ++i)
            if (d0 + i < g.D) {
#pragma unroll
                for (int r = 0; r < R; ++r) atomicAdd(dW + ((int64_t)k * g.D + d0 + i) * R + r, acc[i][r]);
            }
    }
}

static int dt_geom(DtGeom &g, int64_t B, int64_t K, int64_t D, int64_t R, int64_t L, int64_t sb, int64_t sk, int64_t sr, const void *a,
                   const void *b) {
    if (B <= 0 || K <= 0 || D <= 0 || R <= 0 || R > 8 || L <= 0 || D > kDtMaxD) return SS2D_EINVAL;
    // 128-bit accesses along l: every row must start on a 16-byte boundary
    if (L % 4 != 0 || sb % 4 != 0 || sk % 4 != 0 || sr % 4 != 0 || (reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(b) & 15))
        return SS2D_ESTRIDE;
    if (B > 65535 || K * 64 > 65535) return SS2D_EINVAL;
    g.B = B; g.K = K; g.D = D; g.R = R; g.L = L; g.sb = sb; g.sk = sk; g.sr = sr;
    g.dslices = 1; g.dper = (int)D;
    return 0;
}

template <int R> static int dt_fwd_t(const float *dtlr, const float *W, float *out, DtGeom g, cudaStream_t s) {
    const int64_t ltiles = (g.L / 4 + kDtThreads - 1) / kDtThreads;
    // slice D so that small problems still put >= ~4 CTAs on every SM
    int ds = 1;
    while (ds < 8 && ltiles * g.K * g.B * ds < 4 * 148 && g.D % (ds * 2) == 0) ds *= 2;
    g.dslices = ds; g.dper = (int)(g.D / ds);
    dt_proj_fwd_kernel<R><<<dim3((unsigned)ltiles, (unsigned)(g.K * ds), (unsigned)g.B), kDtThreads, g.dper * R * sizeof(float), s>>>(dtlr, W, out, g);
    return (int)cudaGetLastError();
}
template <int R> static int dt_bwd_t(const float *dout, const float *dtlr, const float *W, float *dx, float *dW, DtGeom g, cudaStream_t s) {
    const int64_t ltiles = (g.L / 4 + kDtThreads - 1) / kDtThreads;
    dt_proj_bwd_x_kernel<R><<<dim3((unsigned)ltiles, (unsigned)g.K, (unsigned)g.B), kDtThreads, g.D * R * sizeof(float), s>>>(dout, W, dx, g);
    const int64_t units = g.B * g.K * ((g.D + kDtRows - 1) / kDtRows), wpb = kDtThreads / kWarp;
    dt_proj_bwd_w_kernel<R><<<(unsigned)((units + wpb - 1) / wpb), kDtThreads, 0, s>>>(dout, dtlr, dW, g);
    return (int)cudaGetLastError();
}

}  // namespace ss2d

extern "C" int ss2d_dt_proj_fwd(const float *dtlr, int64_t sb, int64_t sk, int64_t sr, const float *W, float *out, int64_t B, int64_t K,
                                int64_t D, int64_t R, int64_t L, void *stream) {
    using namespace ss2d;
    if (!dtlr || !W || !out) return SS2D_EINVAL;
    DtGeom g;
    if (int rc = dt_geom(g, B, K, D, R, L, sb, sk, sr, dtlr, out)) return rc;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
#define SS2D_DT_FWD(r) dt_fwd_t<r>(dtlr, W, out, g, s)
    switch (R) {
        case 1: return SS2D_DT_FWD(1); case 2: return SS2D_DT_FWD(2); case 3: return SS2D_DT_FWD(3); case 4: return SS2D_DT_FWD(4);
        case 5: return SS2D_DT_FWD(5); case 6: return SS2D_DT_FWD(6); case 7: return SS2D_DT_FWD(7); default: return SS2D_DT_FWD(8);
    }
}

extern "C" int ss2d_dt_proj_bwd(const float *dout, const float *dtlr, int64_t sb, int64_t sk, int64_t sr, const float *W, float *d_dtlr,
                                float *dW, int64_t B, int64_t K, int64_t D, int64_t R, int64_t L, void *stream) {
    using namespace ss2d;
    if (!dout || !dtlr || !W || !d_dtlr || !dW) return SS2D_EINVAL;
    DtGeom g;
    if (int rc = dt_geom(g, B, K, D, R, L, sb, sk, sr, dtlr, dout)) return rc;
    if (reinterpret_cast<uintptr_t>(d_dtlr) & 15) return SS2D_ESTRIDE;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
#define SS2D_DT_BWD(r) dt_bwd_t<r>(dout, dtlr, W, d_dtlr, dW, g, s)
    switch (R) {
        case 1: return SS2D_DT_BWD(1); case 2: return SS2D_DT_BWD(2); case 3: return SS2D_DT_BWD(3); case 4: return SS2D_DT_BWD(4);
        case 5: return SS2D_DT_BWD(5); case 6: return SS2D_DT_BWD(6); case 7: return SS2D_DT_BWD(7); default: return SS2D_DT_BWD(8);
    }
}
