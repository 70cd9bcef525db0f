// ss2d_dwconv.cu — depthwise 3x3 conv (pad 1) + bias + SiLU, channels-last in -> channels-first out (sm_100a).
//
// Replaces three separate passes of SS2D.forwardv2 (reference: ITS/models/vmamba_layers.py:591-594 with the conv
// of :460-469): `x.permute(0,3,1,2).contiguous()` (ATen copy), `self.conv2d(x)` (cuDNN depthwise) and
// `self.act(x)` (SiLU) — one kernel reads the in_proj output's x-half in its native (B,H,W,2*d_inner) layout and
// writes (B,d_inner,H,W).  HBM-bound: algorithmic bytes = 4*B*H*W*C read + 4*B*C*H*W written (the 3x3 halo
// re-reads hit L2).  Lanes walk channels on the read side and pixels on the write side; the 32x32
// (channel, pixel) tile is turned through shared memory so both sides are 128-byte coalesced.
//
// Backward: pass 1 recomputes s = conv+bias and writes dpre = dout * silu'(s) channels-last (scratch);
// pass 2 produces dxin (correlation with the flipped taps) and reduces dweight / dbias per block before
// one atomicAdd per (channel, tap).
#include "ss2d_common.cuh"
#include "../../include/ss2d_b200.h"

namespace ss2d {

constexpr int kCw = 32;   // channels per block (lanes on the channels-last side)
constexpr int kPw = 32;   // pixels (along W) per block
constexpr int kTy = 8;

struct DwGeom {
    int B, C, H, W, tiles_w, tiles_c;
    int64_t cstride;
};

__device__ __forceinline__ void dw_decode(const DwGeom &g, int &b, int &h, int &w0, int &c0) {
    int id = blockIdx.x;
    c0 = (id % g.tiles_c) * kCw; id /= g.tiles_c;
    w0 = (id % g.tiles_w) * kPw; id /= g.tiles_w;
    h = id % g.H; b = id / g.H;
}

// Stage rows h-1..h+1, pixels w0-1..w0+kPw, channels c0..c0+31 of a channels-last tensor (zero padded).
__device__ __forceinline__ void dw_stage(float (*tile)[kPw + 2][kCw], const float *__restrict__ src, int64_t cstride,
                                         const DwGeom &g, int b, int h, int w0, int c0) {
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int c = c0 + tx;
    constexpr int kIt = (kPw + 2 + kTy - 1) / kTy;
    float v[3][kIt];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int hh = h - 1 + r;
#pragma unroll
        for (int k = 0; k < kIt; ++k) {
            const int wi = ty + k * kTy, ww = w0 - 1 + wi;
            v[r][k] = 0.f;
            if (wi < kPw + 2 && hh >= 0 && hh < g.H && ww >= 0 && ww < g.W && c < g.C)
                v[r][k] = __ldg(src + (((int64_t)b * g.H + hh) * g.W + ww) * cstride + c);
        }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int k = 0; k < kIt; ++k)
            if (ty + k * kTy < kPw + 2) tile[r][ty + k * kTy][tx] = v[r][k];
}

// Forward: one block = 32 channels x 32 pixels x kTh rows.  The kTh+2 input rows are staged once (row re-read
// factor (kTh+2)/kTh instead of 3), every thread produces 4*kTh outputs, and the (channel, pixel) tiles are turned
// through shared memory so that the channels-first stores are 128-byte coalesced.
constexpr int kTh = 4;
__global__ void __launch_bounds__(kCw *kTy) dwconv_silu_fwd_kernel(const float *__restrict__ xin, const float *__restrict__ weight,
                                                                  const float *__restrict__ bias, float *__restrict__ out,
                                                                  const DwGeom g) {
    __shared__ float tin[kTh + 2][kPw + 2][kCw];
    __shared__ float tout[kTh][kCw][kPw + 1];
    int id = blockIdx.x;
    const int c0 = (id % g.tiles_c) * kCw; id /= g.tiles_c;
    const int w0 = (id % g.tiles_w) * kPw; id /= g.tiles_w;
    const int tiles_h = (g.H + kTh - 1) / kTh;
    const int h0 = (id % tiles_h) * kTh, b = id / tiles_h;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int c = c0 + tx;
    float wgt[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) wgt[q] = c < g.C ? weight[c * 9 + q] : 0.f;
    const float bv = (bias && c < g.C) ? bias[c] : 0.f;
    {   // all (kTh+2) x 5 loads of a thread are issued before the first use (memory-level parallelism)
        constexpr int kIt = (kPw + 2 + kTy - 1) / kTy;
        float v[kTh + 2][kIt];
#pragma unroll
        for (int r = 0; r < kTh + 2; ++r) {
            const int hh = h0 - 1 + r;
#pragma unroll
            for (int k = 0; k < kIt; ++k) {
                const int wi = ty + k * kTy, ww = w0 - 1 + wi;
                v[r][k] = 0.f;
                if (wi < kPw + 2 && hh >= 0 && hh < g.H && ww >= 0 && ww < g.W && c < g.C)
                    v[r][k] = __ldg(xin + (((int64_t)b * g.H + hh) * g.W + ww) * g.cstride + c);
            }
        }
#pragma unroll
        for (int r = 0; r < kTh + 2; ++r)
#pragma unroll
            for (int k = 0; k < kIt; ++k)
                if (ty + k * kTy < kPw + 2) tin[r][ty + k * kTy][tx] = v[r][k];
    }
    __syncthreads();
#pragma unroll
    for (int hr = 0; hr < kTh; ++hr) {
#pragma unroll
        for (int k = 0; k < kPw / kTy; ++k) {
            const int wi = ty + k * kTy;
            float s = bv;
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int j = 0; j < 3; ++j) s = fmaf(wgt[r * 3 + j], tin[hr + r][wi + j][tx], s);
            tout[hr][tx][wi] = s * sigmoidf_fast(s);
        }
    }
    __syncthreads();
#pragma unroll
    for (int hr = 0; hr < kTh; ++hr) {
        const int h = h0 + hr;
#pragma unroll
        for (int k = 0; k < kCw / kTy; ++k) {  // lanes now walk pixels: contiguous in (B,C,H,W)
            const int cc = ty + k * kTy, w = w0 + tx;
            if (h < g.H && c0 + cc < g.C && w < g.W) out[(((int64_t)b * g.C + c0 + cc) * g.H + h) * g.W + w] = tout[hr][cc][tx];
        }
    }
}

// pass 1 of the backward: dpre (channels-last, dense C) = dout * silu'(conv + bias)
__global__ void __launch_bounds__(kCw *kTy) dwconv_silu_dpre_kernel(const float *__restrict__ xin, const float *__restrict__ weight,
                                                                   const float *__restrict__ bias, const float *__restrict__ dout,
                                                                   float *__restrict__ dpre, const DwGeom g) {
    __shared__ float tin[3][kPw + 2][kCw];
    __shared__ float tg[kCw][kPw + 1];
    int b, h, w0, c0;
    dw_decode(g, b, h, w0, c0);
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int c = c0 + tx;
    float wgt[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) wgt[q] = c < g.C ? weight[c * 9 + q] : 0.f;
    const float bv = (bias && c < g.C) ? bias[c] : 0.f;
    dw_stage(tin, xin, g.cstride, g, b, h, w0, c0);
    for (int cc = ty; cc < kCw; cc += kTy) {
        const int w = w0 + tx;
        tg[cc][tx] = (c0 + cc < g.C && w < g.W) ? __ldg(dout + (((int64_t)b * g.C + c0 + cc) * g.H + h) * g.W + w) : 0.f;
    }
    __syncthreads();
    for (int wi = ty; wi < kPw; wi += kTy) {
        float s = bv;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int j = 0; j < 3; ++j) s = fmaf(wgt[r * 3 + j], tin[r][wi + j][tx], s);
        const float sg = sigmoidf_fast(s);
        const int w = w0 + wi;
        if (c < g.C && w < g.W) dpre[(((int64_t)b * g.H + h) * g.W + w) * g.C + c] = tg[tx][wi] * sg * (1.f + s * (1.f - sg));
    }
}

// pass 2: dxin = corr(dpre, flipped taps); dweight[c][tap] += sum dpre * xin(shifted); dbias[c] += sum dpre.
// One block walks kGradRows rows x all W tiles of its 32-channel slab, so the dweight / dbias partial sums stay in
// registers across ~32 tiles and only one atomicAdd per (channel, tap) per block reaches L2 (the first version
// issued one per tile: 31 M atomics onto 1920 addresses at the training shape).
constexpr int kGradRows = 8;
__global__ void __launch_bounds__(kCw *kTy) dwconv_silu_grad_kernel(const float *__restrict__ xin, const float *__restrict__ weight,
                                                                   const float *__restrict__ dpre, float *__restrict__ dxin,
                                                                   int64_t dx_cstride, float *__restrict__ dweight,
                                                                   float *__restrict__ dbias, const DwGeom g) {
    __shared__ float tin[3][kPw + 2][kCw];
    __shared__ float tdp[3][kPw + 2][kCw];
    __shared__ float red[kTy][10][kCw];
    int id = blockIdx.x;
    const int c0 = (id % g.tiles_c) * kCw; id /= g.tiles_c;
    const int hblocks = (g.H + kGradRows - 1) / kGradRows;
    const int hb = id % hblocks, b = id / hblocks;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int c = c0 + tx;
    float wgt[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) wgt[q] = c < g.C ? weight[c * 9 + q] : 0.f;
    float acc[10];
#pragma unroll
    for (int q = 0; q < 10; ++q) acc[q] = 0.f;
    const int h_end = min(g.H, (hb + 1) * kGradRows);
    for (int h = hb * kGradRows; h < h_end; ++h) {
        for (int wt = 0; wt < g.tiles_w; ++wt) {
            const int w0 = wt * kPw;
            __syncthreads();  // previous tile fully consumed
            dw_stage(tin, xin, g.cstride, g, b, h, w0, c0);
            dw_stage(tdp, dpre, g.C, g, b, h, w0, c0);
            __syncthreads();
#pragma unroll
            for (int k = 0; k < kPw / kTy; ++k) {
                const int wi = ty + k * kTy, w = w0 + wi;
                // input pixel (h, w) was read by output pixel (h+1-r, w+1-j) through tap (r, j)
                float s = 0.f;
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int j = 0; j < 3; ++j) s = fmaf(wgt[r * 3 + j], tdp[2 - r][wi + 2 - j][tx], s);
                if (c < g.C && w < g.W) dxin[(((int64_t)b * g.H + h) * g.W + w) * dx_cstride + c] = s;
                const float gp = tdp[1][wi + 1][tx];  // dpre at output pixel (h, w); zero outside the image
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int j = 0; j < 3; ++j) acc[r * 3 + j] = fmaf(gp, tin[r][wi + j][tx], acc[r * 3 + j]);
                acc[9] += gp;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 10; ++q) red[ty][q][tx] = acc[q];
    __syncthreads();
    for (int q = ty; q < 10; q += kTy) {
        float s = 0.f;
#pragma unroll
        for (int t = 0; t < kTy; ++t) s += red[t][q][tx];
        if (c < g.C) {
            if (q < 9) atomicAdd(dweight + c * 9 + q, s);
            else if (dbias) atomicAdd(dbias + c, s);
        }
    }
}

static int dw_geom(DwGeom &g, int64_t cstride, int64_t B, int64_t C, int64_t H, int64_t W) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || cstride < C) return SS2D_EINVAL;
    g.B = (int)B; g.C = (int)C; g.H = (int)H; g.W = (int)W; g.cstride = cstride;
    g.tiles_w = (int)((W + kPw - 1) / kPw); g.tiles_c = (int)((C + kCw - 1) / kCw);
    if ((int64_t)g.tiles_w * g.tiles_c * H * B > 0x7fffffffLL) return SS2D_EINVAL;
    return 0;
}

}  // namespace ss2d

extern "C" int ss2d_dwconv_silu_fwd(const float *xin, int64_t cstride, const float *weight, const float *bias, float *out,
                                    int64_t batch, int64_t C, int64_t H, int64_t W, void *stream) {
    using namespace ss2d;
    if (!xin || !weight || !out) return SS2D_EINVAL;
    DwGeom g;
    if (int rc = dw_geom(g, cstride, batch, C, H, W)) return rc;
    const unsigned grid = (unsigned)((int64_t)g.tiles_w * g.tiles_c * ((H + kTh - 1) / kTh) * batch);
    dwconv_silu_fwd_kernel<<<grid, dim3(kCw, kTy), 0, reinterpret_cast<cudaStream_t>(stream)>>>(xin, weight, bias, out, g);
    return (int)cudaGetLastError();
}

extern "C" int ss2d_dwconv_silu_bwd(const float *xin, int64_t cstride, const float *weight, const float *bias, const float *dout,
                                    float *dpre_scratch, float *dxin, int64_t dx_cstride, float *dweight, float *dbias,
                                    int64_t batch, int64_t C, int64_t H, int64_t W, void *stream) {
    using namespace ss2d;
    if (!xin || !weight || !dout || !dpre_scratch || !dxin || !dweight || dx_cstride < C) return SS2D_EINVAL;
    DwGeom g;
    if (int rc = dw_geom(g, cstride, batch, C, H, W)) return rc;
    const unsigned grid = (unsigned)((int64_t)g.tiles_w * g.tiles_c * H * batch);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    dwconv_silu_dpre_kernel<<<grid, dim3(kCw, kTy), 0, s>>>(xin, weight, bias, dout, dpre_scratch, g);
    const unsigned grid2 = (unsigned)((int64_t)g.tiles_c * ((H + kGradRows - 1) / kGradRows) * batch);
    dwconv_silu_grad_kernel<<<grid2, dim3(kCw, kTy), 0, s>>>(xin, weight, dpre_scratch, dxin, dx_cstride, dweight, dbias, g);
    return (int)cudaGetLastError();
}
